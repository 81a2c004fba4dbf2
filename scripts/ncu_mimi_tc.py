"""Profiling driver for the tensor-core codec path: warm-up, then ONE encode + decode of N clips x S seconds between
cudaProfilerStart/Stop (run ncu with --profile-from-start off).  Usage: python scripts/ncu_mimi_tc.py [clips] [seconds]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from csm_mlx_b200.mimi import Mimi
from csm_mlx_b200.random_init import random_mimi_weights
from tests.workloads import synthetic_audio

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
dev = torch.device("cuda", 0)
mimi = Mimi(32, device=dev).load_pytorch_weights(random_mimi_weights())
clips = torch.stack([synthetic_audio(100 + i, secs) for i in range(n)])[:, None].to(dev)
codes = mimi.encode(clips)
mimi.decode(codes)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
codes = mimi.encode(clips)
audio = mimi.decode(codes)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok", float(audio.abs().mean()))

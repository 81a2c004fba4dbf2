"""Profiling driver for the codec (BASELINE.json configs[4] shape, scaled to one GPU): encode + decode of B clips of
S seconds.  Usage: python scripts/ncu_mimi.py [B] [seconds]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from csm_mlx_b200.mimi import Mimi
from csm_mlx_b200.random_init import random_mimi_weights
from tests.workloads import synthetic_audio

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
S = float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
dev = torch.device("cuda", 0)
mimi = Mimi(32, device=dev).load_pytorch_weights(random_mimi_weights())
clips = torch.stack([synthetic_audio(100 + i, S) for i in range(B)])[:, None].to(dev)
codes = mimi.encode(clips)
audio = mimi.decode(codes)
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
codes = mimi.encode(clips)
e1.record()
audio = mimi.decode(codes)
e2.record()
torch.cuda.synchronize()
print(f"B={B} S={S}: encode {B * S / (e0.elapsed_time(e1) / 1e3):.0f} audio-s/s, decode {B * S / (e1.elapsed_time(e2) / 1e3):.0f} audio-s/s, "
      f"codes {tuple(codes.shape)} audio {tuple(audio.shape)}")

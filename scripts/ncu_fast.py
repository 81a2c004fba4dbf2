"""Profiling driver for the fused batched frame chain (csrc/batch_frame.cu): B sequences, prefill + first frame on the
per-op path, then N fast frames without a CUDA graph so that ncu sees the individual launches.
Usage: python scripts/ncu_fast.py [B] [frames]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from csm_mlx_b200 import CSM, csm_1b, tokenizers
from csm_mlx_b200.random_init import random_csm_weights
from csm_mlx_b200.runtime import LMState, SamplerSpec
from tests.workloads import prompt_ids

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
model = CSM(csm_1b(), device=dev).load_weights(random_csm_weights())
spec = SamplerSpec()
prompts = [tokenizers.tokenize_text_segment(prompt_ids(21 + i, 8 + i % 9), 0) for i in range(B)]
st = LMState(model, B, max_len=64)
st.prefill([p[0] for p in prompts], [p[1] for p in prompts])
frame = torch.zeros((B, 32), device=dev, dtype=torch.int32)
st.sample_c0(frame, spec)
st.depth_decode(frame, spec)
prev = frame
for _ in range(N):
    nxt = torch.zeros((B, 32), device=dev, dtype=torch.int32)
    st.decode_frame(prev, nxt, spec)
    prev = nxt
torch.cuda.synchronize()
st.check_status()
print("ok", prev.sum().item())

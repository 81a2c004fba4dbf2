"""Profiling driver for the prompt-row path (csmb_prefill_fast): prefill of a T-row prompt, CUDA-event timing, and one
profiled pass between cudaProfilerStart/Stop.  Usage: python scripts/ncu_prefill.py [rows]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from csm_mlx_b200 import CSM, csm_1b
from csm_mlx_b200.random_init import random_csm_weights
from csm_mlx_b200.runtime import LMState
from oracle import lm as olm
from tests.workloads import prompt_ids

T = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda", 0)
model = CSM(csm_1b(), device=dev).load_weights(random_csm_weights())
gen = torch.Generator().manual_seed(5)
if T <= 24:
    tok, mask = olm.text_rows(prompt_ids(3, T - 2))
else:
    t1 = olm.text_rows(prompt_ids(3, 10))
    a1 = olm.audio_rows(torch.randint(0, 2048, (32, T - 13), generator=gen))
    tok, mask = torch.cat([t1[0], a1[0]]), torch.cat([t1[1], a1[1]])
tok, mask = tok.int(), mask
st = LMState(model, 1, max_len=T + 8)
staged = st.stage_prefill([tok], [mask])
for _ in range(3):
    st.reset()
    st.run_prefill(staged)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    st.reset()
    st.run_prefill(staged)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print("checksum h_last / c0_logits:", st.h_last.double().sum().item().hex(), st.c0_logits.double().sum().item().hex(), flush=True)
print(f"prefill rows={int(tok.shape[0])}: {ms:.3f} ms (GPU, staged inputs)  {2 * 973.1e6 * tok.shape[0] / (ms * 1e-3) / 1e12:.1f} TFLOP/s", flush=True)
torch.cuda.cudart().cudaProfilerStart()
st.reset()
st.run_prefill(staged)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
st.check_status()

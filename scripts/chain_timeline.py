"""Developer tool (run under gpurun): in-kernel timeline of one chain frame-step.  Needs the debug build
  CSMB_OUT=$PWD/csm_mlx_b200/libcsm_b200_tl.so CSMB_OBJ_DIR=$PWD/csm_mlx_b200/csrc/.obj_tl bash csm_mlx_b200/csrc/build.sh -DCSMB_TIMELINE
Block (0,0,0) of every chain kernel records %globaltimer at entry, after griddepcontrol.wait, at a mid point (Linears:
accumulator complete; attention: q/k/v staged; sampler: logits summed) and at its end.  Per kernel class this prints
  lead   = wait returned - entry            (how long the CTA was resident before its producer finished)
  body   = end - wait returned              (the kernel's own dependent path, block 0)
  first  = (Linears) first pipeline stage complete - wait returned
  mid    = mid point - wait returned
  gap    = next kernel's wait returned - this kernel's end   (tail of the other CTAs + the kernel boundary)
  period = next kernel's wait returned - this kernel's wait returned
Usage: python scripts/chain_timeline.py B [steps]"""
import ctypes as C
import os
import sys
from collections import defaultdict

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["CSMB_LIB_PATH"] = os.environ.get("TL_LIB") or os.path.join(HERE, "csm_mlx_b200", "libcsm_b200_tl.so")
sys.path.insert(0, HERE)
import numpy as np
import torch

from csm_mlx_b200 import CSM, csm_1b, tokenizers, _lib
from csm_mlx_b200.random_init import random_csm_weights
from csm_mlx_b200.runtime import LMState, SamplerSpec
from tests.workloads import prompt_ids

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda", 0)
model = CSM(csm_1b(), device=dev).load_weights(random_csm_weights())
spec = SamplerSpec()
prompts = [tokenizers.tokenize_text_segment(prompt_ids(21 + i, 8 + i % 9), 0) for i in range(B)]
st = LMState(model, B, max_len=64, row_invariant=(B == 1))
st.prefill([p[0] for p in prompts], [p[1] for p in prompts])
frame = torch.zeros((B, 32), device=dev, dtype=torch.int32)
st.sample_c0(frame, spec)
st.depth_decode(frame, spec)
prev = frame
for _ in range(3):
    prev = st.decode_frame_graphed(prev, spec)
torch.cuda.synchronize()
lib = C.CDLL(os.environ["CSMB_LIB_PATH"])
lib.csmb_debug_timeline_read.argtypes = [C.c_void_p, C.c_int, C.c_int]
lib.csmb_debug_timeline_read(None, 0, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(STEPS):
    prev = st.decode_frame_graphed(prev, spec)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / STEPS
W = 6
buf = np.zeros((16384, W), dtype=np.uint64)
n = lib.csmb_debug_timeline_read(buf.ctypes.data, 16384, 1)
tl = buf[:n].astype(np.int64)
per_step = n // STEPS
print(f"B={B}: {ms:.3f} ms per frame-step (instrumented build), {n} kernels recorded = {per_step} per step")
names = {1: "linear", 2: "gate|up+swiglu", 3: "resid_norm", 4: "attn", 5: "sample_embed", 6: "frame_embed"}
# resolution of the timer
d = np.diff(np.sort(tl[:, 1:5][tl[:, 1:5] > 0].ravel()))
print("globaltimer: smallest non-zero step", int(d[d > 0].min()), "ns")
agg = defaultdict(lambda: defaultdict(list))
order = np.argsort(tl[:, 2], kind="stable")   # by wait-return time
tl = tl[order]
for i in range(n - 1):
    kind, grid = int(tl[i, 0] & 0xff), int(tl[i, 0] >> 8)
    idx_in_step = i % per_step
    stack = "backbone" if idx_in_step < 1 + 16 * 7 + 2 else "decoder"
    key = (stack, names.get(kind, str(kind)), grid)
    t_in, t_w, t_m, t_e = tl[i, 1], tl[i, 2], tl[i, 3], tl[i, 4]
    nxt_w = tl[i + 1, 2]
    a = agg[key]
    a["lead"].append(t_w - t_in)
    a["body"].append(t_e - t_w)
    if t_m > 0:
        a["mid"].append(t_m - t_w)
    if tl[i, 5] > 0:
        a["first"].append(tl[i, 5] - t_w)
    a["gap"].append(nxt_w - t_e)
    a["period"].append(nxt_w - t_w)
print(f"{'stack':9s} {'kernel':16s} {'grid':>5s} {'count/step':>10s} {'lead':>7s} {'first':>7s} {'mid':>7s} {'body':>7s} {'gap':>7s} {'period':>7s} {'us/step':>8s}")
tot = 0.0
for key in sorted(agg, key=lambda k: -np.sum(agg[k]["period"])):
    a = agg[key]
    cnt = len(a["period"]) / STEPS
    mean = lambda x: float(np.mean(x)) / 1e3 if len(x) else float("nan")
    per = float(np.sum(a["period"])) / STEPS / 1e3
    tot += per
    print(f"{key[0]:9s} {key[1]:16s} {key[2]:5d} {cnt:10.1f} {mean(a['lead']):7.2f} {mean(a['first']):7.2f} {mean(a['mid']):7.2f} {mean(a['body']):7.2f} "
          f"{mean(a['gap']):7.2f} {mean(a['period']):7.2f} {per:8.1f}")
print(f"sum of periods: {tot / 1e3:.3f} ms per step")

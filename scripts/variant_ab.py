"""A/B of frame-kernel build variants: `python scripts/variant_ab.py a.so b.so ...` runs, per library (own process,
CSMB_LIB_PATH), the batch-1 frame loop of BASELINE configs[0]/[1] (greedy, short context), checks the first 25 frames
against tests/golden/cfg1_lm.npz and prints the device-timed ms per frame (3 repetitions of 100 frames)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, os, numpy as np, torch
sys.path.insert(0, %r)
from csm_mlx_b200 import CSM, csm_1b, tokenizers
from csm_mlx_b200.random_init import random_csm_weights
from csm_mlx_b200.runtime import LMState, SamplerSpec
from tests.workloads import cfg1_prompt_ids
dev = torch.device("cuda", 0)
model = CSM(csm_1b(), device=dev).load_weights(random_csm_weights())
spec = SamplerSpec()
tok, mask = tokenizers.tokenize_text_segment(cfg1_prompt_ids(), 0)
gold = np.load(os.path.join(%r, "tests", "golden", "cfg1_lm.npz"))["tokens"]
st = LMState(model, 1, max_len=512)
st.prefill([tok], [mask])
frames = [st.first_frame_fused(spec)]
for _ in range(24):
    frames.append(st.decode_frame_fused(frames[-1], spec))
torch.cuda.synchronize(); st.check_status()
got = torch.cat(frames).cpu().numpy()
ok = bool((got == gold[:25]).all())
res = []
frame = frames[-1]
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100):
        frame = st.decode_frame_fused(frame, spec)
    e1.record(); torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / 100)
st.check_status()
print("RESULT", os.environ.get("CSMB_LIB_PATH"), "golden25", ok, " ".join(f"{r:.4f}" for r in res), flush=True)
''' % (ROOT, ROOT)

for lib in sys.argv[1:]:
    env = dict(os.environ, CSMB_LIB_PATH=os.path.abspath(lib))
    try:  # a variant that hangs must not eat the GPU budget: the frame kernel's own bounded waits give up after ~2 s
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=150)
    except subprocess.TimeoutExpired:
        print(f"TIMEOUT {lib}", flush=True)
        continue
    out = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
    print(out[0] if out else f"FAILED {lib}: {r.stderr[-800:]}", flush=True)

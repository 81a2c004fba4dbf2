"""Developer timing script (run under gpurun): the persistent frame kernel on a weight-only FP8 model, with the decoder's ring
copies asking L2 to keep k/8 of their lines (csmb_frame_opts.flags bits 4..6) — does the 115 MB decoder stay in the 126 MB L2
across the 31 depth steps?  Usage: python scripts/fp8_frame_trial.py [flags ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from csm_mlx_b200 import CSM, csm_1b, quantize, tokenizers
from csm_mlx_b200.random_init import random_csm_weights
from csm_mlx_b200.runtime import LMState, SamplerSpec
from tests.workloads import cfg1_prompt_ids

dev = torch.device("cuda", 0)
which = os.environ.get("FP8_TRIAL_MODEL", "e4m3")
model = CSM(csm_1b(), device=dev).load_weights(random_csm_weights())
if which == "e4m3":
    quantize(model)
spec = SamplerSpec()
tok, mask = tokenizers.tokenize_text_segment(cfg1_prompt_ids(), 0)
ref = None
for flags in [int(x, 0) for x in (sys.argv[1:] or ["0", "0x40", "0x60", "0x70"])]:
    os.environ["CSMB_FRAME_FLAGS"] = str(flags)
    st = LMState(model, 1, max_len=200)
    st.prefill([tok], [mask])
    f = st.first_frame_fused(spec)
    frames = [f.clone()]
    for _ in range(5):
        f = st.decode_frame_fused(f, spec)
        frames.append(f.clone())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = int(os.environ.get("FP8_TRIAL_FRAMES", "40"))
    for _ in range(n):
        f = st.decode_frame_fused(f, spec)
    e1.record()
    torch.cuda.synchronize()
    st.check_status()
    toks = torch.cat(frames).cpu()
    if ref is None:
        ref = toks
    print(f"{which} k_frame flags={flags:#x}: {e0.elapsed_time(e1) / n:.3f} ms per frame, tokens_same={bool(torch.equal(ref, toks))}", flush=True)

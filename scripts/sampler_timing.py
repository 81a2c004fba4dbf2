"""ms per fused frame (k_frame, batch 1) for the fused samplers: greedy, temperature, top-k, top-p, min-p, all three."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from csm_mlx_b200 import CSM, csm_1b, tokenizers
from csm_mlx_b200.random_init import random_csm_weights
from csm_mlx_b200.runtime import LMState, SamplerSpec
from tests.workloads import cfg1_prompt_ids

dev = torch.device("cuda", 0)
model = CSM(csm_1b(), device=dev).load_weights(random_csm_weights())
tok, mask = tokenizers.tokenize_text_segment(cfg1_prompt_ids(), 0)
for name, spec in (("greedy", SamplerSpec()), ("temperature 0.8", SamplerSpec(temperature=0.8, seed=1)),
                   ("temp 0.8 top-k 50 (README)", SamplerSpec(temperature=0.8, top_k=50, seed=1)),
                   ("temp 0.8 top-p 0.9", SamplerSpec(temperature=0.8, top_p=0.9, seed=1)),
                   ("temp 0.8 min-p 0.05", SamplerSpec(temperature=0.8, min_p=0.05, seed=1)),
                   ("temp 0.8 top-k 50 top-p 0.9 min-p 0.02", SamplerSpec(temperature=0.8, top_k=50, top_p=0.9, min_p=0.02, seed=1))):
    st = LMState(model, 1, max_len=256)
    assert st.fused_supported(spec), name
    st.prefill([tok], [mask])
    frame = st.first_frame_fused(spec)
    for _ in range(5):
        frame = st.decode_frame_fused(frame, spec)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(60):
        frame = st.decode_frame_fused(frame, spec)
    e1.record(); torch.cuda.synchronize()
    st.check_status()
    print(f"{name:42s} {e0.elapsed_time(e1) / 60:.3f} ms/frame", flush=True)

"""Minimal driver for profiling the persistent frame kernel under ncu: prefill + N fused frames of csm_1b."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from csm_mlx_b200 import CSM, csm_1b, tokenizers
from csm_mlx_b200.random_init import random_csm_weights
from csm_mlx_b200.runtime import LMState, SamplerSpec
from tests.workloads import cfg1_prompt_ids

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
model = CSM(csm_1b(), device=dev).load_weights(random_csm_weights())
tok, mask = tokenizers.tokenize_text_segment(cfg1_prompt_ids(), 0)
st = LMState(model, 1, max_len=256)
st.prefill([tok], [mask])
spec = SamplerSpec()
frame = torch.zeros((1, 32), device=dev, dtype=torch.int32)
st.sample_c0(frame, spec); st.depth_decode(frame, spec)
for _ in range(n):
    frame = st.decode_frame_fused(frame, spec)
torch.cuda.synchronize()
print("frames ok", frame[0, :4].tolist(), "status", int(st.frame_status.item()))

import sys, os, time, torch
sys.path.insert(0, "/root/repo")
from csm_mlx_b200 import CSM, csm_1b
from csm_mlx_b200.random_init import random_csm_weights
from csm_mlx_b200.runtime import LMState, SamplerSpec
from oracle import lm as olm
from tests.workloads import prompt_ids
dev = torch.device("cuda", 0)
model = CSM(csm_1b(), device=dev).load_weights(random_csm_weights())
spec = SamplerSpec()
for n_audio in (0, 150, 380, 1180):
    gen = torch.Generator().manual_seed(5)
    t1 = olm.text_rows(prompt_ids(3, 8))
    parts_t, parts_m = [t1[0]], [t1[1]]
    if n_audio:
        a1 = olm.audio_rows(torch.randint(0, 2048, (32, n_audio), generator=gen))
        parts_t.append(a1[0]); parts_m.append(a1[1])
    tok, mask = torch.cat(parts_t).int(), torch.cat(parts_m)
    st = LMState(model, 1, max_len=tok.shape[0] + 64)
    st.prefill([tok], [mask])
    frame = st.first_frame_fused(spec)
    for _ in range(3):
        frame = st.decode_frame_fused(frame, spec)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        frame = st.decode_frame_fused(frame, spec)
    e1.record(); torch.cuda.synchronize()
    st.check_status()
    print(f"S~{tok.shape[0] + 20}: {e0.elapsed_time(e1) / 30:.3f} ms/frame  tokens {frame[0, :4].tolist()}", flush=True)

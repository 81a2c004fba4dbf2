"""Developer bring-up script (run under gpurun): exercises every product path once against the oracle and prints
max errors / timings.  Not a test; tests/ holds the assertions."""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from csm_mlx_b200 import CSM, csm_1b
from csm_mlx_b200 import generation, tokenizers
from csm_mlx_b200.mimi import Mimi
from csm_mlx_b200.random_init import random_csm_weights, random_mimi_weights
from csm_mlx_b200.runtime import LMState, SamplerSpec
from oracle import lm as olm, mimi as omimi

torch.manual_seed(0)
dev = torch.device("cuda", 0)
SECTIONS = sys.argv[1:] or ["ops", "lm", "mimi", "time"]


def section(name):
    def deco(fn):
        if name in SECTIONS:
            print(f"\n===== {name} =====", flush=True)
            try:
                fn()
            except Exception:
                traceback.print_exc()
        return fn
    return deco


t0 = time.time()
W = random_csm_weights()
print("weights", time.time() - t0, flush=True)
model = CSM(csm_1b(), device=dev).load_weights(W)
torch.cuda.synchronize()
print("loaded", time.time() - t0, flush=True)
g = torch.Generator().manual_seed(7)
ids = [128000] + torch.randint(0, 128000, (8,), generator=g).tolist() + [128001]
ptok, pmask = tokenizers.tokenize_text_segment(ids, 0)


@section("ops")
def _ops():
    from csm_mlx_b200 import _lib
    lib = _lib.lib()
    for (R, N, K) in [(1, 3072, 2048), (1, 2048, 2048), (1, 16384, 2048), (1, 2048, 8192), (1, 2051, 2048), (2, 1536, 1024),
                      (1, 1024, 1024), (1, 16384, 1024), (1, 1024, 8192), (2, 1024, 2048), (1, 2051, 1024), (10, 3072, 2048), (5, 2051, 1024)]:
        x = torch.randn(R, K, device=dev)
        w = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
        y = torch.zeros(R, N, device=dev)
        _lib.check(lib.csmb_linear(x.data_ptr(), K, w.data_ptr(), y.data_ptr(), N, R, N, K, 0, 0, _lib.stream_ptr(dev)))
        ref = x.double() @ w.double().t()
        print(f"linear R={R} N={N} K={K} maxerr {(y.double() - ref).abs().max().item():.3e} ref {ref.abs().max().item():.2f}")
        # timing
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ws = [(torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16) for _ in range(max(2, int(400e6 / (N * K * 2))))]
        e0.record()
        for wi in ws:
            lib.csmb_linear(x.data_ptr(), K, wi.data_ptr(), y.data_ptr(), N, R, N, K, 0, 0, _lib.stream_ptr(dev))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / len(ws)
        print(f"   {ms * 1e3:.1f} us  {N * K * 2 / ms / 1e6:.0f} GB/s")


@section("lm")
def _lm():
    orc = olm.OracleCSM(olm.CSM_1B, W)
    st = LMState(model, 1, max_len=64)
    st.prefill([ptok], [pmask])
    torch.cuda.synchronize()
    ocache = orc.new_backbone_cache()
    forced = torch.randint(0, 2051, (1, 32), generator=g)
    tr = {}
    olm.generate_frame(orc, ptok[None].long(), pmask[None], ocache, forced=forced, trace=tr)
    print("h_last err", (st.h_last.cpu() - tr["h"]).abs().max().item(), "scale", tr["h"].abs().max().item())
    print("c0 logits err", (st.c0_logits.cpu() - tr["logits"][0]).abs().max().item(), "scale", tr["logits"][0].abs().max().item())
    frame = torch.zeros((1, 32), device=dev, dtype=torch.int32)
    lg = torch.zeros((1, 32, 2051), device=dev)
    fdev = forced.to(dev, torch.int32).contiguous()
    st.depth_decode(frame, SamplerSpec(), logits_out=lg, forced=fdev)
    torch.cuda.synchronize()
    errs = [(lg[0, i].cpu() - tr["logits"][i][0]).abs().max().item() for i in range(1, 32)]
    print("depth logits max err", max(errs), "frame==forced", bool((frame.cpu()[:, 1:] == forced[:, 1:].int()).all()))
    # free-running greedy
    t1 = time.time()
    (toks,) = generation.generate_tokens(model, [(ptok, pmask)], 6, temperature=0.0)
    torch.cuda.synchronize()
    print("gpu 6 frames", time.time() - t1)
    t1 = time.time()
    otoks = olm.generate_tokens(orc, ptok.long(), pmask, 6)
    print("oracle 6 frames", time.time() - t1)
    print("tokens equal", bool((toks.long() == otoks).all()), (toks.long() != otoks).sum().item())
    print(toks[:2, :8], otoks[:2, :8])


@section("mimi")
def _mimi():
    MW = random_mimi_weights()
    mimi = Mimi(32, device=dev).load_pytorch_weights(MW)
    codes = torch.randint(0, 2048, (1, 32, 25))
    a_or = omimi.decode(codes, MW)
    a = mimi.decode(codes.to(dev)).cpu()
    err = (a - a_or)
    print("decode shape", a.shape, "max err", err.abs().max().item(), "snr dB", (10 * torch.log10(a_or.pow(2).mean() / err.pow(2).mean())).item())
    # batch 2 + long (window)
    codes2 = torch.randint(0, 2048, (2, 32, 140))
    a2_or = omimi.decode(codes2, MW)
    a2 = mimi.decode(codes2.to(dev)).cpu()
    err = a2 - a2_or
    print("decode B2 F140 snr dB", (10 * torch.log10(a2_or.pow(2).mean() / err.pow(2).mean())).item())
    # streaming
    stream = mimi.new_decode_stream(1, use_graph=True)
    chunks = [stream.step(codes[:, :, i:i + 1].to(dev)).cpu().clone() for i in range(25)]
    s = torch.cat(chunks, -1)
    err = s - a_or
    print("stream snr dB", (10 * torch.log10(a_or.pow(2).mean() / err.pow(2).mean())).item(), "graph", stream.graph is not None)
    # encode
    t = torch.arange(24000 * 5) / 24000.
    gg = torch.Generator().manual_seed(11)
    audio = (0.3 * torch.sin(2 * torch.pi * 220 * t) + 0.2 * torch.sin(2 * torch.pi * 3300 * t) + 0.05 * torch.randn(t.shape, generator=gg)).clamp(-1, 1)
    for n in (120000, 120000 - 700):
        c_or = omimi.encode(audio[None, None, :n], MW)
        c = mimi.encode(audio[None, None, :n].to(dev)).cpu()
        print("encode", n, c.shape, c_or.shape, "agree", (c.long() == c_or).float().mean().item(), "cb0", (c[0, 0].long() == c_or[0, 0]).float().mean().item())
    tokenizers.set_audio_tokenizer(mimi)
    # timing
    for F in (25, 125):
        cc = torch.randint(0, 2048, (1, 32, F), device=dev)
        mimi.decode(cc); torch.cuda.synchronize()
        t1 = time.time(); mimi.decode(cc); torch.cuda.synchronize()
        print(f"decode F={F} {1e3 * (time.time() - t1):.2f} ms")
    t1 = time.time()
    for i in range(50):
        stream.step(codes[:, :, :1].to(dev))
    torch.cuda.synchronize()
    print(f"stream step {1e3 * (time.time() - t1) / 50:.3f} ms")
    x = audio[None, None].to(dev)
    mimi.encode(x); torch.cuda.synchronize()
    t1 = time.time(); mimi.encode(x); torch.cuda.synchronize()
    print(f"encode 5s {1e3 * (time.time() - t1):.2f} ms")


@section("time")
def _time():
    st = LMState(model, 1, max_len=256)
    st.prefill([ptok], [pmask])
    spec = SamplerSpec()
    frame = torch.zeros((1, 32), device=dev, dtype=torch.int32)
    st.sample_c0(frame, spec); st.depth_decode(frame, spec)
    prev = frame
    for _ in range(3):
        nf = torch.zeros_like(prev); st.decode_frame(prev, nf, spec); prev = nf
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t1 = time.time(); e0.record()
    n = 20
    for _ in range(n):
        nf = torch.zeros_like(prev); st.decode_frame(prev, nf, spec); prev = nf
    e1.record(); torch.cuda.synchronize()
    print(f"decode_frame eager: {e0.elapsed_time(e1) / n:.3f} ms/frame (wall {1e3 * (time.time() - t1) / n:.3f})")
    # CUDA graph of one frame
    sprev = prev.clone(); sout = torch.zeros_like(prev)
    gr = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        st.decode_frame(sprev, sout, spec)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(gr):
        st.decode_frame(sprev, sout, spec)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        gr.replay(); sprev.copy_(sout)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"decode_frame graph: {ms:.3f} ms/frame -> {0.08 / (ms / 1e3):.1f} audio-s/s, {9.1067e9 / ms / 1e6:.0f} GB/s algorithmic")


@section("fused")
def _fused():
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "cfg1_lm.npz"))
    st = LMState(model, 1, max_len=1024)
    st.prefill([ptok], [pmask])
    spec = SamplerSpec()
    frame = torch.zeros((1, 32), device=dev, dtype=torch.int32)
    st.sample_c0(frame, spec); st.depth_decode(frame, spec)
    frames = [frame]
    t1 = time.time()
    for i in range(24):
        frames.append(st.decode_frame_fused(frames[-1], spec))
    torch.cuda.synchronize()
    print("fused 24 frames wall", time.time() - t1, "status", int(st.frame_status.item()))
    toks = torch.cat(frames).cpu().numpy()
    eq = (toks == g["tokens"])
    print("tokens equal golden:", bool(eq.all()), "first mismatch frame", (None if eq.all() else int(np.argwhere(~eq)[0][0])), "cb", (None if eq.all() else int(np.argwhere(~eq)[0][1])))
    print(toks[1, :8], g["tokens"][1, :8])
    # timing
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prev = frames[-1]
    for _ in range(3):
        prev = st.decode_frame_fused(prev, spec)
    torch.cuda.synchronize()
    n = 40
    e0.record()
    for _ in range(n):
        prev = st.decode_frame_fused(prev, spec)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"fused frame: {ms:.3f} ms -> {0.08 / (ms / 1e3):.1f} audio-s/s, {9.1067e9 / ms / 1e6:.0f} GB/s algorithmic ({9.1067e9 / ms / 1e6 / 6557.8:.3f} of peak), status {int(st.frame_status.item())}")

    # in-kernel phase timers
    from csm_mlx_b200 import _lib
    prof = torch.zeros((148, 16), device=dev, dtype=torch.int64)
    names = ["poll", "load", "gemv", "fin", "datt", "batt", "sample", "merge", "wait"]

    def prof_run(tag):
        for _ in range(3):
            st.decode_frame_fused(frames[-1], spec)
        torch.cuda.synchronize()
        e0.record()
        pv = frames[-1]
        for _ in range(20):
            pv = st.decode_frame_fused(pv, spec)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        _lib.lib().csmb_debug_set_frame_prof(prof.data_ptr())
        st.decode_frame_fused(pv, spec)
        torch.cuda.synchronize()
        _lib.lib().csmb_debug_set_frame_prof(None)
        pr = prof.cpu().double() / 1.9e3
        print(f"{tag}: {ms:.3f} ms/frame ({9.1067e9 / ms / 1e6 / 6557.8:.3f} of peak) status {int(st.frame_status.item())} | "
              + ", ".join(f"{n} {pr[:, i].mean():.0f}" for i, n in enumerate(names)) + f" | total {pr[:, 12].mean():.0f} phases {int(prof[0, 13])}")
    for fl in (0, 1, 8):
        _lib.lib().csmb_debug_set_frame_flags(fl)
        prof_run(f"dbg flags {fl}")
    _lib.lib().csmb_debug_set_frame_flags(0)
    for (pm, pi) in ((16, 700), (24, 700), (32, 700), (24, 350), (24, 1400), (48, 500)):
        _lib.lib().csmb_debug_set_frame_prefetch(pm, pi)
        prof_run(f"prefetch {pm} stages / {pi} cyc")
        print("   prefetches issued per CTA:", float(prof[:, 14].double().mean()))
    _lib.lib().csmb_debug_set_frame_prefetch(0, 700)
    (toks2,) = generation.generate_tokens(model, [(ptok, pmask)], 25, temperature=0.0)
    print("generate_tokens (fused) equals golden:", bool((toks2.numpy() == g["tokens"]).all()))


@section("batch")
def _batch():
    from tests.workloads import prompt_ids
    for B in (16, 64):
        prompts = [tokenizers.tokenize_text_segment(prompt_ids(21 + i, 8 + i % 9), 0) for i in range(B)]
        st = LMState(model, B, max_len=64)
        t1 = time.time()
        st.prefill([p[0] for p in prompts], [p[1] for p in prompts])
        spec = SamplerSpec()
        frame = torch.zeros((B, 32), device=dev, dtype=torch.int32)
        st.sample_c0(frame, spec); st.depth_decode(frame, spec)
        torch.cuda.synchronize()
        print(f"B={B} prefill+first frame {1e3 * (time.time() - t1):.1f} ms")
        prev = frame
        for _ in range(3):
            prev = st.decode_frame_graphed(prev, spec)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            prev = st.decode_frame_graphed(prev, spec)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"B={B} decode frame (graph): {ms:.2f} ms -> {B * 0.08 / (ms / 1e3):.0f} audio-s/s, {9.1067e9 / ms / 1e6 / 6557.8:.3f} of HBM roofline")


@section("fast")
def _fast():
    """Fused batched frame chain (csrc/batch_frame.cu) vs the per-op batched path: tokens and timing."""
    from csm_mlx_b200 import _lib
    from tests.workloads import prompt_ids
    spec = SamplerSpec()
    for B in (3, 16, 64):
        prompts = [tokenizers.tokenize_text_segment(prompt_ids(21 + i, 8 + i % 9), 0) for i in range(B)]
        frames = {}
        for mode in ("old", "fast"):
            os.environ["CSMB_DISABLE_FAST"] = "1" if mode == "old" else "0"
            st = LMState(model, B, max_len=64)
            st.prefill([p[0] for p in prompts], [p[1] for p in prompts])
            frame = torch.zeros((B, 32), device=dev, dtype=torch.int32)
            st.sample_c0(frame, spec); st.depth_decode(frame, spec)
            out = [frame.clone()]
            prev = frame
            for _ in range(3):
                nxt = torch.zeros((B, 32), device=dev, dtype=torch.int32)
                st.decode_frame(prev, nxt, spec)
                out.append(nxt.clone()); prev = nxt
            torch.cuda.synchronize()
            st.check_status()
            frames[mode] = torch.stack(out).cpu()
            for _ in range(3):
                prev = st.decode_frame_graphed(prev, spec)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 10
            e0.record()
            for _ in range(n):
                prev = st.decode_frame_graphed(prev, spec)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            print(f"B={B} {mode}: {ms:.2f} ms/frame-step -> {B * 0.08 / (ms / 1e3):.0f} audio-s/s, "
                  f"{9.1067e9 / ms / 1e6 / 6557.8:.3f} of HBM roofline", flush=True)
            del st
        same = torch.equal(frames["old"], frames["fast"])
        nd = int((frames["old"] != frames["fast"]).sum())
        print(f"B={B} tokens equal: {same} (differing entries: {nd} of {frames['old'].numel()})", flush=True)
    os.environ["CSMB_DISABLE_FAST"] = "0"
    # tuning knobs at B = 64 (graph replay timing)
    B = 64
    prompts = [tokenizers.tokenize_text_segment(prompt_ids(21 + i, 8 + i % 9), 0) for i in range(B)]
    for (mk, pdl, ctas) in ((4, 1, 148), (4, 0, 148), (2, 1, 148), (8, 1, 148), (4, 1, 296), (2, 1, 296), (1, 1, 148)):
        _lib.lib().csmb_debug_set_fast_frame(mk, pdl, ctas)
        st = LMState(model, B, max_len=64)
        st.prefill([p[0] for p in prompts], [p[1] for p in prompts])
        frame = torch.zeros((B, 32), device=dev, dtype=torch.int32)
        st.sample_c0(frame, spec); st.depth_decode(frame, spec)
        prev = frame
        for _ in range(3):
            prev = st.decode_frame_graphed(prev, spec)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            prev = st.decode_frame_graphed(prev, spec)
        e1.record(); torch.cuda.synchronize()
        st.check_status()
        print(f"B=64 min_kblocks={mk} pdl={pdl} ctas={ctas}: {e0.elapsed_time(e1) / 10:.2f} ms/frame-step", flush=True)
        del st
    _lib.lib().csmb_debug_set_fast_frame(4, 1, 148)


@section("persist")
def _persist():
    """Persistent batched frame kernel (csrc/batch_persist.cu) vs the kernel chain: tokens and timing."""
    from csm_mlx_b200 import _lib
    from tests.workloads import prompt_ids
    spec = SamplerSpec()
    for B in (3, 16, 64):
        prompts = [tokenizers.tokenize_text_segment(prompt_ids(21 + i, 8 + i % 9), 0) for i in range(B)]
        frames = {}
        for mode in ("chain", "persist"):
            os.environ["CSMB_ENABLE_PERSIST"] = "0" if mode == "chain" else "1"
            st = LMState(model, B, max_len=64)
            st.prefill([p[0] for p in prompts], [p[1] for p in prompts])
            frame = torch.zeros((B, 32), device=dev, dtype=torch.int32)
            st.sample_c0(frame, spec); st.depth_decode(frame, spec)
            out = [frame.clone()]
            prev = frame
            try:
                for _ in range(3):
                    nxt = torch.zeros((B, 32), device=dev, dtype=torch.int32)
                    st.decode_frame(prev, nxt, spec)
                    out.append(nxt.clone()); prev = nxt
                torch.cuda.synchronize()
                st.check_status()
            except Exception as e:
                print(f"B={B} {mode}: FAILED {e}", flush=True)
                frames[mode] = None
                continue
            frames[mode] = torch.stack(out).cpu()
            for _ in range(3):
                prev = st.decode_frame_graphed(prev, spec)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 10
            e0.record()
            for _ in range(n):
                prev = st.decode_frame_graphed(prev, spec)
            e1.record(); torch.cuda.synchronize()
            st.check_status()
            ms = e0.elapsed_time(e1) / n
            print(f"B={B} {mode}: {ms:.2f} ms/frame-step -> {B * 0.08 / (ms / 1e3):.0f} audio-s/s, "
                  f"{9.1067e9 / ms / 1e6 / 6557.8:.3f} of HBM roofline", flush=True)
            del st
        if frames["chain"] is not None and frames["persist"] is not None:
            nd = int((frames["chain"] != frames["persist"]).sum())
            print(f"B={B} tokens equal: {nd == 0} (differing entries: {nd} of {frames['chain'].numel()})", flush=True)
            if nd:
                d = (frames["chain"] != frames["persist"])
                print("  first differing (frame, seq, cb):", d.nonzero()[:5].tolist(), flush=True)
    os.environ["CSMB_ENABLE_PERSIST"] = "1"
    B = 64
    prompts = [tokenizers.tokenize_text_segment(prompt_ids(21 + i, 8 + i % 9), 0) for i in range(B)]
    # phase timers
    prof = torch.zeros((148, 2, 12), device=dev, dtype=torch.int64)
    st = LMState(model, B, max_len=64)
    st.prefill([p[0] for p in prompts], [p[1] for p in prompts])
    frame = torch.zeros((B, 32), device=dev, dtype=torch.int32)
    st.sample_c0(frame, spec); st.depth_decode(frame, spec)
    prev = frame
    for _ in range(2):
        nxt = torch.zeros((B, 32), device=dev, dtype=torch.int32); st.decode_frame(prev, nxt, spec); prev = nxt
    _lib.lib().csmb_debug_set_frame_batch_prof(prof.data_ptr())
    nxt = torch.zeros((B, 32), device=dev, dtype=torch.int32); st.decode_frame(prev, nxt, spec); prev = nxt
    torch.cuda.synchronize()
    _lib.lib().csmb_debug_set_frame_batch_prof(None)
    names = ["gemm", "prefetch", "barrier", "attn", "norm", "swiglu", "sample", "embed", "accwait", "epi", "-", "-"]
    pr = prof.cpu().double() / 1.965e3  # us at 1965 MHz
    for th, label in ((0, "thread 0 (producer)"), (1, "thread 128 (epilogue warp)")):
        print(f"  {label}: mean over CTAs [us/frame]: " + ", ".join(f"{n} {pr[:, th, i].mean():.0f}" for i, n in enumerate(names))
              + f" | total {pr[:, th].sum(1).mean():.0f}", flush=True)
        print(f"     CTA0: " + ", ".join(f"{n} {pr[0, th, i]:.0f}" for i, n in enumerate(names)), flush=True)
        print(f"     CTA147: " + ", ".join(f"{n} {pr[147, th, i]:.0f}" for i, n in enumerate(names)), flush=True)
    del st
    for mk in (2, 4):
        _lib.lib().csmb_debug_set_frame_batch(mk)
        st = LMState(model, B, max_len=64)
        st.prefill([p[0] for p in prompts], [p[1] for p in prompts])
        frame = torch.zeros((B, 32), device=dev, dtype=torch.int32)
        st.sample_c0(frame, spec); st.depth_decode(frame, spec)
        prev = frame
        for _ in range(3):
            prev = st.decode_frame_graphed(prev, spec)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            prev = st.decode_frame_graphed(prev, spec)
        e1.record(); torch.cuda.synchronize()
        st.check_status()
        print(f"B=64 persist min_kblocks={mk}: {e0.elapsed_time(e1) / 10:.2f} ms/frame-step", flush=True)
        del st
    os.environ["CSMB_ENABLE_PERSIST"] = "0"


@section("overlap")
def _overlap():
    """Where do the ~0.16 ms per frame between k_frame alone and the streaming loop go?"""
    mimi = Mimi(32, device=dev).load_pytorch_weights(random_mimi_weights())
    tokenizers.set_audio_tokenizer(mimi)
    spec = SamplerSpec()
    st = LMState(model, 1, max_len=512)
    codec = mimi.new_decode_stream(1)
    main = torch.cuda.current_stream(dev)
    side = torch.cuda.Stream(dev)
    ev = torch.cuda.Event()
    sink = torch.zeros((1, 1920), device=dev)

    def run(mode, n=100):
        st.reset(); codec.reset()
        st.prefill([ptok], [pmask])
        frame = st.first_frame_fused(spec)
        for _ in range(3):
            frame = st.decode_frame_fused(frame, spec)
            codec.step(frame.reshape(1, 32, 1))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            if mode == "lm":
                pass
            elif mode == "serial":
                codec.step(frame.reshape(1, 32, 1))
            elif mode in ("overlap", "sync-only", "copy-only"):
                ev.record(main)
                side.wait_event(ev)
                frame.record_stream(side)
                with torch.cuda.stream(side):
                    if mode == "overlap":
                        a = codec.step(frame.reshape(1, 32, 1))
                        sink.copy_(a.reshape(1, -1))
                    elif mode == "copy-only":
                        sink[:, :32].copy_(frame.float())
            frame = st.decode_frame_fused(frame, spec)
        main.wait_stream(side)
        e1.record(); torch.cuda.synchronize()
        print(f"  {mode:10s}: {e0.elapsed_time(e1) / n:.3f} ms/frame", flush=True)

    for rep in range(2):
        for mode in ("lm", "sync-only", "copy-only", "overlap", "serial"):
            run(mode)
    st.check_status()


@section("diverge")
def _diverge():
    """64 x 125 frames through the engine (fused chain) vs each utterance alone (persistent batch-1 kernel)."""
    from csm_mlx_b200 import serving
    from tests.workloads import prompt_ids
    mimi = Mimi(32, device=dev).load_pytorch_weights(random_mimi_weights())
    tokenizers.set_audio_tokenizer(mimi)
    eng = serving.Engine(model, max_batch=64, max_len=160)
    prompts = [tokenizers.tokenize_text_segment(prompt_ids(21 + i, 8 + i % 9), 0) for i in range(64)]
    rids = [eng.submit_prompt(t, m, 125) for t, m in prompts]
    eng.run()
    firsts = []
    for i in range(64):
        (single,) = generation.generate_tokens(model, [prompts[i]], 125, temperature=0.0)
        got = eng.tokens(rids[i])
        neq = (single != got).any(dim=1).nonzero()
        firsts.append(int(neq[0]) if len(neq) else None)
    same = sum(f is None for f in firsts)
    print(f"identical over 125 frames: {same} of 64; first differing frame of the others: {sorted(f for f in firsts if f is not None)}", flush=True)


@section("epi")
def _epi():
    """What do the Linear epilogues of the fused chain cost?  (timing only: flags 1/3 produce wrong results)"""
    from csm_mlx_b200 import _lib
    from tests.workloads import prompt_ids
    spec = SamplerSpec()
    for B in (16, 64):
        prompts = [tokenizers.tokenize_text_segment(prompt_ids(21 + i, 8 + i % 9), 0) for i in range(B)]
        for flags in (0, 1, 3, 0):
            _lib.lib().csmb_debug_set_fast_frame_flags(flags)
            st = LMState(model, B, max_len=64)
            st.prefill([p[0] for p in prompts], [p[1] for p in prompts])
            frame = torch.zeros((B, 32), device=dev, dtype=torch.int32)
            st.sample_c0(frame, spec); st.depth_decode(frame, spec)
            prev = frame
            for _ in range(3):
                prev = st.decode_frame_graphed(prev, spec)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                prev = st.decode_frame_graphed(prev, spec)
            e1.record(); torch.cuda.synchronize()
            print(f"B={B} flags={flags}: {e0.elapsed_time(e1) / 10:.2f} ms/frame-step", flush=True)
            del st
    _lib.lib().csmb_debug_set_fast_frame_flags(0)


@section("lanes")
def _lanes():
    """Do G independent groups of 64/G sequences, each a CUDA-graph replay of the fused chain on its own stream, beat
    one group of 64?  (Every kernel of the chain is latency-bound, so concurrent chains might overlap.)"""
    from tests.workloads import prompt_ids
    spec = SamplerSpec()
    for G in (1, 2, 4):
        B = 64 // G
        states, prevs, streams = [], [], [torch.cuda.Stream(dev) for _ in range(G)]
        for gi in range(G):
            prompts = [tokenizers.tokenize_text_segment(prompt_ids(21 + gi * B + i, 8 + i % 9), 0) for i in range(B)]
            st = LMState(model, B, max_len=64)
            st.prefill([p[0] for p in prompts], [p[1] for p in prompts])
            frame = torch.zeros((B, 32), device=dev, dtype=torch.int32)
            st.sample_c0(frame, spec); st.depth_decode(frame, spec)
            prev = frame
            for _ in range(2):
                prev = st.decode_frame_graphed(prev, spec)
            states.append(st); prevs.append(prev)
        torch.cuda.synchronize()

        def step_all():
            for gi in range(G):
                with torch.cuda.stream(streams[gi]):
                    prevs[gi] = states[gi].decode_frame_graphed(prevs[gi], spec)

        for s in streams:
            s.wait_stream(torch.cuda.current_stream(dev))
        for _ in range(2):
            step_all()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in streams:
            s.wait_event(e0)
        n = 8
        for _ in range(n):
            step_all()
        for s in streams:
            torch.cuda.current_stream(dev).wait_stream(s)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"G={G} groups x B={B}: {ms:.2f} ms per step of all 64 sequences -> {64 * 0.08 / (ms / 1e3):.0f} audio-s/s", flush=True)
        del states


@section("gu")
def _gu():
    """SwiGLU fused into the gate|up Linear's epilogue (default) vs the separate k_swiglu_split launch (debug flag 4):
    tokens must be identical; timing of the CUDA-graph replay."""
    from csm_mlx_b200 import _lib
    from tests.workloads import prompt_ids
    spec = SamplerSpec()
    for B in (3, 64, 130):
        prompts = [tokenizers.tokenize_text_segment(prompt_ids(21 + i, 8 + i % 9), 0) for i in range(B)]
        frames = {}
        for flags in (4, 0):
            _lib.lib().csmb_debug_set_fast_frame_flags(flags)
            st = LMState(model, B, max_len=64)
            st.prefill([p[0] for p in prompts], [p[1] for p in prompts])
            frame = torch.zeros((B, 32), device=dev, dtype=torch.int32)
            st.sample_c0(frame, spec); st.depth_decode(frame, spec)
            out = [frame.clone()]
            prev = frame
            for _ in range(3):
                nxt = torch.zeros((B, 32), device=dev, dtype=torch.int32)
                st.decode_frame(prev, nxt, spec)
                out.append(nxt.clone()); prev = nxt
            torch.cuda.synchronize()
            st.check_status()
            frames[flags] = torch.stack(out).cpu()
            for _ in range(3):
                prev = st.decode_frame_graphed(prev, spec)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                prev = st.decode_frame_graphed(prev, spec)
            e1.record(); torch.cuda.synchronize()
            st.check_status()
            ms = e0.elapsed_time(e1) / 10
            print(f"B={B} {'separate swiglu' if flags else 'fused epilogue '}: {ms:.2f} ms/frame-step -> {B * 0.08 / (ms / 1e3):.0f} audio-s/s", flush=True)
            del st
        print(f"B={B} tokens identical: {torch.equal(frames[0], frames[4])}", flush=True)
    _lib.lib().csmb_debug_set_fast_frame_flags(0)


@section("setup")
def _setup():
    """Where does the first-chunk time of stream_generate go?"""
    mimi = Mimi(32, device=dev).load_pytorch_weights(random_mimi_weights())
    tokenizers.set_audio_tokenizer(mimi)
    spec = SamplerSpec(temperature=0.0)

    def tick(label, fn):
        torch.cuda.synchronize(); t1 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
        print(f"  {label}: {1e3 * (time.perf_counter() - t1):.2f} ms", flush=True)
        return r

    for rep in range(3):
        print("rep", rep)
        sess = tick("session (LMState alloc)", lambda: generation._Session(model, [(ptok, pmask)], 125, spec, None, None))
        codec = tick("codec stream alloc", lambda: mimi.new_decode_stream(1))
        f0 = tick("first step (prefill + frame 0 per-op)", sess.step)
        f1 = tick("second step (fused; workspace alloc)", sess.step)
        f2 = tick("third step", sess.step)
        tick("codec step 0 (eager warm)", lambda: codec.step(f0.reshape(1, 32, 1)))
        tick("codec step 1 (capture + replay)", lambda: codec.step(f1.reshape(1, 32, 1)))
        tick("codec step 2 (replay)", lambda: codec.step(f2.reshape(1, 32, 1)))
    for rep in range(3):
        t1 = time.perf_counter(); lat = []; last = t1
        for ch in generation.stream_generate(model, ids, 0, [], max_audio_length_ms=2000, temperature=0.0):
            now = time.perf_counter(); lat.append(1e3 * (now - last)); last = now
        print("stream_generate chunk gaps ms:", [round(x, 2) for x in lat[:6]], "total", round(1e3 * (last - t1), 1))

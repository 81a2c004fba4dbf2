#!/usr/bin/env python
"""Benchmark of the CSM speech-token generation hot path (BASELINE.json metric: generated audio-seconds per second).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path, one process per GPU
    python bench.py --impl reference --steps K --warmup W     # CPU arm: the oracle port on the host cores

Workload at every N: BASELINE.json configs[3] — csm_1b bf16, 64 independent utterances of 10 s (8-16 prompt rows + 125
frames each), greedy, seeded random-init weights, request-sharded over the N GPUs (utterance i -> rank i mod N; 64 / N per
GPU; utterances share nothing, so there is no data-path collective — NCCL only gathers the tokens and timings).  A "step" =
the whole 64-utterance job: continuous-batching engine (admission prefill + 125 frame-steps of the fused tcgen05 chain) +
batched Mimi decode of every utterance.  The job is fixed, so the scaling is STRONG.

  value  device-timed (CUDA events, max over ranks): prompts staged in HBM before the timed region, audio left on the device
  e2e    the same job through the public serving API with host buffers: prompt rows submitted from host memory, every
         frame's tokens copied to pinned host memory, the decoded audio copied back to the host, inside the timed region
  roofline        the frame-step of the chain (the dominant kernel group: one CUDA-graph replay = 1 047 launches of this
                  library) at this rank's batch: algorithmic bytes / CUDA-event time, against MEASURED_PEAKS.json
  other_configs   latency path (configs[1]: batch 1 streaming through the persistent frame kernel, with its own roofline,
                  p50 / p90 inter-chunk latency and time to first chunk), configs[2] (context prefill), configs[4] (codec,
                  128 x 60 s sharded over the ranks), batch sweep of the chain
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# stdout carries exactly ONE line, the JSON record: libraries that chat on file descriptor 1 (NCCL prints its version
# banner there) are sent to stderr, and the record is written to a private duplicate of the original stdout.
_RECORD_OUT = None


def claim_stdout() -> None:
    """Called once by the command-line entry (never on import)."""
    global _RECORD_OUT
    if _RECORD_OUT is None:
        sys.stdout.flush()
        _RECORD_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(record: dict) -> None:
    out = _RECORD_OUT if _RECORD_OUT is not None else sys.stdout
    out.write(json.dumps(record) + "\n")
    out.flush()


import torch  # noqa: E402

FRAME_S = 0.08
SECONDS = 10.0
N_REQ = 64
WORKLOAD = "configs[3]: csm_1b bf16, 64 independent utterances x 10 s, request-sharded across the GPUs (continuous batching + Mimi decode)"
LATENCY_WORKLOAD = "configs[1]: csm_1b bf16 batch=1 streaming generation, 10 s audio, single B200 (latency path)"
# Sum of all 64 x 125 x 32 greedy tokens of the job.  The engine's numeric path does not depend on the batch size or on the
# sharding (csm_mlx_b200/serving.py, "Batch invariance"), so this constant holds at every --gpus N; a different value means the
# tokens changed.
CFG4_TOKENS_CHECKSUM = 261764229  # profiles/r02_bench_1gpu.json; identical at --gpus 1, 2, 4, 8 (re-baselined with the 2-issuer Linears)


def weight_bytes(proj_table: bool = False) -> float:
    """SURVEY.md §8(d) / BASELINE.md §4: bf16 weight bytes streamed per frame-step (independent of the batch).  With the
    projected-embedding table the chain no longer streams the projection matrix for depth steps 2..31 (30 x 4 194 304 B)."""
    w = 1_946_292_224 + 8_400_896 + 31 * 230_711_296
    return w - (30 * 4_194_304 if proj_table else 0)


def lm_algorithmic_bytes(prompt_rows: float, frames: int, batch: int = 1, proj_table: bool = False) -> float:
    """Weights once per step for the whole batch + per-sequence KV / embedding terms, averaged over the utterance."""
    s_avg = prompt_rows + (frames - 1) / 2.0
    per_seq = 32_768 * 2 * s_avg + 32_768 * 2 + 63 * 4096  # fp32 KV here: 64 KiB per token read + one token written
    return weight_bytes(proj_table) + batch * per_seq


class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU every 100 ms during the timed region (pynvml)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._t = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None
            return self
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def _run(self):
        nv = self._nv
        names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake_slowdown": 0x80, "sw_power_cap": 0x4}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.1)

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cfg4_prompt(i: int):
    """Utterance i of configs[3]: BOS + (8 + i mod 9) seeded ids + EOS, speaker 0."""
    from csm_mlx_b200 import tokenizers
    from tests.workloads import prompt_ids

    return tokenizers.tokenize_text_segment(prompt_ids(21 + i, 8 + i % 9), 0)


# ------------------------------------------------------------------------------------------------- CPU arm
def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.lower().startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def oracle_utterance(orc, mimi_w, tok, mask, frames: int) -> float:
    """Seconds for prefill + `frames` greedy frames + Mimi decode of them with the CPU oracle (one utterance of the job)."""
    from oracle import lm as olm, mimi as omimi

    t0 = time.perf_counter()
    toks = olm.generate_tokens(orc, tok.long(), mask, frames)
    omimi.decode(toks.t()[None].clamp(max=2047), mimi_w)
    return time.perf_counter() - t0


def cpu_sample(frames: int, warm: bool = True):
    """(audio-s/s, seconds, description) of the oracle port on all host cores over utterance 0 of the job."""
    from csm_mlx_b200.random_init import random_csm_weights, random_mimi_weights
    from oracle import lm as olm

    torch.set_num_threads(os.cpu_count())
    orc = olm.OracleCSM(olm.CSM_1B, random_csm_weights())
    mimi_w = random_mimi_weights()
    tok, mask = cfg4_prompt(0)
    if warm:
        oracle_utterance(orc, mimi_w, tok, mask, 1)
    return orc, mimi_w, tok, mask


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    frames = args.ref_frames
    orc, mimi_w, tok, mask = cpu_sample(frames, warm=False)
    for _ in range(args.warmup):
        oracle_utterance(orc, mimi_w, tok, mask, 1)
    times = [oracle_utterance(orc, mimi_w, tok, mask, frames) for _ in range(args.steps)]
    t = statistics.mean(times)
    v = frames * FRAME_S / t
    sample = (f"one utterance of the 64 (utterance 0: prefill of {int(tok.shape[0])} rows + {frames} greedy frames + Mimi decode) per "
              f"step; the job is 64 such independent utterances, so its CPU throughput is this number")
    line = {
        "impl": "reference", "metric": "audio_seconds_per_second", "value": v, "unit": "audio-s/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample, "sampling": "greedy", "frames_per_utterance": frames,
                   "note": "reference dependencies (mlx, mlx_lm, moshi_mlx) are not installable here; this is the "
                           "oracle port of the reference path in PyTorch CPU fp32 on all host cores"},
        "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": torch.get_num_threads(), "cpu_model": cpu_model(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------- helpers (GPU arm)
def ev():
    return torch.cuda.Event(enable_timing=True)


def timed(dev, fn, n=1):
    torch.cuda.synchronize(dev)
    a, b = ev(), ev()
    a.record()
    r = None
    for _ in range(n):
        r = fn()
    b.record()
    torch.cuda.synchronize(dev)
    return a.elapsed_time(b) / n, r


def chain_step_ms(model, dev, B: int, n: int = 10):
    """CUDA-event time of one frame-step of the fused chain (graph replay) for B sequences in lock-step."""
    from csm_mlx_b200.runtime import LMState, SamplerSpec

    spec = SamplerSpec(temperature=0.0)
    prompts = [cfg4_prompt(i) for i in range(B)]
    st = LMState(model, B, max_len=64, row_invariant=True)
    t_pre, _ = timed(dev, lambda: st.prefill([p[0] for p in prompts], [p[1] for p in prompts]))
    frame = torch.zeros((B, 32), device=dev, dtype=torch.int32)
    st.sample_c0(frame, spec)
    st.depth_decode(frame, spec)
    state = {"f": frame}
    for _ in range(3):
        state["f"] = st.decode_frame_graphed(state["f"], spec)

    def step():
        state["f"] = st.decode_frame_graphed(state["f"], spec)
    t_step, _ = timed(dev, step, n)
    st.check_status()
    launches = st.graph_launches
    del st
    return t_step, t_pre, launches


def fp8_config(model, dev):
    """Weight-only FP8 (csm_mlx_b200.quantize, the nn.quantize analogue of the reference README): a second, quantised copy of
    the model on the row-based GEMV path, batch 1, against the bf16 model on the same per-op path and on the frame kernel."""
    from csm_mlx_b200 import CSM, csm_1b, quantize, tokenizers
    from csm_mlx_b200.random_init import random_csm_weights
    from csm_mlx_b200.runtime import LMState, SamplerSpec
    from tests.workloads import cfg1_prompt_ids

    tok, mask = tokenizers.tokenize_text_segment(cfg1_prompt_ids(), 0)
    spec = SamplerSpec(temperature=0.0)

    def per_op_ms(m):
        st = LMState(m, 1, max_len=tok.shape[0] + 64)
        st.prefill([tok], [mask])
        fr = torch.zeros((1, 32), device=dev, dtype=torch.int32)
        st.sample_c0(fr, spec)
        st.depth_decode(fr, spec)
        state = {"f": fr}
        os.environ["CSMB_DISABLE_FAST"] = "1"     # the bf16 model too on the row-based path (CUDA graph of csmb_decode_frame)
        try:
            for _ in range(3):
                state["f"] = st.decode_frame_graphed(state["f"], spec)

            def one():
                state["f"] = st.decode_frame_graphed(state["f"], spec)
            ms, _ = timed(dev, one, 20)
        finally:
            os.environ.pop("CSMB_DISABLE_FAST", None)
        st.check_status()
        return ms

    q = quantize(CSM(csm_1b(), device=dev).load_weights(random_csm_weights()))
    lin_bytes = lambda m: sum(t.numel() * t.element_size() for stck in (m.backbone, m.decoder)
                              for ts in (stck.wqkv, stck.wo, stck.wgu, stck.wdown) for t in ts)
    ms_q, ms_b = per_op_ms(q), per_op_ms(model)
    bytes_q = lin_bytes(q)

    def frame_kernel_ms(m):
        st = LMState(m, 1, max_len=tok.shape[0] + 160)
        if not st.fused_supported(spec):
            return None
        st.prefill([tok], [mask])
        state = {"f": st.first_frame_fused(spec)}
        for _ in range(3):
            state["f"] = st.decode_frame_fused(state["f"], spec)

        def one():
            state["f"] = st.decode_frame_fused(state["f"], spec)
        ms, _ = timed(dev, one, 40)
        st.check_status()
        return ms

    fk_q, fk_b = frame_kernel_ms(q), frame_kernel_ms(model)
    alg_q = lm_algorithmic_bytes(int(tok.shape[0]), 32) - 9_106_743_296 / 2     # e4m3 bytes (+ 0.2 % fp32 scales, not counted)
    del q
    torch.cuda.empty_cache()
    return {"format": "E4M3 bytes + one fp32 scale per output channel (include/csm_b200.h CSMB_WEIGHTS_E4M3); embeddings, norms bf16 / fp32",
            "linear_weight_bytes": {"bf16": lin_bytes(model), "e4m3": bytes_q}, "path": "batch 1: the persistent frame kernel, and the row-based GEMV kernels in a CUDA graph (csmb_decode_frame)",
            "row_based_ms_per_frame": {"e4m3": ms_q, "bf16": ms_b},
            "frame_kernel_ms_per_frame": {"e4m3": fk_q, "bf16": fk_b, "kernel": "csmb::k_frame (persistent whole-frame kernel; e4m3: the same units at one byte per weight)"},
            "roofline": {"bound": "hbm", "kernel": "csmb::k_frame on the e4m3 model", "algorithmic_bytes_per_frame": alg_q,
                         "achieved": alg_q / ((fk_q or ms_q) * 1e-3) / 1e9, "unit": "GB/s", "frac": alg_q / ((fk_q or ms_q) * 1e-3) / 1e9 / 6557.8},
            "note": "parity: tests/test_quantization.py (oracle on the dequantised weights); the tensor-core chain declines a quantised model (DESIGN.md §8)"}


def latency_path(model, mimi, dev, lib, steps: int):
    """BASELINE.json configs[1]: batch 1, 10 s, streaming, through the persistent frame kernel (round 1's headline)."""
    from csm_mlx_b200 import generation, tokenizers
    from csm_mlx_b200.runtime import LMState, SamplerSpec
    from tests.workloads import cfg1_prompt_ids

    ids = cfg1_prompt_ids()
    tok, mask = tokenizers.tokenize_text_segment(ids, 0)
    frames = int(SECONDS / FRAME_S)
    spec = SamplerSpec(temperature=0.0)
    st = LMState(model, 1, max_len=tok.shape[0] + frames + 1)
    fused = st.fused_supported(spec) and os.environ.get("CSMB_DISABLE_FUSED", "0") != "1"
    next_frame = (lambda fr: st.decode_frame_fused(fr, spec)) if fused else (lambda fr: st.decode_frame_graphed(fr, spec))

    def e2e_step():
        lat = []
        t0 = time.perf_counter()
        last = t0
        n = 0
        for _chunk in generation.stream_generate(model, ids, 0, [], max_audio_length_ms=SECONDS * 1000, temperature=0.0):
            now = time.perf_counter()
            lat.append(now - last)
            last = now
            n += 1
        torch.cuda.synchronize(dev)
        return time.perf_counter() - t0, lat, n

    for _ in range(2):
        e2e_step()
    times, lats, firsts = [], [], []
    for _ in range(steps):
        t, lat, n = e2e_step()
        assert n == frames, n
        times.append(t)
        firsts.append(lat[0])
        lats += lat[1:]
    # the frame kernel alone
    st.prefill([tok], [mask])
    fr = torch.zeros((1, 32), device=dev, dtype=torch.int32)
    st.sample_c0(fr, spec)
    st.depth_decode(fr, spec)
    state = {"f": fr}
    for _ in range(3):
        state["f"] = next_frame(state["f"])

    def one():
        state["f"] = next_frame(state["f"])
    frame_ms, _ = timed(dev, one, 60)
    st.check_status()
    alg = lm_algorithmic_bytes(int(tok.shape[0]), frames)
    lats_ms = sorted(1e3 * x for x in lats)
    pct = lambda p: lats_ms[min(len(lats_ms) - 1, int(p * len(lats_ms)))] if lats_ms else None
    t_e2e = statistics.mean(times)
    return {"workload": LATENCY_WORKLOAD, "e2e_audio_s_per_s": frames * FRAME_S / t_e2e, "e2e_ms": 1e3 * t_e2e,
            "api": "csm_mlx.stream_generate (host prompt, every chunk + tokens copied to pinned host memory)",
            "latency_ms": {"p50": pct(0.5), "p90": pct(0.9), "first_chunk": 1e3 * statistics.median(firsts),
                           "what": "time between successive stream_generate chunks on the host; first_chunk = call -> first "
                                   "1 920 samples on the host (prompt assembly + prefill + first frame + the one-frame lookahead + codec step)"},
            "roofline": {"bound": "hbm", "kernel": "csmb::k_frame (persistent whole-frame kernel)" if fused else "per-op graph",
                         "ms_per_frame": frame_ms, "algorithmic_bytes_per_frame": alg, "achieved": alg / (frame_ms * 1e-3) / 1e9,
                         "unit": "GB/s"}}


def context_config(model, mimi, dev):
    """BASELINE.json configs[2]: 2-segment context (2 x 5 s synthetic audio -> Mimi encode) + new text -> prefill -> frames."""
    from csm_mlx_b200 import Segment, generation, tokenizers
    from csm_mlx_b200.runtime import LMState, SamplerSpec
    from tests.workloads import synthetic_audio

    spec = SamplerSpec(temperature=0.0)
    tokenizers.set_text_tokenizer(tokenizers.SyntheticTextTokenizer())
    try:
        clips = [synthetic_audio(11, 5.0), synthetic_audio(12, 5.0)]
        [mimi.encode(c[None, None].to(dev)) for c in clips]
        t_enc, _ = timed(dev, lambda: [mimi.encode(c[None, None].to(dev)) for c in clips])
        segs = [Segment(i, "context sentence number %d" % i, clips[i]) for i in range(2)]
        prompt = generation._build_prompt(model, "and now the answer", 0, segs)
        rows = int(prompt[0].shape[0])
        st = LMState(model, 1, max_len=rows + 40)
        st.prefill([prompt[0]], [prompt[1]])
        st.reset()
        t_pre, _ = timed(dev, lambda: st.prefill([prompt[0]], [prompt[1]]))
        frame = torch.zeros((1, 32), device=dev, dtype=torch.int32)
        st.sample_c0(frame, spec)
        st.depth_decode(frame, spec)
        state = {"f": frame}

        def step1():
            state["f"] = st.decode_frame_fused(state["f"], spec)
        for _ in range(3):
            step1()
        t_f, _ = timed(dev, step1, 20)
        st.check_status()
        del st
        # near-maximum context: a 2 030-row prompt (text + audio rows), GPU time of the prompt pass with staged inputs
        from oracle import lm as olm
        from tests.workloads import prompt_ids

        gen = torch.Generator().manual_seed(77)
        t1 = olm.text_rows(prompt_ids(9, 10))
        a1 = olm.audio_rows(torch.randint(0, 2048, (32, 2017), generator=gen))
        tok_l, mask_l = torch.cat([t1[0], a1[0]]).int(), torch.cat([t1[1], a1[1]])
        stl = LMState(model, 1, max_len=int(tok_l.shape[0]) + 8)
        staged = stl.stage_prefill([tok_l], [mask_l])
        stl.run_prefill(staged)

        def long_pass():
            stl.reset()
            stl.run_prefill(staged)
        t_long, _ = timed(dev, long_pass, 2)
        stl.check_status()
        rows_l = int(tok_l.shape[0])
        del stl
        # the same conversation through the serving engine: the step that admits a turn, context KV prefilled (first turn:
        # prefix-cache miss) vs copied from the KV prefix cache (next turn: hit); both include the frame-step itself
        from csm_mlx_b200 import serving

        eng = serving.Engine(model, max_batch=1, max_len=rows + 48)

        def admission_step(text):
            eng.submit(text, 0, segs, max_audio_length_ms=160)
            t, _ = timed(dev, eng.step)
            eng.run()
            return t
        admission_step("a first turn to warm up")                     # graph capture, allocations
        eng.kv_cache = serving.KVPrefixCache()
        t_miss = admission_step("and now the answer")
        t_hit = admission_step("and one more answer")
        kv_stats = (eng.kv_cache.hits, eng.kv_cache.misses, eng.kv_cache.nbytes)
        eng.state.check_status()
        del eng

        # time to the first 80 ms chunk of a turn through stream_generate (host clock): context seen for the first time
        # (Mimi encode of both clips + prompt pass over all rows) vs the next turn (codes and KV prefix from the model's caches)
        def first_chunk_ms(text):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            g = generation.stream_generate(model, text, 0, segs, max_audio_length_ms=400, temperature=0.0)
            next(g)
            t = 1e3 * (time.perf_counter() - t0)
            g.close()
            torch.cuda.synchronize(dev)
            return t
        first_chunk_ms("warm up the streaming path")
        cold, warm = [], []
        for i in range(3):
            generation.set_conversation_cache(model)
            cold.append(first_chunk_ms("a turn with a context never seen %d" % i))
            warm.append(first_chunk_ms("the following turn of it %d" % i))
        generation.set_conversation_cache(model)
        return {"prompt_rows": rows, "mimi_encode_ms_total": t_enc, "prefill_ms": t_pre,
                "engine_admission_step_ms": {"kv_prefix_miss": t_miss, "kv_prefix_hit": t_hit, "cache_hits_misses_bytes": kv_stats},
                "stream_first_chunk_ms": {"context_first_seen": statistics.median(cold), "next_turn_same_context": statistics.median(warm),
                                          "what": "call -> first 1 920 samples on the host through stream_generate with the 2 x 5 s context"},
                "prefill_tflops": 2 * 973.1e6 * rows / (t_pre * 1e-3) / 1e12, "ms_per_frame": t_f,
                "prefill_long": {"prompt_rows": rows_l, "ms": t_long, "tflops": 2 * 973.1e6 * rows_l / (t_long * 1e-3) / 1e12},
                "note": "prefill = host staging + backbone over all prompt rows on the chain's kernels (csmb_prefill_fast: one tcgen05 launch per "
                        "Linear, fused norms, tiled attention) + c0 head; TFLOP/s = 2 x 973.1 M x rows / time (fp32-equivalent: every "
                        "product runs twice, bf16 hi + lo); frame kernel with 2-chunk attention (S > 128)"}
    finally:
        tokenizers.set_text_tokenizer(None)


def codec_config(mimi, dev, world, rank, barrier, max_over_ranks):
    """BASELINE.json configs[4]: 128 clips x 60 s through the codec (encode -> codes -> decode), clip i -> rank i mod N."""
    from csm_mlx_b200.sharding import shard_indices
    from tests.workloads import synthetic_audio

    n_clips, secs = 128, 60.0
    mine = shard_indices(n_clips, rank, world)
    clips = torch.stack([synthetic_audio(100 + i, secs) for i in mine])[:, None].to(dev)
    warm = clips[:2]
    mimi.decode(mimi.encode(warm))   # warm-up at the measured pass shape: buffers come from the caching allocator afterwards
    barrier()
    t_e, codes = timed(dev, lambda: mimi.encode(clips))
    t_d, audio = timed(dev, lambda: mimi.decode(codes))
    t_e, t_d = max_over_ranks(t_e / 1e3), max_over_ranks(t_d / 1e3)
    total = n_clips * secs
    frames = int(codes.shape[2])
    csum = int(codes.long().sum())
    del clips, audio
    if rank != 0:
        return None
    return {"clips": n_clips, "seconds_per_clip": secs, "clips_per_gpu": len(mine), "frames": frames,
            "encode_audio_s_per_s": total / t_e, "decode_audio_s_per_s": total / t_d,
            "encode_tflops": 0.46e9 * 12.5 * total / t_e / 1e12, "decode_tflops": 0.43e9 * 12.5 * total / t_d / 1e12,
            "rank0_codes_checksum": csum,
            "note": "tensor-core path (csrc/mimi_tc.cu: persistent tcgen05 GEMM on bf16 hi+lo planes, three MMAs per K step); "
                    "TFLOP/s = algorithmic fp32-equivalent FLOPs (BASELINE.md §4) / time; max over ranks"}


# ------------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch.distributed as dist

    from csm_mlx_b200 import CSM, _lib, csm_1b, serving, tokenizers
    from csm_mlx_b200.mimi import Mimi
    from csm_mlx_b200.random_init import random_csm_weights, random_mimi_weights
    from csm_mlx_b200.sharding import gather_ragged, shard_indices

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a B200: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    model = CSM(csm_1b(), device=dev).load_weights(random_csm_weights())
    mimi = Mimi(32, device=dev).load_pytorch_weights(random_mimi_weights())
    tokenizers.set_audio_tokenizer(mimi)
    frames = int(SECONDS / FRAME_S)
    mine = shard_indices(N_REQ, rank, world)
    prompts = {i: cfg4_prompt(i) for i in mine}
    eng = serving.Engine(model, max_batch=max(1, len(mine)), max_len=32 + frames + 2)
    has_table = model.proj_table() is not None and os.environ.get("CSMB_NO_PROJ_TABLE", "0") != "1"

    # ---- one job.  device=True: CUDA-event timed, audio stays in HBM; device=False: wall clock, host buffers in and out
    def job(device: bool):
        torch.cuda.synchronize(dev)
        c0 = lib.csmb_debug_launch_count()
        e0, e1 = ev(), ev()
        t0 = time.perf_counter()
        e0.record()
        rids = [eng.submit_prompt(prompts[i][0], prompts[i][1], frames) for i in mine]
        steps0 = eng.steps
        eng.run()
        audio = eng.audio(rids, to_host=not device)
        e1.record()
        torch.cuda.synchronize(dev)
        wall = time.perf_counter() - t0
        toks = [eng.tokens(r) for r in rids]
        eager = int(lib.csmb_debug_launch_count() - c0)
        n_steps = eng.steps - steps0
        return {"event_s": e0.elapsed_time(e1) / 1e3, "wall_s": wall, "tokens": toks, "audio": audio, "eager_launches": eager,
                "engine_steps": n_steps}

    for _ in range(max(1, args.warmup)):
        job(True)
    barrier()
    clocks = ClockSampler(local).start()
    dev_times, last = [], None
    for _ in range(args.steps):
        barrier()
        last = job(True)
        dev_times.append(last["event_s"])
    barrier()
    clk = clocks.stop()
    # untimed: two results alive at once, as in the timed loop below (this one and the previous one), so that the caching
    # pinned-host allocator owns both 61 MB result blocks before the clock starts (steady state of a serving process)
    warm_a = job(False)
    warm_b = job(False)
    del warm_a, warm_b
    e2e_times = []
    for _ in range(args.steps):
        barrier()
        r = job(False)
        e2e_times.append(r["wall_s"])
        last_e2e = r
    barrier()
    t_dev = max_over_ranks(statistics.mean(dev_times))
    t_e2e = max_over_ranks(statistics.mean(e2e_times))
    # the only use of the interconnect on this path: gather every rank's tokens to rank 0 over NCCL
    gathered = gather_ragged([x.to(dev) for x in last["tokens"]], N_REQ, device=dev)
    d2h_local = sum(int(t.numel()) * 4 for t in last_e2e["tokens"]) + sum(int(a.numel()) * 4 for a in last_e2e["audio"])
    h2d_local = sum(int(prompts[i][0].numel()) * 4 + int(prompts[i][1].numel()) + 4 * (2 * int(prompts[i][0].shape[0]) + 3) for i in mine) \
        + sum(int(t.numel()) * 4 for t in last_e2e["tokens"])   # prompt rows + maps at admission; codes of the batched Mimi decode
    io = torch.tensor([float(h2d_local), float(d2h_local)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(io)

    # ---- the dominant kernel group alone: one frame-step of the chain at this rank's batch (CUDA events, graph replay)
    B_rank = max(1, len(mine))
    step_ms, _, graph_launches = chain_step_ms(model, dev, B_rank)
    step_ms = max_over_ranks(step_ms)

    other = None
    if not args.no_extras:
        other = {}
        other["cfg5_codec_128x60s_sharded"] = codec_config(mimi, dev, world, rank, barrier, max_over_ranks)
        if rank == 0 and world == 1:
            other["latency_path"] = latency_path(model, mimi, dev, lib, args.steps)
            other["context_2x5s"] = context_config(model, mimi, dev)
            other["fp8_weight_only"] = fp8_config(model, dev)
            sweep = {}
            for Bs in (1, 8, 16, 32, 128, 256):
                t, _, _ = chain_step_ms(model, dev, Bs, 6)
                sweep[str(Bs)] = {"ms_per_frame_step": t, "audio_s_per_s": Bs * FRAME_S / (t / 1e3),
                                  "roofline_frac": lm_algorithmic_bytes(12, 20, Bs, has_table) / (t / 1e3) / 1e9 / 6557.8}
            other["chain_batch_sweep_one_gpu"] = sweep

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    total_frames = sum(int(x.shape[0]) for x in gathered)
    audio_s = total_frames * FRAME_S
    checksum = int(sum(int(x.long().sum()) for x in gathered))
    if CFG4_TOKENS_CHECKSUM is not None and os.environ.get("CSMB_BENCH_NO_CHECKSUM", "0") != "1":
        assert checksum == CFG4_TOKENS_CHECKSUM, (checksum, CFG4_TOKENS_CHECKSUM, "greedy tokens of the job changed")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg = lm_algorithmic_bytes(12.0, frames, B_rank, has_table)
    achieved = alg / (step_ms * 1e-3) / 1e9

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        orc, mimi_w, tok0, mask0 = cpu_sample(args.ref_frames)
        nfr = args.cpu_frames
        tc = oracle_utterance(orc, mimi_w, tok0, mask0, nfr)
        cpu = {"value": nfr * FRAME_S / tc, "unit": "audio-s/s", "cores": torch.get_num_threads(), "cpu_model": cpu_model(), "kind": "port",
               "sample": f"utterance 0 of the 64: prefill ({int(tok0.shape[0])} rows) + {nfr} greedy frames + Mimi decode, oracle PyTorch-CPU fp32 "
                         f"on all host cores ({tc:.1f} s); `--impl reference` runs the full 125 frames"}

    line = {
        "metric": "audio_seconds_per_second", "value": audio_s / t_dev, "unit": "audio-s/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "requests": N_REQ, "per_gpu": len(mine), "frames_per_utterance": frames,
                   "frames_total": total_frames, "prompt_rows": "10..18 (8 + i mod 9 ids + BOS + EOS)", "sampling": "greedy",
                   "weights": "bf16 in HBM (seeded random init, 3.1 GB per GPU, replicated)", "activations": "fp32 accumulation (bf16 hi+lo operand planes on the tensor cores)",
                   "parallelism": f"request sharding: utterance i -> rank i mod {world}, one continuous-batching engine per GPU, no data-path collective",
                   "l2": "inputs larger than L2: 9.0 GB of weights streamed per frame-step vs 126 MB L2",
                   "engine_steps_per_job_rank0": last["engine_steps"]},
        "e2e": {"value": audio_s / t_e2e, "unit": "audio-s/s", "h2d_bytes_per_step": int(io[0]), "d2h_bytes_per_step": int(io[1]),
                "ms_per_step": 1e3 * t_e2e, "api": "csm_mlx_b200.serving.Engine.submit_prompt / run / tokens / audio (host tensors in, host tensors out)"},
        "gpu_launches": int(last["eager_launches"] + last["engine_steps"] * graph_launches),
        "gpu_launches_note": f"rank 0, one job: {last['engine_steps']} engine steps x {graph_launches} kernels of libcsm_b200 per CUDA-graph replay of the "
                             f"chain + {last['eager_launches']} eagerly launched (admission prefill, Mimi decode)",
        "clocks": clk,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum of ONE frame-step (all 1 047 launches) at 64 sequences from ncu
                     # launch lists of scripts/ncu_fast.py (profiles/r02b_chain_ncu.md): 7.72 GB with --cache-control none (L2
                     # warm, a kernel's replays partly hit L2), 12.46 GB with ncu's default cache flush before every launch (every
                     # partial / plane / prefetched weight re-read from DRAM); the live step lies between, at the algorithmic bytes
                     "traffic": 7.715e9 if B_rank == 64 else None, "traffic_cold_caches": 12.46e9 if B_rank == 64 else None,
                     "kernel": f"frame-step of the fused chain (csrc/batch_frame.cu) at {B_rank} sequences: one CUDA-graph replay = {graph_launches} launches "
                               "(csmb::k_gemm_part_t tcgen05 linears + fused element-wise kernels)",
                     "ms_per_frame_step": step_ms, "algorithmic_bytes_per_frame_step": alg,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                     "note": "bytes = bf16 weights streamed once per step for the whole batch (projection matrix only in depth step 1: "
                             "projected-embedding table) + per-sequence KV / embedding terms; traffic: per frame-step like `achieved`, see profiles/r02b_chain_ncu.md"},
        "cpu_baseline": cpu,
        "tokens_checksum": checksum,
        "tokens_checksum_expected": CFG4_TOKENS_CHECKSUM,
        "other_configs": other,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-frames", type=int, default=125, help="frames per step of the CPU arm (125 = the job's full utterance)")
    ap.add_argument("--cpu-frames", type=int, default=25, help="frames of the cpu_baseline sample inside the CUDA arm's run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the informational measurements of the other BASELINE configs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    claim_stdout()
    main()

"""Developer timing / A-B script for round 2 (run under gpurun): prints one line per measurement.  Not a test.
Usage: python scripts/r2_trial.py SECTION [SECTION ...]     sections: chain lanes
Switches travel through the environment variables runtime.LMState reads per call (csmb_chain_opts)."""
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from csm_mlx_b200 import CSM, csm_1b, tokenizers
from csm_mlx_b200.random_init import random_csm_weights
from csm_mlx_b200.runtime import LMState, SamplerSpec
from tests.workloads import prompt_ids

dev = torch.device("cuda", 0)
SECTIONS = sys.argv[1:]
model = CSM(csm_1b(), device=dev).load_weights(random_csm_weights())
spec = SamplerSpec()


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


def setenv(**kw):
    for k, v in kw.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)


def warm_state(B, seed0=21, row_invariant=False):
    prompts = [tokenizers.tokenize_text_segment(prompt_ids(seed0 + i, 8 + i % 9), 0) for i in range(B)]
    st = LMState(model, B, max_len=64, row_invariant=row_invariant)
    st.prefill([p[0] for p in prompts], [p[1] for p in prompts])
    frame = torch.zeros((B, 32), device=dev, dtype=torch.int32)
    st.sample_c0(frame, spec)
    st.depth_decode(frame, spec)
    return st, frame


def time_graph_steps(st, prev, n=10):
    for _ in range(3):
        prev = st.decode_frame_graphed(prev, spec)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        prev = st.decode_frame_graphed(prev, spec)
    e1.record()
    torch.cuda.synchronize()
    st.check_status()
    return e0.elapsed_time(e1) / n, prev


@section("chain")
def _chain():
    """Frame-step of the fused chain (CUDA-graph replay) over batch size x {projected-embedding table} x {smem budget}."""
    base = {}
    for B in (int(x) for x in os.environ.get("R2_BATCHES", "1,2,8,16,32,64,128").split(",")):
        for tab, smem in (((0, 0), (1, 0), (1, 96)) if os.environ.get("R2_FULL") else ((1, 0),)):
            for flags in (int(x) for x in os.environ.get("R2_FLAGS", "0").split(",")):
                setenv(CSMB_NO_PROJ_TABLE=None if tab else 1, CSMB_CHAIN_SMEM_KB=smem or None, CSMB_CHAIN_FLAGS=flags or None)
                st, frame = warm_state(B, row_invariant=(B == 1))
                ms, last = time_graph_steps(st, frame)
                key = (B,)
                if key not in base:
                    base[key] = last.clone()
                same = bool(torch.equal(base[key], last))
                print(f"chain B={B:3d} table={tab} smem_kb={smem or 200} flags={flags}: {ms:.3f} ms/step  {B * 0.08 / (ms / 1e3):7.0f} audio-s/s  "
                      f"frac {8.981e9 / (ms * 1e-3) / 6557.8e9:.3f}  tokens_same_as_first_variant={same}", flush=True)
                del st
    setenv(CSMB_NO_PROJ_TABLE=None, CSMB_CHAIN_SMEM_KB=None, CSMB_CHAIN_FLAGS=None)


@section("lanes")
def _lanes():
    """G independent lanes of 64/G sequences, each a CUDA-graph replay of the chain on its own stream.  With a <= 100 KiB
    shared-memory budget two Linear CTAs fit one SM, so the lanes' latency-bound kernels can overlap."""
    total = int(os.environ.get("R2_LANES_TOTAL", "64"))
    for smem in (0, 96):
        setenv(CSMB_CHAIN_SMEM_KB=smem or None)
        for G in (1, 2, 4):
            B = total // G
            if B < 1:
                continue
            streams = [torch.cuda.Stream(dev) for _ in range(G)]
            states, prevs = [], []
            for gi in range(G):
                st, frame = warm_state(B, seed0=21 + gi * B)
                prev = frame
                for _ in range(2):
                    prev = st.decode_frame_graphed(prev, spec)
                states.append(st)
                prevs.append(prev)
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
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            print(f"lanes smem_kb={smem or 200} G={G} x B={B}: {ms:.3f} ms per step of all {total} -> {total * 0.08 / (ms / 1e3):.0f} audio-s/s",
                  flush=True)
            del states
    setenv(CSMB_CHAIN_SMEM_KB=None)

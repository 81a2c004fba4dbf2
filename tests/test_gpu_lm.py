"""LM parity on the GPU: csm_1b with the seeded random-init bf16 checkpoint, against the committed golden
vectors and the live oracle (fp32 math on the same bf16-rounded weights).

Stated tolerances: the product keeps activations in fp32 and accumulates in fp32, so it differs from the oracle
only by summation order.  Gates: |Δh| ≤ 1e-4, |Δlogit| ≤ 1e-4 (logit std ≈ 0.6-0.9; measured ≈ 1e-5);
greedy tokens identical for all 25 frames of BASELINE.json configs[0] (oracle's smallest top-2 margin 5.8e-4)."""
import os

import numpy as np
import pytest
import torch

from csm_mlx_b200 import generation, tokenizers
from csm_mlx_b200.runtime import LMState, SamplerSpec
from oracle import lm as olm
from oracle import sampling as osamp
from tests.conftest import GOLDEN
from tests.workloads import cfg1_prompt_ids, prompt_ids

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _prompt():
    return tokenizers.tokenize_text_segment(cfg1_prompt_ids(), 0)


def test_prefill_and_frame0_logits_vs_golden(model_1b, device):
    g = np.load(os.path.join(GOLDEN, "cfg1_lm.npz"))
    tok, mask = _prompt()
    st = LMState(model_1b, 1, max_len=64)
    st.prefill([tok], [mask])
    assert float((st.h_last.cpu()[0] - torch.from_numpy(g["h_last_f0"])).abs().max()) < TOL
    forced = torch.from_numpy(g["tokens"][:1]).to(device)
    frame = torch.zeros((1, 32), device=device, dtype=torch.int32)
    lg = torch.zeros((1, 32, 2051), device=device)
    st.depth_decode(frame, SamplerSpec(), logits_out=lg, forced=forced.contiguous())
    gl = torch.from_numpy(g["logits_f0"])
    assert float((st.c0_logits.cpu()[0] - gl[0]).abs().max()) < TOL
    assert float((lg.cpu()[0, 1:] - gl[1:]).abs().max()) < TOL
    assert torch.equal(frame.cpu()[0, 1:], forced.cpu()[0, 1:])


def test_greedy_tokens_cfg1_identical_to_golden(model_1b):
    """BASELINE.json configs[0]: 25 greedy frames, token-exact."""
    g = np.load(os.path.join(GOLDEN, "cfg1_lm.npz"))
    (toks,) = generation.generate_tokens(model_1b, [_prompt()], 25, temperature=0.0)
    assert toks.shape == (25, 32)
    assert np.array_equal(toks.numpy(), g["tokens"])


def test_teacher_forced_logits_all_frames_vs_oracle(model_1b, oracle_1b, device):
    """8 frames, teacher-forced with random tokens (includes ids ≥ 2048): every step's logits within tolerance."""
    gen = torch.Generator().manual_seed(123)
    forced = torch.randint(0, 2051, (8, 32), generator=gen)
    tok, mask = _prompt()
    traces = []
    olm.generate_tokens(oracle_1b, tok.long(), mask, 8, traces=traces, forced=forced)
    st = LMState(model_1b, 1, max_len=64)
    st.prefill([tok], [mask])
    worst = 0.0
    for f in range(8):
        fr = forced[f:f + 1].to(device, torch.int32).contiguous()
        frame = torch.zeros((1, 32), device=device, dtype=torch.int32)
        lg = torch.zeros((1, 32, 2051), device=device)
        st.depth_decode(frame, SamplerSpec(), logits_out=lg, forced=fr)
        ref = torch.stack([l[0] for l in traces[f]["logits"]])
        worst = max(worst, float((st.c0_logits.cpu()[0] - ref[0]).abs().max()), float((lg.cpu()[0, 1:] - ref[1:]).abs().max()))
        st.backbone_step(fr)
    assert worst < TOL


def test_prefill_on_chain_kernels_vs_per_op_and_oracle(model_1b, oracle_1b, device, monkeypatch):
    """csmb_prefill_fast (prompt rows through the chain's kernels: one tcgen05 launch per Linear, fused norms, per-row
    attention; the default) against the per-op kernels (CSMB_DISABLE_PREFILL_FAST=1) and the oracle, on a ragged batch of
    three prompts (text rows, and text + audio rows): last hidden rows, c0 logits and the KV cache within TOL (hidden rows
    against the oracle: relative L2); a row's result
    is independent of the rows it is batched with (bit-identical alone and in the batch)."""
    gen = torch.Generator().manual_seed(19)
    t1 = olm.text_rows(prompt_ids(31, 9))
    a1 = olm.audio_rows(torch.randint(0, 2048, (32, 40), generator=gen))
    t2 = olm.text_rows(prompt_ids(32, 5))
    prompts = [(t1[0].int(), t1[1]), (torch.cat([t1[0], a1[0], t2[0]]).int(), torch.cat([t1[1], a1[1], t2[1]])), (t2[0].int(), t2[1])]

    def run(ps):
        st = LMState(model_1b, len(ps), max_len=96)
        st.prefill([p[0] for p in ps], [p[1] for p in ps])
        torch.cuda.synchronize()
        st.check_status()
        return st.h_last.cpu().clone(), st.c0_logits.cpu().clone(), st.kv_pool.cpu().clone()

    h_f, lg_f, kv_f = run(prompts)
    h_1, lg_1, kv_1 = run(prompts[1:2])
    assert torch.equal(h_f[1], h_1[0]) and torch.equal(lg_f[1], lg_1[0])
    monkeypatch.setenv("CSMB_DISABLE_PREFILL_FAST", "1")
    h_p, lg_p, kv_p = run(prompts)
    rel = lambda a, b: float((a - b).norm() / b.norm())
    assert rel(h_f, h_p) < TOL and float((lg_f - lg_p).abs().max()) < 2 * TOL   # hidden / KV entries reach |x| ~ 5: relative L2
    assert rel(kv_f, kv_p) < TOL
    for i, (tok, mask) in enumerate(prompts):
        trace = {}
        olm.generate_frame(oracle_1b, tok.long()[None], mask[None], oracle_1b.new_backbone_cache(), trace=trace)
        # hidden rows reach |h| ~ 5: SURVEY.md §8d states their gate as relative L2 (1e-4); logits (|z| < 1) as absolute
        ref_h = trace["h"][0]
        assert float((h_f[i] - ref_h).norm() / ref_h.norm()) < TOL, i
        # tensor-core path: activations enter every Linear as bf16 hi + lo (2^-17 relative each); over 16 layers that adds
        # up to ~1e-4 on logits of magnitude ~1 (measured 1.1e-4).  Stated gate for this path: 2e-4 (bf16 would be 4e-3).
        assert float((lg_f[i] - trace["logits"][0][0]).abs().max()) < 2 * TOL, i


def test_batched_ragged_prompts_match_single(model_1b):
    """3 utterances with different prompt lengths in lock-step == each generated alone (request batching)."""
    prompts = [tokenizers.tokenize_text_segment(prompt_ids(21 + i, n), i) for i, n in enumerate((8, 13, 16))]
    batched = generation.generate_tokens(model_1b, prompts, 4, temperature=0.0)
    for p, b in zip(prompts, batched):
        (single,) = generation.generate_tokens(model_1b, [p], 4, temperature=0.0)
        assert torch.equal(single, b)


def test_mixed_text_audio_prompt_vs_oracle(model_1b, oracle_1b):
    """Context-style prompt: text rows, audio rows (+ zero EOS row), text rows — greedy tokens equal the oracle's."""
    gen = torch.Generator().manual_seed(9)
    t1 = olm.text_rows(prompt_ids(1, 6))
    a1 = olm.audio_rows(torch.randint(0, 2048, (32, 20), generator=gen))
    t2 = olm.text_rows(prompt_ids(2, 5))
    tok = torch.cat([t1[0], a1[0], t2[0]])
    mask = torch.cat([t1[1], a1[1], t2[1]])
    exp = olm.generate_tokens(oracle_1b, tok, mask, 3)
    (got,) = generation.generate_tokens(model_1b, [(tok.int(), mask)], 3, temperature=0.0)
    assert torch.equal(got.long(), exp)


def test_sampled_generation_matches_oracle_with_same_noise(model_1b, oracle_1b):
    """temperature 0.9 + top-k 50: device Gumbel/Philox sampling reproduces the oracle's definition token for token."""
    tok, mask = _prompt()
    spec = SamplerSpec(temperature=0.9, top_k=50, seed=1234)
    n_prompt = tok.shape[0]
    state = {"f": 0}

    def sampler(logits, i):
        pos = n_prompt - 1 + state["f"]
        t = osamp.sample(logits[0].numpy(), 0.9, seed=1234, draw=pos * 32 + i, row=0, top_k=50)
        if i == 31:
            state["f"] += 1
        return torch.tensor([t])

    exp = olm.generate_tokens(oracle_1b, tok.long(), mask, 3, sampler=sampler)
    (got,) = generation.generate_tokens(model_1b, [(tok, mask)], 3, sampler=spec)
    assert torch.equal(got.long(), exp)


def test_logits_processors_and_host_sampler_path(model_1b):
    """logit_bias forces c0; a foreign python sampler runs the per-codebook host path (README.md:120-122 usage)."""
    from csm_mlx_b200 import make_logits_processors

    procs = make_logits_processors(logit_bias={123: 1e4})
    (toks,) = generation.generate_tokens(model_1b, [_prompt()], 2, temperature=0.0, logits_processors=procs)
    assert toks[:, 0].tolist() == [123, 123]
    (ref,) = generation.generate_tokens(model_1b, [_prompt()], 2, temperature=0.0)
    (host,) = generation.generate_tokens(model_1b, [_prompt()], 2, sampler=lambda lg: torch.argmax(lg, dim=-1))
    assert torch.equal(host, ref)


def test_eos_stops_generation(model_1b, device):
    """An all-zero frame ends the utterance (generation.py:151-152): bias every head towards token 0."""
    import copy

    m = copy.copy(model_1b)
    saved = (m.codebook0_head.weight, m._audio_head_t)
    try:
        # logits = W h; make row 0 of each head a large multiple of the average direction is not possible in
        # general, so instead zero every head: all logits tie at 0 -> argmax picks index 0 -> EOS immediately.
        m.codebook0_head = type(model_1b.codebook0_head)()
        m.codebook0_head.weight = torch.zeros_like(saved[0])
        m._audio_head_t = torch.zeros_like(saved[1])
        m._desc = None
        (toks,) = generation.generate_tokens(m, [_prompt()], 5, temperature=0.0)
        assert toks.shape == (0, 32)
    finally:
        model_1b._desc = None


def test_inputs_too_long_raises(model_1b):
    tok, mask = tokenizers.tokenize_text_segment(prompt_ids(5, 1998), 0)  # 2000 rows
    with pytest.raises(ValueError, match="Inputs too long"):
        generation.generate_tokens(model_1b, [(tok, mask)], 125)


def test_fused_frame_kernel_equals_per_op_path(model_1b, monkeypatch):
    """The persistent frame kernel (csmb_frame_b1) and the per-op CUDA-graph path produce the same greedy tokens."""
    (fused,) = generation.generate_tokens(model_1b, [_prompt()], 6, temperature=0.0)
    monkeypatch.setenv("CSMB_DISABLE_FUSED", "1")
    (per_op,) = generation.generate_tokens(model_1b, [_prompt()], 6, temperature=0.0)
    assert torch.equal(fused, per_op)


def test_fused_frame_kernel_temperature_sampling_matches_oracle(model_1b, oracle_1b):
    """temperature 0.8, no filters: the frame kernel's in-kernel Gumbel/Philox sampling == the oracle's definition."""
    tok, mask = _prompt()
    n_prompt = tok.shape[0]
    state = {"f": 0}

    def sampler(logits, i):
        pos = n_prompt - 1 + state["f"]
        t = osamp.sample(logits[0].numpy(), 0.8, seed=77, draw=pos * 32 + i, row=0)
        if i == 31:
            state["f"] += 1
        return torch.tensor([t])

    exp = olm.generate_tokens(oracle_1b, tok.long(), mask, 4, sampler=sampler)
    (got,) = generation.generate_tokens(model_1b, [(tok, mask)], 4, sampler=SamplerSpec(temperature=0.8, seed=77))
    assert torch.equal(got.long(), exp)


def test_long_context_multi_chunk_attention_vs_oracle(model_1b, oracle_1b):
    """150-row prompt (> 128 keys: the frame kernel's multi-chunk attention + merge path; prefill through the tiled
    linear): greedy tokens equal the oracle's."""
    gen = torch.Generator().manual_seed(31)
    t1 = olm.text_rows(prompt_ids(3, 8))
    a1 = olm.audio_rows(torch.randint(0, 2048, (32, 128), generator=gen))
    t2 = olm.text_rows(prompt_ids(4, 9))
    tok = torch.cat([t1[0], a1[0], t2[0]])
    mask = torch.cat([t1[1], a1[1], t2[1]])
    assert tok.shape[0] == 150
    exp = olm.generate_tokens(oracle_1b, tok, mask, 3)
    (got,) = generation.generate_tokens(model_1b, [(tok.int(), mask)], 3, temperature=0.0)
    assert torch.equal(got.long(), exp)


def test_batch16_tensor_core_path_matches_single(model_1b):
    """16 utterances in lock-step: every linear runs on the tcgen05 path (rows >= 16); each must equal the
    utterance generated alone through the batch-1 kernels."""
    prompts = [tokenizers.tokenize_text_segment(prompt_ids(40 + i, 8 + (i % 5)), i % 3) for i in range(16)]
    batched = generation.generate_tokens(model_1b, prompts, 3, temperature=0.0)
    for i in (0, 7, 15):
        (single,) = generation.generate_tokens(model_1b, [prompts[i]], 3, temperature=0.0)
        assert torch.equal(single, batched[i]), i


def test_batch64_fused_chain_matches_single(model_1b):
    """BASELINE.json configs[3] shape: 64 utterances in lock-step through the fused kernel chain of
    csrc/batch_frame.cu (tcgen05 linears on bf16 hi+lo planes, fused partial-sum kernels, programmatic dependent
    launch); utterances must equal the ones generated alone by the batch-1 persistent kernel."""
    prompts = [tokenizers.tokenize_text_segment(prompt_ids(21 + i, 8 + (i % 9)), i % 4) for i in range(64)]
    batched = generation.generate_tokens(model_1b, prompts, 4, temperature=0.0)
    for i in (0, 31, 63):
        (single,) = generation.generate_tokens(model_1b, [prompts[i]], 4, temperature=0.0)
        assert torch.equal(single, batched[i]), i


@pytest.mark.parametrize("spec", [SamplerSpec(temperature=0.0), SamplerSpec(temperature=0.8, seed=5),
                                  SamplerSpec(temperature=0.9, top_k=50, seed=7), SamplerSpec(temperature=0.9, top_k=20, min_p=0.03, seed=8),
                                  SamplerSpec(temperature=0.8, top_p=0.9, seed=9), SamplerSpec(temperature=1.1, top_k=64, top_p=0.7, seed=10)])
def test_fused_chain_equals_per_op_batched_path(model_1b, monkeypatch, spec):
    """Same batch through csmb_decode_frame_fast and through the per-op csmb_decode_frame (CSMB_DISABLE_FAST=1):
    identical tokens, greedy and with in-kernel Gumbel/Philox temperature sampling."""
    prompts = [tokenizers.tokenize_text_segment(prompt_ids(60 + i, 6 + i), i % 2) for i in range(5)]
    fast = generation.generate_tokens(model_1b, prompts, 4, sampler=spec)
    monkeypatch.setenv("CSMB_DISABLE_FAST", "1")
    slow = generation.generate_tokens(model_1b, prompts, 4, sampler=spec)
    for a, b in zip(fast, slow):
        assert torch.equal(a, b)


def test_fused_chain_filtered_sampler_falls_back(model_1b):
    """min-p with min_tokens_to_keep > 1 (it needs the sorted order) is not fused: the batched path must route to the
    per-op frame (csmb_decode_frame) and still work."""
    from csm_mlx_b200.runtime import LMState

    st = LMState(model_1b, 3, max_len=32)
    assert st.fast_supported(SamplerSpec(temperature=0.0))
    assert st.fast_supported(SamplerSpec(temperature=0.7))
    assert st.fast_supported(SamplerSpec(temperature=0.7, top_k=50))
    assert st.fast_supported(SamplerSpec(temperature=0.7, top_p=0.9))
    assert not st.fast_supported(SamplerSpec(temperature=0.7, min_p=0.05, min_tokens_to_keep=2))
    prompts = [tokenizers.tokenize_text_segment(prompt_ids(70 + i, 7), 0) for i in range(3)]
    toks = generation.generate_tokens(model_1b, prompts, 2, sampler=SamplerSpec(temperature=0.7, min_p=0.05, min_tokens_to_keep=2, seed=3))
    assert all(t.shape == (2, 32) for t in toks)


@pytest.mark.parametrize("B", [5, 130])
def test_chain_swiglu_epilogue_equals_separate_launch(model_1b, device, monkeypatch, B):
    """The gate|up Linear with SwiGLU in its epilogue (k_gemm_part_t<true>: 64 gate rows + the 64 matching up rows per
    UMMA tile, planes written from the drained pipeline stages) against the separate k_swiglu_split launch (debug flag
    4): identical tokens for a ragged row count (5 rows in a 16-row token tile) and for 130 sequences (first depth
    step: 260 rows = three 128-row token tiles)."""
    from csm_mlx_b200 import _lib
    from csm_mlx_b200.runtime import LMState

    spec = SamplerSpec(temperature=0.0)
    prompts = [tokenizers.tokenize_text_segment(prompt_ids(21 + i, 8 + i % 9), 0) for i in range(B)]

    def run(flags):
        monkeypatch.setenv("CSMB_CHAIN_FLAGS", str(flags))   # csmb_chain_opts.flags of the states created from here on
        st = LMState(model_1b, B, max_len=48)
        assert st.fast_supported(spec)
        st.prefill([p[0] for p in prompts], [p[1] for p in prompts])
        frame = torch.zeros((B, 32), device=device, dtype=torch.int32)
        st.sample_c0(frame, spec)
        st.depth_decode(frame, spec)
        out, prev = [frame.clone()], frame
        for _ in range(2):
            nxt = torch.zeros((B, 32), device=device, dtype=torch.int32)
            st.decode_frame(prev, nxt, spec)
            out.append(nxt.clone())
            prev = nxt
        torch.cuda.synchronize()
        st.check_status()
        return torch.stack(out).cpu()

    separate, fused = run(4), run(0)
    assert torch.equal(separate, fused)
    assert int((fused[1:] != fused[:-1]).sum()) > 0  # frames differ from step to step: the loop really decoded


@pytest.mark.parametrize("flags", [8, 16, 32 | 64 | 128 | 256, 65536 | 131072])
def test_chain_staged_kernels_are_bit_identical_to_the_plain_ones(model_1b, device, monkeypatch, flags):
    """The shared-memory staged kernels of the chain (k_attn_decode_small: the depth decoder's whole cache requested at once
    with cp.async; k_attn_decode_chunked: the backbone's cache in 64-position chunks; cp.async staged split-K partial sums in
    the norm / sampling kernels) and its L2 hints do the round-1 kernels' sums in the same order: identical tokens against
    csmb_chain_opts.flags 8 (plain attention kernel), 16 (plain partial sums), the hints switched off, and 65536 | 131072 (every
    Linear with an even number of n-tiles as CTA pairs: one tcgen05.mma.cta_group::2 of M = 256 per K step) — for ragged
    prompts whose caches cross the 64- and 128-position chunk boundaries while decoding (3 sequences, 6 frames)."""
    spec = SamplerSpec(temperature=0.0)
    lens = (58, 61, 125)   # + BOS, EOS: 60 / 63 / 127 rows -> positions 60..65, 63..68, 127..132 while decoding
    prompts = [tokenizers.tokenize_text_segment(prompt_ids(300 + i, n), i) for i, n in enumerate(lens)]

    def run(f):
        monkeypatch.setenv("CSMB_CHAIN_FLAGS", str(f))
        st = LMState(model_1b, len(prompts), max_len=160, row_invariant=True)
        assert st.fast_supported(spec)
        st.prefill([p[0] for p in prompts], [p[1] for p in prompts])
        frame = torch.zeros((len(prompts), 32), device=device, dtype=torch.int32)
        st.sample_c0(frame, spec)
        st.depth_decode(frame, spec)
        out, prev = [frame.clone()], frame
        for _ in range(6):
            nxt = torch.zeros_like(prev)
            st.decode_frame(prev, nxt, spec)
            out.append(nxt.clone())
            prev = nxt
        torch.cuda.synchronize()
        st.check_status()
        return torch.stack(out).cpu()

    shipped, plain = run(0), run(flags)
    assert torch.equal(shipped, plain)
    assert int((shipped[1:] != shipped[:-1]).sum()) > 0


def test_chain_chunked_attention_long_cache_equals_plain_kernel(model_1b, device, monkeypatch):
    """k_attn_decode_chunked over many 64-position chunks (caches of 705 and 1 300 positions, a 9-row one beside them) against
    the round-1 kernel (flag 8): identical tokens over 3 frames."""
    spec = SamplerSpec(temperature=0.0)
    prompts = [tokenizers.tokenize_text_segment(prompt_ids(400 + i, n), 0) for i, n in enumerate((703, 7, 1298))]

    def run(f):
        monkeypatch.setenv("CSMB_CHAIN_FLAGS", str(f))
        st = LMState(model_1b, len(prompts), max_len=1312, row_invariant=True)
        assert st.fast_supported(spec)
        st.prefill([p[0] for p in prompts], [p[1] for p in prompts])
        frame = torch.zeros((len(prompts), 32), device=device, dtype=torch.int32)
        st.sample_c0(frame, spec)
        st.depth_decode(frame, spec)
        out, prev = [frame.clone()], frame
        for _ in range(3):
            nxt = torch.zeros_like(prev)
            st.decode_frame(prev, nxt, spec)
            out.append(nxt.clone())
            prev = nxt
        torch.cuda.synchronize()
        st.check_status()
        return torch.stack(out).cpu()

    assert torch.equal(run(0), run(8))


def test_chain_projected_embedding_table_is_bit_exact(model_1b, monkeypatch):
    """csmb_build_proj_table: depth steps >= 2 read projection(embed_audio(cb, token)) rows from a table built with the
    chain's own projection Linear instead of running that Linear (60 launches less per frame-step): identical tokens,
    greedy and sampled, ragged batch of 20."""
    prompts = [tokenizers.tokenize_text_segment(prompt_ids(80 + i, 7 + i % 6), i % 3) for i in range(20)]
    spec = SamplerSpec(temperature=0.9, top_k=50, seed=11)
    with_table = generation.generate_tokens(model_1b, prompts, 6, temperature=0.0)
    sampled_t = generation.generate_tokens(model_1b, prompts[:5], 3, sampler=spec)
    assert model_1b._proj_table is not None and tuple(model_1b._proj_table.shape) == (32, 2051, 1024)
    monkeypatch.setenv("CSMB_NO_PROJ_TABLE", "1")
    without = generation.generate_tokens(model_1b, prompts, 6, temperature=0.0)
    sampled_w = generation.generate_tokens(model_1b, prompts[:5], 3, sampler=spec)
    for a, b in zip(with_table + sampled_t, without + sampled_w):
        assert torch.equal(a, b)


def test_cfg2_full_length_two_implementations_agree(model_1b, monkeypatch):
    """BASELINE.json configs[1] at full size (125 frames = 10 s): the persistent frame kernel and the per-op CUDA-graph
    path are independent implementations of the frame; they must agree on all 4 000 tokens, the first 25 frames must
    equal the oracle's golden tokens, and a second run must reproduce the first bit for bit."""
    g = np.load(os.path.join(GOLDEN, "cfg1_lm.npz"))
    (fused,) = generation.generate_tokens(model_1b, [_prompt()], 125, temperature=0.0)
    (again,) = generation.generate_tokens(model_1b, [_prompt()], 125, temperature=0.0)
    monkeypatch.setenv("CSMB_DISABLE_FUSED", "1")
    (per_op,) = generation.generate_tokens(model_1b, [_prompt()], 125, temperature=0.0)
    assert fused.shape == (125, 32)
    assert torch.equal(fused, again)
    assert torch.equal(fused, per_op)
    assert np.array_equal(fused.numpy()[:25], g["tokens"])


@pytest.mark.parametrize("spec", [SamplerSpec(temperature=0.9, top_k=50, seed=21), SamplerSpec(temperature=0.7, min_p=0.05, seed=22),
                                  SamplerSpec(temperature=1.1, top_k=8, min_p=0.02, seed=23), SamplerSpec(temperature=0.9, top_p=0.9, seed=24),
                                  SamplerSpec(temperature=1.2, top_k=40, top_p=0.8, min_p=0.01, seed=25)])
def test_fused_frame_kernel_topk_minp_equals_per_op_sampler(model_1b, monkeypatch, spec):
    """top-k / top-p / min-p inside the persistent frame kernel (exact k-th-largest search by bisection on the bit pattern
    of exp(logit - max); nucleus by the same bisection on exact integer masses) selects exactly the tokens of the per-op
    sampler (k_sample_filtered: sorted probabilities), so with the same Philox noise both paths generate the same frames
    (the README / CLI sampler, cli/generate.py:168-174)."""
    from csm_mlx_b200.runtime import LMState

    assert LMState(model_1b, 1, max_len=32).fused_supported(spec)
    assert not LMState(model_1b, 1, max_len=32).fused_supported(SamplerSpec(temperature=0.9, min_p=0.1, min_tokens_to_keep=2))
    (fused,) = generation.generate_tokens(model_1b, [_prompt()], 6, sampler=spec)
    monkeypatch.setenv("CSMB_DISABLE_FUSED", "1")
    (per_op,) = generation.generate_tokens(model_1b, [_prompt()], 6, sampler=spec)
    assert torch.equal(fused, per_op)


def test_near_maximum_context_fused_equals_per_op(model_1b, monkeypatch):
    """Edge of the position range (2 048, attention.py:38): a 2 030-row prompt (prefill through the tensor-core linear in
    row tiles, 127 KV pages) followed by 6 frames up to position 2 036 — the frame kernel's 16-chunk attention + merge
    against the per-op kernels, token for token; one more frame than the window allows is refused like the reference."""
    gen = torch.Generator().manual_seed(77)
    t1 = olm.text_rows(prompt_ids(9, 10))
    a1 = olm.audio_rows(torch.randint(0, 2048, (32, 2005), generator=gen))
    t2 = olm.text_rows(prompt_ids(10, 10))
    tok = torch.cat([t1[0], a1[0], t2[0]]).int()
    mask = torch.cat([t1[1], a1[1], t2[1]])
    assert tok.shape[0] == 2030
    (fused,) = generation.generate_tokens(model_1b, [(tok, mask)], 6, temperature=0.0)
    monkeypatch.setenv("CSMB_DISABLE_FUSED", "1")
    (per_op,) = generation.generate_tokens(model_1b, [(tok, mask)], 6, temperature=0.0)
    assert fused.shape == (6, 32) and torch.equal(fused, per_op)
    with pytest.raises(ValueError, match="Inputs too long"):
        generation.generate_tokens(model_1b, [(tok, mask)], 18, temperature=0.0)

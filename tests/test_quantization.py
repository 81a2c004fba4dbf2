"""Weight-only FP8 (csm_mlx_b200.quantize — what /root/reference README.md:92-128 does with mlx.nn.quantize(csm)).

CPU: the quantiser (per-output-channel scales, E4M3 round-to-nearest, blob layout = include/csm_b200.h).
GPU: csmb_linear_e4m3 against float64 on the dequantised matrix; the quantised csm_1b against the oracle run on exactly the
dequantised weights (model.parameters()): teacher-forced logits within 1e-4 over 4 frames x 32 codebooks, greedy tokens
identical up to a first difference, which must be a near-tie of the oracle (top-2 margin < 1e-3)."""
import ctypes as C

import numpy as np
import pytest
import torch

from csm_mlx_b200 import _lib, quantization as qz


def test_quantiser_is_exact_on_e4m3_values_and_bounded_elsewhere():
    g = torch.Generator().manual_seed(5)
    w = torch.randn((37, 64), generator=g) * 0.02
    q, s = qz.quantize_rows_e4m3(w)
    assert q.dtype == torch.uint8 and q.shape == w.shape and s.shape == (37,)
    d = qz.dequantize_rows_e4m3(q, s)
    # the row maximum maps to +-448 exactly; every other entry is within half an e4m3 step (2^-4 relative) of its value
    assert torch.allclose(d.abs().amax(1), w.abs().amax(1), rtol=1e-6)
    assert float(((d - w).abs() / (w.abs() + s[:, None] * 2 ** -9)).max()) <= 2 ** -4 + 1e-6
    # a matrix that already consists of scale * e4m3 values survives the round trip bit for bit
    q2, s2 = qz.quantize_rows_e4m3(d)
    assert torch.equal(q2, q) and torch.equal(qz.dequantize_rows_e4m3(q2, s2), d)
    # zero rows do not divide by zero
    qz0, sz0 = qz.quantize_rows_e4m3(torch.zeros((2, 16)))
    assert torch.equal(qz.dequantize_rows_e4m3(qz0, sz0), torch.zeros((2, 16)))


def test_blob_layout_matches_the_header():
    n, k = 37, 64
    g = torch.Generator().manual_seed(6)
    q, s = qz.quantize_rows_e4m3(torch.randn((n, k), generator=g))
    blob = qz.pack_blob(q, s)
    assert blob.numel() == qz.blob_bytes(n, k) == 256 + 2560 and blob.numel() % 256 == 0
    assert blob.numel() == _lib.lib().csmb_e4m3_blob_bytes(n, k)   # host-side helper of the library: no device needed
    q2, s2 = qz.unpack_blob(blob, n, k)
    assert torch.equal(q2, q) and torch.equal(s2, s)
    with pytest.raises(NotImplementedError):
        qz.quantize(object(), bits=4)


@pytest.mark.gpu
@pytest.mark.parametrize("R", [1, 3, 8, 21])
def test_linear_e4m3_matches_float64(device, R):
    n, k = 203, 1024     # ragged feature count: the last block has 3 live warps
    g = torch.Generator().manual_seed(40 + R)
    w = torch.randn((n, k), generator=g) * 0.02
    x = torch.randn((R, k), generator=g)
    q, s = qz.quantize_rows_e4m3(w)
    blob = qz.pack_blob(q, s).to(device)
    xd = x.to(device)
    y0 = torch.randn((R, n), generator=g)
    ref = x.double() @ qz.dequantize_rows_e4m3(q, s).double().t()
    for acc in (0, 1):
        y = y0.clone().to(device)
        _lib.check(_lib.lib().csmb_linear_e4m3(xd.data_ptr(), k, blob.data_ptr(), y.data_ptr(), n, R, n, k, acc, 0,
                                               _lib.stream_ptr(device)))
        want = ref + (y0.double() if acc else 0)
        assert float((y.cpu().double() - want).abs().max()) < 1e-5 * float(want.abs().max() + 1)


@pytest.fixture(scope="module")
def quantized_pair(csm_weights, device):
    """(quantised CSM on the GPU, oracle holding exactly its dequantised weights)."""
    from csm_mlx_b200 import CSM, csm_1b, quantize
    from oracle import lm as olm

    model = quantize(CSM(csm_1b(), device=device).load_weights(csm_weights))
    assert model.quantized and model.proj_table() is None
    weights = {k: v.detach().to("cpu", torch.float32) for k, v in model.parameters().items()}
    return model, olm.OracleCSM(olm.CSM_1B, weights)


@pytest.mark.gpu
def test_quantized_model_halves_the_linear_bytes_and_declines_the_fused_kernels(quantized_pair, model_1b, csm_weights):
    from csm_mlx_b200.runtime import LMState, SamplerSpec

    model, _ = quantized_pair
    lin_bytes = lambda m: sum(t.numel() * t.element_size() for st in (m.backbone, m.decoder)
                              for ts in (st.wqkv, st.wo, st.wgu, st.wdown) for t in ts)
    assert 0.50 <= lin_bytes(model) / lin_bytes(model_1b) <= 0.51       # e4m3 bytes + one fp32 scale per output channel
    st = LMState(model, 1, max_len=48)
    # the batch-1 frame kernel reads e4m3 blobs; the batched tensor-core chain (and the prompt pass on its kernels) declines them
    assert st.fused_supported(SamplerSpec()) and not st.fast_supported(SamplerSpec()) and not st._prefill_fast_ok()
    st2 = LMState(model, 3, max_len=48)
    assert not st2.fast_supported(SamplerSpec())
    with pytest.raises(RuntimeError):
        model.load_weights(csm_weights)


@pytest.mark.gpu
def test_quantized_teacher_forced_logits_vs_oracle_on_dequantised_weights(quantized_pair, device):
    from csm_mlx_b200 import tokenizers
    from csm_mlx_b200.runtime import LMState, SamplerSpec
    from oracle import lm as olm
    from tests.workloads import cfg1_prompt_ids

    model, oracle = quantized_pair
    gen = torch.Generator().manual_seed(77)
    forced = torch.randint(0, 2051, (4, 32), generator=gen)
    tok, mask = tokenizers.tokenize_text_segment(cfg1_prompt_ids(), 0)
    traces = []
    olm.generate_tokens(oracle, tok.long(), mask, 4, traces=traces, forced=forced)
    st = LMState(model, 1, max_len=48)
    st.prefill([tok], [mask])
    worst = 0.0
    for f in range(4):
        fr = forced[f:f + 1].to(device, torch.int32).contiguous()
        frame = torch.zeros((1, 32), device=device, dtype=torch.int32)
        lg = torch.zeros((1, 32, 2051), device=device)
        st.depth_decode(frame, SamplerSpec(), logits_out=lg, forced=fr)
        ref = torch.stack([l[0] for l in traces[f]["logits"]])
        worst = max(worst, float((st.c0_logits.cpu()[0] - ref[0]).abs().max()), float((lg.cpu()[0, 1:] - ref[1:]).abs().max()))
        st.backbone_step(fr)
    assert worst < 1e-4, worst


@pytest.mark.gpu
def test_quantized_frame_kernel_equals_row_based_path(quantized_pair, monkeypatch):
    """The persistent frame kernel on e4m3 blobs (k_frame: the same units at one byte per weight, scale on the finished
    dot product) and the row-based GEMV path (CSMB_DISABLE_FUSED=1) are two implementations of the quantised frame: same
    greedy tokens over 10 frames; a batch of 3 (row-based path only) reproduces the single utterance."""
    from csm_mlx_b200 import generation, tokenizers
    from tests.workloads import cfg1_prompt_ids, prompt_ids

    model, _ = quantized_pair
    p0 = tokenizers.tokenize_text_segment(cfg1_prompt_ids(), 0)
    (fused,) = generation.generate_tokens(model, [p0], 10, temperature=0.0)
    monkeypatch.setenv("CSMB_DISABLE_FUSED", "1")
    (rows,) = generation.generate_tokens(model, [p0], 10, temperature=0.0)
    assert torch.equal(fused, rows)
    others = [tokenizers.tokenize_text_segment(prompt_ids(500 + i, 7 + i), 1) for i in range(2)]
    batch = generation.generate_tokens(model, [p0] + others, 4, temperature=0.0)
    assert torch.equal(batch[0], rows[:4])


@pytest.mark.gpu
def test_quantized_greedy_generation_vs_oracle(quantized_pair):
    from csm_mlx_b200 import generation, tokenizers
    from oracle import lm as olm
    from tests.workloads import cfg1_prompt_ids

    model, oracle = quantized_pair
    tok, mask = tokenizers.tokenize_text_segment(cfg1_prompt_ids(), 0)
    frames = 12
    (got,) = generation.generate_tokens(model, [(tok, mask)], frames, temperature=0.0)
    traces = []
    exp = olm.generate_tokens(oracle, tok.long(), mask, frames, traces=traces)
    got, exp = got.long(), exp.long()
    assert got.shape == exp.shape == (frames, 32)
    if not torch.equal(got, exp):
        # free-running sequences diverge after a flipped near-tie: identical up to there, and there the oracle's own
        # top-2 margin must be below 1e-3 (float-level difference of the two summation orders)
        diff = (got != exp).nonzero()
        f, c = int(diff[0, 0]), int(diff[0, 1])
        assert torch.equal(got[:f], exp[:f]) and torch.equal(got[f, :c], exp[f, :c])
        top2 = traces[f]["logits"][c][0].topk(2).values
        assert float(top2[0] - top2[1]) < 1e-3, (f, c, top2)
    assert f"{model.backbone.wqkv[0].dtype}" == "torch.uint8"
    # the bf16 model gives other tokens: the mode really changes the weights
    assert np.unique(got.numpy()).size > 32


@pytest.mark.gpu
def test_quantized_model_through_the_engine(quantized_pair, mimi_gpu):
    """serving.Engine with a weight-only FP8 model: the chain declines it, the engine serves the requests on the row-based
    path (continuous batching unchanged); every request's tokens equal the utterance generated alone."""
    from csm_mlx_b200 import generation, tokenizers
    from csm_mlx_b200.serving import Engine
    from tests.workloads import prompt_ids

    model, _ = quantized_pair
    prompts = [tokenizers.tokenize_text_segment(prompt_ids(600 + i, 6 + 2 * i), i % 2) for i in range(3)]
    eng = Engine(model, max_batch=2, max_len=64)          # 3 requests through 2 slots: the third is admitted mid-run
    ids = [eng.submit_prompt(t, m, 4 + i) for i, (t, m) in enumerate(prompts)]
    eng.run()
    for i, rid in enumerate(ids):
        (alone,) = generation.generate_tokens(model, [prompts[i]], 4 + i, temperature=0.0)
        assert torch.equal(alone, eng.tokens(rid)), i
    wav = eng.audio(ids[:1])[0]
    assert wav.shape == (4 * 1920,) and bool(torch.isfinite(wav).all())

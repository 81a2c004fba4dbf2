"""Host-side logic of the product package, no GPU: API surface, frame assembly, RoPE table, guards, samplers."""
import os
import re
import wave

import numpy as np
import pytest
import torch

import csm_mlx
from csm_mlx_b200 import CSM, Segment, csm_1b, csm_tiny, tokenizers
from csm_mlx_b200.attention import llama3_rope_table
from csm_mlx_b200.random_init import csm_param_shapes, mimi_param_shapes, random_csm_weights
from oracle import lm as olm
from oracle import sampling as osamp


def test_export_surface_matches_reference():
    # /root/reference/csm_mlx/__init__.py:6-16
    assert csm_mlx.__all__ == ["generate", "stream_generate", "CSM", "csm_1b", "Segment", "CSMDataset", "CSMTrainer",
                               "TrainArgs", "load_adapters"]
    from csm_mlx.generation import generate, generate_frame, stream_generate  # noqa: F401
    from csm_mlx.tokenizers import decode_audio, tokenize_audio, tokenize_segment, tokenize_text_segment  # noqa: F401
    from csm_mlx.utils import read_audio, write_audio  # noqa: F401
    with pytest.raises(NotImplementedError):
        csm_mlx.CSMTrainer()
    with pytest.raises(FileNotFoundError):
        csm_mlx.load_adapters(None, "/nonexistent/adapter")


def test_model_args_and_param_tree():
    a = csm_1b()
    assert (a.backbone_name, a.decoder_name, a.n_text_vocab, a.n_audio_vocab, a.n_audio_codebooks) == \
        ("1b", "100m", 128256, 2051, 32)
    shapes = {n: s for n, s, _ in csm_param_shapes()}
    assert shapes["audio_head"] == (31, 1024, 2051)
    assert shapes["backbone.layers.15.mlp.down_proj.weight"] == (2048, 8192)
    assert shapes["decoder.layers.3.self_attn.k_proj.weight"] == (256, 1024)
    n = sum(int(np.prod(s)) for s in shapes.values())
    assert abs(n - 1552.8e6) < 1e6
    m = CSM(a)
    assert len(m.backbone.layers) == 16 and len(m.decoder.layers) == 4
    assert m.backbone.args.max_position_embeddings is None
    n_mimi = sum(int(np.prod(s)) for _, s, _ in mimi_param_shapes())
    assert abs(n_mimi - 96.15e6) < 0.1e6


def test_load_weights_strictness():
    m = CSM(csm_tiny())
    W = random_csm_weights(csm_tiny(), seed=1)
    bad = dict(W)
    bad["nope"] = torch.zeros(1)
    with pytest.raises(ValueError, match="not in model"):
        m.load_weights(bad)
    miss = dict(W)
    del miss["audio_head"]
    with pytest.raises(ValueError, match="Missing"):
        m.load_weights(miss)
    shp = dict(W)
    shp["projection.weight"] = torch.zeros(3, 3)
    with pytest.raises(ValueError, match="shape"):
        m.load_weights(shp)
    with pytest.raises(ValueError, match="Unsupported"):
        m.load_weights("weights.bin")
    with pytest.raises(RuntimeError, match="not loaded"):
        m.parameters()


def test_random_init_is_deterministic_and_fills_everything():
    a, b = random_csm_weights(csm_tiny(), seed=3), random_csm_weights(csm_tiny(), seed=3)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert float(a["audio_head"].float().abs().max()) > 0  # the reference leaves audio_head at zeros (models.py:65-67)


@pytest.mark.parametrize("hd", [64, 128])
def test_rope_table_matches_oracle_bit_for_bit(hd):
    t = llama3_rope_table(hd, 500_000.0, 32.0, 2048)
    o = olm.rope_table(hd, 500_000.0, 32.0, 2048)
    assert t.shape == (2048, hd // 2, 2) and torch.equal(t, o)
    # attention.py:102-117: high frequencies untouched, low ones divided by 32
    base = 1.0 / (500_000.0 ** (torch.arange(0, hd, 2).float() / hd))
    f = olm.rope_scaled_freqs(hd, 500_000.0, 32.0)
    lo, hi = (14, 18) if hd == 64 else (28, 35)
    assert torch.equal(f[: lo + 1], base[: lo + 1])
    torch.testing.assert_close(f[hi:], base[hi:] / 32.0)


def test_text_and_audio_rows():
    tok, mask = tokenizers.tokenize_text_segment([128000, 5, 6, 128001], speaker=0)
    assert tok.shape == (4, 33) and tok.dtype == torch.int32
    assert tok[:, 32].tolist() == [128000, 5, 6, 128001] and int(tok[:, :32].abs().sum()) == 0
    assert mask[:, 32].all() and not mask[:, :32].any()
    otok, omask = olm.text_rows([128000, 5, 6, 128001])
    assert torch.equal(tok.long(), otok) and torch.equal(mask, omask)

    class FakeMimi:
        def encode(self, x):
            assert x.shape == (1, 1, 4000)
            return torch.arange(32 * 3, dtype=torch.int32).reshape(1, 32, 3) + 1

    tokenizers.set_audio_tokenizer(FakeMimi())
    try:
        at, am = tokenizers.tokenize_audio(torch.zeros(4000))
        assert at.shape == (4, 33) and int(at[3].abs().sum()) == 0  # appended EOS frame (tokenizers.py:73-75)
        assert am[:, :32].all() and not am[:, 32].any()
        oat, oam = olm.audio_rows(torch.arange(32 * 3).reshape(32, 3) + 1)
        assert torch.equal(at.long(), oat) and torch.equal(am, oam)
        seg = Segment(1, "x", torch.zeros(4000))
        tokenizers.set_text_tokenizer(tokenizers.SyntheticTextTokenizer())
        st, sm = tokenizers.tokenize_segment(seg)
        n_text = len(tokenizers.SyntheticTextTokenizer().encode("[1]x"))
        assert st.shape == (n_text + 4, 33) and st[0, 32] == 128000 and st[n_text - 1, 32] == 128001
    finally:
        tokenizers.set_audio_tokenizer(None)
        tokenizers.set_text_tokenizer(None)


def test_segment_contract(tmp_path):
    s = Segment(0, "hi", None, None)  # positional order of cli/generate.py:186-189
    with pytest.raises(ValueError):
        _ = s.audio
    s.audio = torch.ones(3)
    assert s.audio.shape == (3,)
    # audio_path is read + resampled to 24 kHz mono on access (segment.py:23-28, utils.py:9-21)
    p = tmp_path / "a.wav"
    sr = 48000
    sig = (np.sin(2 * np.pi * 440 * np.arange(sr) / sr) * 0.5 * 32767).astype("<i2")
    with wave.open(str(p), "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(sr)
        w.writeframes(np.stack([sig, sig], 1).tobytes())
    a = Segment(0, "hi", None, p).audio
    assert a.dtype == torch.float32 and a.shape == (24000,) and 0.3 < float(a.abs().max()) < 0.6
    from csm_mlx_b200.utils import read_audio, write_audio
    q = tmp_path / "b.wav"
    write_audio(a, q, 24000)
    b = read_audio(q, 24000)
    assert b.shape == a.shape and float((a - b).abs().max()) < 1e-3


def test_prompt_length_guard():
    from csm_mlx_b200.generation import _check_length

    m = CSM(csm_1b())
    _check_length(m, 10, 125)
    with pytest.raises(ValueError, match="Inputs too long"):
        _check_length(m, 2048 - 125, 125)  # generation.py:132-137


def test_no_cpu_fallback():
    from csm_mlx_b200 import _lib

    with pytest.raises(_lib.CsmbError):
        _lib.require_device(torch.device("cpu"))
    with pytest.raises(_lib.CsmbError):
        _lib.ptr(torch.zeros(4))
    if not torch.cuda.is_available():
        m = CSM(csm_tiny())
        with pytest.raises(Exception):
            m.load_weights(random_csm_weights(csm_tiny(), seed=1))  # must not silently run on the CPU


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for base in ("csm_mlx_b200", "csm_mlx"):
        for dirpath, _, files in os.walk(os.path.join(root, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f


def test_philox_known_answers():
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]  # Random123 kat_vectors, philox4x32 10 rounds
    for ctr, key, out in kat:
        r = osamp.philox4x32_10(np.array(ctr, dtype=np.uint32), key)
        assert tuple(int(x) for x in r) == out


def test_nucleus_by_bisection_equals_sorted_definition():
    """The fused samplers (frame_kernel.cu::sample_token, batch_frame.cu::k_sample_embed) do not sort: they find the
    largest bit pattern T whose integer mass sum(q : e >= T) still reaches top_p * sum(q) by a 31-step greedy bisection,
    and keep e >= T.  Restated here in numpy and compared with the sorted definition of oracle/sampling.py (the one
    k_sample_filtered implements), including ties, one-hot and flat distributions."""
    rng = np.random.default_rng(3)

    def by_bisection(logits, top_p):
        e = np.exp((logits - logits.max()).astype(np.float32), dtype=np.float32)
        bits = e.view(np.uint32).astype(np.int64)
        q = np.floor(e.astype(np.float64) * 4294967296.0).astype(np.uint64)
        need = float(np.float32(top_p)) * float(int(q.sum(dtype=np.uint64)))
        T = 0
        for bit in range(30, -1, -1):
            cand = T | (1 << bit)
            if float(int(q[bits >= cand].sum(dtype=np.uint64))) >= need:
                T = cand
        return bits >= T

    cases = [rng.standard_normal(2051).astype(np.float32) * s for s in (0.3, 2.0, 6.0)]
    cases.append(np.zeros(2051, dtype=np.float32))                                   # flat: everything ties
    one_hot = np.full(2051, -60.0, dtype=np.float32); one_hot[77] = 0.0
    cases.append(one_hot)
    tied = rng.standard_normal(2051).astype(np.float32); tied[::3] = tied[0]          # large tie groups
    cases.append(tied)
    for lg in cases:
        for top_p in (0.05, 0.5, 0.9, 0.999):
            assert np.array_equal(by_bisection(lg, top_p), osamp.keep_mask(lg, top_p=top_p)), top_p


def test_oracle_sampler_filters():
    lg = np.log(np.array([0.5, 0.2, 0.15, 0.1, 0.05], dtype=np.float32))
    assert osamp.keep_mask(lg, top_k=2).tolist() == [True, True, False, False, False]
    assert osamp.keep_mask(lg, top_p=0.6).tolist() == [True, True, False, False, False]   # 0 < .6, .5 < .6, .7 !< .6
    assert osamp.keep_mask(lg, min_p=0.25).tolist() == [True, True, True, False, False]   # p >= .125
    assert osamp.keep_mask(lg, min_p=0.9, min_keep=3).tolist() == [True, True, True, False, False]
    assert osamp.sample(lg, 0.0) == 0
    draws = [osamp.sample(lg, 1.0, seed=1, draw=d) for d in range(400)]
    freq = np.bincount(draws, minlength=5) / 400
    assert abs(freq[0] - 0.5) < 0.1 and abs(freq[1] - 0.2) < 0.08
    assert set(osamp.sample(lg, 1.0, seed=1, draw=d, top_k=2) for d in range(100)) == {0, 1}


def test_sesame_checkpoint_key_names_are_accepted():
    """the original sesame/csm-1b (torchtune) names map onto the mlx names one to one."""
    from csm_mlx_b200.models import _expected_shapes, normalize_checkpoint_keys

    W = random_csm_weights(csm_tiny(), seed=2)
    ren = {}
    for k, v in W.items():
        k2 = (k.replace(".self_attn.q_proj.", ".attn.q_proj.").replace(".self_attn.k_proj.", ".attn.k_proj.")
              .replace(".self_attn.v_proj.", ".attn.v_proj.").replace(".self_attn.o_proj.", ".attn.output_proj.")
              .replace(".mlp.gate_proj.", ".mlp.w1.").replace(".mlp.up_proj.", ".mlp.w3.").replace(".mlp.down_proj.", ".mlp.w2.")
              .replace(".input_layernorm.weight", ".sa_norm.scale").replace(".post_attention_layernorm.weight", ".mlp_norm.scale")
              .replace(".norm.weight", ".norm.scale"))
        ren["model." + k2] = v
    assert set(ren) != set(W)
    back = normalize_checkpoint_keys(ren)
    assert set(back) == set(_expected_shapes(csm_tiny())) == set(W)
    assert all(back[k] is W[k] for k in W)
    assert normalize_checkpoint_keys(W) is W


def test_merge_lora_math_and_errors():
    """finetune/utils.py:84-108 folded: W' = W + scale * (lora_a @ lora_b)^T, fp32."""
    from csm_mlx_b200.adapters import merge_lora

    g = torch.Generator().manual_seed(0)
    W = {"backbone.layers.0.self_attn.q_proj.weight": torch.randn(12, 8, generator=g), "other.weight": torch.randn(3, 3, generator=g)}
    a, b = torch.randn(8, 2, generator=g), torch.randn(2, 12, generator=g)
    out = merge_lora(W, {"backbone.layers.0.self_attn.q_proj.lora_a": a, "backbone.layers.0.self_attn.q_proj.lora_b": b}, 0.5)
    x = torch.randn(5, 8, generator=g)
    ref = x @ W["backbone.layers.0.self_attn.q_proj.weight"].t() + 0.5 * ((x @ a) @ b)      # LoRALinear forward
    assert torch.allclose(x @ out["backbone.layers.0.self_attn.q_proj.weight"].t(), ref, atol=1e-5)
    assert out["other.weight"] is W["other.weight"]
    with pytest.raises(ValueError, match="not a Linear"):
        merge_lora(W, {"nope.lora_a": a, "nope.lora_b": b}, 1.0)
    with pytest.raises(ValueError, match="shapes"):
        merge_lora(W, {"backbone.layers.0.self_attn.q_proj.lora_a": a.t().contiguous(), "backbone.layers.0.self_attn.q_proj.lora_b": b}, 1.0)


def test_cli_parser_matches_reference_defaults():
    """cli/generate.py:72-160: option names and defaults of `csm-mlx generate`."""
    from csm_mlx_b200.cli.generate import build_parser, main

    a = build_parser().parse_args(["Hello from Sesame.", "-o", "out.wav"])
    assert (a.model, a.weight, a.speaker, a.max_audio_length, a.temperature, a.top_k, a.top_p, a.min_p,
            a.min_tokens_to_keep, a.adapter) == ("1b", "senstella/csm-1b-mlx", 0, 10000, 0.8, 50, None, None, 1, None)
    b = build_parser().parse_args(["hi", "--output", "o.wav", "-s", "2", "-l", "2000", "--temp", "0", "-k", "0", "-is", "0", "1",
                                   "-ia", "a.wav", "b.wav", "-it", "x", "y"])
    assert b.speaker == 2 and b.input_speakers == [0, 1] and b.input_audios == ["a.wav", "b.wav"] and b.input_texts == ["x", "y"]
    assert main(["hi", "-o", "o.wav", "-ia", "a.wav"]) == 1      # mismatched context lists: cli/generate.py:157-165

"""The oracle reproduces the committed golden vectors (tests/golden, made by scripts/make_golden.py)."""
import os

import numpy as np
import torch

from oracle import lm as olm, mimi as omimi
from tests.conftest import GOLDEN
from tests.workloads import cfg1_prompt_ids, synthetic_audio, tiny_prompt


def test_tiny_lm_golden():
    from csm_mlx_b200.models import csm_tiny
    from csm_mlx_b200.random_init import random_csm_weights

    g = np.load(os.path.join(GOLDEN, "tiny_lm.npz"))
    orc = olm.OracleCSM(olm.TINY, random_csm_weights(csm_tiny(), seed=99, std=0.08))
    tok, mask = tiny_prompt()
    traces = []
    toks = olm.generate_tokens(orc, tok, mask, 8, traces=traces)
    assert np.array_equal(toks.numpy(), g["tokens"])
    np.testing.assert_allclose(traces[0]["h"][0].numpy(), g["h_last_f0"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(torch.stack([l[0] for l in traces[0]["logits"]]).numpy(), g["logits_f0"], rtol=0, atol=1e-5)


def test_cfg1_lm_golden(oracle_1b):
    """BASELINE.json configs[0]: csm_1b random-init greedy, 10 prompt rows, 25 frames (2 s)."""
    g = np.load(os.path.join(GOLDEN, "cfg1_lm.npz"))
    assert list(g["prompt_ids"]) == cfg1_prompt_ids()
    tok, mask = olm.text_rows(cfg1_prompt_ids())
    traces = []
    toks = olm.generate_tokens(oracle_1b, tok, mask, 25, traces=traces)
    assert toks.shape == (25, 32)
    assert np.array_equal(toks.numpy(), g["tokens"])
    np.testing.assert_allclose(traces[0]["h"][0].numpy(), g["h_last_f0"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(torch.stack([l[0] for l in traces[0]["logits"]]).numpy(), g["logits_f0"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(traces[24]["logits"][31][0].numpy(), g["logits_f24_c31"], rtol=0, atol=2e-5)


def test_mimi_golden(mimi_weights):
    g = np.load(os.path.join(GOLDEN, "mimi.npz"))
    audio = omimi.decode(torch.from_numpy(g["decode_codes"]).long(), mimi_weights)[0, 0]
    np.testing.assert_allclose(audio[:9600].numpy(), g["decode_head"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(audio[-1920:].numpy(), g["decode_tail"], rtol=0, atol=1e-4)
    assert abs(float(audio.double().pow(2).sum()) - float(g["decode_sumsq"])) / float(g["decode_sumsq"]) < 1e-5
    clip = synthetic_audio(11, 5.0)
    enc = omimi.encode(clip[None, None], mimi_weights)[0]
    assert enc.shape == (32, 63)
    assert (enc.numpy() == g["encode_codes"]).mean() > 0.999
    enc_r = omimi.encode(clip[None, None, : 120000 - 700], mimi_weights)[0]
    assert enc_r.shape == (32, 63)
    assert (enc_r.numpy() == g["encode_codes_ragged"]).mean() > 0.999


def test_mimi_streaming_equals_offline(mimi_weights):
    g = torch.Generator().manual_seed(5)
    codes = torch.randint(0, 2048, (2, 32, 6), generator=g)
    off = omimi.decode(codes, mimi_weights)
    sd = omimi.StreamingDecoder(mimi_weights)
    st = torch.cat([sd.decode_step(codes[:, :, i:i + 1]) for i in range(6)], -1)
    assert st.shape == off.shape == (2, 1, 6 * 1920)
    assert (st - off).abs().max() < 1e-4


def test_mimi_edge_shapes(mimi_weights):
    assert omimi.encode(torch.zeros(1, 1, 1921), mimi_weights).shape == (1, 32, 2)
    assert omimi.encode(torch.zeros(1, 1, 1920), mimi_weights).shape == (1, 32, 1)
    assert omimi.decode(torch.zeros(1, 32, 1, dtype=torch.long), mimi_weights).shape == (1, 1, 1920)

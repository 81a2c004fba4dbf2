"""Golden vectors produced by the REFERENCE'S OWN MODULES (csm_mlx/{attention,models,generation,tokenizers,segment}.py
from /root/reference, unmodified) running over the mlx stand-in of oracle/mlx_shim.py — see
scripts/make_reference_golden.py, which wrote tests/golden/reference_cfg1.npz in the build container.

CPU part: the oracle restatement and the product's host logic (RoPE tables, frame assembly, error text) against those
vectors.  GPU part: the product's kernels against them.  Tolerances: tokens and index work bit-exact; fp32 hidden
states / logits |Δ| ≤ 2e-5 oracle-vs-reference (both torch fp32, different op order; measured 5e-6) and ≤ 1e-4
product-vs-reference (the LM tolerance stated in tests/test_gpu_lm.py)."""
import os

import numpy as np
import pytest
import torch

from csm_mlx_b200 import generation, tokenizers
from csm_mlx_b200.attention import llama3_rope_table
from oracle import lm as olm
from tests.conftest import GOLDEN
from tests.workloads import cfg1_prompt_ids


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(GOLDEN, "reference_cfg1.npz"))


def test_oracle_golden_tokens_equal_the_reference_run(ref):
    """The oracle's committed 25-frame greedy golden (cfg1_lm.npz, what every GPU test is checked against) is token for
    token what the reference's generate() produced; hidden state and all 32 logit rows of frame 0 agree to 2e-5."""
    g = np.load(os.path.join(GOLDEN, "cfg1_lm.npz"))
    assert np.array_equal(ref["tokens"], g["tokens"])
    assert np.abs(ref["h_last_f0"] - g["h_last_f0"]).max() < 2e-5
    assert np.abs(ref["logits_f0"] - g["logits_f0"]).max() < 2e-5
    assert np.abs(ref["logits_f24"][31] - g["logits_f24_c31"]).max() < 2e-5


def test_oracle_frames_vs_reference(ref, oracle_1b):
    """Live oracle, 2 frames of BASELINE configs[0]: tokens identical, frame-0 logits within 2e-5."""
    tok = torch.from_numpy(ref["prompt_tokens"]).long()
    mask = torch.from_numpy(ref["prompt_mask"]).bool()
    traces = []
    toks = olm.generate_tokens(oracle_1b, tok, mask, 2, traces=traces)
    assert np.array_equal(toks.numpy(), ref["tokens"][:2])
    lg = torch.stack([l[0] for l in traces[0]["logits"]]).numpy()
    assert np.abs(lg - ref["logits_f0"]).max() < 2e-5
    assert np.abs(traces[0]["h"][0].numpy() - ref["h_last_f0"]).max() < 2e-5


def test_context_prompt_and_frames_vs_reference(ref, oracle_1b):
    """Segment context: the reference's tokenize_segment rows (text rows, audio rows, zero EOS row) are what the
    product's frame assembly builds from the same ids / codes, and the oracle continues them with the same tokens."""

    class FakeMimi:
        def encode(self, audio):
            return torch.from_numpy(ref["ctx_audio_codes"])[None]

    class FakeText:
        def encode(self, s):
            return {"[1]a context sentence": ref["ctx_ids"].tolist(), "[0]and now the answer": ref["ctx_new_ids"].tolist()}[s]

    from csm_mlx_b200 import Segment

    tokenizers.set_audio_tokenizer(FakeMimi())
    tokenizers.set_text_tokenizer(FakeText())
    try:
        seg = Segment(1, "a context sentence", torch.zeros(1920 * 4 + 7))
        st, sm = tokenizers.tokenize_segment(seg)
        assert np.array_equal(st.numpy(), ref["ctx_segment_tokens"])
        assert np.array_equal(sm.numpy().astype(np.int32), ref["ctx_segment_mask"].astype(np.int32))
        nt, nm = tokenizers.tokenize_text_segment("and now the answer", 0)
    finally:
        tokenizers.set_audio_tokenizer(None)
        tokenizers.set_text_tokenizer(None)
    tok, mask = torch.cat([st, nt]).long(), torch.cat([sm, nm]).bool()
    toks = olm.generate_tokens(oracle_1b, tok, mask, 3)
    assert np.array_equal(toks.numpy(), ref["ctx_tokens"])


def test_prompt_rows_rope_tables_and_error_text_vs_reference(ref):
    tok, mask = tokenizers.tokenize_text_segment(cfg1_prompt_ids(), 0)
    assert np.array_equal(tok.numpy(), ref["prompt_tokens"])
    assert np.array_equal(mask.numpy().astype(np.int32), ref["prompt_mask"])
    for name, hd in (("b", 64), ("d", 128)):
        for table in (llama3_rope_table(hd, 500_000.0, 32.0, 2048), olm.rope_table(hd, 500_000.0, 32.0, 2048)):
            t = torch.as_tensor(table).reshape(2048, hd // 2, 2)
            assert np.array_equal(t[:64].numpy(), ref[f"rope_{name}_head"]), name
            assert np.array_equal(t[-8:].numpy(), ref[f"rope_{name}_tail"]), name
            assert abs(float(t.double().sum()) - float(ref[f"rope_{name}_sum"])) < 1e-6
    with pytest.raises(ValueError) as e:
        generation._check_length(type("M", (), {"backbone": type("B", (), {"args": type("A", (), {"max_position_embeddings": None})()})()})(),
                                 2040, 12)
    assert str(e.value) == str(ref["too_long_message"])


def test_logits_processor_run_vs_reference(ref, oracle_1b):
    """generation.py:44-49, 59-60: a codebook-0 logits processor (+1e4 on token 123) with the c0 history; the oracle
    reproduces the reference's two frames and sees the same history shapes ((0,) then (1, B, 1))."""
    seen = []

    def bias(history, logits):
        seen.append(tuple(history.shape))
        logits = logits.clone()
        logits[:, 123] += 1e4
        return logits

    tok = torch.from_numpy(ref["prompt_tokens"]).long()
    mask = torch.from_numpy(ref["prompt_mask"]).bool()
    toks = olm.generate_tokens(oracle_1b, tok, mask, 2, logits_processors=[bias])
    assert np.array_equal(toks.numpy(), ref["bias_tokens"])
    want = [tuple(int(x) for x in row[: (1 if i == 0 else 3)]) for i, row in enumerate(ref["bias_history_shapes"])]
    assert seen == want
    # stream_generate of the reference: one 1920-sample chunk per frame, the same tokens as generate(), state reset twice
    assert ref["stream_chunks"].tolist() == [3, 1920, 2]
    assert np.array_equal(ref["stream_tokens"], ref["tokens"][:3])


def test_fixture_is_reproducible_from_the_reference_tree():
    """Where the reference tree is mounted (the build container), re-run the reference's own modules over the mlx
    stand-in for 2 frames and compare everything with the committed fixture; skipped on the GPU box, which has no
    /root/reference."""
    import subprocess
    import sys

    if not os.path.isdir("/root/reference/csm_mlx"):
        pytest.skip("no reference tree here")
    root = os.path.dirname(GOLDEN.rstrip("/")).rsplit("/tests", 1)[0]
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "make_reference_golden.py"), "--frames", "2", "--check"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    assert "identical" in r.stdout


def test_live_reference_vs_oracle_on_small_models():
    """Where the reference tree is mounted: six random small checkpoints with ragged text / audio / EOS-row prompts, six
    greedy frames each through the reference's own generate_frame (over the mlx stand-in) and through the oracle —
    identical tokens, frame-0 logits within 2e-5 (scripts/reference_live_check.py)."""
    import subprocess
    import sys

    if not os.path.isdir("/root/reference/csm_mlx"):
        pytest.skip("no reference tree here")
    root = GOLDEN.rsplit("/tests", 1)[0]
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "reference_live_check.py"), "6"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    assert "reference == oracle" in r.stdout


@pytest.mark.gpu
def test_product_vs_reference_run(ref, model_1b, device):
    """The CUDA path against the reference's own run: 25 greedy frames token-exact, frame-0 hidden state and logits
    within the stated 1e-4, and the context-segment continuation token-exact."""
    from csm_mlx_b200.runtime import LMState, SamplerSpec

    tok = torch.from_numpy(ref["prompt_tokens"]).int()
    mask = torch.from_numpy(ref["prompt_mask"]).bool()
    (toks,) = generation.generate_tokens(model_1b, [(tok, mask)], 25, temperature=0.0)
    assert np.array_equal(toks.numpy(), ref["tokens"])
    st = LMState(model_1b, 1, max_len=64)
    st.prefill([tok], [mask])
    assert float((st.h_last.cpu()[0] - torch.from_numpy(ref["h_last_f0"])).abs().max()) < 1e-4
    forced = torch.from_numpy(ref["tokens"][:1]).to(device).contiguous()
    frame = torch.zeros((1, 32), device=device, dtype=torch.int32)
    lg = torch.zeros((1, 32, 2051), device=device)
    st.depth_decode(frame, SamplerSpec(), logits_out=lg, forced=forced)
    rl = torch.from_numpy(ref["logits_f0"])
    assert float((st.c0_logits.cpu()[0] - rl[0]).abs().max()) < 1e-4
    assert float((lg.cpu()[0, 1:] - rl[1:]).abs().max()) < 1e-4
    ctx = torch.cat([torch.from_numpy(ref["ctx_segment_tokens"]),
                     tokenizers.tokenize_text_segment(ref["ctx_new_ids"].tolist(), 0)[0]]).int()
    cmask = torch.cat([torch.from_numpy(ref["ctx_segment_mask"]).bool(),
                       tokenizers.tokenize_text_segment(ref["ctx_new_ids"].tolist(), 0)[1]])
    (ctoks,) = generation.generate_tokens(model_1b, [(ctx, cmask)], 3, temperature=0.0)
    assert np.array_equal(ctoks.numpy(), ref["ctx_tokens"])

    def bias(history, logits):
        logits = logits.clone()
        logits[:, 123] += 1e4
        return logits

    (btoks,) = generation.generate_tokens(model_1b, [(tok, mask)], 2, temperature=0.0, logits_processors=[bias])
    assert np.array_equal(btoks.numpy(), ref["bias_tokens"])

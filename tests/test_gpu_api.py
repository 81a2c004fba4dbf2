"""End-to-end through the public csm_mlx API on the GPU (README.md:29-55 usage of the reference)."""
import os

import numpy as np
import pytest
import torch

import csm_mlx
from csm_mlx import Segment, generate, stream_generate
from csm_mlx_b200 import generation, tokenizers
from oracle import lm as olm
from oracle import mimi as omimi
from tests.conftest import snr_db
from tests.workloads import cfg1_prompt_ids, synthetic_audio

pytestmark = pytest.mark.gpu


def test_generate_cfg1_audio_matches_oracle_pipeline(model_1b, mimi_gpu, oracle_1b, mimi_weights):
    """generate() == oracle LM tokens -> oracle Mimi decode, for 5 frames of configs[0] (ids ≥ 2048 clamped)."""
    audio = generate(model_1b, cfg1_prompt_ids(), 0, [], max_audio_length_ms=400, temperature=0.0)
    assert audio.dtype == torch.float32 and audio.shape == (5 * 1920,) and audio.device.type == "cpu"
    assert np.asarray(audio).shape == (9600,)
    tok, mask = olm.text_rows(cfg1_prompt_ids())
    toks = olm.generate_tokens(oracle_1b, tok, mask, 5)
    ref = omimi.decode(toks.t()[None].clamp(max=2047), mimi_weights)[0, 0]
    assert snr_db(ref, audio) > 80


def test_stream_generate_chunks_concat_to_generate(model_1b, mimi_gpu):
    ids = cfg1_prompt_ids()
    full = generate(model_1b, ids, 0, [], max_audio_length_ms=640, temperature=0.0)
    chunks = list(stream_generate(model_1b, ids, 0, [], max_audio_length_ms=640, temperature=0.0))
    assert len(chunks) == 8 and all(c.shape == (1920,) and c.dtype == torch.float32 for c in chunks)
    assert snr_db(full, torch.cat(chunks)) > 80


def test_generate_with_segment_context(model_1b, mimi_gpu, oracle_1b, mimi_weights):
    """BASELINE.json configs[2] in miniature: one context Segment (1 s of synthetic audio, Mimi-encoded on the GPU)
    then 3 generated frames; tokens equal the oracle fed with the oracle's own encode of the same audio."""
    tokenizers.set_text_tokenizer(tokenizers.SyntheticTextTokenizer())
    try:
        clip = synthetic_audio(12, 1.0)
        seg = Segment(speaker=1, text="context line", audio=clip)
        prompt = generation._build_prompt(model_1b, "hello there", 0, [seg])
        (got,) = generation.generate_tokens(model_1b, [prompt], 3, temperature=0.0)
        tt = olm.text_rows(tokenizers.SyntheticTextTokenizer().encode("[1]context line"))
        aa = olm.audio_rows(omimi.encode(clip[None, None], mimi_weights)[0])
        t2 = olm.text_rows(tokenizers.SyntheticTextTokenizer().encode("[0]hello there"))
        otok = torch.cat([tt[0], aa[0], t2[0]])
        omask = torch.cat([tt[1], aa[1], t2[1]])
        assert prompt[0].shape == otok.shape and torch.equal(prompt[1], omask)
        # codec parity is asserted on its own (identical codes, or refereed float-level near-ties of the RVQ search) ...
        from tests.codec_referee import assert_codes_match

        n_text = tt[0].shape[0]
        n_audio = aa[0].shape[0] - 1                      # the last audio row is the all-zero EOS frame
        got_codes = prompt[0][n_text:n_text + n_audio, :32].t()[None]
        assert_codes_match(got_codes, otok[n_text:n_text + n_audio, :32].t()[None], clip[None, None], mimi_weights)
        # ... and the LM is compared on the prompt rows the product actually built
        exp = olm.generate_tokens(oracle_1b, prompt[0].long(), prompt[1], 3)
        assert torch.equal(got.long(), exp)
    finally:
        tokenizers.set_text_tokenizer(None)


def test_conversation_caches_do_not_change_what_generate_returns(model_1b, mimi_gpu, monkeypatch):
    """Two turns of a conversation through generate(): the second turn takes the context segments' codes from the model's
    ContextCache and the context rows' backbone KV from its KVPrefixCache (generation.py:108-121 recomputes both per call).
    Audio identical, sample for sample, to the same two calls with CSMB_DISABLE_CONV_CACHE=1."""
    tokenizers.set_text_tokenizer(tokenizers.SyntheticTextTokenizer())
    try:
        ctx = [Segment(0, "so how was the trip", synthetic_audio(21, 1.1)), Segment(1, "long, but fine", synthetic_audio(22, 0.9))]
        turns = ["tell me more about it", "and what happens next"]
        generation.set_conversation_cache(model_1b)                      # fresh caches
        cached = [generate(model_1b, t, 1, ctx, max_audio_length_ms=400, temperature=0.0) for t in turns]
        segs, kv = model_1b.__dict__["_conv_cache"]
        assert (segs.hits, segs.misses) == (2, 2) and (kv.hits, kv.misses) == (1, 1)
        chunks = list(stream_generate(model_1b, turns[1], 1, ctx, max_audio_length_ms=400, temperature=0.0))
        assert kv.hits == 2 and snr_db(cached[1], torch.cat(chunks)) > 80
        monkeypatch.setenv("CSMB_DISABLE_CONV_CACHE", "1")
        plain = [generate(model_1b, t, 1, ctx, max_audio_length_ms=400, temperature=0.0) for t in turns]
        assert (kv.hits, kv.misses) == (2, 1)
        for a, b in zip(cached, plain):
            assert a.shape == (5 * 1920,) and torch.equal(a, b)
        assert not torch.equal(cached[0], cached[1])
    finally:
        generation.set_conversation_cache(model_1b)
        tokenizers.set_text_tokenizer(None)


def test_sampler_argument_forms(model_1b, mimi_gpu):
    ids = cfg1_prompt_ids()
    a = generate(model_1b, ids, 0, [], max_audio_length_ms=160, sampler=csm_mlx.make_sampler(temp=0.8, top_k=50, seed=3))
    b = generate(model_1b, ids, 0, [], max_audio_length_ms=160, sampler=csm_mlx.make_sampler(temp=0.8, top_k=50, seed=3))
    c = generate(model_1b, ids, 0, [], max_audio_length_ms=160, temperature=0.8, seed=4)
    assert torch.equal(a, b) and a.shape == c.shape == (3840,) and not torch.equal(a, c)
    s = csm_mlx.make_sampler(temp=0.0)
    assert int(s(torch.tensor([0.1, 0.9, 0.3], device=model_1b.device))) == 1


def test_generate_frame_signature(model_1b):
    """generate_frame(model, tokens, *, temperature, token_mask, cache, …) -> (B,32) int32 (generation.py:21-31,92)."""
    tok, mask = tokenizers.tokenize_text_segment(cfg1_prompt_ids(), 0)
    cache = generation.make_cache(model_1b, 1, 64)
    f0 = generation.generate_frame(model_1b, tok[None], temperature=0.0, token_mask=mask[None], cache=cache)
    assert f0.shape == (1, 32) and f0.dtype == torch.int32
    nxt = torch.cat([f0.cpu(), torch.zeros(1, 1, dtype=torch.int32)], 1)[:, None]
    nmask = torch.cat([torch.ones(1, 32), torch.zeros(1, 1)], 1)[:, None].bool()
    f1 = generation.generate_frame(model_1b, nxt, temperature=0.0, token_mask=nmask, cache=cache)
    (ref,) = generation.generate_tokens(model_1b, [(tok, mask)], 2, temperature=0.0)
    assert torch.equal(torch.cat([f0, f1]).cpu(), ref)


def test_generate_batch(model_1b, mimi_gpu):
    from tests.workloads import prompt_ids

    texts = [prompt_ids(21 + i, 8 + i) for i in range(3)]
    audios, frames = csm_mlx.generate_batch(model_1b, texts, [0, 1, 2], max_audio_length_ms=240, temperature=0.0,
                                            return_tokens=True)
    assert len(audios) == 3 and all(a.shape == (3 * 1920,) for a in audios)
    single = generate(model_1b, texts[1], 1, [], max_audio_length_ms=240, temperature=0.0)
    assert snr_db(single, audios[1]) > 80


def test_pooled_state_and_codec_reuse_is_deterministic(model_1b, mimi_gpu, monkeypatch):
    """Back-to-back utterances reuse the pooled LM state / codec stream (KV pages, workspaces, captured graphs):
    a different utterance in between must not leak into the next one, with and without the codec on its own stream."""
    ids_a, ids_b = cfg1_prompt_ids(), [128000, 11, 22, 33, 44, 128001]
    first = torch.cat(list(stream_generate(model_1b, ids_a, 0, [], max_audio_length_ms=800, temperature=0.0)))
    other = torch.cat(list(stream_generate(model_1b, ids_b, 3, [], max_audio_length_ms=400, temperature=0.0)))
    again = torch.cat(list(stream_generate(model_1b, ids_a, 0, [], max_audio_length_ms=800, temperature=0.0)))
    assert first.shape == (10 * 1920,) and other.shape == (5 * 1920,)
    assert torch.equal(first, again)
    assert len(model_1b.__dict__.get("_lm_pool", [])) >= 1 and len(mimi_gpu.__dict__.get("_stream_pool", [])) >= 1
    monkeypatch.setenv("CSMB_DISABLE_OVERLAP", "1")
    serial = torch.cat(list(stream_generate(model_1b, ids_a, 0, [], max_audio_length_ms=800, temperature=0.0)))
    assert torch.equal(first, serial)


def test_abandoned_stream_releases_resources(model_1b, mimi_gpu):
    """A consumer that stops iterating early (generator closed) still orders the codec stream and returns the state."""
    gen = stream_generate(model_1b, cfg1_prompt_ids(), 0, [], max_audio_length_ms=2000, temperature=0.0)
    head = [next(gen) for _ in range(3)]
    gen.close()
    full = list(stream_generate(model_1b, cfg1_prompt_ids(), 0, [], max_audio_length_ms=400, temperature=0.0))
    assert all(torch.equal(a, b) for a, b in zip(head, full))


def test_load_adapters_folds_lora_into_dense_weights(device, tmp_path):
    """load_adapters (finetune/utils.py:84-108) on the generation path: the LoRA pair is merged, W' = W + s (A B)^T,
    and generation runs on the merged weights; a `full` adapter is a non-strict weight load."""
    import json

    from safetensors.torch import save_file

    from csm_mlx_b200 import CSM, csm_tiny, load_adapters
    from csm_mlx_b200.random_init import random_csm_weights

    args = csm_tiny()
    W = random_csm_weights(args, seed=5)
    m = CSM(args, device=device).load_weights(W)
    name = "decoder.layers.0.mlp.down_proj"
    w0 = m.parameters()[name + ".weight"].float().cpu()
    g = torch.Generator().manual_seed(1)
    a, b = 0.05 * torch.randn(w0.shape[1], 4, generator=g), 0.05 * torch.randn(4, w0.shape[0], generator=g)
    d = tmp_path / "adapter"
    d.mkdir()
    save_file({name + ".lora_a": a, name + ".lora_b": b}, str(d / "adapters.safetensors"))
    (d / "adapter_config.json").write_text(json.dumps({"fine_tune_type": "lora", "lora_parameters": {"rank": 4, "scale": 2.0}}))
    load_adapters(m, str(d))
    w1 = m.parameters()[name + ".weight"].float().cpu()
    want = (w0 + 2.0 * (a @ b).t()).to(torch.bfloat16).float()
    assert torch.equal(w1, want)
    other = m.parameters()["decoder.layers.0.mlp.up_proj.weight"].float().cpu()
    assert torch.equal(other, W["decoder.layers.0.mlp.up_proj.weight"].to(torch.bfloat16).float())
    d2 = tmp_path / "full"
    d2.mkdir()
    save_file({"projection.weight": torch.zeros_like(W["projection.weight"])}, str(d2 / "adapters.safetensors"))
    (d2 / "adapter_config.json").write_text(json.dumps({"fine_tune_type": "full"}))
    load_adapters(m, str(d2))
    assert float(m.parameters()["projection.weight"].float().abs().max()) == 0.0


def test_cli_generate_end_to_end(tmp_path):
    """`csm-mlx generate TEXT -o out.wav` (cli/generate.py:72-202) on synthetic assets, with one context segment read
    back from a WAV file: a 24 kHz mono file of max_audio_length comes out."""
    import subprocess
    import sys

    from csm_mlx_b200.utils import read_audio, write_audio
    from tests.workloads import synthetic_audio

    ctx = tmp_path / "ctx.wav"
    write_audio(synthetic_audio(5, 1.0), ctx, 24000)
    out = tmp_path / "out.wav"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "csm_mlx_b200.cli.generate", "Hello from Sesame.", "-o", str(out), "-l", "400",
                        "--temp", "0.8", "-k", "50", "--seed", "3", "--synthetic-assets", "-is", "1", "-ia", str(ctx), "-it",
                        "a context sentence"], cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Success! Audio saved to" in r.stdout
    audio = read_audio(out, 24000)
    assert audio.shape == (5 * 1920,) and bool(torch.isfinite(audio).all())


def test_zero_length_request_returns_empty_audio(model_1b, mimi_gpu, capsys):
    """max_audio_length_ms below one frame: no frame is generated, the warning of generation.py:163-165 is printed and
    a (0,) float32 array comes back; stream_generate yields nothing."""
    out = generate(model_1b, cfg1_prompt_ids(), 0, [], max_audio_length_ms=79, temperature=0.0)
    assert out.shape == (0,) and out.dtype == torch.float32
    assert "No samples generated" in capsys.readouterr().out
    assert list(stream_generate(model_1b, cfg1_prompt_ids(), 0, [], max_audio_length_ms=79, temperature=0.0)) == []


def test_engine_tokens_invariant_to_batch_size(model_1b, mimi_gpu):
    """The reference generates one utterance at a time (generation.py:139-161), so an utterance's tokens cannot depend on
    its neighbours.  Same property here: the same 16 utterances x 125 frames (BASELINE.json configs[3] prompts) through
    engines of 4, 8, 16 and 64 slots — different waves, admissions and row counts in every Linear — and through the
    request sharding of a 2-GPU job (requests i mod 2 on engines of 8 slots) give identical tokens, bit for bit; the first
    three also equal the utterance served alone by a 1-slot engine."""
    from csm_mlx_b200 import serving
    from csm_mlx_b200.sharding import shard_indices
    from tests.workloads import prompt_ids

    frames = 125
    prompts = [tokenizers.tokenize_text_segment(prompt_ids(21 + i, 8 + i % 9), 0) for i in range(16)]

    def through(max_batch, idx):
        eng = serving.Engine(model_1b, max_batch=max_batch, max_len=32 + frames + 2)
        rids = [eng.submit_prompt(prompts[i][0], prompts[i][1], frames) for i in idx]
        eng.run()
        return {i: eng.tokens(r) for i, r in zip(idx, rids)}

    ref = through(16, list(range(16)))
    assert all(t.shape == (frames, 32) for t in ref.values())
    for B in (4, 8, 64):
        got = through(B, list(range(16)))
        for i in range(16):
            assert torch.equal(ref[i], got[i]), (B, i, int((ref[i] != got[i]).nonzero()[0][0]))
    for rank in range(2):
        mine = shard_indices(16, rank, 2)
        got = through(len(mine), mine)
        for i in mine:
            assert torch.equal(ref[i], got[i]), ("shard", rank, i)
    for i in range(3):
        alone = through(1, [i])
        assert torch.equal(ref[i], alone[i]), ("alone", i)

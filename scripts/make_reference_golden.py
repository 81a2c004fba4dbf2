"""Generates tests/golden/reference_cfg1.npz by RUNNING THE REFERENCE'S OWN MODULES from /root/reference
(attention.py, models.py, generation.py, tokenizers.py, segment.py, config.py — unmodified) over the mlx stand-in of
oracle/mlx_shim.py.  Run once in the build container (the reference tree does not exist on the GPU box); the fixture
is committed and checked by tests/test_reference_golden.py against the oracle and, on the GPU, against the product.

    python scripts/make_reference_golden.py

What is real reference code here: Llama3ScaledRoPE + Attention, the CSM parameter tree, embed_tokens / embed_audio,
generate_frame (mask-multiply + sum, last-position slice, codebook-0 head, the 31-step depth loop with a fresh decoder
cache, audio_head[i-1], embedding offsets), the generate() driver (prompt assembly, "inputs too long" guard, EOS test,
next-input construction) and the frame assembly of tokenizers.py.  Supplied by the stand-in: the mlx array API, the
mlx_lm Llama block / KV cache / attention primitive.  Patched out: the text tokenizer (not downloadable: a fake that
returns BASELINE configs[0]'s stand-in ids) and the Mimi codec (a fake that records the codes it is asked to decode
and returns seeded codes from encode).
"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import transformers  # noqa: E402,F401  (its is_mlx_available() probe must run before the stand-in is registered)
from transformers import AutoTokenizer, LlamaTokenizer  # noqa: E402,F401

from oracle import mlx_shim  # noqa: E402

mx = mlx_shim.install("/root/reference")
ref_attention = importlib.import_module("csm_mlx.attention")
ref_models = importlib.import_module("csm_mlx.models")
ref_tokenizers = importlib.import_module("csm_mlx.tokenizers")
ref_segment = importlib.import_module("csm_mlx.segment")
ref_generation = importlib.import_module("csm_mlx.generation")

from csm_mlx_b200.random_init import random_csm_weights  # noqa: E402
from tests.workloads import cfg1_prompt_ids, prompt_ids  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
torch.set_num_threads(os.cpu_count() or 8)


class FakeTextTokenizer:
    """encode("[speaker]text") -> ids.  Texts are registered up front; anything else is an error."""

    def __init__(self):
        self.table = {}

    def encode(self, s):
        return list(self.table[s])


class FakeMimi:
    """decode records the (1, K, F) codes; encode returns seeded (1, K, T) codes, T = ceil(samples / 1920)."""

    def __init__(self):
        self.decoded = []

    def decode(self, codes):
        self.decoded.append(torch.as_tensor(codes).clone())
        return mx.zeros((1, 1, 1920 * codes.shape[-1]))

    def reset_state(self):
        self.resets = getattr(self, "resets", 0) + 1

    def decode_step(self, codes):
        self.steps = getattr(self, "steps", [])
        self.steps.append(torch.as_tensor(codes).clone())      # (1, K, 1) per frame
        return mx.zeros((1, 1, 1920))

    def encode(self, audio):
        n = int(audio.shape[-1])
        t = -(-n // 1920)
        g = torch.Generator().manual_seed(n)
        return mx.array(torch.randint(0, 2048, (1, 32, t), generator=g).to(torch.int32))


text_tok, mimi = FakeTextTokenizer(), FakeMimi()
ref_tokenizers.get_text_tokenizer = lambda: text_tok
ref_tokenizers.get_audio_tokenizer = lambda n_audio_codebooks=32: mimi
ref_generation.get_audio_tokenizer = ref_tokenizers.get_audio_tokenizer


def build_model():
    W = random_csm_weights()                       # bf16-rounded values (the product's HBM copy), used in fp32 here
    model = ref_models.CSM(ref_models.csm_1b())
    model.load_weights([(k, v.to(torch.float32)) for k, v in W.items()])
    return model


def record(model):
    """Wrap the two heads so that hidden states / logits of every frame are captured without touching the loop."""
    rec = {"h": [], "c0": [], "ci": []}
    head = model.codebook0_head

    class Rec(mlx_shim.Module):
        def __call__(self, x):
            y = head(x)
            rec["h"].append(torch.as_tensor(x).clone())
            rec["c0"].append(torch.as_tensor(y).clone())
            return y

    model.codebook0_head = Rec()
    mm = mx.matmul

    def matmul(a, b):
        y = mm(a, b)
        rec["ci"].append(torch.as_tensor(y).clone())
        return y

    ref_generation.mx.matmul = matmul
    return rec


def main():
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=25, help="greedy frames of BASELINE configs[0] to generate")
    ap.add_argument("--check", action="store_true",
                    help="do not write: compare with the committed fixture (tokens of the generated frames, RoPE tables, "
                         "prompt / context rows, frame-0 logits, guard text) and exit non-zero on any difference")
    args = ap.parse_args()
    model = build_model()
    out = {}
    # ---- RoPE tables as the reference builds them (attention.py:57-117), both head sizes of csm_1b
    for name, stack in (("b", model.backbone), ("d", model.decoder)):
        cache = torch.as_tensor(stack.layers[0].self_attn.rope._cache)
        out[f"rope_{name}_head"] = cache[:64].numpy().astype(np.float32)          # first 64 positions
        out[f"rope_{name}_tail"] = cache[-8:].numpy().astype(np.float32)          # positions 2040..2047
        out[f"rope_{name}_sum"] = np.float64(cache.double().sum().item())
        out[f"rope_{name}_theta"] = torch.as_tensor(stack.layers[0].self_attn.rope._theta).numpy().astype(np.float32)

    # ---- BASELINE.json configs[0]: "[0]Hello from Sesame.", speaker 0, no context, greedy, 25 frames
    text_tok.table["[0]Hello from Sesame."] = cfg1_prompt_ids()
    rec = record(model)
    ref_generation.generate(model, "Hello from Sesame.", 0, [], max_audio_length_ms=80 * args.frames, temperature=0)
    codes = mimi.decoded[-1]                                  # (1, 32, F): mx.stack(samples).transpose(1, 2, 0)
    tokens = codes[0].t().contiguous()
    assert tokens.shape == (args.frames, 32), tokens.shape
    out["tokens"] = tokens.numpy().astype(np.int32)
    out["h_last_f0"] = rec["h"][0][0].numpy().astype(np.float32)
    out["logits_f0"] = torch.cat([rec["c0"][0], *rec["ci"][:31]]).numpy().astype(np.float32)         # (32, 2051)
    if args.frames >= 25:
        out["logits_f24"] = torch.cat([rec["c0"][24], *rec["ci"][24 * 31:25 * 31]]).numpy().astype(np.float32)
    tp, mp = ref_tokenizers.tokenize_text_segment("Hello from Sesame.", 0)
    out["prompt_tokens"] = torch.as_tensor(tp).numpy().astype(np.int32)
    out["prompt_mask"] = torch.as_tensor(mp).numpy().astype(np.int32)

    # ---- a context segment (text + audio rows + EOS row, tokenizers.py:61-102) followed by new text, 3 frames
    ctx_ids, new_ids = prompt_ids(1, 6), prompt_ids(2, 5)
    text_tok.table["[1]a context sentence"] = ctx_ids
    text_tok.table["[0]and now the answer"] = new_ids
    audio = mx.zeros((1920 * 4 + 7,))                         # 5 codec frames (the last one partial)
    seg = ref_segment.Segment(1, "a context sentence", audio)
    st, sm = ref_tokenizers.tokenize_segment(seg, n_audio_codebooks=32)
    out["ctx_segment_tokens"] = torch.as_tensor(st).numpy().astype(np.int32)
    out["ctx_segment_mask"] = torch.as_tensor(sm).numpy().astype(np.int32)
    out["ctx_audio_codes"] = torch.as_tensor(mimi.encode(mx.expand_dims(mx.expand_dims(audio, 0), 0))[0]).numpy().astype(np.int32)
    rec2 = {"h": [], "c0": [], "ci": []}
    rec.update(rec2)
    ref_generation.generate(model, "and now the answer", 0, [seg], max_audio_length_ms=240, temperature=0)
    out["ctx_tokens"] = mimi.decoded[-1][0].t().contiguous().numpy().astype(np.int32)                 # (3, 32)
    out["ctx_new_ids"] = np.array(new_ids, dtype=np.int64)
    out["ctx_ids"] = np.array(ctx_ids, dtype=np.int64)

    # ---- stream_generate (generation.py:181-258): one chunk per frame, codes handed to decode_step as (1, K, 1)
    mimi.steps, mimi.resets = [], 0
    chunks = list(ref_generation.stream_generate(model, "Hello from Sesame.", 0, [], max_audio_length_ms=240, temperature=0))
    out["stream_chunks"] = np.array([len(chunks), int(chunks[0].shape[0]), mimi.resets], dtype=np.int64)
    out["stream_tokens"] = torch.cat([c[0, :, 0][None] for c in mimi.steps]).numpy().astype(np.int32)      # (3, 32)

    # ---- logits processors on codebook 0 with the c0 history (generation.py:44-49, 59-60): a +1e4 bias on token 123
    seen = []

    def bias(history, logits):
        seen.append(tuple(torch.as_tensor(history).shape))
        logits = torch.as_tensor(logits).clone()
        logits[:, 123] += 1e4
        return mx.array(logits)

    ref_generation.generate(model, "Hello from Sesame.", 0, [], max_audio_length_ms=160, temperature=0, logits_processors=[bias])
    out["bias_tokens"] = mimi.decoded[-1][0].t().contiguous().numpy().astype(np.int32)                 # (2, 32)
    out["bias_history_shapes"] = np.array([list(s) + [0] * (3 - len(s)) for s in seen], dtype=np.int64)

    # ---- the guard of generation.py:131-137
    try:
        text_tok.table["[0]long"] = [1] * 2040
        ref_generation.generate(model, "long", 0, [], max_audio_length_ms=1000, temperature=0)
        out["too_long_message"] = np.array("")
    except ValueError as e:
        out["too_long_message"] = np.array(str(e))

    if args.check:
        ref = np.load(os.path.join(OUT, "reference_cfg1.npz"))
        bad = []
        for k, v in out.items():
            want = ref[k][: args.frames] if k == "tokens" else ref[k]
            same = (str(v) == str(want)) if v.dtype.kind in "US" else (
                np.array_equal(v, want) if v.dtype.kind in "iu" else np.allclose(v, want, rtol=0, atol=1e-6))
            if not same:
                bad.append(k)
        print("check against the committed fixture:", "identical" if not bad else f"DIFFERENT: {bad}")
        raise SystemExit(1 if bad else 0)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "reference_cfg1.npz"), **out)
    print("reference_cfg1.npz:", {k: getattr(v, "shape", None) for k, v in out.items()})
    print("first frame:", out["tokens"][0, :8], "guard:", str(out["too_long_message"]))


if __name__ == "__main__":
    main()

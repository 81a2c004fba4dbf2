"""Generates tests/golden/*.npz with the CPU oracle (run once, here; the files are committed).

The reference ships no golden vectors and cannot be run as shipped in this image (mlx / mlx_lm / moshi_mlx are not
installable), so these are ORACLE outputs; the oracle itself is pinned against the reference's own modules run over
an mlx stand-in (scripts/make_reference_golden.py -> tests/golden/reference_cfg1.npz) and, for the third-party
pieces, against the independent HF implementations in tests/test_oracle_vs_hf.py.  Inputs follow SURVEY.md §8d (cfg 1 prompt, seeds).

    python scripts/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from csm_mlx_b200.random_init import random_csm_weights, random_mimi_weights  # noqa: E402
from csm_mlx_b200.models import csm_tiny  # noqa: E402
from oracle import lm as olm, mimi as omimi  # noqa: E402
from tests.workloads import cfg1_prompt_ids, synthetic_audio, tiny_prompt  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def lm_1b():
    W = random_csm_weights()
    orc = olm.OracleCSM(olm.CSM_1B, W)
    tok, mask = olm.text_rows(cfg1_prompt_ids())
    traces = []
    toks = olm.generate_tokens(orc, tok, mask, 25, traces=traces)
    tr0 = traces[0]
    margins = []
    for tr in traces:
        for lg in tr["logits"]:
            t2 = torch.topk(lg[0], 2).values
            margins.append(float(t2[0] - t2[1]))
    np.savez_compressed(
        os.path.join(OUT, "cfg1_lm.npz"),
        prompt_ids=np.array(cfg1_prompt_ids(), dtype=np.int64),
        tokens=toks.numpy().astype(np.int32),                                # (25, 32) greedy frames
        h_last_f0=tr0["h"][0].numpy().astype(np.float32),                   # (2048,) backbone output, frame 0
        layer_rms_f0=np.array([float(h.pow(2).mean().sqrt()) for h in tr0["backbone_layers"]], dtype=np.float32),
        logits_f0=torch.stack([l[0] for l in tr0["logits"]]).numpy().astype(np.float32),  # (32, 2051)
        logits_f1_c0=traces[1]["logits"][0][0].numpy().astype(np.float32),
        logits_f24_c31=traces[24]["logits"][31][0].numpy().astype(np.float32),
        min_margin=np.float32(min(margins)),
    )
    print("cfg1_lm: tokens", toks.shape, "min top-2 margin", min(margins))


def lm_tiny():
    args = csm_tiny()
    W = random_csm_weights(args, seed=99, std=0.08)
    orc = olm.OracleCSM(olm.TINY, W)
    tok, mask = tiny_prompt()
    traces = []
    toks = olm.generate_tokens(orc, tok, mask, 8, traces=traces)
    np.savez_compressed(os.path.join(OUT, "tiny_lm.npz"), tokens=toks.numpy().astype(np.int32),
                        h_last_f0=traces[0]["h"][0].numpy(), logits_f0=torch.stack([l[0] for l in traces[0]["logits"]]).numpy())
    print("tiny_lm:", toks.shape)


def mimi():
    MW = random_mimi_weights()
    g = torch.Generator().manual_seed(77)
    codes = torch.randint(0, 2048, (1, 32, 25), generator=g)
    audio = omimi.decode(codes, MW)[0, 0]
    clip = synthetic_audio(11, 5.0)
    enc = omimi.encode(clip[None, None], MW)[0]
    enc_ragged = omimi.encode(clip[None, None, : 120000 - 700], MW)[0]
    np.savez_compressed(
        os.path.join(OUT, "mimi.npz"),
        decode_codes=codes.numpy().astype(np.int32),
        decode_head=audio[:9600].numpy().astype(np.float32),
        decode_tail=audio[-1920:].numpy().astype(np.float32),
        decode_sum=np.float64(audio.double().sum()), decode_sumsq=np.float64(audio.double().pow(2).sum()),
        encode_codes=enc.numpy().astype(np.int32),                 # (32, 63)
        encode_codes_ragged=enc_ragged.numpy().astype(np.int32),   # (32, 63) from 119 300 samples
    )
    print("mimi: decode", audio.shape, "encode", enc.shape)


if __name__ == "__main__":
    which = sys.argv[1:] or ["lm_1b", "lm_tiny", "mimi"]
    for w in which:
        globals()[w]()

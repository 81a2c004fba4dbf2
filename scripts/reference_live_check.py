"""Live cross-check of the oracle against the REFERENCE'S OWN MODULES (run over oracle/mlx_shim.py) on a small model
configuration and several seeds: for every seed a random checkpoint, a ragged prompt (text rows, audio rows + EOS row,
text rows) and 6 greedy frames through the reference's `generate_frame` loop and through `oracle.lm`, compared token for
token, plus frame-0 logits.  Needs the reference tree (build container only); exits non-zero on any difference.

    python scripts/reference_live_check.py [n_seeds]
"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import transformers  # noqa: E402,F401  (its mlx probe must run before the stand-in is registered)

from oracle import mlx_shim  # noqa: E402

mx = mlx_shim.install("/root/reference")
ref_config = importlib.import_module("csm_mlx.config")
ref_models = importlib.import_module("csm_mlx.models")
ref_generation = importlib.import_module("csm_mlx.generation")

from oracle import lm as olm  # noqa: E402

LlamaArgs = mlx_shim.LlamaArgs
ROPE = {"factor": 32.0, "high_freq_factor": 4.0, "low_freq_factor": 1.0, "original_max_position_embeddings": 8192,
        "rope_type": "llama3"}
T = olm.TINY
# the reference only defines "1b" / "100m": register a small pair with the same structure (config.py:3-45)
ref_config.BACKBONE_CONFIGURATION["tiny"] = LlamaArgs(
    model_type="llama", vocab_size=T.n_text_vocab, num_hidden_layers=T.backbone.n_layers, num_attention_heads=T.backbone.n_heads,
    num_key_value_heads=T.backbone.n_kv_heads, head_dim=T.backbone.head_dim, intermediate_size=T.backbone.d_ff,
    hidden_size=T.backbone.d_model, rms_norm_eps=1e-5, rope_scaling=dict(ROPE), rope_theta=500_000.0)
ref_config.DECODER_CONFIGURATION["tiny"] = LlamaArgs(
    model_type="llama", vocab_size=T.n_text_vocab, num_hidden_layers=T.decoder.n_layers, num_attention_heads=T.decoder.n_heads,
    num_key_value_heads=T.decoder.n_kv_heads, head_dim=T.decoder.head_dim, intermediate_size=T.decoder.d_ff,
    hidden_size=T.decoder.d_model, rms_norm_eps=1e-5, rope_scaling=dict(ROPE), rope_theta=500_000.0)


def random_weights(seed: int):
    g = torch.Generator().manual_seed(seed)
    W = {}

    def lin(name, out, inp):
        W[name] = torch.randn(out, inp, generator=g) * 0.08

    for stack, cfg in (("backbone", T.backbone), ("decoder", T.decoder)):
        for l in range(cfg.n_layers):
            p = f"{stack}.layers.{l}."
            lin(p + "self_attn.q_proj.weight", cfg.n_heads * cfg.head_dim, cfg.d_model)
            lin(p + "self_attn.k_proj.weight", cfg.n_kv_heads * cfg.head_dim, cfg.d_model)
            lin(p + "self_attn.v_proj.weight", cfg.n_kv_heads * cfg.head_dim, cfg.d_model)
            lin(p + "self_attn.o_proj.weight", cfg.d_model, cfg.n_heads * cfg.head_dim)
            lin(p + "mlp.gate_proj.weight", cfg.d_ff, cfg.d_model)
            lin(p + "mlp.up_proj.weight", cfg.d_ff, cfg.d_model)
            lin(p + "mlp.down_proj.weight", cfg.d_model, cfg.d_ff)
            W[p + "input_layernorm.weight"] = 1 + 0.05 * torch.randn(cfg.d_model, generator=g)
            W[p + "post_attention_layernorm.weight"] = 1 + 0.05 * torch.randn(cfg.d_model, generator=g)
        W[f"{stack}.norm.weight"] = 1 + 0.05 * torch.randn(cfg.d_model, generator=g)
    lin("text_embeddings.weight", T.n_text_vocab, T.backbone.d_model)
    lin("audio_embeddings.weight", T.n_audio_vocab * T.n_audio_codebooks, T.backbone.d_model)
    lin("projection.weight", T.decoder.d_model, T.backbone.d_model)
    lin("codebook0_head.weight", T.n_audio_vocab, T.backbone.d_model)
    W["audio_head"] = torch.randn(T.n_audio_codebooks - 1, T.decoder.d_model, T.n_audio_vocab, generator=g) * 0.08
    return W


def ragged_prompt(seed: int):
    g = torch.Generator().manual_seed(1000 + seed)
    ncb = T.n_audio_codebooks
    n_text1, n_audio, n_text2 = 2 + seed % 3, 3 + seed % 4, 1 + seed % 2
    rows = n_text1 + n_audio + 1 + n_text2
    tok = torch.zeros((rows, ncb + 1), dtype=torch.int64)
    mask = torch.zeros((rows, ncb + 1), dtype=torch.bool)
    tok[:n_text1, -1] = torch.randint(0, T.n_text_vocab, (n_text1,), generator=g)
    mask[:n_text1, -1] = True
    a0 = n_text1
    tok[a0:a0 + n_audio, :-1] = torch.randint(0, T.n_audio_vocab - 3, (n_audio, ncb), generator=g)
    mask[a0:a0 + n_audio + 1, :-1] = True                      # audio rows + the all-zero EOS row
    t0 = a0 + n_audio + 1
    tok[t0:, -1] = torch.randint(0, T.n_text_vocab, (n_text2,), generator=g)
    mask[t0:, -1] = True
    return tok, mask


def reference_frames(W, tok, mask, frames):
    """The frame loop of generation.py:139-161 around the reference's own generate_frame, with its own KVCache."""
    args = ref_models.ModelArgs(backbone_name="tiny", decoder_name="tiny", n_text_vocab=T.n_text_vocab,
                                n_audio_vocab=T.n_audio_vocab, n_audio_codebooks=T.n_audio_codebooks)
    model = ref_models.CSM(args)
    model.load_weights(list(W.items()))
    rec = []
    head = model.codebook0_head

    class Rec(mlx_shim.Module):
        def __call__(self, x):
            y = head(x)
            rec.append(torch.as_tensor(y).clone())
            return y

    model.codebook0_head = Rec()
    cache = [ref_generation.KVCache() for _ in model.backbone.layers]
    inp = mx.expand_dims(mx.array(tok.to(torch.int32)), 0)
    msk = mx.expand_dims(mx.array(mask), 0)
    out = []
    for _ in range(frames):
        sample = ref_generation.generate_frame(model, inp, temperature=0, token_mask=msk, cache=cache)
        out.append(torch.as_tensor(sample)[0].clone())
        inp = mx.expand_dims(mx.concat([sample, mx.zeros((1, 1))], axis=1), 1).astype(mx.int32)        # generation.py:156-161
        msk = mx.expand_dims(mx.concat([mx.ones_like(sample), mx.zeros((1, 1))], axis=1), 1).astype(mx.bool_)
    return torch.stack(out), rec[0][0]


def main():
    n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    frames = 6
    worst = 0.0
    for seed in range(n_seeds):
        W = random_weights(seed)
        tok, mask = ragged_prompt(seed)
        ref_toks, ref_c0 = reference_frames(W, tok, mask, frames)
        orc = olm.OracleCSM(T, W)
        traces = []
        toks = olm.generate_tokens(orc, tok, mask, frames, traces=traces)
        if toks.shape[0] != frames or not torch.equal(toks.to(torch.int64), ref_toks.to(torch.int64)):
            print(f"seed {seed}: TOKENS DIFFER\n reference {ref_toks.tolist()}\n oracle    {toks.tolist()}")
            raise SystemExit(1)
        d = float((traces[0]["logits"][0][0] - ref_c0).abs().max())
        worst = max(worst, d)
        if d > 2e-5:
            print(f"seed {seed}: c0 logits differ by {d}")
            raise SystemExit(1)
    print(f"reference == oracle on {n_seeds} seeds x {frames} frames (ragged prompts); max |dlogit| {worst:.2e}")


if __name__ == "__main__":
    main()

"""Pins the oracle against independent implementations importable in this image (HF transformers):
``MimiModel`` for the codec, ``CsmBackboneModel`` / ``CsmDepthDecoderForCausalLM`` for the LM.
The reference itself (mlx / mlx_lm / moshi_mlx) cannot be installed here — see oracle/__init__.py."""
import warnings

import pytest
import torch

from oracle import lm as olm, mimi as omimi
from tests.hf_map import _perm_rows, build_hf_mimi
from tests.workloads import synthetic_audio

warnings.filterwarnings("ignore")


def test_mimi_decode_vs_hf(mimi_weights):
    hf = build_hf_mimi(mimi_weights)
    g = torch.Generator().manual_seed(0)
    codes = torch.randint(0, 2048, (1, 32, 140), generator=g)  # 280 transformer steps > the 250-step window
    with torch.no_grad():
        a_hf = hf.decode(codes).audio_values
    a_or = omimi.decode(codes, mimi_weights)
    assert a_hf.shape == a_or.shape == (1, 1, 140 * 1920)
    assert (a_hf - a_or).abs().max() < 1e-4 * a_or.abs().max()


def test_mimi_encode_vs_hf(mimi_weights):
    hf = build_hf_mimi(mimi_weights)
    audio = synthetic_audio(11, 12.0)[None, None, : 24000 * 12 - 700]  # ragged: not a multiple of 1920
    with torch.no_grad():
        c_hf = hf.encode(audio).audio_codes
    c_or = omimi.encode(audio, mimi_weights)
    assert c_hf.shape == c_or.shape == (1, 32, 150)
    assert (c_hf == c_or).float().mean() > 0.999


def _rand_lm_weights(cfg, g):
    def rnd(*s, std=0.02):
        return torch.empty(*s).normal_(0, std, generator=g)

    W = {"text_embeddings.weight": rnd(cfg.n_text_vocab, 2048), "audio_embeddings.weight": rnd(2051 * 32, 2048),
         "projection.weight": rnd(1024, 2048), "codebook0_head.weight": rnd(2051, 2048), "audio_head": rnd(31, 1024, 2051)}
    for name, c in (("backbone", cfg.backbone), ("decoder", cfg.decoder)):
        for l in range(c.n_layers):
            p = f"{name}.layers.{l}."
            W[p + "self_attn.q_proj.weight"] = rnd(c.n_heads * c.head_dim, c.d_model)
            W[p + "self_attn.k_proj.weight"] = rnd(c.n_kv_heads * c.head_dim, c.d_model)
            W[p + "self_attn.v_proj.weight"] = rnd(c.n_kv_heads * c.head_dim, c.d_model)
            W[p + "self_attn.o_proj.weight"] = rnd(c.d_model, c.n_heads * c.head_dim)
            W[p + "mlp.gate_proj.weight"] = rnd(c.d_ff, c.d_model)
            W[p + "mlp.up_proj.weight"] = rnd(c.d_ff, c.d_model)
            W[p + "mlp.down_proj.weight"] = rnd(c.d_model, c.d_ff)
            W[p + "input_layernorm.weight"] = 1 + rnd(c.d_model, std=0.05)
            W[p + "post_attention_layernorm.weight"] = 1 + rnd(c.d_model, std=0.05)
        W[f"{name}.norm.weight"] = 1 + rnd(c.d_model, std=0.05)
    return W


def _hf_llama_sd(W, name, c, prefix):
    sd = {}
    for l in range(c.n_layers):
        p, hp = f"{name}.layers.{l}.", f"{prefix}layers.{l}."
        sd[hp + "self_attn.q_proj.weight"] = _perm_rows(W[p + "self_attn.q_proj.weight"], c.n_heads, c.head_dim)
        sd[hp + "self_attn.k_proj.weight"] = _perm_rows(W[p + "self_attn.k_proj.weight"], c.n_kv_heads, c.head_dim)
        for k in ("self_attn.v_proj", "self_attn.o_proj", "mlp.gate_proj", "mlp.up_proj", "mlp.down_proj",
                  "input_layernorm", "post_attention_layernorm"):
            sd[hp + k + ".weight"] = W[p + k + ".weight"]
    sd[prefix + "norm.weight"] = W[f"{name}.norm.weight"]
    return sd


def test_lm_vs_hf_csm():
    """Real csm_1b head sizes (hd 64 / 128, GQA 32:8 / 8:2, Llama-3 RoPE scaling), 2 + 2 layers, small text vocab."""
    from transformers import CsmConfig, CsmDepthDecoderConfig
    from transformers.models.csm.modeling_csm import CsmBackboneModel, CsmDepthDecoderForCausalLM

    rope = {"rope_type": "llama3", "rope_theta": 5e5, "factor": 32.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
            "original_max_position_embeddings": 8192}
    cfg = olm.CSMCfg(olm.LlamaCfg(2, 2048, 32, 8, 64, 8192), olm.LlamaCfg(2, 1024, 8, 2, 128, 8192), n_text_vocab=1000)
    g = torch.Generator().manual_seed(5)
    W = _rand_lm_weights(cfg, g)
    orc = olm.OracleCSM(cfg, W)

    bc = CsmConfig(num_hidden_layers=2, rope_parameters=rope, text_vocab_size=1000)
    bc._attn_implementation = "eager"
    hb = CsmBackboneModel(bc).eval()
    sd = _hf_llama_sd(W, "backbone", cfg.backbone, "")
    sd["embed_tokens.embed_audio_tokens.weight"] = hb.state_dict()["embed_tokens.embed_audio_tokens.weight"]
    hb.load_state_dict(sd, strict=False)

    tok = torch.zeros(1, 11, 33, dtype=torch.int64)
    mask = torch.zeros(1, 11, 33, dtype=torch.bool)
    tok[0, :5, 32] = torch.randint(0, 1000, (5,), generator=g)
    mask[0, :5, 32] = True
    tok[0, 5:, :32] = torch.randint(0, 2051, (6, 32), generator=g)
    mask[0, 5:, :32] = True
    x = orc.embed_frames(tok, mask)
    cache = orc.new_backbone_cache()
    h_or = olm.llama_forward(x, orc.W, "backbone", cfg.backbone, orc.rope_b, cache)
    with torch.no_grad():
        h_hf = hb(inputs_embeds=x, position_ids=torch.arange(11)[None]).last_hidden_state
    assert (h_or - h_hf).abs().max() < 5e-5
    # one cached decode step == HF full recompute of 12 positions
    tok2 = torch.zeros(1, 1, 33, dtype=torch.int64)
    tok2[0, 0, :32] = torch.randint(0, 2051, (32,), generator=g)
    m2 = torch.zeros(1, 1, 33, dtype=torch.bool)
    m2[0, 0, :32] = True
    x2 = orc.embed_frames(tok2, m2)
    h2_or = olm.llama_forward(x2, orc.W, "backbone", cfg.backbone, orc.rope_b, cache)
    with torch.no_grad():
        h2_hf = hb(inputs_embeds=torch.cat([x, x2], 1), position_ids=torch.arange(12)[None]).last_hidden_state[:, -1:]
    assert (h2_or - h2_hf).abs().max() < 5e-5

    dc = CsmDepthDecoderConfig(num_hidden_layers=2, rope_parameters=rope)
    dc._attn_implementation = "eager"
    hd = CsmDepthDecoderForCausalLM(dc).eval()
    sd = _hf_llama_sd(W, "decoder", cfg.decoder, "model.")
    sd["model.embed_tokens.weight"] = W["audio_embeddings.weight"]
    sd["model.inputs_embeds_projector.weight"] = W["projection.weight"]
    sd["codebooks_head.weight"] = W["audio_head"]
    hd.load_state_dict(sd, strict=True)
    forced = torch.randint(0, 2051, (1, 32), generator=g)
    tr = {}
    olm.generate_frame(orc, tok, mask, orc.new_backbone_cache(), forced=forced, trace=tr)
    ids = torch.cat([torch.zeros(1, 1, dtype=torch.int64), forced[:, :31]], 1)
    with torch.no_grad():
        lg_hf = hd(input_ids=ids, backbone_last_hidden_state=tr["h"]).logits  # (1, 31, 2051): positions 1..31
    assert lg_hf.shape == (1, 31, 2051)
    err = max((tr["logits"][i][0] - lg_hf[0, i - 1]).abs().max().item() for i in range(1, 32))
    assert err < 5e-5

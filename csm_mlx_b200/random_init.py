"""Deterministic synthetic checkpoints (there is no network for real ones).

The reference constructs ``audio_head`` as zeros (``/root/reference/csm_mlx/models.py:65-67``) and
``MimiModel`` style codecs start with all-zero codebooks, so "random init" must explicitly fill
every tensor or codebooks 1..31 all argmax to 0 (SURVEY.md hazard H1).  This writes every tensor
of the parameter tree in SURVEY.md §3.4 with a seeded CPU generator, so the same bytes are
reproduced on any host with the same torch build.

Deviation from SURVEY.md §8d, on purpose: norm weights are 1 + 0.05·N(0,1) instead of exactly 1,
so a kernel that forgot the norm weight cannot pass parity.
"""

from __future__ import annotations

from typing import Dict

import torch

from .config import BACKBONE_CONFIGURATION, DECODER_CONFIGURATION, LlamaArgs


def _llama_shapes(prefix: str, a: LlamaArgs):
    d, hd = a.hidden_size, a.head_dim
    for l in range(a.num_hidden_layers):
        p = f"{prefix}.layers.{l}."
        yield p + "self_attn.q_proj.weight", (a.num_attention_heads * hd, d), "lin"
        yield p + "self_attn.k_proj.weight", (a.num_key_value_heads * hd, d), "lin"
        yield p + "self_attn.v_proj.weight", (a.num_key_value_heads * hd, d), "lin"
        yield p + "self_attn.o_proj.weight", (d, a.num_attention_heads * hd), "lin"
        yield p + "mlp.gate_proj.weight", (a.intermediate_size, d), "lin"
        yield p + "mlp.up_proj.weight", (a.intermediate_size, d), "lin"
        yield p + "mlp.down_proj.weight", (d, a.intermediate_size), "lin"
        yield p + "input_layernorm.weight", (d,), "norm"
        yield p + "post_attention_layernorm.weight", (d,), "norm"
    yield f"{prefix}.norm.weight", (d,), "norm"


def csm_param_shapes(backbone_name: str = "1b", decoder_name: str = "100m", n_text_vocab: int = 128_256,
                     n_audio_vocab: int = 2051, n_audio_codebooks: int = 32):
    """(name, shape, kind) for every parameter of CSM (models.py:32-77), in a fixed order."""
    b, d = BACKBONE_CONFIGURATION[backbone_name], DECODER_CONFIGURATION[decoder_name]
    db = b.num_attention_heads * b.head_dim
    dd = d.num_attention_heads * d.head_dim
    yield "text_embeddings.weight", (n_text_vocab, db), "lin"
    yield "audio_embeddings.weight", (n_audio_vocab * n_audio_codebooks, db), "lin"
    yield "projection.weight", (dd, db), "lin"
    yield "codebook0_head.weight", (n_audio_vocab, db), "lin"
    yield "audio_head", (n_audio_codebooks - 1, dd, n_audio_vocab), "lin"
    yield from _llama_shapes("backbone", b)
    yield from _llama_shapes("decoder", d)


def random_csm_weights(args=None, seed: int = 1234, std: float = 0.02,
                       dtype: torch.dtype = torch.bfloat16) -> Dict[str, torch.Tensor]:
    """All CSM tensors ~ N(0, std²) (norms 1 + 0.05·N(0,1), kept fp32), rounded to ``dtype``."""
    kw = {}
    if args is not None:
        kw = dict(backbone_name=args.backbone_name, decoder_name=args.decoder_name,
                  n_text_vocab=args.n_text_vocab, n_audio_vocab=args.n_audio_vocab,
                  n_audio_codebooks=args.n_audio_codebooks)
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shape, kind in csm_param_shapes(**kw):
        t = torch.empty(shape, dtype=torch.float32)
        if kind == "norm":
            t.normal_(0.0, 0.05, generator=g).add_(1.0)
            out[name] = t
        else:
            t.normal_(0.0, std, generator=g)
            out[name] = t.to(dtype)
    return out


# ----------------------------------------------------------------------------- Mimi codec
MIMI_RATIOS = (8, 6, 5, 4)
MIMI_DIM, MIMI_FILTERS, MIMI_FF, MIMI_LAYERS = 512, 64, 2048, 8
MIMI_CODEBOOK_DIM, MIMI_BINS = 256, 2048


def mimi_param_shapes(n_q: int = 32):
    """(name, shape, kind) for the Mimi codec in the moshi checkpoint key layout (weight-norm-free).

    kinds: conv/convtr (weight, fan-in scaled), bias, ln_w, ln_b, scale (LayerScale), cb (embedding_sum),
    usage (cluster_usage).
    """
    # SEANet encoder: conv k7; [res, ELU, conv k=2r s=r] for r in 4,5,6,8; ELU; conv k3
    def conv(name, cout, cin, k):
        yield name + ".weight", (cout, cin, k), "conv"
        yield name + ".bias", (cout,), "bias"

    yield from conv("encoder.model.0.conv.conv", MIMI_FILTERS, 1, 7)
    c, idx = MIMI_FILTERS, 1
    for r in reversed(MIMI_RATIOS):
        yield from conv(f"encoder.model.{idx}.block.1.conv.conv", c // 2, c, 3)
        yield from conv(f"encoder.model.{idx}.block.3.conv.conv", c, c // 2, 1)
        yield from conv(f"encoder.model.{idx + 2}.conv.conv", 2 * c, c, 2 * r)
        c, idx = 2 * c, idx + 3
    yield from conv(f"encoder.model.{idx + 1}.conv.conv", MIMI_DIM, c, 3)
    # SEANet decoder: conv k7; [ELU, convT k=2r s=r, res] for r in 8,6,5,4; ELU; conv k3
    yield from conv("decoder.model.0.conv.conv", c, MIMI_DIM, 7)
    idx = 1
    for r in MIMI_RATIOS:
        yield f"decoder.model.{idx + 1}.convtr.convtr.weight", (c, c // 2, 2 * r), "convtr"
        yield f"decoder.model.{idx + 1}.convtr.convtr.bias", (c // 2,), "bias"
        c //= 2
        yield from conv(f"decoder.model.{idx + 2}.block.1.conv.conv", c // 2, c, 3)
        yield from conv(f"decoder.model.{idx + 2}.block.3.conv.conv", c, c // 2, 1)
        idx += 3
    yield from conv(f"decoder.model.{idx + 1}.conv.conv", 1, c, 3)
    for tr in ("encoder_transformer", "decoder_transformer"):
        for l in range(MIMI_LAYERS):
            p = f"{tr}.transformer.layers.{l}."
            yield p + "self_attn.in_proj_weight", (3 * MIMI_DIM, MIMI_DIM), "conv"
            yield p + "self_attn.out_proj.weight", (MIMI_DIM, MIMI_DIM), "conv"
            yield p + "linear1.weight", (MIMI_FF, MIMI_DIM), "conv"
            yield p + "linear2.weight", (MIMI_DIM, MIMI_FF), "conv"
            for n in ("norm1", "norm2"):
                yield p + n + ".weight", (MIMI_DIM,), "ln_w"
                yield p + n + ".bias", (MIMI_DIM,), "ln_b"
            yield p + "layer_scale_1.scale", (MIMI_DIM,), "scale"
            yield p + "layer_scale_2.scale", (MIMI_DIM,), "scale"
    yield "downsample.conv.conv.conv.weight", (MIMI_DIM, MIMI_DIM, 4), "conv"
    yield "upsample.convtr.convtr.convtr.weight", (MIMI_DIM, 1, 4), "convtr_dw"
    for group, n in (("rvq_first", 1), ("rvq_rest", n_q - 1)):
        yield f"quantizer.{group}.input_proj.weight", (MIMI_CODEBOOK_DIM, MIMI_DIM, 1), "conv"
        yield f"quantizer.{group}.output_proj.weight", (MIMI_DIM, MIMI_CODEBOOK_DIM, 1), "conv"
        for i in range(n):
            p = f"quantizer.{group}.vq.layers.{i}._codebook."
            yield p + "embedding_sum", (MIMI_BINS, MIMI_CODEBOOK_DIM), "cb"
            yield p + "cluster_usage", (MIMI_BINS,), "usage"


def random_mimi_weights(seed: int = 4321, n_q: int = 32) -> Dict[str, torch.Tensor]:
    """fp32 Mimi weights: convs/linears N(0, 1/fan_in), small biases, LayerScale ≈ 0.1 (larger than the
    0.01 initial value so the transformer branches matter in parity tests), codebooks N(0,1) with
    cluster_usage ~ U(0.5, 2)."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shape, kind in mimi_param_shapes(n_q):
        t = torch.empty(shape, dtype=torch.float32)
        if kind == "conv":
            t.normal_(0.0, 1.0, generator=g).mul_((shape[1] * (shape[2] if len(shape) > 2 else 1)) ** -0.5)
        elif kind == "convtr":  # (cin, cout, k): each output sample sums cin * k/stride taps
            t.normal_(0.0, 1.0, generator=g).mul_((shape[0] * 2) ** -0.5)
        elif kind == "convtr_dw":
            t.normal_(0.0, 1.0, generator=g).mul_(0.5)
        elif kind == "bias" or kind == "ln_b":
            t.normal_(0.0, 0.05, generator=g)
        elif kind == "ln_w":
            t.normal_(0.0, 0.05, generator=g).add_(1.0)
        elif kind == "scale":
            t.normal_(0.0, 0.02, generator=g).add_(0.1)
        elif kind == "cb":
            t.normal_(0.0, 1.0, generator=g)
        elif kind == "usage":
            t.uniform_(0.5, 2.0, generator=g)
        out[name] = t
    return out

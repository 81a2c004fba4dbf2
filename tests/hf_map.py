"""Helpers that copy oracle-layout weights into the independent HF implementations used to pin the oracle.

HF differences that the mapping absorbs:
* HF applies RoPE in rotate-half form, moshi/csm_mlx in adjacent-pair form ⇒ permute the rows of every
  q_proj / k_proj head from [0,1,2,…] to [0,2,4,…,1,3,5,…].
* HF Mimi splits moshi's fused ``in_proj_weight`` into q/k/v and renames most modules.
"""

from __future__ import annotations

import re
from typing import Dict

import torch


def _perm_rows(w: torch.Tensor, n_heads: int, head_dim: int) -> torch.Tensor:
    """adjacent-pair → rotate-half row order, per head."""
    idx = torch.cat([torch.arange(0, head_dim, 2), torch.arange(1, head_dim, 2)])
    return w.reshape(n_heads, head_dim, -1)[:, idx, :].reshape(n_heads * head_dim, -1)


def mimi_to_hf(W: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    out: Dict[str, torch.Tensor] = {}
    for k, v in W.items():
        m = re.match(r"(encoder|decoder)\.model\.(\d+)\.(block\.\d+\.)?(conv\.conv|convtr\.convtr)\.(weight|bias)", k)
        if m:
            side, n, blk, _, wb = m.groups()
            out[f"{side}.layers.{n}.{blk or ''}conv.{wb}"] = v
            continue
        m = re.match(r"(encoder|decoder)_transformer\.transformer\.layers\.(\d+)\.(.+)", k)
        if m:
            side, l, rest = m.groups()
            p = f"{side}_transformer.layers.{l}."
            if rest == "self_attn.in_proj_weight":
                q, kk, vv = v.chunk(3, dim=0)
                out[p + "self_attn.q_proj.weight"] = _perm_rows(q, 8, 64)
                out[p + "self_attn.k_proj.weight"] = _perm_rows(kk, 8, 64)
                out[p + "self_attn.v_proj.weight"] = vv
            else:
                rest = (rest.replace("self_attn.out_proj", "self_attn.o_proj")
                        .replace("linear1", "mlp.fc1").replace("linear2", "mlp.fc2")
                        .replace("norm1", "input_layernorm").replace("norm2", "post_attention_layernorm")
                        .replace("layer_scale_1", "self_attn_layer_scale").replace("layer_scale_2", "mlp_layer_scale"))
                out[p + rest] = v
            continue
        if k == "downsample.conv.conv.conv.weight":
            out["downsample.conv.weight"] = v
            continue
        if k == "upsample.convtr.convtr.convtr.weight":
            out["upsample.conv.weight"] = v
            continue
        m = re.match(r"quantizer\.(rvq_first|rvq_rest)\.(.+)", k)
        if m:
            g, rest = m.groups()
            hg = "semantic_residual_vector_quantizer" if g == "rvq_first" else "acoustic_residual_vector_quantizer"
            rest = (rest.replace("vq.layers", "layers").replace("_codebook", "codebook")
                    .replace("embedding_sum", "embed_sum"))
            out[f"quantizer.{hg}.{rest}"] = v
            continue
        raise KeyError(k)
    return out


def build_hf_mimi(W: Dict[str, torch.Tensor]):
    from transformers import MimiConfig, MimiModel

    cfg = MimiConfig()
    cfg._attn_implementation = "eager"
    model = MimiModel(cfg).eval()
    sd = model.state_dict()
    mapped = mimi_to_hf(W)
    missing = [k for k in sd if k not in mapped and not k.endswith("initialized")]
    extra = [k for k in mapped if k not in sd]
    assert not missing and not extra, (missing[:5], extra[:5])
    for k in sd:
        if k.endswith("initialized"):
            mapped[k] = torch.ones_like(sd[k])
    model.load_state_dict(mapped, strict=True)
    # the codebook property caches embed_sum / usage lazily; make sure nothing stale is cached
    for mod in model.modules():
        if hasattr(mod, "_embed"):
            mod._embed = None
    return model

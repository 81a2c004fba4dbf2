"""``load_adapters`` for the generation path: fold fine-tuned adapters into the dense weights at load time.

The reference wraps the target Linears in ``mlx_lm.tuner.lora.LoRALinear`` modules and keeps the low-rank pair at
run time (``/root/reference/csm_mlx/finetune/utils.py:16-81, 84-108``): y = x W^T + scale * ((x A) B), with
``lora_a`` (in, r) and ``lora_b`` (r, out) stored in ``adapters.safetensors`` under ``<module>.lora_a`` /
``<module>.lora_b`` and ``scale`` in ``adapter_config.json["lora_parameters"]``.  The hot path here streams each
weight matrix once per step, so the adapter is merged instead: W' = W + scale * (A B)^T — what ``mlx_lm fuse``
produces — and the kernels stay unchanged.  ``fine_tune_type == "full"`` is a plain non-strict weight load;
DoRA needs the column norms of W' at run time and is not supported.  Training itself is out of scope.
"""

from __future__ import annotations

import json
import os
from typing import Dict

import torch

_TARGETS = ("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.o_proj",
            "mlp.gate_proj", "mlp.up_proj", "mlp.down_proj")


def merge_lora(weights: Dict[str, torch.Tensor], adapters: Dict[str, torch.Tensor], scale: float) -> Dict[str, torch.Tensor]:
    """Returns ``weights`` with every (``<p>.lora_a``, ``<p>.lora_b``) pair of ``adapters`` folded into ``<p>.weight``
    in fp32: W' = W + scale * (lora_a @ lora_b)^T.  Unknown prefixes raise ``ValueError``."""
    out = dict(weights)
    prefixes = sorted({k[: -len(".lora_a")] for k in adapters if k.endswith(".lora_a")})
    for p in prefixes:
        p_base = p[: -len(".linear")] if p.endswith(".linear") else p
        name = p_base + ".weight"
        if name not in out:
            raise ValueError(f"adapter targets {p_base}, which is not a Linear of this model")
        if p + ".lora_b" not in adapters:
            raise ValueError(f"adapter has {p}.lora_a but no {p}.lora_b")
        a = adapters[p + ".lora_a"].to(torch.float32)
        b = adapters[p + ".lora_b"].to(torch.float32)
        w = out[name].to(torch.float32)
        if a.shape[0] != w.shape[1] or b.shape[1] != w.shape[0] or a.shape[1] != b.shape[0]:
            raise ValueError(f"adapter shapes {tuple(a.shape)} x {tuple(b.shape)} do not match {name} {tuple(w.shape)}")
        out[name] = w + float(scale) * (a @ b).t()
    return out


def load_adapters(model, adapter_path: str):
    """finetune/utils.py:84-108: ``adapter_path`` holds ``adapter_config.json`` and ``adapters.safetensors``.
    The model must already hold its base weights; returns the model with the adapter folded in."""
    from safetensors.torch import load_file

    if not os.path.isdir(adapter_path):
        raise FileNotFoundError(f"The adapter path does not exist: {adapter_path}")
    with open(os.path.join(adapter_path, "adapter_config.json")) as fid:
        config = json.load(fid)
    tensors = load_file(os.path.join(adapter_path, "adapters.safetensors"))
    kind = config.get("fine_tune_type", "lora")
    if kind == "full":
        return model.load_weights(tensors, strict=False)
    if kind != "lora":
        raise NotImplementedError(f"fine_tune_type={kind!r}: only 'lora' and 'full' adapters can be folded into dense weights")
    scale = float(config.get("lora_parameters", {}).get("scale", 1.0))
    base = {k: v.detach().to("cpu", torch.float32) for k, v in model.parameters().items()}
    merged = merge_lora(base, {k: v for k, v in tensors.items() if k.endswith((".lora_a", ".lora_b"))}, scale)
    rest = {k: v for k, v in tensors.items() if not k.endswith((".lora_a", ".lora_b")) and k in base}
    merged.update(rest)
    return model.load_weights(merged, strict=True)

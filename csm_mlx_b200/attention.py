"""Llama-3 scaled RoPE table, host side.

Restates ``/root/reference/csm_mlx/attention.py:57-117`` (``Llama3ScaledRoPE.rope_init``,
``build_rope_cache``, ``apply_scaling``): fp32 throughout, table of ``max_seq_len`` positions holding
(cos, sin) per adjacent-pair frequency.  The rotation itself (attention.py:119-177) and the GQA attention
(attention.py:207-253) run in libcsm_b200.so (``csmb_rope_kv_append`` / ``csmb_attention`` and the fused
frame kernel); the table is uploaded once per model.
"""

from __future__ import annotations

import math

import torch


def llama3_scaled_freqs(head_dim: int, base: float, scale_factor: float, low_freq_factor: int = 1,
                        high_freq_factor: int = 4, old_context_len: int = 8192) -> torch.Tensor:
    freqs = 1.0 / (base ** (torch.arange(0, head_dim, 2)[: head_dim // 2].to(torch.float32) / head_dim))
    wavelen = 2 * math.pi / freqs
    smooth = (old_context_len / wavelen - low_freq_factor) / (high_freq_factor - low_freq_factor)
    blended = (1 - smooth) * freqs / scale_factor + smooth * freqs
    out = torch.where(wavelen < old_context_len / high_freq_factor, freqs,
                      torch.where(wavelen > old_context_len / low_freq_factor, freqs / scale_factor, blended))
    return out.to(torch.float32)


def llama3_rope_table(head_dim: int, base: float, scale_factor: float, max_seq_len: int = 2048) -> torch.Tensor:
    """(max_seq_len, head_dim/2, 2) fp32: [..., 0] = cos(p·θ'), [..., 1] = sin(p·θ')."""
    theta = llama3_scaled_freqs(head_dim, base, scale_factor)
    idx_theta = torch.einsum("i,j->ij", torch.arange(max_seq_len, dtype=torch.float32), theta).to(torch.float32)
    return torch.stack([torch.cos(idx_theta), torch.sin(idx_theta)], dim=-1).contiguous()

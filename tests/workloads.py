"""Synthetic inputs shared by tests, bench.py and scripts/make_golden.py."""
from __future__ import annotations

import math
from typing import List, Tuple

import torch


def cfg1_prompt_ids() -> List[int]:
    """Stand-in for "[0]Hello from Sesame.": BOS + 8 seeded ids + EOS (the Llama tokenizer is not available offline)."""
    g = torch.Generator().manual_seed(7)
    return [128000] + torch.randint(0, 128000, (8,), generator=g).tolist() + [128001]


def prompt_ids(seed: int, n: int) -> List[int]:
    g = torch.Generator().manual_seed(seed)
    return [128000] + torch.randint(0, 128000, (n,), generator=g).tolist() + [128001]


def synthetic_audio(seed: int, seconds: float, sr: int = 24000) -> torch.Tensor:
    """0.3·sin(2π·220t) + 0.2·sin(2π·3300t) + 0.05·N(0,1), clipped to [-1, 1]."""
    n = int(round(seconds * sr))
    t = torch.arange(n, dtype=torch.float32) / sr
    g = torch.Generator().manual_seed(seed)
    x = 0.3 * torch.sin(2 * math.pi * 220 * t) + 0.2 * torch.sin(2 * math.pi * 3300 * t) \
        + 0.05 * torch.randn(n, generator=g)
    return x.clamp(-1.0, 1.0)


def tiny_prompt() -> Tuple[torch.Tensor, torch.Tensor]:
    """Mixed text + audio rows for the tiny configuration (4 codebooks, text vocab 512, audio vocab 67)."""
    g = torch.Generator().manual_seed(3)
    tok = torch.zeros((9, 5), dtype=torch.int64)
    mask = torch.zeros((9, 5), dtype=torch.bool)
    tok[:4, 4] = torch.randint(0, 512, (4,), generator=g)
    mask[:4, 4] = True
    tok[4:8, :4] = torch.randint(0, 67, (4, 4), generator=g)
    mask[4:8, :4] = True
    tok[8, 4] = 5
    mask[8, 4] = True
    return tok, mask

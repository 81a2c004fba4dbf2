"""mlx_lm-style samplers and logits processors.

The reference's README/CLI pass ``sampler=make_sampler(temp, top_p, min_p, min_tokens_to_keep, top_k)`` from
``mlx_lm.sample_utils`` (``/root/reference/README.md:43-52``, ``csm_mlx/cli/generate.py:168-174``) and logits
processors from ``make_logits_processors`` (README.md:120-122).  mlx_lm is not vendored; this restates the
call surface.  ``make_sampler`` returns a ``DeviceSampler``: callable on a ``(B,V)`` logits tensor like the
mlx_lm closure, and recognised by ``generate`` so that sampling stays on the GPU inside the frame kernels.

Semantics: temp 0 -> argmax; otherwise top-k, then top-p (0<p<1), then min-p (keeping at least
``min_tokens_to_keep``) mask tokens on softmax(logits), then categorical(logits / temp) (Gumbel-max, Philox).
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Dict, List, Optional

import torch

from . import _lib
from .runtime import SamplerSpec


class DeviceSampler:
    def __init__(self, spec: SamplerSpec):
        self.spec = spec
        self._calls = 0
        self._own_seed = spec.seed if spec.seed is not None else int.from_bytes(os.urandom(8), "little")

    def __call__(self, logits: torch.Tensor) -> torch.Tensor:
        """(B,V) or (V,) device logits -> (B,) / () int32 ids."""
        squeeze = logits.dim() == 1
        lg = logits.reshape(-1, logits.shape[-1]).to(torch.float32).contiguous()
        dev = _lib.require_device(lg.device)
        out = torch.empty((lg.shape[0],), device=lg.device, dtype=torch.int32)
        s = SamplerSpec(**{**self.spec.__dict__, "seed": self._own_seed}).to_c()
        _lib.check(_lib.lib().csmb_sample(lg.data_ptr(), lg.shape[1], out.data_ptr(), 1, lg.shape[0], lg.shape[1],
                                          C.byref(s), self._calls, None, 0, dev, _lib.stream_ptr(lg.device)))
        self._calls += 1
        return out[0] if squeeze else out


def make_sampler(temp: float = 0.0, top_p: float = 0.0, min_p: float = 0.0, min_tokens_to_keep: int = 1,
                 top_k: int = -1, seed: Optional[int] = None) -> DeviceSampler:
    """``seed=None`` (default): a fresh Philox key is drawn from the OS for every ``generate`` call that uses the sampler,
    like ``mx.random.categorical`` without an explicit key; pass a seed for reproducible utterances."""
    return DeviceSampler(SamplerSpec(temperature=float(temp), top_k=int(top_k) if top_k and top_k > 0 else 0,
                                     top_p=float(top_p), min_p=float(min_p),
                                     min_tokens_to_keep=int(min_tokens_to_keep),
                                     seed=None if seed is None else int(seed)))


def make_logits_processors(logit_bias: Optional[Dict[int, float]] = None, repetition_penalty: Optional[float] = None,
                           repetition_context_size: Optional[int] = 20) -> List[Callable]:
    """Callables ``(token_history, logits) -> logits`` applied to the codebook-0 logits only
    (generation.py:44-49).  ``token_history`` is the stacked c0 history ``(n, B, 1)`` or an empty tensor."""
    procs: List[Callable] = []
    if logit_bias:
        idx = torch.tensor(list(logit_bias.keys()), dtype=torch.long)
        val = torch.tensor(list(logit_bias.values()), dtype=torch.float32)

        def bias_proc(_, logits):
            logits = logits.clone()
            logits[:, idx.to(logits.device)] += val.to(logits.device)
            return logits

        procs.append(bias_proc)
    if repetition_penalty and repetition_penalty != 0:
        if repetition_penalty < 0:
            raise ValueError("repetition_penalty must be a non-negative float")

        def rep_proc(tokens, logits):
            if tokens.numel() == 0:
                return logits
            hist = tokens.reshape(tokens.shape[0], -1)[-repetition_context_size:].to(logits.device).long()  # (n,B)
            logits = logits.clone()
            for b in range(logits.shape[0]):
                ids = hist[:, b]
                sel = logits[b, ids]
                logits[b, ids] = torch.where(sel < 0, sel * repetition_penalty, sel / repetition_penalty)
            return logits

        procs.append(rep_proc)
    return procs

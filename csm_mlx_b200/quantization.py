"""Weight-only FP8 for the Linear layers of a loaded CSM: the B200 analogue of ``mlx.nn.quantize(csm)``
(/root/reference README.md:92-128 quantises the model in place before ``generate``).

Format (include/csm_b200.h, CSMB_WEIGHTS_E4M3): a matrix ``W[N][K]`` becomes one blob — ``N`` fp32 per-output-channel scales
(``max|W[n]| / 448``), padded to 256 bytes, then ``N * K`` OCP E4M3 bytes: ``W[n][k] ≈ scale[n] * e4m3[n][k]``.  Embedding
tables and norm weights stay as they are (they are gathered / applied, not streamed as matrices).  A quantised model is served
by the batch-1 frame kernel (``csmb_frame_b1``: a second instantiation of the persistent kernel for one-byte weights) and by the
row-based GEMV kernels (``csmb_backbone_forward`` / ``csmb_depth_decode`` / ``csmb_decode_frame``); both widen e4m3 to fp32 in
registers, accumulate in fp32 and apply the scale to the finished dot product.  The tensor-core chain declines it.
"""
from __future__ import annotations

from typing import Tuple

import torch

E4M3_MAX = 448.0


def quantize_rows_e4m3(w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """``w`` (N, K) any float dtype → (uint8 (N, K) e4m3 bit patterns, fp32 (N,) scales); round to nearest even."""
    w32 = w.to(torch.float32)
    scale = (w32.abs().amax(dim=1) / E4M3_MAX).clamp_min(torch.finfo(torch.float32).tiny)
    q = (w32 / scale[:, None]).clamp_(-E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn)
    return q.view(torch.uint8), scale


def dequantize_rows_e4m3(q: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """fp32 (N, K) = scale[n] * e4m3[n][k] — exactly the numbers the kernels multiply with."""
    return q.view(torch.float8_e4m3fn).to(torch.float32) * scale.to(torch.float32)[:, None]


def blob_bytes(n: int, k: int) -> int:
    return ((n * 4 + 255) & ~255) + ((n * k + 255) & ~255)


def pack_blob(q: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """One uint8 tensor in the layout of include/csm_b200.h (scales, pad, bytes, pad), on the device of ``q``."""
    n, k = q.shape
    out = torch.zeros((blob_bytes(n, k),), dtype=torch.uint8, device=q.device)
    out[: n * 4] = scale.to(torch.float32).contiguous().view(torch.uint8).reshape(-1)
    off = (n * 4 + 255) & ~255
    out[off: off + n * k] = q.contiguous().reshape(-1)
    return out


def unpack_blob(blob: torch.Tensor, n: int, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    off = (n * 4 + 255) & ~255
    scale = blob[: n * 4].view(torch.float32).clone()
    return blob[off: off + n * k].reshape(n, k), scale


def quantize(model, bits: int = 8, group_size=None):
    """``nn.quantize(csm)`` for this package: quantises ``model`` in place (weight-only FP8 E4M3, one scale per output
    channel) and returns it.  ``bits`` other than 8 / a ``group_size`` are not offered (E4M3 with per-channel scales is the
    one format the kernels read)."""
    if bits != 8 or group_size is not None:
        raise NotImplementedError("csm_mlx_b200.quantize: weight-only FP8 (bits=8, per-output-channel scales) only")
    model.quantize_weights()
    return model

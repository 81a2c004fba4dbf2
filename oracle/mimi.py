"""Codec half of the oracle: Mimi ("mimi_202407", 32 codebooks) encode / decode / streaming decode.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference calls ``moshi_mlx.models.mimi.Mimi``
(``/root/reference/csm_mlx/tokenizers.py:6,14-21,70,150``; ``generation.py:224-225,251,258``), an
un-vendored dependency (pyproject pin ``moshi-mlx>=0.2.3``) that cannot be installed here.  This file
restates the published architecture (kyutai moshi ``mimi_202407``) with
plain torch CPU ops and is pinned against HF ``transformers`` ``MimiModel`` (an independent
implementation of the same codec) in ``tests/test_oracle_vs_hf.py``.

Weights: flat dict in the moshi checkpoint key layout (weight-norm-free), e.g.
``encoder.model.0.conv.conv.weight``, ``decoder.model.2.convtr.convtr.weight``,
``decoder_transformer.transformer.layers.0.self_attn.in_proj_weight``,
``quantizer.rvq_rest.vq.layers.3._codebook.embedding_sum``.
"""

from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

SAMPLE_RATE = 24_000
FRAME = 1920
RATIOS = (8, 6, 5, 4)  # decoder order; the encoder uses them reversed
N_FILTERS = 64
DIM = 512
N_LAYERS = 8
N_HEADS = 8
HEAD_DIM = 64
FF = 2048
CONTEXT = 250
ROPE_BASE = 10_000.0
CODEBOOK_DIM = 256
BINS = 2048
LN_EPS = 1e-5


# ----------------------------------------------------------------------------- convolutions
def conv1d(x: Tensor, w: Tensor, b: Optional[Tensor], stride: int = 1, pad_mode: str = "constant") -> Tensor:
    """Causal conv (moshi StreamingConv1d, offline): left pad k-stride, right pad to a whole window."""
    k = w.shape[-1]
    pad_total = k - stride
    length = x.shape[-1]
    n_frames = math.ceil((length - k + pad_total) / stride + 1) - 1
    extra = n_frames * stride + k - pad_total - length
    x = F.pad(x, (pad_total, extra), mode=pad_mode)
    return F.conv1d(x, w, b, stride=stride)


def convtr1d(x: Tensor, w: Tensor, b: Optional[Tensor], stride: int, groups: int = 1) -> Tensor:
    """Causal transposed conv (moshi StreamingConvTranspose1d, offline): drop the last k-stride samples."""
    k = w.shape[-1]
    y = F.conv_transpose1d(x, w, b, stride=stride, groups=groups)
    return y[..., : y.shape[-1] - (k - stride)]


def resblock(x: Tensor, W: Dict[str, Tensor], p: str) -> Tensor:
    """SEANetResnetBlock, one residual layer, dilation 1, true_skip: x + conv1(ELU(conv3(ELU(x))))."""
    h = conv1d(F.elu(x), W[p + ".block.1.conv.conv.weight"], W[p + ".block.1.conv.conv.bias"])
    h = conv1d(F.elu(h), W[p + ".block.3.conv.conv.weight"], W[p + ".block.3.conv.conv.bias"])
    return x + h


def seanet_encoder(x: Tensor, W: Dict[str, Tensor]) -> Tensor:
    """(B,1,N) → (B,512,N/960): conv k7; [res, ELU, conv k=2r s=r] for r in 4,5,6,8; ELU; conv k3."""
    h = conv1d(x, W["encoder.model.0.conv.conv.weight"], W["encoder.model.0.conv.conv.bias"])
    idx = 1
    for r in reversed(RATIOS):
        h = resblock(h, W, f"encoder.model.{idx}")
        h = conv1d(F.elu(h), W[f"encoder.model.{idx + 2}.conv.conv.weight"],
                   W[f"encoder.model.{idx + 2}.conv.conv.bias"], stride=r)
        idx += 3
    return conv1d(F.elu(h), W[f"encoder.model.{idx + 1}.conv.conv.weight"], W[f"encoder.model.{idx + 1}.conv.conv.bias"])


def seanet_decoder(z: Tensor, W: Dict[str, Tensor]) -> Tensor:
    """(B,512,T) → (B,1,960·T): conv k7; [ELU, convT k=2r s=r, res] for r in 8,6,5,4; ELU; conv k3."""
    h = conv1d(z, W["decoder.model.0.conv.conv.weight"], W["decoder.model.0.conv.conv.bias"])
    idx = 1
    for r in RATIOS:
        h = convtr1d(F.elu(h), W[f"decoder.model.{idx + 1}.convtr.convtr.weight"],
                     W[f"decoder.model.{idx + 1}.convtr.convtr.bias"], stride=r)
        h = resblock(h, W, f"decoder.model.{idx + 2}")
        idx += 3
    return conv1d(F.elu(h), W[f"decoder.model.{idx + 1}.conv.conv.weight"], W[f"decoder.model.{idx + 1}.conv.conv.bias"])


# ----------------------------------------------------------------------------- transformer
def rope_adjacent(x: Tensor, positions: Tensor) -> Tensor:
    """moshi rope: adjacent pairs, θ_i = 10000^(-2i/D), fp32.  x (B,H,T,D), positions (T,)."""
    D = x.shape[-1]
    freqs = torch.exp(torch.arange(D // 2, dtype=torch.float32) * (-math.log(ROPE_BASE) * 2 / D))
    ang = positions.to(torch.float32)[:, None] * freqs[None, :]  # (T, D/2)
    cos, sin = torch.cos(ang), torch.sin(ang)
    xs = x.to(torch.float32).reshape(*x.shape[:-1], D // 2, 2)
    xr, xi = xs[..., 0], xs[..., 1]
    return torch.stack([xr * cos - xi * sin, xr * sin + xi * cos], dim=-1).flatten(-2)


class TransformerState:
    """Streaming state: per-layer K/V (all positions kept; the window is applied by the mask) + offset."""

    def __init__(self) -> None:
        self.k: List[Optional[Tensor]] = [None] * N_LAYERS
        self.v: List[Optional[Tensor]] = [None] * N_LAYERS
        self.offset = 0


def transformer(x: Tensor, W: Dict[str, Tensor], p: str, state: Optional[TransformerState] = None) -> Tensor:
    """8 pre-LayerNorm layers, d512, 8 heads, GELU FF 2048, LayerScale, RoPE, causal window 250.  x (B,T,512)."""
    B, T, _ = x.shape
    off = state.offset if state is not None else 0
    qpos = torch.arange(off, off + T)
    for l in range(N_LAYERS):
        lp = f"{p}.transformer.layers.{l}."
        n = F.layer_norm(x, (DIM,), W[lp + "norm1.weight"], W[lp + "norm1.bias"], LN_EPS)
        qkv = F.linear(n, W[lp + "self_attn.in_proj_weight"]).reshape(B, T, 3, N_HEADS, HEAD_DIM)
        q, k, v = (qkv[:, :, i].transpose(1, 2) for i in range(3))  # (B,H,T,hd)
        q, k = rope_adjacent(q, qpos), rope_adjacent(k, qpos)
        if state is not None:
            if state.k[l] is not None:
                k = torch.cat([state.k[l], k], dim=2)
                v = torch.cat([state.v[l], v], dim=2)
            state.k[l], state.v[l] = k[:, :, -CONTEXT:], v[:, :, -CONTEXT:]
        S = k.shape[2]
        kpos = torch.arange(off + T - S, off + T)
        delta = qpos[:, None] - kpos[None, :]
        allowed = (delta >= 0) & (delta < CONTEXT)
        s = torch.matmul(q, k.transpose(2, 3)) * (HEAD_DIM ** -0.5)
        s = s.masked_fill(~allowed, float("-inf"))
        a = torch.matmul(torch.softmax(s, dim=-1), v).transpose(1, 2).reshape(B, T, DIM)
        x = x + W[lp + "layer_scale_1.scale"] * F.linear(a, W[lp + "self_attn.out_proj.weight"])
        n = F.layer_norm(x, (DIM,), W[lp + "norm2.weight"], W[lp + "norm2.bias"], LN_EPS)
        f = F.linear(F.gelu(F.linear(n, W[lp + "linear1.weight"])), W[lp + "linear2.weight"])
        x = x + W[lp + "layer_scale_2.scale"] * f
    if state is not None:
        state.offset += T
    return x


# ----------------------------------------------------------------------------- quantiser
def codebook(W: Dict[str, Tensor], group: str, i: int) -> Tensor:
    p = f"quantizer.{group}.vq.layers.{i}._codebook."
    return W[p + "embedding_sum"] / W[p + "cluster_usage"].clamp(min=1e-5)[:, None]


def rvq_decode(codes: Tensor, W: Dict[str, Tensor]) -> Tensor:
    """codes (B,K,F) → (B,512,F): out_proj_first(C0[c0]) + out_proj_rest(Σ_{k≥1} Ck[ck])."""
    K = codes.shape[1]
    sem = F.embedding(codes[:, 0], codebook(W, "rvq_first", 0)).transpose(1, 2)  # (B,256,F)
    out = F.conv1d(sem, W["quantizer.rvq_first.output_proj.weight"])
    if K > 1:
        ac = sum(F.embedding(codes[:, k], codebook(W, "rvq_rest", k - 1)) for k in range(1, K)).transpose(1, 2)
        out = out + F.conv1d(ac, W["quantizer.rvq_rest.output_proj.weight"])
    return out


def _rvq_encode_group(x: Tensor, W: Dict[str, Tensor], group: str, n_q: int) -> List[Tensor]:
    r = F.conv1d(x, W[f"quantizer.{group}.input_proj.weight"]).transpose(1, 2)  # (B,F,256)
    out = []
    for i in range(n_q):
        C = codebook(W, group, i)
        # argmin ‖r−c‖² = argmin (‖c‖² − 2 r·c); first index on ties
        d = (C * C).sum(-1)[None, None, :] - 2.0 * torch.matmul(r, C.t())
        idx = d.argmin(dim=-1)
        out.append(idx)
        r = r - F.embedding(idx, C)
    return out


def rvq_encode(x: Tensor, W: Dict[str, Tensor], n_q: int = 32) -> Tensor:
    """latent (B,512,F) → codes (B,n_q,F); semantic and acoustic groups both start from x."""
    idx = _rvq_encode_group(x, W, "rvq_first", 1)
    if n_q > 1:
        idx += _rvq_encode_group(x, W, "rvq_rest", n_q - 1)
    return torch.stack(idx, dim=1)


# ----------------------------------------------------------------------------- pipelines
def encode(audio: Tensor, W: Dict[str, Tensor], n_q: int = 32) -> Tensor:
    """(B,1,N) fp32 → (B,n_q,ceil(N/1920)) int64."""
    h = seanet_encoder(audio.to(torch.float32), W)
    h = transformer(h.transpose(1, 2), W, "encoder_transformer").transpose(1, 2)
    h = conv1d(h, W["downsample.conv.conv.conv.weight"], None, stride=2, pad_mode="replicate")
    return rvq_encode(h, W, n_q)


def encode_latent(audio: Tensor, W: Dict[str, Tensor]) -> Tensor:
    """(B,1,N) fp32 → (B,512,F): the 12.5 Hz latent the quantiser sees (``encode`` without ``rvq_encode``)."""
    h = seanet_encoder(audio.to(torch.float32), W)
    h = transformer(h.transpose(1, 2), W, "encoder_transformer").transpose(1, 2)
    return conv1d(h, W["downsample.conv.conv.conv.weight"], None, stride=2, pad_mode="replicate")


def rvq_disagreement_margins(latent: Tensor, W: Dict[str, Tensor], ref: Tensor, other: Tensor) -> List[tuple]:
    """Referee for two code tensors (B,K,F) of the same latent (B,512,F).  A nearest-neighbour search is index work, but
    its input is floating point: an implementation that sums in another order may pick the other of two (almost)
    equidistant codewords, after which the rest of that frame's residual chain legitimately differs.  For every frame and
    chain (semantic: codebook 0; acoustic: codebooks 1..K-1) this walks the chain in float64 along ``ref``'s codes to the
    FIRST codebook where ``other`` differs and returns (b, f, k, |d_ref - d_other| / |r|^2) there: the relative distance gap
    of the two candidates, which the caller bounds (a genuine near-tie) — disagreements past that point are not examined."""
    out = []
    B, K, F_ = ref.shape
    lat = latent.to(torch.float64)
    for group, k0, n in (("rvq_first", 0, 1), ("rvq_rest", 1, K - 1)):
        if n <= 0:
            continue
        wp = W[f"quantizer.{group}.input_proj.weight"].to(torch.float64)[:, :, 0]         # (256,512)
        cbs = [codebook(W, group, i).to(torch.float64) for i in range(n)]
        for b in range(B):
            for f in range(F_):
                if bool((ref[b, k0:k0 + n, f] == other[b, k0:k0 + n, f]).all()):
                    continue
                r = wp @ lat[b, :, f]
                for i in range(n):
                    a_, o_ = int(ref[b, k0 + i, f]), int(other[b, k0 + i, f])
                    if a_ != o_:
                        da, do = (r - cbs[i][a_]).pow(2).sum(), (r - cbs[i][o_]).pow(2).sum()
                        out.append((b, f, k0 + i, float((da - do).abs() / r.pow(2).sum().clamp_min(1e-300))))
                        break
                    r = r - cbs[i][a_]
    return out


def decode_latent(codes: Tensor, W: Dict[str, Tensor]) -> Tensor:
    """codes → (B,512,2F) transformer input (after RVQ decode and the depthwise ×2 upsample)."""
    z = rvq_decode(codes, W)
    return convtr1d(z, W["upsample.convtr.convtr.convtr.weight"], None, stride=2, groups=DIM)


def decode(codes: Tensor, W: Dict[str, Tensor]) -> Tensor:
    """(B,K,F) int → (B,1,1920·F) fp32."""
    z = decode_latent(codes, W)
    z = transformer(z.transpose(1, 2), W, "decoder_transformer").transpose(1, 2)
    return seanet_decoder(z, W)


# ----------------------------------------------------------------------------- streaming decode
class _SConv:
    """Streaming causal conv, stride 1: carries the last k-1 inputs (zeros before the first call)."""

    def __init__(self, w: Tensor, b: Optional[Tensor]):
        self.w, self.b, self.prev = w, b, None

    def __call__(self, x: Tensor) -> Tensor:
        k = self.w.shape[-1]
        if self.prev is None:
            self.prev = x.new_zeros(x.shape[0], x.shape[1], k - 1)
        xin = torch.cat([self.prev, x], dim=-1)
        self.prev = xin[..., xin.shape[-1] - (k - 1):]
        return F.conv1d(xin, self.w, self.b)


class _SConvTr:
    """Streaming transposed conv: overlap-add with a carried tail of k-stride partial sums."""

    def __init__(self, w: Tensor, b: Optional[Tensor], stride: int, groups: int = 1):
        self.w, self.b, self.stride, self.groups, self.tail = w, b, stride, groups, None

    def __call__(self, x: Tensor) -> Tensor:
        k = self.w.shape[-1]
        y = F.conv_transpose1d(x, self.w, None, stride=self.stride, groups=self.groups)
        if self.tail is not None:
            y[..., : k - self.stride] += self.tail
        n = x.shape[-1] * self.stride
        self.tail = y[..., n:].clone()
        out = y[..., :n]
        return out if self.b is None else out + self.b[None, :, None]


class StreamingDecoder:
    """``reset_state`` + ``decode_step`` (generation.py:224-225, 249-256): one (B,K,1) frame → (B,1,1920)."""

    def __init__(self, W: Dict[str, Tensor]):
        self.W = W
        self.reset_state()

    def reset_state(self) -> None:
        W = self.W
        self.up = _SConvTr(W["upsample.convtr.convtr.convtr.weight"], None, 2, groups=DIM)
        self.tstate = TransformerState()
        self.first = _SConv(W["decoder.model.0.conv.conv.weight"], W["decoder.model.0.conv.conv.bias"])
        self.stages = []
        idx = 1
        for r in RATIOS:
            p = f"decoder.model.{idx + 2}"
            self.stages.append((
                _SConvTr(W[f"decoder.model.{idx + 1}.convtr.convtr.weight"],
                         W[f"decoder.model.{idx + 1}.convtr.convtr.bias"], r),
                _SConv(W[p + ".block.1.conv.conv.weight"], W[p + ".block.1.conv.conv.bias"]),
                _SConv(W[p + ".block.3.conv.conv.weight"], W[p + ".block.3.conv.conv.bias"]),
            ))
            idx += 3
        self.last = _SConv(W[f"decoder.model.{idx + 1}.conv.conv.weight"], W[f"decoder.model.{idx + 1}.conv.conv.bias"])

    def decode_step(self, codes: Tensor) -> Tensor:
        z = self.up(rvq_decode(codes, self.W))
        z = transformer(z.transpose(1, 2), self.W, "decoder_transformer", self.tstate).transpose(1, 2)
        h = self.first(z)
        for tr, c1, c2 in self.stages:
            h = tr(F.elu(h))
            h = h + c2(F.elu(c1(F.elu(h))))
        return self.last(F.elu(h))

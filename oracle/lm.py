"""LM half of the oracle: CSM backbone + depth decoder, frame loop, generate driver.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Plain PyTorch CPU, fp32 arithmetic, written
for readability; every function cites the reference lines it restates (paths relative to
/root/reference).  Weights come in as a flat ``dict[str, Tensor]`` keyed exactly like the
reference parameter tree, e.g. ``backbone.layers.3.mlp.gate_proj.weight``.
"""

from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------- config
@dataclass(frozen=True)
class LlamaCfg:
    """csm_mlx/config.py:3-45 — only the fields the hot path reads."""

    n_layers: int
    d_model: int
    n_heads: int
    n_kv_heads: int
    head_dim: int
    d_ff: int
    eps: float = 1e-5
    rope_theta: float = 500_000.0
    rope_factor: float = 32.0  # rope_scaling["factor"]; the other keys are ignored (attention.py:201-205)


@dataclass(frozen=True)
class CSMCfg:
    """csm_mlx/models.py:12-28."""

    backbone: LlamaCfg
    decoder: LlamaCfg
    n_text_vocab: int = 128_256
    n_audio_vocab: int = 2051
    n_audio_codebooks: int = 32
    max_seq_len: int = 2048  # RoPE table length, attention.py:38


BACKBONE_1B = LlamaCfg(16, 2048, 32, 8, 64, 8192)
DECODER_100M = LlamaCfg(4, 1024, 8, 2, 128, 8192)
CSM_1B = CSMCfg(BACKBONE_1B, DECODER_100M)

# A small configuration with the same structure, for CPU-speed tests.
TINY = CSMCfg(
    LlamaCfg(2, 128, 4, 2, 32, 256),
    LlamaCfg(2, 64, 2, 1, 32, 128),
    n_text_vocab=512,
    n_audio_vocab=67,
    n_audio_codebooks=4,
    max_seq_len=256,
)


# ----------------------------------------------------------------------------- RoPE
def rope_scaled_freqs(head_dim: int, base: float, scale_factor: float,
                      low_freq_factor: int = 1, high_freq_factor: int = 4,
                      old_context_len: int = 8192) -> Tensor:
    """attention.py:57-69 (base frequencies) + :94-117 (Llama-3 scaling), all fp32."""
    freqs = 1.0 / (base ** (torch.arange(0, head_dim, 2)[: head_dim // 2].to(torch.float32) / head_dim))
    low_freq_wavelen = old_context_len / low_freq_factor
    high_freq_wavelen = old_context_len / high_freq_factor
    out = []
    for freq in freqs:  # 0-dim fp32 tensors, like the mx scalars in the reference loop (:105)
        wavelen = 2 * math.pi / freq
        if wavelen < high_freq_wavelen:
            out.append(freq)
        elif wavelen > low_freq_wavelen:
            out.append(freq / scale_factor)
        else:
            smooth = (old_context_len / wavelen - low_freq_factor) / (high_freq_factor - low_freq_factor)
            out.append((1 - smooth) * freq / scale_factor + smooth * freq)
    return torch.stack(out).to(torch.float32)


def rope_table(head_dim: int, base: float, scale_factor: float, max_seq_len: int = 2048) -> Tensor:
    """attention.py:81-92 — (max_seq_len, head_dim/2, 2) fp32 table of (cos, sin)."""
    theta = rope_scaled_freqs(head_dim, base, scale_factor)
    seq_idx = torch.arange(max_seq_len, dtype=torch.float32)
    idx_theta = torch.einsum("i,j->ij", seq_idx, theta).to(torch.float32)
    return torch.stack([torch.cos(idx_theta), torch.sin(idx_theta)], dim=-1)


def apply_rope(x: Tensor, table: Tensor, offset: int) -> Tensor:
    """attention.py:119-177 — x (B,T,H,hd); rotate ADJACENT pairs (x[2i], x[2i+1]) in fp32."""
    T = x.shape[1]
    rc = table[offset: offset + T]  # (T, hd/2, 2)
    if rc.shape[0] != T:
        raise ValueError("RoPE table exhausted")  # the reference would mis-broadcast here
    xs = x.to(torch.float32).reshape(*x.shape[:-1], -1, 2)
    rc = rc.reshape(1, T, 1, xs.shape[3], 2)
    out = torch.stack(
        [xs[..., 0] * rc[..., 0] - xs[..., 1] * rc[..., 1],
         xs[..., 1] * rc[..., 0] + xs[..., 0] * rc[..., 1]], dim=-1)
    return out.flatten(3).to(x.dtype)


# ----------------------------------------------------------------------------- Llama block
class KVCache:
    """mlx_lm.models.cache.KVCache restated: growable (B,Hkv,S,hd) K/V, ``offset``."""

    def __init__(self) -> None:
        self.keys: Optional[Tensor] = None
        self.values: Optional[Tensor] = None
        self.offset = 0

    def update_and_fetch(self, k: Tensor, v: Tensor):
        if self.keys is None:
            self.keys, self.values = k, v
        else:
            self.keys = torch.cat([self.keys, k], dim=2)
            self.values = torch.cat([self.values, v], dim=2)
        self.offset = self.keys.shape[2]
        return self.keys, self.values


def rms_norm(x: Tensor, w: Tensor, eps: float) -> Tensor:
    """mlx nn.RMSNorm (mx.fast.rms_norm): statistics in fp32."""
    x32 = x.to(torch.float32)
    return (x32 * torch.rsqrt(x32.pow(2).mean(-1, keepdim=True) + eps)) * w.to(torch.float32)


def attention(x: Tensor, W: Dict[str, Tensor], prefix: str, cfg: LlamaCfg, table: Tensor,
              cache: KVCache, trace: Optional[dict] = None) -> Tensor:
    """attention.py:207-253."""
    B, T, _ = x.shape
    q = F.linear(x, W[prefix + "q_proj.weight"]).reshape(B, T, cfg.n_heads, cfg.head_dim)
    k = F.linear(x, W[prefix + "k_proj.weight"]).reshape(B, T, cfg.n_kv_heads, cfg.head_dim)
    v = F.linear(x, W[prefix + "v_proj.weight"]).reshape(B, T, cfg.n_kv_heads, cfg.head_dim)
    offset = cache.offset  # read BEFORE the append (:227-228)
    q = apply_rope(q, table, offset).transpose(1, 2)
    k = apply_rope(k, table, offset).transpose(1, 2)
    v = v.transpose(1, 2)
    k, v = cache.update_and_fetch(k, v)
    rep = cfg.n_heads // cfg.n_kv_heads
    k = k.repeat_interleave(rep, dim=1)  # mx.repeat(axis=1): q-head j ↔ kv-head j // rep (:242-245)
    v = v.repeat_interleave(rep, dim=1)
    S = k.shape[2]
    scores = torch.matmul(q, k.transpose(2, 3)) * (cfg.head_dim ** -0.5)
    if T > 1:  # mlx_lm creates the causal mask iff T > 1
        qpos = torch.arange(offset, offset + T).unsqueeze(1)
        kpos = torch.arange(S).unsqueeze(0)
        scores = scores.masked_fill(kpos > qpos, float("-inf"))
    p = torch.softmax(scores.to(torch.float32), dim=-1)
    o = torch.matmul(p, v).transpose(1, 2).reshape(B, T, cfg.n_heads * cfg.head_dim)
    return F.linear(o, W[prefix + "o_proj.weight"])


def llama_forward(x: Tensor, W: Dict[str, Tensor], name: str, cfg: LlamaCfg, table: Tensor,
                  caches: Sequence[KVCache], trace: Optional[List[Tensor]] = None) -> Tensor:
    """mlx_lm LlamaModel with Identity embeddings (models.py:50-51, 70-77): pre-norm blocks."""
    h = x
    for l in range(cfg.n_layers):
        p = f"{name}.layers.{l}."
        n = rms_norm(h, W[p + "input_layernorm.weight"], cfg.eps)
        h = h + attention(n, W, p + "self_attn.", cfg, table, caches[l])
        n = rms_norm(h, W[p + "post_attention_layernorm.weight"], cfg.eps)
        g = F.linear(n, W[p + "mlp.gate_proj.weight"])
        u = F.linear(n, W[p + "mlp.up_proj.weight"])
        h = h + F.linear(F.silu(g) * u, W[p + "mlp.down_proj.weight"])
        if trace is not None:
            trace.append(h.clone())
    return rms_norm(h, W[f"{name}.norm.weight"], cfg.eps)


# ----------------------------------------------------------------------------- CSM
class OracleCSM:
    """Holds fp32 weights + RoPE tables; mirrors the pieces of models.py:31-92 the path uses."""

    def __init__(self, cfg: CSMCfg, weights: Dict[str, Tensor]):
        self.cfg = cfg
        self.W = {k: v.to(torch.float32) for k, v in weights.items()}
        b, d = cfg.backbone, cfg.decoder
        self.rope_b = rope_table(b.head_dim, b.rope_theta, b.rope_factor, cfg.max_seq_len)
        self.rope_d = rope_table(d.head_dim, d.rope_theta, d.rope_factor, cfg.max_seq_len)

    # models.py:79-80
    def embed_audio(self, codebook: int, tokens: Tensor) -> Tensor:
        return F.embedding(tokens + codebook * self.cfg.n_audio_vocab, self.W["audio_embeddings.weight"])

    # models.py:82-92 + generation.py:32-36
    def embed_frames(self, tokens: Tensor, mask: Tensor) -> Tensor:
        ncb = self.cfg.n_audio_codebooks
        text = F.embedding(tokens[:, :, -1], self.W["text_embeddings.weight"]).unsqueeze(-2)
        atok = tokens[:, :, :-1] + self.cfg.n_audio_vocab * torch.arange(ncb)
        audio = F.embedding(atok, self.W["audio_embeddings.weight"])
        e = torch.cat([audio, text], dim=-2)  # (B,T,33,D)
        return (e * mask.unsqueeze(-1).to(e.dtype)).sum(-2)

    def new_backbone_cache(self) -> List[KVCache]:
        return [KVCache() for _ in range(self.cfg.backbone.n_layers)]


def generate_frame(model: OracleCSM, tokens: Tensor, token_mask: Tensor, cache: List[KVCache], *,
                   sampler: Optional[Callable[[Tensor, int], Tensor]] = None,
                   logits_processors: Optional[List[Callable[[Tensor, Tensor], Tensor]]] = None,
                   c0_history: Optional[List[Tensor]] = None,
                   forced: Optional[Tensor] = None,
                   trace: Optional[dict] = None) -> Tensor:
    """generation.py:21-92.  tokens/mask (B,T,33) → (B,32) int64.

    ``sampler(logits (B,V), codebook_index) -> (B,)``; None = greedy argmax (temperature 0).
    ``forced`` (B,32): teacher forcing — the returned/propagated samples are ``forced`` while the
    logits recorded in ``trace`` are the model's own (used for logits-tolerance parity tests).
    ``trace`` collects: 'backbone_layers' (list of (B,T,D)), 'h' (B,D), 'logits' (list of 32 (B,V)).
    """
    cfg, W = model.cfg, model.W
    B = tokens.shape[0]
    x = model.embed_frames(tokens, token_mask)
    layer_trace = [] if trace is not None else None
    h = llama_forward(x, W, "backbone", cfg.backbone, model.rope_b, cache, layer_trace)
    last = h[:, -1, :]
    c0_logits = F.linear(last, W["codebook0_head.weight"])
    if logits_processors:
        for proc in logits_processors:
            hist = torch.stack(c0_history, 0) if c0_history else torch.zeros((0,))
            c0_logits = proc(hist, c0_logits)
    logits_all = [c0_logits]

    def pick(logits: Tensor, i: int) -> Tensor:
        if forced is not None:
            return forced[:, i].to(torch.int64)
        if sampler is None:
            return torch.argmax(logits, dim=-1)
        return sampler(logits, i).to(torch.int64)

    c0 = pick(c0_logits, 0).unsqueeze(-1)  # (B,1)
    if c0_history is not None:
        c0_history.append(c0)
    dec_in = torch.cat([last.unsqueeze(1), model.embed_audio(0, c0)], dim=1)  # (B,2,D)
    frame = torch.zeros((B, cfg.n_audio_codebooks), dtype=torch.int64)
    frame[:, :1] = c0
    dcache = [KVCache() for _ in range(cfg.decoder.n_layers)]  # fresh per frame (:70)
    for i in range(1, cfg.n_audio_codebooks):
        dh = llama_forward(F.linear(dec_in, W["projection.weight"]), W, "decoder", cfg.decoder,
                           model.rope_d, dcache)
        ci_logits = torch.matmul(dh[:, -1, :], W["audio_head"][i - 1])  # (in,out) layout (:79)
        logits_all.append(ci_logits)
        ci = pick(ci_logits, i).unsqueeze(-1)
        dec_in = model.embed_audio(i, ci)
        frame[:, i: i + 1] = ci
    if trace is not None:
        trace["backbone_layers"] = layer_trace
        trace["h"] = last
        trace["logits"] = logits_all
    return frame


def next_frame_input(sample: Tensor):
    """generation.py:156-161 — (B,32) → tokens (B,1,33) with text slot 0, mask [1×32, 0]."""
    B = sample.shape[0]
    tok = torch.cat([sample, torch.zeros((B, 1), dtype=sample.dtype)], dim=1).unsqueeze(1)
    mask = torch.cat([torch.ones_like(sample), torch.zeros((B, 1), dtype=sample.dtype)], dim=1).unsqueeze(1)
    return tok, mask.bool()


def generate_tokens(model: OracleCSM, prompt_tokens: Tensor, prompt_mask: Tensor, max_audio_frames: int, *,
                    sampler=None, logits_processors=None, traces: Optional[list] = None,
                    forced: Optional[Tensor] = None) -> Tensor:
    """generation.py:120-161 for one utterance: prompt (T,33) → (F,32) int64 (stops at all-zero frame)."""
    cfg = model.cfg
    inp, mask = prompt_tokens.unsqueeze(0).to(torch.int64), prompt_mask.unsqueeze(0)
    if inp.shape[1] >= cfg.max_seq_len - max_audio_frames:  # :132-137
        raise ValueError(
            f"Inputs too long ({inp.shape[1]}), must be below max_seq_len - max_audio_frames: "
            f"{cfg.max_seq_len - max_audio_frames}")
    cache = model.new_backbone_cache()
    c0_history: List[Tensor] = []
    samples = []
    for f in range(max_audio_frames):
        tr = {} if traces is not None else None
        s = generate_frame(model, inp, mask, cache, sampler=sampler, logits_processors=logits_processors,
                           c0_history=c0_history, trace=tr,
                           forced=None if forced is None else forced[f: f + 1])
        if traces is not None:
            traces.append(tr)
        if not bool(s.any()):
            break
        samples.append(s[0])
        inp, mask = next_frame_input(s)
    if not samples:
        return torch.zeros((0, cfg.n_audio_codebooks), dtype=torch.int64)
    return torch.stack(samples)


# ----------------------------------------------------------------------------- frame assembly
def text_rows(ids: Sequence[int], n_codebooks: int = 32):
    """tokenizers.py:43-58 given already-tokenised ids (BOS…EOS included)."""
    n = len(ids)
    tok = torch.zeros((n, n_codebooks + 1), dtype=torch.int64)
    mask = torch.zeros((n, n_codebooks + 1), dtype=torch.bool)
    tok[:, -1] = torch.tensor(list(ids), dtype=torch.int64)
    mask[:, -1] = True
    return tok, mask


def audio_rows(codes: Tensor):
    """tokenizers.py:61-85 — codes (K,F) → (F+1, K+1) rows with an all-zero EOS frame appended."""
    K, Fr = codes.shape
    codes = torch.cat([codes.to(torch.int64), torch.zeros((K, 1), dtype=torch.int64)], dim=1)
    tok = torch.zeros((Fr + 1, K + 1), dtype=torch.int64)
    mask = torch.zeros((Fr + 1, K + 1), dtype=torch.bool)
    tok[:, :-1] = codes.t()
    mask[:, :-1] = True
    return tok, mask

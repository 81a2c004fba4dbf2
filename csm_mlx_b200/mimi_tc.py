"""Batch-scale Mimi encode / decode on the tensor cores: host orchestration of csrc/mimi_tc.cu.

What ``moshi_mlx.models.mimi.Mimi.encode`` / ``.decode`` do for the reference (``/root/reference/csm_mlx/tokenizers.py:61-85``
encode of context audio, ``:148-150`` decode; BASELINE.json configs[4]: 128 clips x 60 s).  Every SEANet convolution, every
transformer Linear and the RVQ nearest-neighbour search is one call of the persistent tcgen05 GEMM ``csmb_gemm_tc3``;
activations travel between layers as two bf16 planes (x = hi + lo to 2^-17) written by the producing kernel's epilogue with
the consuming layer's ELU already applied, plus an fp32 copy only where a residual connection needs it.  The small
latent-rate pieces (RVQ gather, x2 depthwise upsampler, windowed attention) stay on the fp32 kernels of csrc/mimi.cu.

Clips are processed in sub-batches sized to a memory budget (``CSMB_MIMI_TC_BYTES``, default 12 GiB of activations): a 60 s
clip holds ~2.3 GB of intermediate activations at 24 kHz.  Streaming decode (one frame per call) keeps the weight-streaming
fp32 path of ``mimi.MimiDecodeStream``.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib

FRAME = 1920
DIM, FF, LAYERS, HEADS, CONTEXT = 512, 2048, 8, 8, 250
CB_DIM, BINS = 256, 2048
LN_EPS = 1e-5
BYTES_PER_AUDIO_SECOND = 48e6   # activations alive per second of audio in one encode or decode pass (planes + fp32 copies)


def split_planes_host(w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """fp32 [N][K] -> bf16 hi / lo planes as int16 [N16][K] (N rounded up to 16 with zero rows): w = hi + lo to 2^-17."""
    n, k = w.shape
    n16 = -(-n // 16) * 16
    hi = w.to(torch.bfloat16)
    lo = (w - hi.to(torch.float32)).to(torch.bfloat16)
    out = []
    for p in (hi, lo):
        buf = torch.zeros((n16, k), device=w.device, dtype=torch.bfloat16)
        buf[:n] = p
        out.append(buf.view(torch.int16).contiguous())
    return out[0], out[1]


class _W:
    """Weight of one tensor-core op: planes [N16][K], bias (fp32 or None), taps and C (K = taps * C)."""
    __slots__ = ("hi", "lo", "n", "k", "taps", "c", "bias")

    def __init__(self, w2d: torch.Tensor, taps: int, bias: Optional[torch.Tensor] = None):
        self.hi, self.lo = split_planes_host(w2d)
        self.n, self.k = int(w2d.shape[0]), int(w2d.shape[1])
        self.taps, self.c, self.bias = taps, self.k // taps, bias


class _Act:
    """[B][rows][C] activation: bf16 hi/lo planes (int16 storage) and / or an fp32 copy; ``pad`` leading context rows."""

    def __init__(self, B: int, rows: int, C: int, device, planes: bool = True, raw: bool = False, pad: int = 0,
                 valid: Optional[int] = None):
        self.B, self.rows, self.C, self.pad = B, rows, C, pad
        self.hi = torch.empty((B, rows, C), device=device, dtype=torch.int16) if planes else None
        self.lo = torch.empty((B, rows, C), device=device, dtype=torch.int16) if planes else None
        self.raw = torch.empty((B, rows, C), device=device, dtype=torch.float32) if raw else None
        valid = rows - pad if valid is None else valid
        for t in (self.hi, self.lo):                       # context rows and right padding are zeros (ELU(0) = 0)
            if t is not None:
                if pad:
                    t[:, :pad].zero_()
                if pad + valid < rows:
                    t[:, pad + valid:].zero_()

    def off(self, rows: int, itemsize: int) -> int:
        return rows * self.C * itemsize


class MimiTC:
    def __init__(self, mimi):
        self.m = mimi
        dev = mimi.device
        self.err = torch.zeros((1,), device=dev, dtype=torch.int32)
        m = mimi
        cw = lambda c, taps=None: _W(c.w, c.k if taps is None else taps, c.b)
        # SEANet encoder: [conv0 (fp32 kernel)], then per ratio (res conv3, res conv1, strided conv), last conv
        self.enc_res: List[Tuple[_W, _W]] = []
        self.enc_down: List[Tuple[_W, int]] = []
        for item in m.enc[1:-1]:
            if isinstance(item, tuple):
                self.enc_res.append((cw(item[1]), cw(item[2])))
            else:
                self.enc_down.append((cw(item, taps=2), item.s))    # k = 2s, stride s: two taps over rows of s*Cin
        self.enc_last = cw(m.enc[-1])
        # SEANet decoder
        self.dec_first = cw(m.dec_first)
        self.dec_stages = [(_W(ct.w, 2, ct.b), ct.s, ct.cout, cw(c1), cw(c2)) for (ct, c1, c2) in m.dec_stages]
        self.dec_last = cw(m.dec_last)
        # transformers
        self.tr: Dict[str, list] = {}
        for side, layers in m.tr.items():
            self.tr[side] = [(_W(t.in_proj, 1), _W(t.out_proj, 1), _W(t.lin1, 1), _W(t.lin2, 1), t) for t in layers]
        self.down = cw(m.down, taps=2)
        self.in_proj = [_W(w, 1) for w in m.in_proj]
        self.codebooks = [_W(m.codebooks[k], 1) for k in range(m.n_q)]

    # ------------------------------------------------------------------ one tensor-core op
    def gemm(self, a: _Act, view: int, w: _W, B: int, T: int, *, y32_ptr=0, y_batch=0, ldy=0, planes: Optional[_Act] = None,
             p_row0=0, ldp=None, p_batch=None, plane_act=0, res_ptr=0, r_batch=0, ldr=0, scale=None, act_out=0):
        """One csmb_gemm_tc3 call.  A = ``a``'s planes viewed as rows of ``view`` * a.C values (``view`` = the stride of a
        strided conv, else 1); output row t of batch b reads view rows t .. t + taps - 1 of that batch.  B, T: batches and
        output rows per batch of THIS op (a dense [1][M][C] activation may be read as B batches of T rows)."""
        m = self.m
        g = _lib.Tc3()
        cv = a.C * view
        total_rows = a.B * a.rows
        assert w.c == cv and total_rows % view == 0 and (total_rows // view) % B == 0, (w.c, cv, a.rows, view, B)
        g.a_hi, g.a_lo = a.hi.data_ptr(), a.lo.data_ptr()
        g.a_rows = total_rows // view
        g.lda, g.rpb, g.C, g.taps = cv, (total_rows // view) // B, cv, w.taps
        g.w_hi, g.w_lo, g.w_rows, g.ldw = w.hi.data_ptr(), w.lo.data_ptr(), int(w.hi.shape[0]), w.k
        g.y32, g.y_batch, g.ldy = (y32_ptr or None), y_batch, ldy
        if planes is not None:
            g.y_hi = planes.hi.data_ptr() + p_row0 * planes.C * 2
            g.y_lo = planes.lo.data_ptr() + p_row0 * planes.C * 2
            g.p_batch = planes.rows * planes.C if p_batch is None else p_batch
            g.ldp, g.plane_act = (planes.C if ldp is None else ldp), plane_act
        g.bias = w.bias.data_ptr() if w.bias is not None else None
        g.scale = scale.data_ptr() if scale is not None else None
        g.residual, g.r_batch, g.ldr = (res_ptr or None), r_batch, ldr
        g.B, g.T, g.N, g.act_out = B, T, w.n, act_out
        g.err_flag = self.err.data_ptr()
        _lib.check(_lib.lib().csmb_gemm_tc3(C.byref(g), m.dev_idx, m._st()))

    def check(self) -> None:
        if int(self.err.item()) != 0:
            self.err.zero_()
            raise _lib.CsmbError("csmb_gemm_tc3: a bounded wait inside the tensor-core codec kernel timed out")

    # ------------------------------------------------------------------ transformer (8 layers, in place on x fp32 [B][T][512])
    def transformer(self, side: str, x: torch.Tensor, T: int, final_planes: Optional[_Act] = None, final_row0: int = 0) -> None:
        m, lib = self.m, _lib.lib()
        B = x.shape[0]
        dev = x.device
        xb = x.stride(0)
        h = _Act(B, T, DIM, dev)
        att_p = _Act(B, T, DIM, dev)
        ffp = _Act(B, T, FF, dev)
        qkv = torch.empty((B, T, 3 * DIM), device=dev, dtype=torch.float32)
        layers = self.tr[side]
        for li, (w_in, w_out, w1, w2, t) in enumerate(layers):
            _lib.check(lib.csmb_layernorm_planes(x.data_ptr(), xb, t.n1w.data_ptr(), t.n1b.data_ptr(), h.hi.data_ptr(),
                                                 h.lo.data_ptr(), B, T, DIM, LN_EPS, m.dev_idx, m._st()))
            self.gemm(h, 1, w_in, B, T, y32_ptr=qkv.data_ptr(), y_batch=T * 3 * DIM, ldy=3 * DIM)
            _lib.check(lib.csmb_mimi_attention_planes(qkv.data_ptr(), m.freqs.data_ptr(), att_p.hi.data_ptr(), att_p.lo.data_ptr(),
                                                      B, T, HEADS, CONTEXT, m.dev_idx, m._st()))
            self.gemm(att_p, 1, w_out, B, T, y32_ptr=x.data_ptr(), y_batch=xb, ldy=DIM, res_ptr=x.data_ptr(), r_batch=xb,
                      ldr=DIM, scale=t.ls1)
            _lib.check(lib.csmb_layernorm_planes(x.data_ptr(), xb, t.n2w.data_ptr(), t.n2b.data_ptr(), h.hi.data_ptr(),
                                                 h.lo.data_ptr(), B, T, DIM, LN_EPS, m.dev_idx, m._st()))
            self.gemm(h, 1, w1, B, T, planes=ffp, act_out=1)
            last = li == len(layers) - 1
            self.gemm(ffp, 1, w2, B, T, y32_ptr=x.data_ptr(), y_batch=xb, ldy=DIM, res_ptr=x.data_ptr(), r_batch=xb,
                      ldr=DIM, scale=t.ls2, planes=final_planes if last else None, p_row0=final_row0)

    # ------------------------------------------------------------------ decode
    def clips_per_pass(self, seconds_per_clip: float) -> int:
        budget = float(os.environ.get("CSMB_MIMI_TC_BYTES", 12 * (1 << 30)))
        return max(1, int(budget / (BYTES_PER_AUDIO_SECOND * max(seconds_per_clip, 0.08))))

    def decode(self, codes: torch.Tensor) -> torch.Tensor:
        """(B,K,F) int -> (B,1,1920*F) fp32."""
        B, K, F = codes.shape
        out = torch.empty((B, 1, FRAME * F), device=self.m.device, dtype=torch.float32)
        step = self.clips_per_pass(F * 0.08)
        for b0 in range(0, B, step):
            out[b0:b0 + step] = self._decode_pass(codes[b0:b0 + step].contiguous())
        self.check()
        return out

    def _decode_pass(self, codes: torch.Tensor) -> torch.Tensor:
        m, lib = self.m, _lib.lib()
        dev = m.device
        B, K, F = codes.shape
        T = 2 * F
        st = m._st()
        e = lambda *s: torch.empty(s, device=dev, dtype=torch.float32)
        # RVQ dequantise + output projections + x2 depthwise upsampler: latent-rate, fp32 kernels
        codes = codes.to(device=dev, dtype=torch.int32).contiguous()
        sem, ac, lat = e(B, F, CB_DIM), e(B, F, CB_DIM), e(B, F, DIM)
        _lib.check(lib.csmb_rvq_gather(codes.data_ptr(), m.codebooks.data_ptr(), sem.data_ptr(), ac.data_ptr(), B, m.n_q, F, BINS,
                                       CB_DIM, m.dev_idx, st))
        m._gemm(sem.data_ptr(), F * CB_DIM, CB_DIM, m.out_proj[0], lat.data_ptr(), F * DIM, DIM, B, F, DIM, CB_DIM)
        m._gemm(ac.data_ptr(), F * CB_DIM, CB_DIM, m.out_proj[1], lat.data_ptr(), F * DIM, DIM, B, F, DIM, CB_DIM,
                res=lat.data_ptr(), r_batch=F * DIM, ldr=DIM)
        x = e(B, T, DIM)
        up_prev = torch.zeros((B, DIM), device=dev, dtype=torch.float32)
        _lib.check(lib.csmb_upsample_dw(lat.data_ptr(), up_prev.data_ptr(), m.up_w.data_ptr(), x.data_ptr(), B, F, DIM, m.dev_idx, st))
        # transformer; its last Linear also writes the planes of the first SEANet conv's input (k7: 6 context rows)
        w0 = self.dec_first
        a0 = _Act(B, (w0.taps - 1) + T, DIM, dev, pad=w0.taps - 1)
        self.transformer("decoder_transformer", x, T, final_planes=a0, final_row0=w0.taps - 1)
        # SEANet decoder
        cur = _Act(B, 1 + T, w0.n, dev, pad=1)                         # input of the first transposed conv (row 0 = x[-1] = 0)
        self.gemm(a0, 1, w0, B, T, planes=cur, p_row0=1, plane_act=1)
        t = T
        for i, (wt, s, cout, w1, w2) in enumerate(self.dec_stages):
            t2 = t * s
            res_in = _Act(B, 2 + t2, cout, dev, raw=True, pad=2)      # ResBlock input: fp32 for the skip + ELU planes (k3: 2 context rows)
            self.gemm(cur, 1, wt, B, t, y32_ptr=res_in.raw.data_ptr() + res_in.off(2, 4), y_batch=res_in.rows * cout,
                      ldy=s * cout, planes=res_in, p_row0=2, ldp=s * cout, plane_act=1)
            hbuf = _Act(B, t2, w1.n, dev)
            self.gemm(res_in, 1, w1, B, t2, planes=hbuf, plane_act=1)
            last = i == len(self.dec_stages) - 1
            npad = 2 if last else 1
            nxt = _Act(B, npad + t2, cout, dev, pad=npad)
            self.gemm(hbuf, 1, w2, B, t2, planes=nxt, p_row0=npad, plane_act=1,
                      res_ptr=res_in.raw.data_ptr() + res_in.off(2, 4), r_batch=res_in.rows * cout, ldr=cout)
            cur, t = nxt, t2
            del res_in, hbuf
        audio = e(B, t, 1)
        self.gemm(cur, 1, self.dec_last, B, t, y32_ptr=audio.data_ptr(), y_batch=t, ldy=1)
        return audio.reshape(B, 1, t)

    # ------------------------------------------------------------------ encode
    def encode(self, audio: torch.Tensor) -> torch.Tensor:
        """(B,1,N) fp32 -> (B,n_q,ceil(N/1920)) int32."""
        B, _, N = audio.shape
        F = -(-N // FRAME)
        out = torch.empty((B, self.m.n_q, F), device=self.m.device, dtype=torch.int32)
        step = self.clips_per_pass(N / 24000.0)
        for b0 in range(0, B, step):
            out[b0:b0 + step] = self._encode_pass(audio[b0:b0 + step])
        self.check()
        return out

    def _encode_pass(self, audio: torch.Tensor) -> torch.Tensor:
        m, lib = self.m, _lib.lib()
        dev = m.device
        st = m._st()
        x = audio.to(device=dev, dtype=torch.float32)
        B, _, N = x.shape
        c0 = m.enc[0]
        xin = torch.zeros((B, (c0.k - 1) + N), device=dev, dtype=torch.float32)
        xin[:, c0.k - 1:] = x[:, 0, :]
        t = N
        cur = _Act(B, 2 + t, c0.cout, dev, raw=True, pad=2)          # ResBlock input
        if cur.raw is not None:
            cur.raw[:, :2].zero_()
        _lib.check(lib.csmb_conv_in_planes(xin.data_ptr(), xin.stride(0), c0.w.data_ptr(), c0.b.data_ptr(),
                                           cur.raw.data_ptr() + cur.off(2, 4), cur.hi.data_ptr() + cur.off(2, 2),
                                           cur.lo.data_ptr() + cur.off(2, 2), cur.rows * cur.C, B, N, c0.cout, c0.k, m.dev_idx, st))
        del xin
        for i, ((w1, w2), (wd, s)) in enumerate(zip(self.enc_res, self.enc_down)):
            c = cur.C
            hbuf = _Act(B, t, w1.n, dev)
            self.gemm(cur, 1, w1, B, t, planes=hbuf, plane_act=1)
            # ResBlock output -> ELU -> strided conv (k = 2s): s context rows, rows rounded up to whole windows
            t_out = -(-t // s)
            dn = _Act(B, (t_out + 1) * s, c, dev, pad=s, valid=t)
            self.gemm(hbuf, 1, w2, B, t, planes=dn, p_row0=s, plane_act=1,
                      res_ptr=cur.raw.data_ptr() + cur.off(2, 4), r_batch=cur.rows * c, ldr=c)
            del hbuf
            last = i == len(self.enc_res) - 1
            nxt = _Act(B, 2 + t_out, wd.n, dev, raw=not last, pad=2)   # next ResBlock input, or the last conv's (k3) input
            self.gemm(dn, s, wd, B, t_out, planes=nxt, p_row0=2, plane_act=1,
                      y32_ptr=0 if last else nxt.raw.data_ptr() + nxt.off(2, 4), y_batch=nxt.rows * wd.n, ldy=wd.n)
            cur, t = nxt, t_out
            del dn
        T = t
        lat = torch.empty((B, T, DIM), device=dev, dtype=torch.float32)
        self.gemm(cur, 1, self.enc_last, B, T, y32_ptr=lat.data_ptr(), y_batch=T * DIM, ldy=DIM)
        del cur
        # transformer; its last Linear also writes the (un-activated) planes of the downsampling conv's input
        F = -(-T // 2)
        dbuf = _Act(B, (F + 1) * 2, DIM, dev, pad=2, valid=T)
        self.transformer("encoder_transformer", lat, T, final_planes=dbuf, final_row0=2)
        for p in (dbuf.hi, dbuf.lo):                                  # replicate padding (moshi's 12.5 Hz resampler)
            p[:, :2] = p[:, 2:3]
            if 2 + T < dbuf.rows:
                p[:, 2 + T:] = p[:, 1 + T:2 + T]
        latq = _Act(B, F, DIM, dev)
        self.gemm(dbuf, 2, self.down, B, F, planes=latq)
        # RVQ encode: semantic (1) and acoustic (n_q - 1) chains both start from the latent
        codes = torch.empty((B, m.n_q, F), device=dev, dtype=torch.int32)
        M = B * F
        dots = torch.empty((M, BINS), device=dev, dtype=torch.float32)
        for g, (k0, n) in enumerate(((0, 1), (1, m.n_q - 1))):
            r = torch.empty((M, CB_DIM), device=dev, dtype=torch.float32)
            rp = _Act(1, M, CB_DIM, dev)
            self.gemm(latq, 1, self.in_proj[g], B, F, y32_ptr=r.data_ptr(), y_batch=F * CB_DIM, ldy=CB_DIM, planes=rp,
                      p_batch=F * CB_DIM)
            for i in range(n):
                k = k0 + i
                self.gemm(rp, 1, self.codebooks[k], 1, M, y32_ptr=dots.data_ptr(), y_batch=0, ldy=BINS)
                _lib.check(lib.csmb_rvq_argmin_update_planes(dots.data_ptr(), m.cb_norm2[k].data_ptr(), m.codebooks[k].data_ptr(),
                                                             r.data_ptr(), rp.hi.data_ptr(), rp.lo.data_ptr(), codes.data_ptr(), M,
                                                             BINS, CB_DIM, m.n_q, k, F, m.dev_idx, st))
        return codes

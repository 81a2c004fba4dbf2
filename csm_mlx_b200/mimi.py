"""Mimi codec ("mimi_202407", 32 codebooks) on the GPU: host-side orchestration of the codec kernels.

Stands in for ``moshi_mlx.models.mimi.Mimi`` as the reference uses it
(``/root/reference/csm_mlx/tokenizers.py:14-21`` construction + ``load_pytorch_weights``, ``:70`` encode,
``:150`` decode; ``generation.py:224-225,251,258`` reset_state / decode_step).  Every arithmetic step is a call
into libcsm_b200.so (csrc/mimi.cu); torch is used for device buffers only.  Activations are time-major
``[batch][time][channels]`` fp32 so every conv / transposed conv / linear is the same strided-row GEMM.

Weights come in the moshi checkpoint key layout (weight-norm-free) and are re-laid once at load:
Conv1d ``[Cout][Cin][k] -> [Cout][k*Cin]``; ConvTranspose1d ``[Cin][Cout][2s] -> [s*Cout][2*Cin]`` (phase-major
rows, taps (r+s | r) for input rows (t-1 | t)); codebooks ``embedding_sum / max(cluster_usage, 1e-5)``.
"""

from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib

SAMPLE_RATE = 24_000
FRAME = 1920
RATIOS = (8, 6, 5, 4)
DIM, FILTERS, FF, LAYERS, HEADS = 512, 64, 2048, 8, 8
CONTEXT = 250
CB_DIM, BINS = 256, 2048
LN_EPS = 1e-5


class _Conv:
    __slots__ = ("w", "b", "k", "s", "cin", "cout")

    def __init__(self, w: torch.Tensor, b: Optional[torch.Tensor], stride: int = 1):
        cout, cin, k = w.shape
        self.w = w.permute(0, 2, 1).contiguous().reshape(cout, k * cin)
        self.b, self.k, self.s, self.cin, self.cout = b, k, stride, cin, cout


class _ConvTr:
    __slots__ = ("w", "b", "s", "cin", "cout")

    def __init__(self, w: torch.Tensor, b: Optional[torch.Tensor], stride: int):
        cin, cout, k = w.shape
        assert k == 2 * stride
        s = stride
        # rows (r, co), cols (half, ci): half 0 multiplies x[t-1] with tap r+s, half 1 multiplies x[t] with tap r
        wp = torch.stack([w[:, :, s:], w[:, :, :s]], dim=0)  # (half, ci, co, r)
        self.w = wp.permute(3, 2, 0, 1).contiguous().reshape(s * cout, 2 * cin)
        self.b = None if b is None else b.repeat(s).contiguous()
        self.s, self.cin, self.cout = s, cin, cout


class _TLayer:
    pass


def _g(lib_fn, *args):
    _lib.check(lib_fn(*args))


class Mimi:
    def __init__(self, n_q: int = 32, device=None):
        self.n_q = n_q
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._loaded = False
        self._default_stream: Optional["MimiDecodeStream"] = None

    # ------------------------------------------------------------------ weights
    def load_pytorch_weights(self, file_or_weights, strict: bool = True) -> "Mimi":
        """moshi checkpoint (``tokenizer-e351c8d8-checkpoint125.safetensors`` layout) or a dict with those keys."""
        if isinstance(file_or_weights, dict):
            W = file_or_weights
        else:
            from safetensors.torch import load_file

            W = load_file(str(file_or_weights))
        from .random_init import mimi_param_shapes

        expected = {n: tuple(s) for n, s, _ in mimi_param_shapes(self.n_q)}
        missing = [k for k in expected if k not in W]
        if missing:
            raise ValueError(f"Missing Mimi parameters: {missing[:6]}")
        if strict:
            bad = [k for k in expected if tuple(W[k].shape) != expected[k]]
            if bad:
                raise ValueError(f"Mimi parameter shape mismatch: {bad[:6]}")
        self.dev_idx = _lib.require_device(self.device)
        f = lambda k: W[k].to(device=self.device, dtype=torch.float32).contiguous()
        # SEANet encoder
        self.enc: List[object] = [_Conv(f("encoder.model.0.conv.conv.weight"), f("encoder.model.0.conv.conv.bias"))]
        idx = 1
        for r in reversed(RATIOS):
            self.enc.append(("res", _Conv(f(f"encoder.model.{idx}.block.1.conv.conv.weight"), f(f"encoder.model.{idx}.block.1.conv.conv.bias")),
                             _Conv(f(f"encoder.model.{idx}.block.3.conv.conv.weight"), f(f"encoder.model.{idx}.block.3.conv.conv.bias"))))
            self.enc.append(_Conv(f(f"encoder.model.{idx + 2}.conv.conv.weight"), f(f"encoder.model.{idx + 2}.conv.conv.bias"), stride=r))
            idx += 3
        self.enc.append(_Conv(f(f"encoder.model.{idx + 1}.conv.conv.weight"), f(f"encoder.model.{idx + 1}.conv.conv.bias")))
        # SEANet decoder
        self.dec_first = _Conv(f("decoder.model.0.conv.conv.weight"), f("decoder.model.0.conv.conv.bias"))
        self.dec_stages = []
        idx = 1
        for r in RATIOS:
            self.dec_stages.append((
                _ConvTr(f(f"decoder.model.{idx + 1}.convtr.convtr.weight"), f(f"decoder.model.{idx + 1}.convtr.convtr.bias"), r),
                _Conv(f(f"decoder.model.{idx + 2}.block.1.conv.conv.weight"), f(f"decoder.model.{idx + 2}.block.1.conv.conv.bias")),
                _Conv(f(f"decoder.model.{idx + 2}.block.3.conv.conv.weight"), f(f"decoder.model.{idx + 2}.block.3.conv.conv.bias"))))
            idx += 3
        self.dec_last = _Conv(f(f"decoder.model.{idx + 1}.conv.conv.weight"), f(f"decoder.model.{idx + 1}.conv.conv.bias"))
        # transformers
        self.tr: Dict[str, List[_TLayer]] = {}
        for side in ("encoder_transformer", "decoder_transformer"):
            layers = []
            for l in range(LAYERS):
                p = f"{side}.transformer.layers.{l}."
                t = _TLayer()
                t.in_proj, t.out_proj = f(p + "self_attn.in_proj_weight"), f(p + "self_attn.out_proj.weight")
                t.lin1, t.lin2 = f(p + "linear1.weight"), f(p + "linear2.weight")
                t.n1w, t.n1b, t.n2w, t.n2b = f(p + "norm1.weight"), f(p + "norm1.bias"), f(p + "norm2.weight"), f(p + "norm2.bias")
                t.ls1, t.ls2 = f(p + "layer_scale_1.scale"), f(p + "layer_scale_2.scale")
                layers.append(t)
            self.tr[side] = layers
        self.down = _Conv(f("downsample.conv.conv.conv.weight"), None, stride=2)
        self.up_w = f("upsample.convtr.convtr.convtr.weight").reshape(DIM, 4).contiguous()
        # quantiser
        cbs = []
        for group, n in (("rvq_first", 1), ("rvq_rest", self.n_q - 1)):
            for i in range(n):
                p = f"quantizer.{group}.vq.layers.{i}._codebook."
                cbs.append(f(p + "embedding_sum") / f(p + "cluster_usage").clamp(min=1e-5)[:, None])
        self.codebooks = torch.stack(cbs).contiguous()  # [n_q][bins][256]
        self.cb_norm2 = (self.codebooks * self.codebooks).sum(-1).contiguous()  # [n_q][bins]
        self.in_proj = [f("quantizer.rvq_first.input_proj.weight").reshape(CB_DIM, DIM).contiguous(),
                        f("quantizer.rvq_rest.input_proj.weight").reshape(CB_DIM, DIM).contiguous()]
        self.out_proj = [f("quantizer.rvq_first.output_proj.weight").reshape(DIM, CB_DIM).contiguous(),
                         f("quantizer.rvq_rest.output_proj.weight").reshape(DIM, CB_DIM).contiguous()]
        # RoPE frequencies: theta_i = exp(-ln(10000) * 2i / 64), fp32 (moshi rope, adjacent pairs)
        self.freqs = torch.exp(torch.arange(32, dtype=torch.float32) * (-math.log(10_000.0) * 2 / 64)).to(self.device)
        self._zero_pos = torch.zeros((1,), dtype=torch.int32, device=self.device)
        self._loaded = True
        return self

    # ------------------------------------------------------------------ kernel call helpers
    def _st(self) -> int:
        return _lib.stream_ptr(self.device)

    def _gemm(self, A, a_batch, lda, W, Y, y_batch, ldy, B, T, N, K, bias=None, scale=None, res=None, r_batch=0, ldr=0,
              act_in=0, act_out=0):
        """A, Y, res: integer device addresses (so views with offsets are cheap)."""
        _g(_lib.lib().csmb_gemm_f32, A, a_batch, lda, W.data_ptr(), Y, y_batch, ldy,
           None if bias is None else bias.data_ptr(), None if scale is None else scale.data_ptr(), res, r_batch, ldr,
           B, T, N, K, act_in, act_out, self.dev_idx, self._st())

    def _conv(self, c: _Conv, xbuf: torch.Tensor, t_in: int, ybuf: torch.Tensor, y_pad: int, act_in: int,
              res: Optional[torch.Tensor] = None, res_pad: int = 0) -> int:
        """xbuf [B][(k-s) + t_in + extra][cin] (left context first) -> ybuf[:, y_pad : y_pad + t_out, :cout]."""
        B = xbuf.shape[0]
        t_out = -(-t_in // c.s)
        need = (t_out - 1) * c.s + c.k
        assert xbuf.shape[1] >= need and xbuf.shape[2] == c.cin, (xbuf.shape, need, c.cin)
        self._gemm(xbuf.data_ptr(), xbuf.stride(0), c.s * c.cin, c.w, ybuf.data_ptr() + 4 * y_pad * c.cout,
                   ybuf.stride(0), c.cout, B, t_out, c.cout, c.k * c.cin, bias=c.b, act_in=act_in,
                   res=None if res is None else res.data_ptr() + 4 * res_pad * c.cout,
                   r_batch=0 if res is None else res.stride(0), ldr=c.cout)
        return t_out

    def _convtr(self, c: _ConvTr, xbuf: torch.Tensor, t_in: int, ybuf: torch.Tensor, y_pad: int, act_in: int) -> int:
        """xbuf [B][1 + t_in][cin] (row 0 = x[-1]) -> ybuf[:, y_pad : y_pad + t_in*s, :cout]."""
        B = xbuf.shape[0]
        self._gemm(xbuf.data_ptr(), xbuf.stride(0), c.cin, c.w, ybuf.data_ptr() + 4 * y_pad * c.cout, ybuf.stride(0),
                   c.s * c.cout, B, t_in, c.s * c.cout, 2 * c.cin, bias=c.b, act_in=act_in)
        return t_in * c.s

    def _transformer(self, side: str, xbuf: torch.Tensor, x_pad: int, T: int, cache: torch.Tensor, pos0: torch.Tensor,
                     scratch: Dict[str, torch.Tensor]) -> None:
        """In place on xbuf[:, x_pad:x_pad+T, :512].  cache [L][B][cap][2][H][64]."""
        B = xbuf.shape[0]
        xptr = xbuf.data_ptr() + 4 * x_pad * DIM
        xb = xbuf.stride(0)
        h, qkv, att, ff = scratch["h"], scratch["qkv"], scratch["att"], scratch["ff"]
        cap = cache.shape[2]
        lib = _lib.lib()
        for l, t in enumerate(self.tr[side]):
            _g(lib.csmb_layernorm, xptr, xb, t.n1w.data_ptr(), t.n1b.data_ptr(), h.data_ptr(), B, T, DIM, LN_EPS,
               self.dev_idx, self._st())
            self._gemm(h.data_ptr(), T * DIM, DIM, t.in_proj, qkv.data_ptr(), T * 3 * DIM, 3 * DIM, B, T, 3 * DIM, DIM)
            _g(lib.csmb_mimi_attention, qkv.data_ptr(), cache[l].data_ptr(), self.freqs.data_ptr(), pos0.data_ptr(),
               att.data_ptr(), B, T, HEADS, cap, CONTEXT, self.dev_idx, self._st())
            self._gemm(att.data_ptr(), T * DIM, DIM, t.out_proj, xptr, xb, DIM, B, T, DIM, DIM, scale=t.ls1, res=xptr,
                       r_batch=xb, ldr=DIM)
            _g(lib.csmb_layernorm, xptr, xb, t.n2w.data_ptr(), t.n2b.data_ptr(), h.data_ptr(), B, T, DIM, LN_EPS,
               self.dev_idx, self._st())
            self._gemm(h.data_ptr(), T * DIM, DIM, t.lin1, ff.data_ptr(), T * FF, FF, B, T, FF, DIM, act_out=1)
            self._gemm(ff.data_ptr(), T * FF, FF, t.lin2, xptr, xb, DIM, B, T, DIM, FF, scale=t.ls2, res=xptr, r_batch=xb,
                       ldr=DIM)

    def _tr_scratch(self, B: int, T: int) -> Dict[str, torch.Tensor]:
        e = lambda *s: torch.empty(s, device=self.device, dtype=torch.float32)
        return {"h": e(B, T, DIM), "qkv": e(B, T, 3 * DIM), "att": e(B, T, DIM), "ff": e(B, T, FF)}

    def _require(self):
        if not self._loaded:
            raise RuntimeError("Mimi weights are not loaded: call load_pytorch_weights(...) first")

    # ------------------------------------------------------------------ decode
    def tc(self):
        """The tensor-core (tcgen05) whole-clip path, built on first use (bf16 hi/lo weight planes; mimi_tc.MimiTC)."""
        self._require()
        if self.__dict__.get("_tc") is None:
            from .mimi_tc import MimiTC

            self._tc = MimiTC(self)
        return self._tc

    @staticmethod
    def _use_tc() -> bool:
        """Whole-clip encode / decode run on the tensor cores unless ``CSMB_MIMI_FP32=1`` (the fp32 CUDA-core kernels, kept
        as the streaming path and as an A/B reference)."""
        import os

        return os.environ.get("CSMB_MIMI_FP32", "0") != "1"

    def decode(self, codes: torch.Tensor) -> torch.Tensor:
        """(B,K,F) int -> (B,1,1920*F) fp32 (``Mimi.decode``; tokenizers.py:148-150)."""
        self._require()
        if self._use_tc() and int(codes.shape[2]) > 0:
            return self.tc().decode(codes)
        st = MimiDecodeStream(self, int(codes.shape[0]), max_frames=int(codes.shape[2]), offline=True)
        return st.step(codes)

    def new_decode_stream(self, batch: int = 1, frames_per_step: int = 1, use_graph: bool = True) -> "MimiDecodeStream":
        self._require()
        return MimiDecodeStream(self, batch, max_frames=frames_per_step, offline=False, use_graph=use_graph)

    def acquire_decode_stream(self, batch: int = 1) -> "MimiDecodeStream":
        """A start-of-stream decoder state from the pool (its captured CUDA graph is kept between utterances)."""
        pool = self.__dict__.setdefault("_stream_pool", [])
        for i, st in enumerate(pool):
            if st.m is self and st.B == batch and st.F == 1:
                pool.pop(i)
                st.reset()
                return st
        return self.new_decode_stream(batch)

    def release_decode_stream(self, st: "MimiDecodeStream") -> None:
        pool = self.__dict__.setdefault("_stream_pool", [])
        if len(pool) < 4 and all(x is not st for x in pool):
            pool.append(st)

    def reset_state(self) -> None:
        """generation.py:224-225,258."""
        self._default_stream = None

    def decode_step(self, codes: torch.Tensor) -> torch.Tensor:
        """(B,K,1) -> (B,1,1920) with carried streaming state (generation.py:249-256)."""
        self._require()
        if self._default_stream is None or self._default_stream.B != int(codes.shape[0]):
            self._default_stream = self.new_decode_stream(int(codes.shape[0]))
        return self._default_stream.step(codes).clone()

    # ------------------------------------------------------------------ encode
    def encode(self, audio: torch.Tensor) -> torch.Tensor:
        """(B,1,N) fp32 -> (B,n_q,ceil(N/1920)) int32 (``Mimi.encode``; tokenizers.py:70-72)."""
        self._require()
        if self._use_tc() and int(audio.shape[2]) > 0:
            return self.tc().encode(audio)
        dev = self.device
        x = audio.to(device=dev, dtype=torch.float32)
        B, _, N = x.shape
        z = lambda *s: torch.zeros(s, device=dev, dtype=torch.float32)
        # per-layer output lengths (each conv: ceil(L / stride))
        convs: List[Tuple[str, object]] = []
        for item in self.enc:
            convs.append(("res", item) if isinstance(item, tuple) else ("conv", item))
        # first conv input: [B][6 + N][1]
        c0: _Conv = self.enc[0]
        cur = z(B, (c0.k - c0.s) + N, 1)
        cur[:, c0.k - c0.s:, 0] = x[:, 0, :]
        t = N
        act = 0
        for i, (kind, item) in enumerate(convs):
            nxt_kind, nxt = convs[i + 1] if i + 1 < len(convs) else ("end", None)
            if kind == "conv":
                c: _Conv = item
                t_out = -(-t // c.s)
                if nxt_kind == "res":
                    npad, nextra = 2, 0  # res conv k3 s1
                elif nxt_kind == "conv":
                    nk, ns = nxt.k, nxt.s
                    npad = nk - ns
                    nextra = (-(-t_out // ns) - 1) * ns + nk - npad - t_out
                else:
                    npad, nextra = 0, 0
                out = z(B, npad + t_out + max(nextra, 0), c.cout)
                # right zero padding of `cur` was allocated by the producer; make sure it is long enough
                need = (t_out - 1) * c.s + c.k
                if cur.shape[1] < need:
                    cur = torch.cat([cur, z(B, need - cur.shape[1], cur.shape[2])], 1)
                self._conv(c, cur, t, out, npad, act_in=act)
                cur, t, act = out, t_out, 1  # every later conv is preceded by ELU
            else:
                _, c1, c2 = item
                # cur: [B][2 + t][C] (pad 2).  h = conv3(ELU(x)) -> [B][t][C/2]; y = x + conv1(ELU(h))
                hbuf = z(B, t, c1.cout)
                self._conv(c1, cur, t, hbuf, 0, act_in=1)
                nk, ns = nxt.k, nxt.s
                npad = nk - ns
                nextra = (-(-t // ns) - 1) * ns + nk - npad - t
                out = z(B, npad + t + max(nextra, 0), c2.cout)
                self._conv(c2, hbuf, t, out, npad, act_in=1, res=cur, res_pad=2)
                cur, act = out, 1
        # cur: [B][t][512] latent at 25 Hz
        T = t
        cache = z(LAYERS, B, CONTEXT + T, 2, HEADS, 64)
        self._transformer("encoder_transformer", cur, 0, T, cache, self._zero_pos, self._tr_scratch(B, T))
        # downsample conv k4 s2, replicate padding (left 2 rows = first row; right extra = last row)
        F = -(-T // 2)
        extra = (F - 1) * 2 + 4 - 2 - T
        dbuf = torch.empty((B, 2 + T + max(extra, 0), DIM), device=dev, dtype=torch.float32)
        dbuf[:, 2:2 + T] = cur[:, :T]
        dbuf[:, :2] = cur[:, :1]
        if extra > 0:
            dbuf[:, 2 + T:] = cur[:, T - 1:T]
        lat = z(B, F, DIM)
        self._conv(self.down, dbuf, T, lat, 0, act_in=0)
        # RVQ encode: semantic (1) and acoustic (n_q-1) chains both start from the latent
        codes = torch.empty((B, self.n_q, F), device=dev, dtype=torch.int32)
        M = B * F
        dots = torch.empty((M, BINS), device=dev, dtype=torch.float32)
        lib = _lib.lib()
        for g, (k0, n) in enumerate(((0, 1), (1, self.n_q - 1))):
            r = torch.empty((M, CB_DIM), device=dev, dtype=torch.float32)
            self._gemm(lat.data_ptr(), F * DIM, DIM, self.in_proj[g], r.data_ptr(), F * CB_DIM, CB_DIM, B, F, CB_DIM, DIM)
            for i in range(n):
                k = k0 + i
                self._gemm(r.data_ptr(), 0, CB_DIM, self.codebooks[k], dots.data_ptr(), 0, BINS, 1, M, BINS, CB_DIM)
                _g(lib.csmb_rvq_argmin_update, dots.data_ptr(), self.cb_norm2[k].data_ptr(), self.codebooks[k].data_ptr(),
                   r.data_ptr(), codes.data_ptr(), M, BINS, CB_DIM, self.n_q, k, F, self.dev_idx, self._st())
        return codes


class MimiDecodeStream:
    """Decoder-side streaming state for `B` parallel streams: conv left contexts, transposed-conv previous
    rows, transformer ring KV cache and position.  ``offline=True`` sizes it for one whole-utterance call."""

    def __init__(self, mimi: Mimi, batch: int, max_frames: int, offline: bool, use_graph: bool = False):
        self.m, self.B, self.F = mimi, batch, max_frames
        dev = mimi.device
        z = lambda *s: torch.zeros(s, device=dev, dtype=torch.float32)
        B, F = batch, max_frames
        T = 2 * F
        self.T = T
        self.codes = torch.zeros((B, mimi.n_q, F), device=dev, dtype=torch.int32)
        self.sem, self.ac = z(B, F, CB_DIM), z(B, F, CB_DIM)
        self.lat = z(B, F, DIM)
        self.up_prev = z(B, DIM)
        # transformer input lives inside the first conv's padded buffer (pad 6)
        self.buf0 = z(B, 6 + T, DIM)
        self.cache = z(LAYERS, B, CONTEXT + T, 2, HEADS, 64)
        self.pos = torch.zeros((1,), dtype=torch.int32, device=dev)
        self.scratch = mimi._tr_scratch(B, T)
        # SEANet decoder buffers
        self.bufs = []
        t, c = T, mimi.dec_first.cout
        self.tr_in = z(B, 1 + t, c)  # convtr input (row 0 = previous step's last row)
        for (ct, c1, c2) in mimi.dec_stages:
            t2 = t * ct.s
            res_in = z(B, 2 + t2, ct.cout)     # resblock input, pad 2 for k3
            hbuf = z(B, t2, c1.cout)
            self.bufs.append((res_in, hbuf))
            t = t2
        # outputs of resblocks feed the next convtr (pad 1) or the last conv (pad 2)
        self.next_in = []
        t = T
        for i, (ct, c1, c2) in enumerate(mimi.dec_stages):
            t *= ct.s
            last = i == len(mimi.dec_stages) - 1
            self.next_in.append(z(B, (2 if last else 1) + t, c2.cout))
        self.audio = z(B, t, 1)
        self.n_samples = t
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.use_graph = use_graph and not offline
        self._warm = 0

    def reset(self) -> None:
        """Back to the start-of-stream state (zero conv contexts, position 0); buffers and graph are kept."""
        self.pos.zero_()
        self.up_prev.zero_()
        self.buf0[:, :6].zero_()
        self.tr_in[:, :1].zero_()
        for res_in, _ in self.bufs:
            res_in[:, :2].zero_()
        for i, nxt in enumerate(self.next_in):
            nxt[:, : (2 if i == len(self.next_in) - 1 else 1)].zero_()

    def _run(self) -> None:
        m, B, F, T = self.m, self.B, self.F, self.T
        lib = _lib.lib()
        st = m._st()
        _g(lib.csmb_rvq_gather, self.codes.data_ptr(), m.codebooks.data_ptr(), self.sem.data_ptr(), self.ac.data_ptr(),
           B, m.n_q, F, BINS, CB_DIM, m.dev_idx, st)
        m._gemm(self.sem.data_ptr(), F * CB_DIM, CB_DIM, m.out_proj[0], self.lat.data_ptr(), F * DIM, DIM, B, F, DIM, CB_DIM)
        m._gemm(self.ac.data_ptr(), F * CB_DIM, CB_DIM, m.out_proj[1], self.lat.data_ptr(), F * DIM, DIM, B, F, DIM, CB_DIM,
                res=self.lat.data_ptr(), r_batch=F * DIM, ldr=DIM)
        # x2 upsample into buf0[:, 6:]
        xin = self.buf0.data_ptr() + 4 * 6 * DIM
        # the upsampler writes a dense [B][2F][C] block; with B > 1 the padded buffer is not dense, so go
        # through a dense temp in that case
        if B == 1:
            _g(lib.csmb_upsample_dw, self.lat.data_ptr(), self.up_prev.data_ptr(), m.up_w.data_ptr(), xin, B, F, DIM,
               m.dev_idx, st)
        else:
            tmp = self.scratch["h"]
            _g(lib.csmb_upsample_dw, self.lat.data_ptr(), self.up_prev.data_ptr(), m.up_w.data_ptr(), tmp.data_ptr(), B, F,
               DIM, m.dev_idx, st)
            _g(lib.csmb_copy_rows, tmp.data_ptr(), T * DIM, 0, self.buf0.data_ptr(), self.buf0.stride(0), 6, B, T, DIM, 0,
               m.dev_idx, st)
        _g(lib.csmb_copy_rows, self.lat.data_ptr(), F * DIM, F - 1, self.up_prev.data_ptr(), DIM, 0, B, 1, DIM, 0,
           m.dev_idx, st)
        m._transformer("decoder_transformer", self.buf0, 6, T, self.cache, self.pos, self.scratch)
        _g(lib.csmb_add_int, self.pos.data_ptr(), T, m.dev_idx, st)
        # SEANet decoder
        t = m._conv(m.dec_first, self.buf0, T, self.tr_in, 1, act_in=0)
        _g(lib.csmb_shift_rows, self.buf0.data_ptr(), self.buf0.stride(0), B, T, 6, DIM, m.dev_idx, st)
        cur = self.tr_in
        for i, (ct, c1, c2) in enumerate(m.dec_stages):
            res_in, hbuf = self.bufs[i]
            t_in = t
            t = m._convtr(ct, cur, t_in, res_in, 2, act_in=1)
            _g(lib.csmb_shift_rows, cur.data_ptr(), cur.stride(0), B, t_in, 1, ct.cin, m.dev_idx, st)
            m._conv(c1, res_in, t, hbuf, 0, act_in=1)
            nxt = self.next_in[i]
            last = i == len(m.dec_stages) - 1
            m._conv(c2, hbuf, t, nxt, 2 if last else 1, act_in=1, res=res_in, res_pad=2)
            _g(lib.csmb_shift_rows, res_in.data_ptr(), res_in.stride(0), B, t, 2, c1.cin, m.dev_idx, st)
            cur = nxt
        m._conv(m.dec_last, cur, t, self.audio, 0, act_in=1)
        _g(lib.csmb_shift_rows, cur.data_ptr(), cur.stride(0), B, t, 2, m.dec_last.cin, m.dev_idx, st)

    def step(self, codes: torch.Tensor) -> torch.Tensor:
        """codes (B,K,F) -> (B,1,1920*F).  The returned tensor is a view of an internal buffer that the next
        step overwrites."""
        assert tuple(codes.shape) == (self.B, self.m.n_q, self.F), (codes.shape, (self.B, self.m.n_q, self.F))
        self.codes.copy_(codes.to(torch.int32))
        if self.use_graph:
            if self.graph is None and self._warm >= 1:
                # capture replays exactly the kernel sequence of _run on fixed buffers; positions live on device
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._run()
                self.graph = g
                # capture does not execute: run it now
                self.graph.replay()
            elif self.graph is not None:
                self.graph.replay()
            else:
                self._warm += 1
                self._run()
        else:
            self._run()
        return self.audio.reshape(self.B, 1, self.n_samples)

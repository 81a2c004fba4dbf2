"""Device-side generation state for a batch of utterances: paged KV pools, block tables, scratch, positions.

Plays the role of the per-call ``KVCache`` lists of the reference (``/root/reference/csm_mlx/generation.py:70``
decoder cache re-created every frame, ``:127`` backbone cache per ``generate`` call): the caller's ``generate``
creates one ``LMState`` and drops it at the end.  All arithmetic happens in libcsm_b200.so.
"""

from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import _lib
from .config import MAX_SEQ_LEN
from .models import CSM


@dataclass
class SamplerSpec:
    """Device-side sampler parameters (mlx_lm ``make_sampler`` arguments, cli/generate.py:168-174)."""

    temperature: float = 0.0
    top_k: int = 0
    top_p: float = 0.0
    min_p: float = 0.0
    min_tokens_to_keep: int = 1
    seed: Optional[int] = 0    # None: "draw a fresh key per generate() call" (resolved by generation._resolve_sampler)

    def to_c(self) -> _lib.Sampler:
        return _lib.Sampler(float(self.temperature), int(self.top_k) if self.top_k and self.top_k > 0 else 0,
                            float(self.top_p), float(self.min_p), int(self.min_tokens_to_keep),
                            int(self.seed or 0) & 0xFFFFFFFFFFFFFFFF)


class LMState:
    def __init__(self, model: CSM, batch: int, max_len: int = MAX_SEQ_LEN, row_invariant: bool = False):
        """``row_invariant``: every Linear of the row-based calls runs on the tensor-core path whatever the row count
        (``CSMB_BATCH_ROW_INVARIANT``) and the fused chain is used from one sequence up, so that a sequence's tokens do
        not depend on what it is batched with — the serving engine sets it (generation.py:139-161 is a batch-1 loop)."""
        model._require_loaded()
        self.model = model
        self.row_invariant = bool(row_invariant)
        self.device = model.device
        self.dev_idx = _lib.require_device(self.device)
        self.batch = batch
        self.max_len = min(int(max_len), MAX_SEQ_LEN)
        b, d = model.backbone.args, model.decoder.args
        ncb = model.n_audio_codebooks
        P = _lib.PAGE
        self.pages_per_seq = (self.max_len + P - 1) // P
        n_pages = batch * self.pages_per_seq
        page_floats_b = 2 * b.num_key_value_heads * P * b.head_dim
        self.kv_pool = torch.zeros((b.num_hidden_layers, n_pages, page_floats_b), device=self.device,
                                   dtype=torch.float32)
        self.block_table = torch.arange(n_pages, device=self.device, dtype=torch.int32).reshape(batch, self.pages_per_seq)
        dec_pages = (ncb + P - 1) // P
        page_floats_d = 2 * d.num_key_value_heads * P * d.head_dim
        self.dec_kv_pool = torch.zeros((d.num_hidden_layers, batch * dec_pages, page_floats_d), device=self.device,
                                       dtype=torch.float32)
        self._ws_rows = 0
        self.workspace: Optional[torch.Tensor] = None
        self._ensure_workspace(2 * batch)
        self.pos = torch.zeros((batch,), device=self.device, dtype=torch.int32)  # next position per sequence
        self.pos_host = [0] * batch
        self.h_last = torch.empty((batch, b.hidden_size), device=self.device, dtype=torch.float32)
        self.c0_logits = torch.empty((batch, model.n_audio_vocab), device=self.device, dtype=torch.float32)
        self._seq_iota = torch.arange(batch, device=self.device, dtype=torch.int32)
        # Buffers of the fused chain are created HERE, never lazily inside decode_frame: a torch.zeros under CUDA-graph
        # capture becomes a memset node that every replay would repeat (wiping the sticky error flag and the admission
        # overrides armed for that step).
        self._fast_ws: Optional[torch.Tensor] = None
        self._ovr_x: Optional[torch.Tensor] = None
        self._ovr_flag: Optional[torch.Tensor] = None
        self._ovr_armed = False
        self._chain_opts = None   # (key, _lib.ChainOpts) of the last chain call
        self._fws: Optional[torch.Tensor] = None
        self.frame_status: Optional[torch.Tensor] = None
        self._frame_opts: Optional[_lib.FrameOpts] = None
        self._graph_key = None
        self.graph_launches = 0
        if self._chain_possible():
            self._prepare_chain()

    # ------------------------------------------------------------------ reuse across utterances
    @classmethod
    def acquire(cls, model: CSM, batch: int, max_len: int = MAX_SEQ_LEN) -> "LMState":
        """A rewound state from the model's pool (KV pages, workspaces and captured graphs are kept between
        utterances), or a new one whose reservation is rounded up to 256 positions so later requests fit."""
        need = min(int(max_len), MAX_SEQ_LEN)
        pool = model.__dict__.setdefault("_lm_pool", [])
        for i, st in enumerate(pool):
            if st.model is model and st.batch == batch and st.max_len >= need and not st.row_invariant:
                pool.pop(i)
                st.reset()
                return st
        return cls(model, batch, min(MAX_SEQ_LEN, -(-need // 256) * 256))

    def release(self) -> None:
        """Hand the state back for the next utterance (the caller's stream must be ordered after its last use)."""
        pool = self.model.__dict__.setdefault("_lm_pool", [])
        if len(pool) < 4 and self.kv_pool.numel() * 4 <= (1 << 30) and all(st is not self for st in pool):
            pool.append(self)

    # ------------------------------------------------------------------ plumbing
    def _ensure_workspace(self, rows: int) -> None:
        if rows <= self._ws_rows:
            return
        nbytes = _lib.lib().csmb_lm_workspace_bytes(C.byref(self.model.desc()), rows)
        self.workspace = torch.zeros((nbytes,), device=self.device, dtype=torch.uint8)  # zeroed: sticky error flag inside
        self._ws_rows = rows
        self._graph_key = None  # a graph captured over the per-op path holds pointers into the old workspace

    def _batch_desc(self) -> _lib.Batch:
        b = _lib.Batch()
        b.batch, b.max_pages = self.batch, self.pages_per_seq
        b.kv_pool, b.kv_layer_stride = self.kv_pool.data_ptr(), self.kv_pool.stride(0)
        b.block_table = self.block_table.data_ptr()
        b.dec_kv_pool, b.dec_kv_layer_stride = self.dec_kv_pool.data_ptr(), self.dec_kv_pool.stride(0)
        b.workspace, b.workspace_bytes = self.workspace.data_ptr(), self.workspace.numel()
        b.flags = _lib.BATCH_ROW_INVARIANT if self.row_invariant else 0
        return b

    def _stream(self) -> int:
        return _lib.stream_ptr(self.device)

    # ------------------------------------------------------------------ steps
    def stage_prefill(self, tokens: Sequence[torch.Tensor], masks: Sequence[torch.Tensor]) -> dict:
        """Host -> device staging of prompt rows and their (sequence, position) maps; see ``prefill``."""
        assert len(tokens) == self.batch
        lens = [int(t.shape[0]) for t in tokens]
        for b, n in enumerate(lens):
            if self.pos_host[b] + n > self.max_len:
                raise ValueError("sequence exceeds the KV pages reserved for it")
        nb = dict(non_blocking=True)
        tok = torch.cat([t.to(torch.int32) for t in tokens], 0).contiguous().to(self.device, **nb)
        msk = torch.cat([m.to(torch.uint8) for m in masks], 0).contiguous().to(self.device, **nb)
        seq = torch.cat([torch.full((n,), b, dtype=torch.int32) for b, n in enumerate(lens)]).to(self.device, **nb)
        pos = torch.cat([torch.arange(self.pos_host[b], self.pos_host[b] + n, dtype=torch.int32)
                         for b, n in enumerate(lens)]).to(self.device, **nb)
        ends, acc = [], 0
        for n in lens:
            acc += n
            ends.append(acc - 1)
        last = torch.tensor(ends, dtype=torch.int32).to(self.device, **nb)
        new_pos = torch.tensor([p + n for p, n in zip(self.pos_host, lens)], dtype=torch.int32).to(self.device, **nb)
        return {"tok": tok, "msk": msk, "seq": seq, "pos": pos, "last": last, "lens": lens, "new_pos": new_pos,
                "h2d_bytes": tok.numel() * 4 + msk.numel() + seq.numel() * 4 + pos.numel() * 4 + last.numel() * 4
                + new_pos.numel() * 4}

    def _prefill_fast_ok(self) -> bool:
        """Prompt rows go through the chain's kernels (csmb_prefill_fast: one tcgen05 launch per Linear, fused norms, 8 launches
        per layer) when the model has the chain's shapes; ``CSMB_DISABLE_PREFILL_FAST=1`` keeps the per-op kernels."""
        if os.environ.get("CSMB_DISABLE_PREFILL_FAST", "0") == "1":
            return False
        if getattr(self, "_pf_supported", None) is None:
            s = SamplerSpec().to_c()
            self._pf_supported = bool(_lib.lib().csmb_decode_frame_fast_supported(C.byref(self.model.desc()), C.byref(s)))
        return self._pf_supported

    def _prefill_ws(self, rows: int) -> torch.Tensor:
        if getattr(self, "_pf_ws", None) is None or self._pf_rows < rows:
            rows_alloc = max(rows, 64)
            nbytes = _lib.lib().csmb_prefill_fast_workspace_bytes(C.byref(self.model.desc()), rows_alloc)
            self._pf_ws = torch.zeros((nbytes,), device=self.device, dtype=torch.uint8)   # sticky error flag inside
            self._pf_rows = rows_alloc
        return self._pf_ws

    def _backbone_rows(self, tok, msk, seq, pos, R: int, last, n_last: int, h_last, c0_logits) -> None:
        """model.backbone over R prompt rows (generation.py:34-42 with T > 1) + final norm / c0 head on the listed rows."""
        bd = self._batch_desc()
        if self._prefill_fast_ok():
            ws = self._prefill_ws(R)
            _lib.check(_lib.lib().csmb_prefill_fast(
                C.byref(self.model.desc()), C.byref(bd), tok.data_ptr(), msk.data_ptr(), seq.data_ptr(), pos.data_ptr(), R,
                None if n_last == 0 else last.data_ptr(), n_last, None if n_last == 0 else h_last.data_ptr(),
                None if (n_last == 0 or c0_logits is None) else c0_logits.data_ptr(), ws.data_ptr(), ws.numel(), self.dev_idx,
                self._stream()))
            return
        self._ensure_workspace(max(R, 2 * self.batch))
        bd = self._batch_desc()
        _lib.check(_lib.lib().csmb_backbone_forward(
            C.byref(self.model.desc()), C.byref(bd), tok.data_ptr(), msk.data_ptr(), seq.data_ptr(), pos.data_ptr(), R,
            last.data_ptr(), n_last, h_last.data_ptr(), None if c0_logits is None else c0_logits.data_ptr(), self.dev_idx,
            self._stream()))

    def run_prefill(self, staged: dict) -> None:
        R = int(staged["tok"].shape[0])
        self._backbone_rows(staged["tok"], staged["msk"], staged["seq"], staged["pos"], R, staged["last"], self.batch,
                            self.h_last, self.c0_logits)
        for b, n in enumerate(staged["lens"]):
            self.pos_host[b] += n
        self.pos.copy_(staged["new_pos"])

    def prefill(self, tokens: Sequence[torch.Tensor], masks: Sequence[torch.Tensor]) -> None:
        """Backbone over every sequence's prompt rows ((T_b,33) each); leaves h_last / c0_logits of the last
        row of each sequence and advances positions (generation.py:34-42 with T>1)."""
        self.run_prefill(self.stage_prefill(tokens, masks))

    def sample_c0(self, frame: torch.Tensor, sampler: SamplerSpec, logits: Optional[torch.Tensor] = None) -> None:
        """frame[:,0] = sample(c0 logits)   (generation.py:51-56).  RNG draw index = (pos-1)*n_codebooks."""
        lg = self.c0_logits if logits is None else logits.to(device=self.device, dtype=torch.float32).contiguous()
        ncb = self.model.n_audio_codebooks
        s = sampler.to_c()
        pos_prev = (self.pos - 1).contiguous()  # position of the row that produced these logits
        _lib.check(_lib.lib().csmb_sample(lg.data_ptr(), lg.shape[1], frame.data_ptr(), ncb, self.batch, lg.shape[1],
                                          C.byref(s), 0, pos_prev.data_ptr(), ncb, self.dev_idx, self._stream()))

    def depth_decode(self, frame: torch.Tensor, sampler: SamplerSpec, logits_out: Optional[torch.Tensor] = None,
                     forced: Optional[torch.Tensor] = None, step_begin: int = 1, step_end: Optional[int] = None) -> None:
        """generation.py:56-90: fills frame[:, step_begin:step_end] from h_last and frame[:,0]."""
        ncb = self.model.n_audio_codebooks
        bd = self._batch_desc()
        s = sampler.to_c()
        pos_prev = (self.pos - 1).contiguous()
        _lib.check(_lib.lib().csmb_depth_decode(
            C.byref(self.model.desc()), C.byref(bd), self.h_last.data_ptr(), frame.data_ptr(), C.byref(s), 0,
            pos_prev.data_ptr(), None if logits_out is None else logits_out.data_ptr(),
            None if forced is None else forced.data_ptr(), step_begin, ncb if step_end is None else step_end,
            self.dev_idx, self._stream()))

    def backbone_step(self, prev_frame: torch.Tensor) -> None:
        """One T=1 backbone step from the previous frame (generation.py:156-161 then :34-42)."""
        ncb = self.model.n_audio_codebooks
        tok = torch.cat([prev_frame.to(torch.int32), torch.zeros((self.batch, 1), dtype=torch.int32, device=self.device)], 1)
        msk = torch.cat([torch.ones((self.batch, ncb), dtype=torch.uint8, device=self.device),
                         torch.zeros((self.batch, 1), dtype=torch.uint8, device=self.device)], 1)
        self._check_room()
        bd = self._batch_desc()
        _lib.check(_lib.lib().csmb_backbone_forward(
            C.byref(self.model.desc()), C.byref(bd), tok.contiguous().data_ptr(), msk.contiguous().data_ptr(),
            self._seq_iota.data_ptr(), self.pos.data_ptr(), self.batch, self._seq_iota.data_ptr(), self.batch,
            self.h_last.data_ptr(), self.c0_logits.data_ptr(), self.dev_idx, self._stream()))
        self._advance()

    def _chain_min_batch(self) -> int:
        return 1 if self.row_invariant else int(os.environ.get("CSMB_FAST_MIN_BATCH", "2"))

    def _chain_possible(self) -> bool:
        if os.environ.get("CSMB_DISABLE_FAST", "0") == "1" or self.batch < self._chain_min_batch():
            return False
        s = SamplerSpec().to_c()
        return bool(_lib.lib().csmb_decode_frame_fast_supported(C.byref(self.model.desc()), C.byref(s)))

    def _prepare_chain(self) -> None:
        """Workspace, admission-override buffers and per-call options of the fused chain (eager; see __init__)."""
        if self._fast_ws is not None:
            return
        nbytes = _lib.lib().csmb_decode_frame_fast_workspace_bytes(C.byref(self.model.desc()), self.batch)
        self._fast_ws = torch.zeros((nbytes,), device=self.device, dtype=torch.uint8)  # sticky error flag inside
        d = self.model.backbone.args.hidden_size
        self._ovr_x = torch.zeros((self.batch, d), device=self.device, dtype=torch.float32)
        self._ovr_flag = torch.zeros((self.batch,), device=self.device, dtype=torch.uint8)
        if os.environ.get("CSMB_NO_PROJ_TABLE", "0") != "1":
            self.model.proj_table()   # built once per model (260 MB for csm_1b), outside any capture

    def _chain_opts_now(self):
        """csmb_chain_opts of the next chain call, from the (A/B, debug) environment switches: (key tuple, struct)."""
        no_pdl = 1 if os.environ.get("CSMB_CHAIN_NO_PDL", "0") == "1" else 0
        flags = int(os.environ.get("CSMB_CHAIN_FLAGS", "0"))
        smem_kb = int(os.environ.get("CSMB_CHAIN_SMEM_KB", "0"))
        tab = None if os.environ.get("CSMB_NO_PROJ_TABLE", "0") == "1" else self.model.proj_table()
        key = (no_pdl, flags, smem_kb, tab is not None)
        if self._chain_opts is None or self._chain_opts[0] != key:
            o = _lib.ChainOpts()
            o.no_pdl, o.flags, o.smem_kb = no_pdl, flags, smem_kb
            o.proj_table = tab.data_ptr() if tab is not None else None
            self._chain_opts = (key, o)
        return self._chain_opts

    def fast_supported(self, sampler: SamplerSpec) -> bool:
        """True if the fused kernel chain of csrc/batch_frame.cu covers this model, batch and sampler (tensor-core linears
        from 2 sequences up — from 1 for a row-invariant state; ``CSMB_DISABLE_FAST=1`` / ``CSMB_FAST_MIN_BATCH`` override)."""
        if self._fast_ws is None or os.environ.get("CSMB_DISABLE_FAST", "0") == "1":
            return False
        s = sampler.to_c()
        return bool(_lib.lib().csmb_decode_frame_fast_supported(C.byref(self.model.desc()), C.byref(s)))

    def decode_frame(self, prev_frame: torch.Tensor, frame: torch.Tensor, sampler: SamplerSpec) -> None:
        """Whole frame on device, no host round trip (generation.py:21-92 with T=1)."""
        self._check_room()
        bd = self._batch_desc()
        s = sampler.to_c()
        if self.fast_supported(sampler):
            _lib.check(_lib.lib().csmb_decode_frame_fast_admit(
                C.byref(self.model.desc()), C.byref(bd), prev_frame.data_ptr(), self.pos.data_ptr(), frame.data_ptr(),
                C.byref(s), 0, self._ovr_x.data_ptr(), self._ovr_flag.data_ptr(), C.byref(self._chain_opts_now()[1]),
                self._fast_ws.data_ptr(), self._fast_ws.numel(), self.dev_idx, self._stream()))
        else:
            _lib.check(_lib.lib().csmb_decode_frame(
                C.byref(self.model.desc()), C.byref(bd), prev_frame.data_ptr(), self.pos.data_ptr(), frame.data_ptr(),
                C.byref(s), 0, self.dev_idx, self._stream()))
        self._advance()

    # ------------------------------------------------------------------ admission into a running batch (fused chain)
    def prefill_rows(self, slots: Sequence[int], tokens: Sequence[torch.Tensor], masks: Sequence[torch.Tensor]) -> None:
        """Backbone over the given rows of the given sequence slots only (KV append at their current positions); the other
        sequences are untouched.  generation.py:34-42 restricted to some sequences."""
        lens = [int(t.shape[0]) for t in tokens]
        R = sum(lens)
        if R == 0:
            return
        for s, n in zip(slots, lens):
            if self.pos_host[s] + n > self.max_len:
                raise ValueError("sequence exceeds the KV pages reserved for it")
        nb = dict(non_blocking=True)
        tok = torch.cat([t.to(torch.int32) for t in tokens], 0).contiguous().to(self.device, **nb)
        msk = torch.cat([m.to(torch.uint8) for m in masks], 0).contiguous().to(self.device, **nb)
        seq = torch.cat([torch.full((n,), s, dtype=torch.int32) for s, n in zip(slots, lens)]).to(self.device, **nb)
        pos = torch.cat([torch.arange(self.pos_host[s], self.pos_host[s] + n, dtype=torch.int32)
                         for s, n in zip(slots, lens)]).to(self.device, **nb)
        ends, acc = [], 0
        for n in lens:
            acc += n
            ends.append(max(acc - 1, 0))
        live = [i for i, n in enumerate(lens) if n > 0]
        last = torch.tensor([ends[i] for i in live], dtype=torch.int32).to(self.device, **nb)
        b = self.model.backbone.args
        if self._prefill_fast_ok():
            self._backbone_rows(tok, msk, seq, pos, R, None, 0, None, None)   # only the KV cache is wanted
        else:
            h_tmp = torch.empty((len(live), b.hidden_size), device=self.device, dtype=torch.float32)
            lg_tmp = torch.empty((len(live), self.model.n_audio_vocab), device=self.device, dtype=torch.float32)
            self._backbone_rows(tok, msk, seq, pos, R, last, len(live), h_tmp, lg_tmp)
        for s, n in zip(slots, lens):
            self.pos_host[s] += n

    def export_kv_prefix(self, slot: int, n_rows: int) -> torch.Tensor:
        """A copy of the backbone KV pages that hold positions [0, n_rows) of ``slot`` ((layers, pages, page floats), stream
        ordered).  The prompt pass is row-invariant, so these entries depend on the slot's first n_rows prompt rows only."""
        n_pages = -(-int(n_rows) // _lib.PAGE)
        p0 = slot * self.pages_per_seq
        return self.kv_pool[:, p0:p0 + n_pages].clone()

    def import_kv_prefix(self, slot: int, pages: torch.Tensor) -> None:
        """Pages of ``export_kv_prefix`` (of any slot of any state of the same model) become the first pages of ``slot``."""
        n_pages = int(pages.shape[1])
        if n_pages > self.pages_per_seq or pages.shape[0] != self.kv_pool.shape[0] or pages.shape[2] != self.kv_pool.shape[2]:
            raise ValueError("KV prefix does not fit this state")
        p0 = slot * self.pages_per_seq
        self.kv_pool[:, p0:p0 + n_pages].copy_(pages)

    def arm_admission(self, slots: Sequence[int], tokens: Sequence[torch.Tensor], masks: Sequence[torch.Tensor],
                      known_rows: Optional[Sequence[int]] = None) -> None:
        """Prepare the next fused-chain step so that the sequences in ``slots`` start from their prompts ((T, 33) rows
        each) while every other slot decodes normally: rows 0..T-2 are prefilled now, the embedded last row becomes the
        slot's backbone input of the step (csmb_decode_frame_fast_admit) and its position is set to T-1.
        ``known_rows[i]`` leading rows of request i are already in the slot's KV pages (``import_kv_prefix``) and are skipped."""
        if self._fast_ws is None:
            raise _lib.CsmbError("arm_admission needs the fused chain (unsupported model shape, or disabled)")
        ncb = self.model.n_audio_codebooks
        known = [0] * len(slots) if known_rows is None else [int(k) for k in known_rows]
        for s, k, t in zip(slots, known, tokens):
            if not 0 <= k <= int(t.shape[0]) - 1:
                raise ValueError("known_rows must leave at least the last prompt row")
            self.pos_host[s] = k
        self.prefill_rows(slots, [t[k:-1] for t, k in zip(tokens, known)], [m[k:-1] for m, k in zip(masks, known)])
        n = len(slots)
        last_tok = torch.stack([t[-1].to(torch.int32) for t in tokens]).contiguous().to(self.device)
        last_msk = torch.stack([m[-1].to(torch.uint8) for m in masks]).contiguous().to(self.device)
        d = self.model.backbone.args.hidden_size
        emb = torch.empty((n, d), device=self.device, dtype=torch.float32)
        _lib.check(_lib.lib().csmb_embed_sum(
            last_tok.data_ptr(), last_msk.data_ptr(), self.model.text_embeddings.weight.data_ptr(),
            self.model.audio_embeddings.weight.data_ptr(), emb.data_ptr(), n, d, ncb, self.model.n_audio_vocab,
            self.dev_idx, self._stream()))
        idx = torch.tensor(list(slots), dtype=torch.long, device=self.device)
        self._ovr_x.index_copy_(0, idx, emb)
        self._ovr_flag.index_fill_(0, idx, 1)
        self.pos.index_copy_(0, idx, torch.tensor([self.pos_host[s] for s in slots], dtype=torch.int32, device=self.device))
        self._ovr_armed = True

    def disarm_admission(self) -> None:
        if self._ovr_armed:
            self._ovr_flag.zero_()
            self._ovr_armed = False

    def decode_frame_graphed(self, prev_frame: torch.Tensor, sampler: SamplerSpec) -> torch.Tensor:
        """decode_frame through a CUDA graph captured on first use (fixed buffers; positions live on the device
        and advance inside the graph).  Returns a fresh (B, n_codebooks) int32 tensor."""
        path = "chain" if self.fast_supported(sampler) else "per-op"
        # greedy decoding draws no random numbers: the seed is not part of the key then, so that the graph of a pooled state
        # survives from utterance to utterance
        greedy = sampler.temperature == 0
        key = (sampler.temperature, sampler.top_k, sampler.top_p, sampler.min_p, sampler.min_tokens_to_keep,
               None if greedy else sampler.seed, path, self._chain_opts_now()[0] if path == "chain" else None)
        if self._graph_key != key:
            ncb = self.model.n_audio_codebooks
            self._g_prev = torch.zeros((self.batch, ncb), device=self.device, dtype=torch.int32)
            self._g_out = torch.zeros((self.batch, ncb), device=self.device, dtype=torch.int32)
            self._g_prev.copy_(prev_frame)
            pos_save, host_save = self.pos.clone(), list(self.pos_host)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            n0 = _lib.lib().csmb_debug_launch_count()
            with torch.cuda.graph(g):
                self.decode_frame(self._g_prev, self._g_out, sampler)
            self.graph_launches = int(_lib.lib().csmb_debug_launch_count() - n0)   # kernels of this library per replay
            self.pos.copy_(pos_save)  # capture does not execute, but keep host/device views in lock-step
            self.pos_host = host_save
            self._graph, self._graph_key = g, key
        self._check_room()
        self._g_prev.copy_(prev_frame)
        self._graph.replay()
        self.pos_host = [p + 1 for p in self.pos_host]
        return self._g_out.clone()

    # ------------------------------------------------------------------ fused persistent frame kernel (B = 1)
    def slot_fused_supported(self, sampler: SamplerSpec) -> bool:
        """True if the persistent frame kernel covers this model and sampler for ONE sequence of this state."""
        a, d = self.model.backbone.args, self.model.decoder.args
        shape = (a.hidden_size, a.num_attention_heads, a.num_key_value_heads, a.head_dim, d.hidden_size,
                 d.num_attention_heads, d.num_key_value_heads, d.head_dim) == (2048, 32, 8, 64, 1024, 8, 2, 128)
        # in-kernel samplers of k_frame: greedy; temperature with top-k, top-p and / or min-p (not min-p with min_tokens_to_keep > 1)
        plain = sampler.temperature == 0 or not (sampler.min_p > 0 and sampler.min_tokens_to_keep > 1)
        return bool(shape and plain and 3 <= self.model.n_audio_codebooks <= 32)   # bf16 and weight-only FP8 models alike

    def fused_supported(self, sampler: SamplerSpec) -> bool:
        return self.batch == 1 and self.slot_fused_supported(sampler)

    def _frame_workspace(self) -> None:
        if self._fws is None:
            nbytes = _lib.lib().csmb_frame_workspace_bytes(C.byref(self.model.desc()), self.dev_idx)
            self._fws = torch.zeros((nbytes,), device=self.device, dtype=torch.uint8)
            self.frame_status = torch.zeros((1,), device=self.device, dtype=torch.int32)  # sticky: first abort code ever
            o = _lib.FrameOpts()
            o.ctas = int(os.environ.get("CSMB_FRAME_CTAS", "0"))
            o.flags = int(os.environ.get("CSMB_FRAME_FLAGS", "0"))   # A/B, debug (include/csm_b200.h csmb_frame_opts.flags)
            self._frame_opts = o

    def first_frame_fused(self, sampler: SamplerSpec) -> torch.Tensor:
        """The frame that follows a prefill (codebook-0 head + sampling + depth loop on ``h_last``) in one launch of
        the persistent kernel (csmb_frame_b1_depth).  Returns a fresh (1, n_codebooks) tensor."""
        self._frame_workspace()
        frame = torch.empty((1, self.model.n_audio_codebooks), device=self.device, dtype=torch.int32)
        s = sampler.to_c()
        _lib.check(_lib.lib().csmb_frame_b1_depth(
            C.byref(self.model.desc()), self.h_last.data_ptr(), self.pos.data_ptr(), frame.data_ptr(), C.byref(s), 0,
            C.byref(self._frame_opts), self._fws.data_ptr(), self._fws.numel(), self.frame_status.data_ptr(), self.dev_idx,
            self._stream()))
        return frame

    def decode_frame_fused(self, prev_frame: torch.Tensor, sampler: SamplerSpec) -> torch.Tensor:
        """One whole frame in the persistent kernel (csmb_frame_b1).  Returns a fresh (1, n_codebooks) tensor."""
        self._frame_workspace()
        self._check_room()
        frame = torch.empty((1, self.model.n_audio_codebooks), device=self.device, dtype=torch.int32)
        s = sampler.to_c()
        prev = prev_frame if prev_frame.dtype == torch.int32 and prev_frame.is_contiguous() else prev_frame.to(torch.int32).contiguous()
        _lib.check(_lib.lib().csmb_frame_b1(
            C.byref(self.model.desc()), self.kv_pool.data_ptr(), self.kv_pool.stride(0), self.block_table.data_ptr(),
            prev.data_ptr(), self.pos.data_ptr(), frame.data_ptr(), C.byref(s), 0, C.byref(self._frame_opts),
            self._fws.data_ptr(), self._fws.numel(), self.frame_status.data_ptr(), self.dev_idx, self._stream()))
        self._advance()
        return frame

    def decode_frame_slot(self, slot: int, prev_frame: torch.Tensor, sampler: SamplerSpec) -> torch.Tensor:
        """One frame for sequence ``slot`` ALONE through the persistent frame kernel (csmb_frame_b1_slot) — what a serving
        loop runs while only one of its slots is busy: 3.2 ms instead of the batched chain's 6.7 ms.  ``prev_frame`` is the
        (batch, n_codebooks) frame tensor of the last step (only row ``slot`` is read); returns a fresh tensor of that
        shape with row ``slot`` written (other rows zero).  Every sequence's position advances, like after any frame-step."""
        self._frame_workspace()
        self._check_room()
        ncb = self.model.n_audio_codebooks
        frame = torch.zeros((self.batch, ncb), device=self.device, dtype=torch.int32)
        s = sampler.to_c()
        prev = prev_frame if prev_frame.dtype == torch.int32 and prev_frame.is_contiguous() else prev_frame.to(torch.int32).contiguous()
        _lib.check(_lib.lib().csmb_frame_b1_slot(
            C.byref(self.model.desc()), self.kv_pool.data_ptr(), self.kv_pool.stride(0),
            self.block_table.data_ptr() + slot * self.block_table.stride(0) * 4, prev.data_ptr() + slot * ncb * 4,
            self.pos.data_ptr() + slot * 4, frame.data_ptr() + slot * ncb * 4, C.byref(s), 0, slot,
            C.byref(self._frame_opts), self._fws.data_ptr(), self._fws.numel(), self.frame_status.data_ptr(), self.dev_idx,
            self._stream()))
        self._advance()
        return frame

    def status_word(self) -> torch.Tensor:
        """(1,) int32 DEVICE view of the sticky abort flag of whichever fused path this state runs (zero = healthy): cheap
        enough to ride along with every frame's device-to-host copy, so a failed frame is noticed one frame late."""
        if self.frame_status is not None:
            return self.frame_status
        if self._fast_ws is not None:
            return self._fast_ws[:4].view(torch.int32)
        if getattr(self, "_zero_status", None) is None:
            self._zero_status = torch.zeros((1,), device=self.device, dtype=torch.int32)
        return self._zero_status

    def check_status(self) -> None:
        """Raises if any frame since the state was created reported a timed-out wait (synchronises).  Both flags are sticky:
        the kernels only ever set them, so a failure in an early frame is still seen at the end of an utterance."""
        st = self.frame_status
        if st is not None and int(st.item()) != 0:
            code = int(st.item())
            st.zero_()
            raise _lib.CsmbError(f"persistent frame kernel aborted (code {code})")
        for fw in (self._fast_ws, getattr(self, "_pf_ws", None)):
            if fw is not None and int(fw[:4].view(torch.int32).item()) != 0:
                fw[:4].zero_()
                raise _lib.CsmbError("fused batched frame / prefill: a bounded wait of the tensor-core linear timed out")

    def reset(self) -> None:
        """Rewind every sequence to position 0 (new utterances in the same slots).  KV pages are simply
        overwritten; captured graphs stay valid because no buffer moves."""
        self.pos_host = [0] * self.batch
        self.pos.zero_()

    def _check_room(self) -> None:
        if max(self.pos_host) + 1 > self.max_len:
            raise ValueError("sequence exceeds the KV pages reserved for it")

    def _advance(self) -> None:
        self.pos_host = [p + 1 for p in self.pos_host]
        self.pos.add_(1)

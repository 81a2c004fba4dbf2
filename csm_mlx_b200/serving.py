"""Continuous batching of independent utterances on one GPU, and a cache of context-segment codes.

The reference generates one utterance at a time (``/root/reference/csm_mlx/generation.py:124,156`` assert batch 1)
and its streaming demo feeds requests to it from a thread pool, re-tokenising up to six context segments on every
turn (``run_streaming_csm_mlx.py:102, 830-872, 960-965, 1060-1073``).  SURVEY.md §8e asks for the throughput form of
the same path: requests sharded over GPUs, and *within a GPU* a fixed set of sequence slots over the paged KV pool
that requests join and leave between frames.  This module is that host logic; every arithmetic step still runs in
libcsm_b200.so through ``runtime.LMState`` (prefill / mixed steps on the per-op kernels, steady-state frames on the
fused kernel chain replayed as a CUDA graph) and ``mimi.Mimi``.

* ``Engine``: ``submit()`` queues a request, ``step()`` generates one frame for every occupied slot and admits queued
  requests into free slots — on the fused chain all but the last prompt row of a new request are prefilled on the side
  and that last row is the slot's input of the very step in which the running sequences decode
  (``csmb_decode_frame_fast_admit``), so admission never stalls the batch; with samplers the chain does not fuse, the
  prompts are prefilled in one mixed per-op backbone pass together with the running sequences' rows —; ``run()`` drains
  everything.  EOS (an all-zero frame, generation.py:151-152) and the
  frame budget are checked on pinned host copies one step late, so the GPU never waits for Python; a finished
  sequence frees its slot, whose KV pages are simply overwritten by the next occupant.
* ``ContextCache``: Mimi codes of context audio keyed by content, so a conversation's segments are encoded once
  (tokenizers.py:61-85 is re-run per turn by the reference).
* ``KVPrefixCache``: the backbone KV pages of a prompt's context rows (everything before the text to speak) keyed by those
  rows, so the next turn of a conversation with the same context copies ~64 KiB per row into its slot instead of running
  the prompt pass over them again (generation.py:108-121 re-prefills the whole context every call).  The prompt pass is
  row-invariant, so a request's tokens are the same with a hit, a miss or no cache.

**Batch invariance.**  A request's greedy tokens do not depend on ``max_batch``, on which other requests share its steps,
on when it is admitted or on how a job is sharded over GPUs: the engine's ``LMState`` is ``row_invariant`` (every Linear of
the prompt prefill on the tensor-core path with a split-K factor that is a function of the Linear's shape only, the fused
chain from one sequence up), so one numeric path serves every batch size — the property of the reference's batch-1 loop
(generation.py:139-161).  ``tests/test_gpu_api.py::test_engine_tokens_invariant_to_batch_size`` and the token checksum of
``bench.py`` gate it.  The batch-1 frame kernel (``generate`` / ``stream_generate``'s latency path, CUDA-core fp32 sums) is a
different summation order: against it, argmax near-ties of a random-init model can break differently; ``solo_kernel=True``
opts a lone busy slot into that kernel (3.2 instead of 6.7 ms per frame) at the price of that invariance.
With temperature sampling the random draws are indexed by (seed, position, slot), so they depend on the slot a
request lands in — like any batched sampler.
"""

from __future__ import annotations

import os
from collections import deque
from dataclasses import dataclass, field
from typing import Deque, Dict, List, Optional, Sequence, Tuple, Union

import torch

from .caches import ContextCache, KVPrefixCache  # noqa: F401  (re-exported: the engine's conversation caches)
from .config import MAX_SEQ_LEN
from .generation import _check_length
from .models import CSM
from .runtime import LMState, SamplerSpec
from .segment import Segment
from .tokenizers import get_audio_tokenizer, tokenize_text_segment


@dataclass
class Request:
    rid: int
    tokens: torch.Tensor           # (T, 33) int32 prompt rows
    mask: torch.Tensor             # (T, 33) bool
    max_frames: int
    prefix_rows: int = 0           # leading rows shared with other requests (context segments): KV-prefix cache unit
    frames: List[torch.Tensor] = field(default_factory=list)   # (32,) int32 CPU rows, EOS excluded
    slot: int = -1
    issued: int = 0                # frames enqueued on the GPU so far
    done: bool = False


def plan_admissions(free_slots: Sequence[int], queue_len: int) -> List[int]:
    """Slots that take a queued request this step (lowest slot first, FIFO requests)."""
    return sorted(free_slots)[: max(0, queue_len)]


class Engine:
    """Fixed ``max_batch`` sequence slots over one ``LMState``; see the module docstring."""

    def __init__(self, model: CSM, max_batch: int = 64, max_len: int = MAX_SEQ_LEN, sampler: Optional[SamplerSpec] = None,
                 context_cache: Optional[ContextCache] = None, solo_kernel: bool = False,
                 kv_prefix_cache: Optional[KVPrefixCache] = None):
        self.model, self.B = model, int(max_batch)
        self.spec = sampler if sampler is not None else SamplerSpec(temperature=0.0)
        self.solo_kernel = bool(solo_kernel)
        self.state = LMState(model, self.B, max_len=max_len, row_invariant=True)
        self.ncb = model.n_audio_codebooks
        self.cache = context_cache if context_cache is not None else ContextCache(n_audio_codebooks=self.ncb)
        # only prompts submitted with context rows (prefix_rows > 0) ever touch it; capacity=0 switches it off
        self.kv_cache = kv_prefix_cache if kv_prefix_cache is not None else KVPrefixCache()
        self.queue: Deque[Request] = deque()
        self.slots: List[Optional[Request]] = [None] * self.B
        self.finished: Dict[int, Request] = {}
        self._next_id = 0
        self._prev: Optional[torch.Tensor] = None          # device (B, 32) frame of the last step
        self._pending: Optional[Tuple[int, List[Optional[Request]]]] = None   # (mirror slot, occupants) of the last step
        self._host = [torch.empty((self.B, self.ncb), dtype=torch.int32).pin_memory() for _ in range(2)]
        self._events = [torch.cuda.Event() for _ in range(2)]
        self._flip = 0
        self.steps = self.mixed_steps = self.admissions = self.solo_steps = 0

    # ------------------------------------------------------------------ requests
    def build_prompt(self, text, speaker: int, context: Sequence[Segment]) -> Tuple[torch.Tensor, torch.Tensor]:
        """generation.py:108-121 with cached context tokenisation."""
        tok, mask, _ = self._build_prompt(text, speaker, context)
        return tok, mask

    def _build_prompt(self, text, speaker: int, context: Sequence[Segment]) -> Tuple[torch.Tensor, torch.Tensor, int]:
        toks, masks = [], []
        for seg in context:
            t, m = self.cache.segment_rows(seg)
            toks.append(t)
            masks.append(m)
        n_context = sum(int(t.shape[0]) for t in toks)
        t, m = tokenize_text_segment(text, speaker, n_audio_codebooks=self.ncb)
        toks.append(t)
        masks.append(m)
        return torch.cat(toks, 0).to(torch.int32), torch.cat(masks, 0), n_context

    def submit(self, text: Union[str, Sequence[int]], speaker: int, context: Optional[Sequence[Segment]] = None,
               max_audio_length_ms: float = 90_000) -> int:
        tok, mask, n_context = self._build_prompt(text, speaker, context or [])
        return self.submit_prompt(tok, mask, int(max_audio_length_ms / 80), prefix_rows=n_context)

    def submit_prompt(self, tokens: torch.Tensor, mask: torch.Tensor, max_frames: int, prefix_rows: int = 0) -> int:
        """``prefix_rows``: the first that many rows are context other requests are likely to share (``submit`` passes the
        rows of the context segments): their backbone KV is kept in / taken from the ``KVPrefixCache``."""
        _check_length(self.model, int(tokens.shape[0]), max_frames)           # generation.py:131-137
        if int(tokens.shape[0]) + max_frames + 1 > self.state.max_len:
            raise ValueError("request exceeds the KV pages reserved per slot (Engine(max_len=...))")
        if not 0 <= int(prefix_rows) < int(tokens.shape[0]):
            raise ValueError("prefix_rows must leave at least one prompt row")
        r = Request(self._next_id, tokens.to(torch.int32).cpu(), mask.cpu(), int(max_frames), int(prefix_rows))
        self._next_id += 1
        self.queue.append(r)
        return r.rid

    @property
    def active(self) -> int:
        return sum(r is not None for r in self.slots)

    # ------------------------------------------------------------------ one frame for every occupied slot
    def step(self) -> List[Request]:
        """Enqueues one frame-step, then retires what the previous step finished.  Returns newly finished requests."""
        st = self.state
        newly_done: List[Request] = []
        free = [b for b, r in enumerate(self.slots) if r is None]
        admit = plan_admissions(free, len(self.queue))
        if self.active == 0 and not admit:
            newly_done += self._drain()
            return newly_done
        busy = [b for b, r in enumerate(self.slots) if r is not None]
        if (self.solo_kernel and len(busy) == 1 and not admit and self._prev is not None and self.B > 1
                and st.slot_fused_supported(self.spec) and os.environ.get("CSMB_DISABLE_FUSED", "0") != "1"):
            # opt-in: one running sequence and nothing to admit: its frame through the batch-1 persistent kernel (3.2 ms)
            # instead of a chain step over all slots (6.7 ms); same Philox draws as the chain would use for this slot, but
            # CUDA-core summation order (see "Batch invariance" in the module docstring)
            self._park_idle_slots()
            frame = st.decode_frame_slot(busy[0], self._prev, self.spec)
            self.solo_steps += 1
        elif st.fast_supported(self.spec):
            # fused chain: admitted requests prefill all but their last prompt row now; that row is their input of this
            # step, in which everybody else decodes (csmb_decode_frame_fast_admit) — admission never stalls the batch
            if admit:
                reqs = []
                for b in admit:
                    r = self.queue.popleft()
                    r.slot = b
                    self.slots[b] = r
                    reqs.append(r)
                self._admit_on_chain(admit, reqs)
                self.admissions += len(admit)
            self._park_idle_slots()
            prev = self._prev if self._prev is not None else torch.zeros((self.B, self.ncb), device=st.device, dtype=torch.int32)
            frame = st.decode_frame_graphed(prev, self.spec)
            st.disarm_admission()
        elif admit or self._prev is None:
            frame = self._mixed_step(admit)
        else:
            self._park_idle_slots()
            frame = st.decode_frame_graphed(self._prev, self.spec)
        self.steps += 1
        for r in self.slots:
            if r is not None:
                r.issued += 1
        # publish this step's frame to the host (pinned, asynchronous) and look at the previous one
        slot = self._flip
        self._flip ^= 1
        self._host[slot].copy_(frame, non_blocking=True)
        self._events[slot].record(torch.cuda.current_stream(st.device))
        newly_done += self._drain()
        self._pending = (slot, list(self.slots))
        self._prev = frame
        return newly_done

    def _admit_on_chain(self, slots: List[int], reqs: List[Request]) -> None:
        """arm_admission with the context rows' KV served from / stored into the prefix cache."""
        st, kvc = self.state, self.kv_cache
        use = [kvc.capacity > 0 and r.prefix_rows > 0 for r in reqs]
        ver = int(getattr(self.model, "weights_version", 0))
        keys = [KVPrefixCache.key(r.tokens, r.mask, r.prefix_rows, ver) if u else None for r, u in zip(reqs, use)]
        known = []
        for b, r, k in zip(slots, reqs, keys):
            pages = kvc.get(k) if k is not None else None
            if pages is not None:
                st.import_kv_prefix(b, pages)
                known.append(r.prefix_rows)
            else:
                known.append(0)
        st.arm_admission(slots, [r.tokens for r in reqs], [r.mask for r in reqs], known_rows=known)
        for b, r, k, n in zip(slots, reqs, keys, known):
            if k is not None and n == 0:        # a miss: the rows were just prefilled into the slot's pages (stream ordered)
                kvc.put(k, st.export_kv_prefix(b, r.prefix_rows), r.prefix_rows)

    def _drain(self) -> List[Request]:
        done: List[Request] = []
        if self._pending is None:
            return done
        slot, occupants = self._pending
        self._pending = None
        self._events[slot].synchronize()
        host = self._host[slot]
        for b, r in enumerate(occupants):
            if r is None or r.done:
                continue
            row = host[b]
            if not bool(row.any()):                      # EOS: generation.py:151-152
                r.done = True
            else:
                r.frames.append(row.clone())
                if len(r.frames) >= r.max_frames:
                    r.done = True
            if r.done:
                done.append(r)
                self.finished[r.rid] = r
                if self.slots[b] is r:
                    self.slots[b] = None                 # the frame already in flight for this slot is discarded
        return done

    def _park_idle_slots(self) -> None:
        """Empty slots keep stepping inside the captured graph; pin them at position 0 so they never leave their pages."""
        st = self.state
        idle = [b for b, r in enumerate(self.slots) if r is None]
        if idle:
            idx = torch.tensor(idle, dtype=torch.long, device=st.device)
            st.pos.index_fill_(0, idx, 0)
            for b in idle:
                st.pos_host[b] = 0

    def _mixed_step(self, admit: List[int]) -> torch.Tensor:
        """One backbone pass in which admitted requests prefill their prompts and running sequences take their
        one-row step (generation.py:34-42 for both T > 1 and T = 1), then c0 + the depth loop for all slots."""
        st, ncb = self.state, self.ncb
        self.mixed_steps += 1
        prev_host = None
        if self._prev is not None:
            prev_host = self._prev.to("cpu")             # synchronises; admissions are rare next to frames
        for b in admit:
            r = self.queue.popleft()
            r.slot = b
            self.slots[b] = r
            st.pos_host[b] = 0
        rows, masks = [], []
        one = torch.cat([torch.ones((1, ncb), dtype=torch.bool), torch.zeros((1, 1), dtype=torch.bool)], 1)
        for b in range(self.B):
            r = self.slots[b]
            if r is not None and b in admit:
                rows.append(r.tokens)
                masks.append(r.mask)
            elif r is not None:
                rows.append(torch.cat([prev_host[b:b + 1], torch.zeros((1, 1), dtype=torch.int32)], 1))
                masks.append(one)
            else:                                          # idle slot: a masked-out dummy row at position 0
                st.pos_host[b] = 0
                rows.append(torch.zeros((1, ncb + 1), dtype=torch.int32))
                masks.append(torch.zeros((1, ncb + 1), dtype=torch.bool))
        st.prefill(rows, masks)
        frame = torch.zeros((self.B, ncb), device=st.device, dtype=torch.int32)
        st.sample_c0(frame, self.spec)
        st.depth_decode(frame, self.spec)
        return frame

    # ------------------------------------------------------------------ drivers
    def run(self) -> Dict[int, Request]:
        """Steps until the queue and all slots are empty; returns {request id: Request}."""
        while self.queue or self.active or self._pending is not None:
            self.step()
        self.state.check_status()
        return self.finished

    def tokens(self, rid: int) -> torch.Tensor:
        r = self.finished[rid]
        return torch.stack(r.frames) if r.frames else torch.zeros((0, self.ncb), dtype=torch.int32)

    def audio(self, rids: Sequence[int], to_host: bool = True) -> List[torch.Tensor]:
        """Mimi decode (tokenizers.py:148-150) of finished requests, batched; 1-D float32 tensors (on the host by default,
        views of one device tensor with ``to_host=False``)."""
        mimi = get_audio_tokenizer(self.ncb)
        toks = [self.tokens(r) for r in rids]
        fmax = max((int(t.shape[0]) for t in toks), default=0)
        if fmax == 0:
            return [torch.zeros((0,), dtype=torch.float32) for _ in toks]
        codes = torch.zeros((len(toks), self.ncb, fmax), dtype=torch.int32)
        for i, t in enumerate(toks):
            codes[i, :, : t.shape[0]] = t.t()
        audio = mimi.decode(codes.to(self.model.device))
        if to_host:
            # one device-to-host copy into pinned memory (torch's caching host allocator: the block is reused once the
            # caller drops the previous result); the returned tensors are views of that one buffer
            host = torch.empty(audio.shape, dtype=audio.dtype, pin_memory=True)
            host.copy_(audio, non_blocking=True)
            torch.cuda.current_stream(self.model.device).synchronize()
            return [host[i, 0, : 1920 * int(t.shape[0])] for i, t in enumerate(toks)]
        return [audio[i, 0, : 1920 * int(t.shape[0])] for i, t in enumerate(toks)]

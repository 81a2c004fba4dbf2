"""Conversation caches shared by the serving engine and by ``generate`` / ``stream_generate``.

The reference re-tokenises (Mimi-encodes) and re-prefills every context segment on every call
(``/root/reference/csm_mlx/generation.py:108-121``; its demo passes up to six earlier segments per turn,
``run_streaming_csm_mlx.py:102, 1060-1073``).  Both steps are pure functions of the segment / of the prompt rows and the
weights, so their results are kept:

* ``ContextCache``: (rows, mask) of a context segment's audio, keyed by the samples;
* ``KVPrefixCache``: the backbone KV pages of a prompt's leading (context) rows, keyed by those rows and the model's weights
  version.  The prompt pass is row-invariant (a row's KV entries depend on the rows before it in its own sequence only, never
  on how a call was batched), so a sequence that starts from copied pages generates the same tokens as one that prefilled them.
"""

from __future__ import annotations

import hashlib
from collections import OrderedDict
from typing import Optional, Tuple

import torch

from .segment import Segment
from .tokenizers import tokenize_audio, tokenize_text_segment


class ContextCache:
    """LRU cache of the (rows, mask) tokenisation of context segments, keyed by speaker, text and audio content."""

    def __init__(self, capacity: int = 64, n_audio_codebooks: int = 32):
        self.capacity, self.ncb = capacity, n_audio_codebooks
        self._audio: "OrderedDict[str, Tuple[torch.Tensor, torch.Tensor]]" = OrderedDict()
        self.hits = self.misses = 0

    @staticmethod
    def _key(audio: torch.Tensor) -> str:
        a = audio.detach().to("cpu", torch.float32).contiguous()
        return hashlib.sha1(a.numpy().tobytes()).hexdigest() + f":{a.numel()}"

    def audio_rows(self, audio: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        k = self._key(audio)
        if k in self._audio:
            self._audio.move_to_end(k)
            self.hits += 1
            return self._audio[k]
        self.misses += 1
        rows = tokenize_audio(audio, n_audio_codebooks=self.ncb)
        self._audio[k] = rows
        while len(self._audio) > self.capacity:
            self._audio.popitem(last=False)
        return rows

    def segment_rows(self, seg: Segment) -> Tuple[torch.Tensor, torch.Tensor]:
        """tokenize_segment (tokenizers.py:88-102) with the audio half served from the cache."""
        tt, tm = tokenize_text_segment(seg.text, seg.speaker, n_audio_codebooks=self.ncb)
        at, am = self.audio_rows(seg.audio)
        return torch.cat([tt, at], 0), torch.cat([tm, am], 0)


class KVPrefixCache:
    """LRU cache: sha1 of a prompt's leading rows -> (device KV pages of ``LMState.export_kv_prefix``, number of rows)."""

    def __init__(self, capacity: int = 8, max_bytes: int = 1 << 30):
        self.capacity, self.max_bytes = int(capacity), int(max_bytes)
        self._kv: "OrderedDict[str, Tuple[torch.Tensor, int]]" = OrderedDict()
        self.hits = self.misses = 0

    @staticmethod
    def key(tokens: torch.Tensor, mask: torch.Tensor, n_rows: int, weights_version: int = 0) -> str:
        """``weights_version``: ``CSM.weights_version`` — KV entries are a function of the rows AND the weights."""
        t = tokens[:n_rows].to("cpu", torch.int32).contiguous().numpy().tobytes()
        m = mask[:n_rows].to("cpu", torch.uint8).contiguous().numpy().tobytes()
        return hashlib.sha1(t + m).hexdigest() + f":{int(n_rows)}:{int(weights_version)}"

    @property
    def nbytes(self) -> int:
        return sum(int(p.numel()) * p.element_size() for p, _ in self._kv.values())

    def get(self, key: str) -> Optional[torch.Tensor]:
        hit = self._kv.get(key)
        if hit is None:
            self.misses += 1
            return None
        self._kv.move_to_end(key)
        self.hits += 1
        return hit[0]

    def put(self, key: str, pages: torch.Tensor, n_rows: int) -> None:
        if int(pages.numel()) * pages.element_size() > self.max_bytes:
            return
        self._kv[key] = (pages, int(n_rows))
        self._kv.move_to_end(key)
        while len(self._kv) > self.capacity or self.nbytes > self.max_bytes:
            self._kv.popitem(last=False)

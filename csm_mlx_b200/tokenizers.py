"""Frame assembly: text / audio -> (T, 33) token + mask rows.  Mirrors
``/root/reference/csm_mlx/tokenizers.py:14-102,148-150``.

Row layout (tokenizers.py:43-85): columns 0-31 audio codebooks, column 32 the text id; masks likewise.  A text
segment is one row per token of ``"[{speaker}]{text}"`` wrapped in BOS/EOS; an audio segment is one row per Mimi
frame plus one all-zero EOS row.  Rows are small host tensors; the GPU work (Mimi encode) is in ``mimi.py``.

No tokenizer / codec files exist offline (SURVEY.md hazard H7): ``get_text_tokenizer`` / ``get_audio_tokenizer``
try the reference's sources (HF hub ids in ``config.TOKENIZERS``) and otherwise require the caller to install
one with ``set_text_tokenizer`` / ``set_audio_tokenizer`` (tests and the benchmark install synthetic ones).
"""

from __future__ import annotations

from typing import List, Optional, Sequence, Tuple, Union

import torch

from .config import TOKENIZERS
from .segment import Segment

_text_tokenizer = None
_audio_tokenizers: dict = {}   # n_audio_codebooks -> Mimi (the reference caches per argument with functools.cache)


class SyntheticTextTokenizer:
    """Deterministic stand-in for the Llama-3.2 tokenizer: BOS + one pseudo-random id per UTF-8 byte pair + EOS.
    Only for synthetic benchmarks/tests — it does not produce meaningful Llama ids."""

    bos_token_id, eos_token_id = 128000, 128001

    def __init__(self, vocab: int = 128000):
        self.vocab = vocab

    def encode(self, text: str) -> List[int]:
        data = text.encode("utf-8")
        ids, h = [], 2166136261
        for i in range(0, len(data), 2):
            for byte in data[i:i + 2]:
                h = ((h ^ byte) * 16777619) & 0xFFFFFFFF
            ids.append(h % self.vocab)
        return [self.bos_token_id] + ids + [self.eos_token_id]


def set_text_tokenizer(tok) -> None:
    """Install the object whose ``encode(str) -> list[int]`` (BOS/EOS included) is used for text segments."""
    global _text_tokenizer
    _text_tokenizer = tok


def set_audio_tokenizer(mimi) -> None:
    """Install (or, with None, drop) the codec served by ``get_audio_tokenizer(mimi.n_q)``."""
    if mimi is None:
        _audio_tokenizers.clear()
    else:
        _audio_tokenizers[int(getattr(mimi, "n_q", 32))] = mimi


def get_text_tokenizer():
    """tokenizers.py:24-40: Llama-3.2 tokenizer with a ``BOS $A EOS`` post-processor."""
    global _text_tokenizer
    if _text_tokenizer is None:
        try:
            from tokenizers.processors import TemplateProcessing
            from transformers import AutoTokenizer

            tokenizer = AutoTokenizer.from_pretrained(TOKENIZERS["text"]["repo_id"])
            bos, eos = tokenizer.bos_token, tokenizer.eos_token
            tokenizer._tokenizer.post_processor = TemplateProcessing(
                single=f"{bos}:0 $A:0 {eos}:0",
                pair=f"{bos}:0 $A:0 {eos}:0 {bos}:1 $B:1 {eos}:1",
                special_tokens=[(f"{bos}", tokenizer.bos_token_id), (f"{eos}", tokenizer.eos_token_id)],
            )
            _text_tokenizer = tokenizer
        except Exception as e:  # offline image: no tokenizer files
            raise RuntimeError(
                f"text tokenizer {TOKENIZERS['text']['repo_id']} is not available offline ({type(e).__name__}); "
                "install one with csm_mlx.tokenizers.set_text_tokenizer(...)") from e
    return _text_tokenizer


def get_audio_tokenizer(n_audio_codebooks: int = 32):
    """tokenizers.py:14-21: the process-wide Mimi holding the codec weights, one per codebook count like the reference's
    ``@cache`` (a model with another ``n_audio_codebooks`` never gets a codec of the wrong depth)."""
    n_audio_codebooks = int(n_audio_codebooks)
    if n_audio_codebooks not in _audio_tokenizers:
        try:
            from huggingface_hub import hf_hub_download

            from .mimi import Mimi

            weight = hf_hub_download(**TOKENIZERS["audio"])
            _audio_tokenizers[n_audio_codebooks] = Mimi(n_audio_codebooks).load_pytorch_weights(weight)
        except Exception as e:
            have = sorted(_audio_tokenizers)
            raise RuntimeError(
                f"no Mimi codec with {n_audio_codebooks} codebooks: the weights {TOKENIZERS['audio']['repo_id']} are not "
                f"available offline ({type(e).__name__}) and the installed codecs have {have} codebooks; "
                "install one with csm_mlx.tokenizers.set_audio_tokenizer(...)") from e
    return _audio_tokenizers[n_audio_codebooks]


def tokenize_text_segment(text: Union[str, Sequence[int]], speaker: int, *, n_audio_codebooks: int = 32
                          ) -> Tuple[torch.Tensor, torch.Tensor]:
    """tokenizers.py:43-58.  Extension: ``text`` may already be a sequence of token ids (BOS/EOS included)."""
    if isinstance(text, str):
        ids = list(get_text_tokenizer().encode(f"[{speaker}]{text}"))
    else:
        ids = [int(t) for t in text]
    n = len(ids)
    frame = torch.zeros((n, n_audio_codebooks + 1), dtype=torch.int32)
    mask = torch.zeros((n, n_audio_codebooks + 1), dtype=torch.bool)
    frame[:, -1] = torch.tensor(ids, dtype=torch.int32)
    mask[:, -1] = True
    return frame, mask


def tokenize_audio(audio: torch.Tensor, *, n_audio_codebooks: int = 32) -> Tuple[torch.Tensor, torch.Tensor]:
    """tokenizers.py:61-85: Mimi-encode ``(N,)`` audio, append the all-zero EOS frame."""
    mimi = get_audio_tokenizer(n_audio_codebooks)
    a = torch.as_tensor(audio, dtype=torch.float32)
    codes = mimi.encode(a.reshape(1, 1, -1))[0].to("cpu", torch.int32)  # (K, F)
    codes = torch.cat([codes, torch.zeros((codes.shape[0], 1), dtype=torch.int32)], dim=1)
    n = codes.shape[1]
    frame = torch.zeros((n, n_audio_codebooks + 1), dtype=torch.int32)
    mask = torch.zeros((n, n_audio_codebooks + 1), dtype=torch.bool)
    frame[:, :-1] = codes.t()
    mask[:, :-1] = True
    return frame, mask


def tokenize_segment(segment: Segment, *, n_audio_codebooks: int = 32) -> Tuple[torch.Tensor, torch.Tensor]:
    """tokenizers.py:88-102 -> ((seq_len, 33) int32, (seq_len, 33) bool)."""
    tt, tm = tokenize_text_segment(segment.text, segment.speaker, n_audio_codebooks=n_audio_codebooks)
    at, am = tokenize_audio(segment.audio, n_audio_codebooks=n_audio_codebooks)
    return torch.cat([tt, at], 0), torch.cat([tm, am], 0)


def decode_audio(audio_tokens: torch.Tensor, *, n_audio_codebooks: int = 32) -> torch.Tensor:
    """tokenizers.py:148-150: (B,32,F) -> (B,1,1920*F)."""
    return get_audio_tokenizer(n_audio_codebooks).decode(audio_tokens)

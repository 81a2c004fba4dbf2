"""``Segment`` context record — mirrors ``/root/reference/csm_mlx/segment.py:12-46``.

Positional order ``Segment(speaker, text, audio=None, audio_path=None)`` matters: the reference CLI builds
``Segment(speaker, text, None, path)`` (cli/generate.py:186-189).  ``.audio`` lazily reads + resamples the file
on every access when only a path was given; raises ``ValueError`` if neither is available.
"""

from __future__ import annotations

from pathlib import Path
from typing import Optional

import torch

from .utils import read_audio

SAMPLING_RATE = 24000


class Segment:
    def __init__(self, speaker: int, text: str, audio: Optional[torch.Tensor] = None,
                 audio_path: Optional[Path] = None):
        self.speaker = speaker
        self.text = text
        self._audio = audio
        self.audio_path = audio_path

    # The reference is a dataclass whose hand-written __init__ overrides the generated one, so __post_init__
    # (segment.py:19-21) never runs there; the check effectively happens on first access (:30).
    @property
    def audio(self):
        if self._audio is not None:
            return self._audio
        elif self.audio_path is not None:
            return read_audio(self.audio_path, SAMPLING_RATE)
        raise ValueError("Neither 'audio' nor 'audio_path' is provided")

    @audio.setter
    def audio(self, value):
        self._audio = value

    def __repr__(self) -> str:
        return f"Segment(speaker={self.speaker!r}, text={self.text!r}, audio_path={self.audio_path!r})"

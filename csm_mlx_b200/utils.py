"""Audio file helpers — mirror ``/root/reference/csm_mlx/utils.py:9-27`` (``read_audio`` / ``write_audio``).

The reference uses ``audiofile`` + ``audresample`` (neither exists in this image).  This is host-side I/O, off the
hot path: PCM/float WAV through the standard library + numpy, linear-phase polyphase resampling via scipy.
"""

from __future__ import annotations

import wave
from pathlib import Path

import numpy as np
import torch


def _read_wav(path: str):
    with wave.open(path, "rb") as w:
        n_ch, width, rate, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
        raw = w.readframes(n)
    if width == 2:
        data = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif width == 4:
        data = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif width == 1:
        data = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    elif width == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v & 0x800000, v - 0x1000000, v)
        data = v.astype(np.float32) / 8388608.0
    else:
        raise ValueError(f"unsupported WAV sample width {width}")
    return data.reshape(-1, n_ch).T, rate  # (channels, samples)


def read_audio(filename: Path, sampling_rate: int) -> torch.Tensor:
    """file -> mono float32 ``(audio_length,)`` at ``sampling_rate`` (mean over channels, utils.py:16-19)."""
    signal, sr = _read_wav(str(filename))
    if sr != sampling_rate:
        from math import gcd

        from scipy.signal import resample_poly

        g = gcd(int(sr), int(sampling_rate))
        signal = resample_poly(signal, sampling_rate // g, sr // g, axis=1).astype(np.float32)
    return torch.from_numpy(np.ascontiguousarray(signal.mean(axis=0), dtype=np.float32))


def write_audio(array, filename: Path, sampling_rate: int) -> None:
    """float array ``(n,)`` or ``(channels, n)`` -> 16-bit PCM WAV (utils.py:24-27)."""
    a = array.detach().cpu().numpy() if isinstance(array, torch.Tensor) else np.asarray(array)
    a = np.atleast_2d(a.astype(np.float32))
    pcm = (np.clip(a, -1.0, 1.0) * 32767.0).round().astype("<i2")
    with wave.open(str(filename), "wb") as w:
        w.setnchannels(pcm.shape[0])
        w.setsampwidth(2)
        w.setframerate(int(sampling_rate))
        w.writeframes(pcm.T.tobytes())

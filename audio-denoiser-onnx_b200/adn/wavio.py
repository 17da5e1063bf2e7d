"""PCM16 wav file I/O for the demo CLI (the reference uses pydub / soundfile for this,
`GTCRN/Inference_GTCRN_ONNX.py:272,340`; neither is a dependency here -- stdlib `wave` only).
Outside the hot path."""
from __future__ import annotations

import wave

import numpy as np


def read_wav(path) -> tuple[np.ndarray, int]:
    """-> (int16 samples (channels, n), sample_rate).  PCM16 only (every file under the reference's Test_Examples/)."""
    with wave.open(str(path), "rb") as w:
        if w.getsampwidth() != 2:
            raise ValueError(f"{path}: only 16-bit PCM wav files are supported (sample width {w.getsampwidth()})")
        ch, sr, n = w.getnchannels(), w.getframerate(), w.getnframes()
        data = np.frombuffer(w.readframes(n), dtype="<i2")
    return np.ascontiguousarray(data.reshape(-1, ch).T), sr


def write_wav(path, samples: np.ndarray, sample_rate: int) -> None:
    """samples: int16 (n,) or (channels, n)."""
    a = np.asarray(samples)
    if a.dtype != np.int16:
        raise ValueError("write_wav expects int16 samples")
    if a.ndim == 1:
        a = a.reshape(1, -1)
    with wave.open(str(path), "wb") as w:
        w.setnchannels(a.shape[0])
        w.setsampwidth(2)
        w.setframerate(int(sample_rate))
        w.writeframes(np.ascontiguousarray(a.T).astype("<i2").tobytes())


def to_mono(samples: np.ndarray) -> np.ndarray:
    """(channels, n) int16 -> (n,) int16, channel mean like pydub's set_channels(1)."""
    a = np.asarray(samples)
    if a.ndim == 1 or a.shape[0] == 1:
        return a.reshape(-1)
    return (a.astype(np.int32).sum(axis=0) // a.shape[0]).astype(np.int16)

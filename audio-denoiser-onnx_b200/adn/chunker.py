"""Chunk scheduler: the host loop of `Inference_GTCRN_ONNX.py:276-333` (pad -> fixed
windows -> run -> concatenate -> trim), with the B200 difference that all windows of a
file (or of many files) are stacked into ONE batched run instead of a Python while-loop
of batch-1 runs.  Window/stride/padding arithmetic is the reference's, bit for bit."""
from __future__ import annotations

import numpy as np


def plan_windows(audio_len: int, in_len: int, out_len: int, same_rate: bool = True):
    """Returns (stride_step, num_windows, padded_len) -- Inference_GTCRN_ONNX.py:287-299."""
    stride = in_len
    if audio_len > in_len:
        if in_len != out_len and same_rate:
            stride = out_len                       # :289-290 overlap by in-out samples
        num = int(np.ceil((audio_len - in_len) / stride)) + 1
        total = (num - 1) * stride + in_len
    else:
        num, total = 1, in_len
    return stride, num, total


def split(audio: np.ndarray, in_len: int, out_len: int) -> tuple[np.ndarray, int]:
    """audio (N,) -> windows (num_windows, 1, in_len) (zero-padded tail), stride."""
    n = audio.shape[-1]
    stride, num, total = plan_windows(n, in_len, out_len)
    a = audio.reshape(-1)
    if total > n:
        a = np.concatenate((a, np.zeros(total - n, dtype=a.dtype)))
    idx = np.arange(num)[:, None] * stride + np.arange(in_len)[None, :]
    return np.ascontiguousarray(a[idx]).reshape(num, 1, in_len), stride


def denoise(session, audio: np.ndarray, max_batch: int = 4096) -> np.ndarray:
    """Whole-file drop-in for the reference's run section (:306-332): returns the
    concatenated output trimmed to the input length (`[:audio_len]`, :332)."""
    from .ort_shim import OrtValue

    i = session.get_inputs()[0]
    o = session.get_outputs()[0]
    in_len, out_len = i.shape[-1], o.shape[-1]
    audio = np.asarray(audio).reshape(-1)
    windows, _ = split(audio, in_len, out_len)
    outs = []
    for s in range(0, windows.shape[0], max_batch):
        w = windows[s:s + max_batch]
        vin = OrtValue.ortvalue_from_numpy(w)
        vout = OrtValue.ortvalue_from_numpy(np.zeros((w.shape[0], 1, out_len), dtype=_np_dtype(o.type)))
        b = session.io_binding()
        b.bind_ortvalue_input(i.name, vin)
        b.bind_ortvalue_output(o.name, vout)
        session.run_with_iobinding(b)
        outs.append(vout.numpy())
    return np.concatenate(outs, axis=0).reshape(-1)[: audio.shape[0]]


def _np_dtype(ort_type: str):
    if "int16" in ort_type:
        return np.int16
    if "float16" in ort_type:
        return np.float16
    return np.float32

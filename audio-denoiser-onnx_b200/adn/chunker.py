"""Chunk scheduler: the host loop of `Inference_GTCRN_ONNX.py:276-333` /
`Mel_Band_Roformer/Stereo/Inference_MelBandRoformer_ONNX.py:266-345` (pad -> fixed windows -> run ->
concatenate -> trim), with the B200 difference that all windows of a file (or of many files) are
stacked into ONE batched run instead of a Python while-loop of batch-1 runs.  Window / stride /
padding arithmetic is the reference's, bit for bit.

Batch fold (reference default for Mel-Band-Roformer, `Export_MelBandRoformer.py:46-51,645-650`):
the graph reshapes `(1, C, n*W)` into `n` independent windows `(n*C, 1, W)`, window-major /
channel-minor, and stitches them back.  Here the model already takes `(B, C, W)` windows, so
folding *is* the host-side split with stride W and a zero-padded tail (`:298-300`).
"""
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


def tail_pad(audio: np.ndarray, pad_amount: int, mode: str = "zeros", rng=None) -> np.ndarray:
    """audio (C, N) -> (C, N+pad).  'zeros' (GTCRN :291-298, folded Mel-Band :298-300), 'noise': RMS-matched gaussian
    tail of the un-folded Mel-Band script (:301-303, :309-311), or 'reflect': the signal mirrored about its last sample,
    a single sample repeated, nothing -> zeros (`pad_audio_tail_with_context`, H-GTCRN/Inference_H_GTCRN_ONNX.py:138-153,
    the un-folded H-GTCRN script)."""
    if pad_amount <= 0:
        return audio
    c = audio.shape[0]
    if mode == "zeros":
        block = np.zeros((c, pad_amount), dtype=audio.dtype)
    elif mode == "noise":
        rng = rng or np.random.default_rng()
        ref = audio[:, -pad_amount:] if audio.shape[1] > pad_amount else audio
        ref = ref.astype(np.float32)
        rms = np.sqrt(np.mean(ref * ref, dtype=np.float32), dtype=np.float32)
        block = (rms * rng.normal(loc=0.0, scale=1.0, size=(c, pad_amount))).astype(audio.dtype)
    elif mode == "reflect":
        n = audio.shape[1]
        if n == 0:
            block = np.zeros((c, pad_amount), dtype=audio.dtype)
        elif n == 1:
            block = np.repeat(audio[:, -1:], pad_amount, axis=-1)
        else:
            block = np.pad(audio, ((0, 0), (0, pad_amount)), mode="reflect")[:, n:]
    else:
        raise ValueError(f"unknown tail pad mode {mode!r}")
    return np.concatenate((audio, block), axis=-1)


def _rates(session) -> tuple[int, int, float]:
    """(in_sample_rate, out_sample_rate, input_to_output_scale) of the model file; equal rates when the keys are absent."""
    md = session.get_modelmeta().custom_metadata_map or {}
    in_sr, out_sr = int(md.get("in_sample_rate") or 0), int(md.get("out_sample_rate") or 0)
    if in_sr <= 0 or out_sr <= 0:
        in_sr = out_sr = 1
    scale = md.get("input_to_output_scale")
    return in_sr, out_sr, float(scale) if scale not in (None, "") else out_sr / in_sr


def split(audio: np.ndarray, in_len: int, out_len: int, tail: str = "zeros", rng=None, same_rate: bool = True) -> tuple[np.ndarray, int]:
    """audio (N,) or (C, N) -> windows (num_windows, C, in_len), stride."""
    a = np.asarray(audio)
    if a.ndim == 1:
        a = a.reshape(1, -1)
    n = a.shape[-1]
    stride, num, total = plan_windows(n, in_len, out_len, same_rate)
    a = tail_pad(a, total - n, tail, rng)
    idx = np.arange(num)[:, None] * stride + np.arange(in_len)[None, :]
    w = a[:, idx]                                   # (C, num, in_len)
    return np.ascontiguousarray(w.transpose(1, 0, 2)), stride


def match_channels(audio: np.ndarray, channels: int) -> np.ndarray:
    """Mono -> duplicated channels, extra channels dropped (Inference_MelBandRoformer_ONNX.py:273-287)."""
    a = np.asarray(audio)
    if a.ndim == 1:
        a = a.reshape(1, -1)
    if a.shape[0] < channels:
        a = np.concatenate((a, np.repeat(a[-1:], channels - a.shape[0], axis=0)), axis=0)
    elif a.shape[0] > channels:
        a = a[:channels]
    return a


def denoise(session, audio: np.ndarray, max_batch: int = 4096, tail: str = "zeros", rng=None) -> np.ndarray:
    """Whole-file drop-in for the reference's run section: returns the concatenated output trimmed
    to the input length at the output rate (`audio_len = int(audio_len * OUT / IN)`, `[:audio_len]`, GTCRN :303, :332 /
    Mel-Band :345).  The overlap stride of a model whose output window is shorter than its input window applies only when
    the two sample rates are equal (:288-290).  audio (N,) -> (N',) for mono models, (C, N) -> (C, N') otherwise."""
    from .ort_shim import OrtValue

    i = session.get_inputs()[0]
    o = session.get_outputs()[0]
    chans, in_len, out_len = i.shape[-2], i.shape[-1], o.shape[-1]
    mono_in = np.asarray(audio).ndim == 1
    a = match_channels(audio, chans)
    n = a.shape[-1]
    in_sr, out_sr, _ = _rates(session)
    windows, _ = split(a, in_len, out_len, tail, rng, same_rate=in_sr == out_sr)
    n_out = int(n * out_sr / in_sr)
    outs = []
    for s in range(0, windows.shape[0], max_batch):
        w = np.ascontiguousarray(windows[s:s + max_batch])
        vin = OrtValue.ortvalue_from_numpy(w)
        vout = OrtValue.ortvalue_from_numpy(np.zeros((w.shape[0], o.shape[-2], out_len), dtype=_np_dtype(o.type)))
        b = session.io_binding()
        b.bind_ortvalue_input(i.name, vin)
        b.bind_ortvalue_output(o.name, vout)
        session.run_with_iobinding(b)
        outs.append(vout.numpy())
    y = np.concatenate(outs, axis=0)                # (num, C, out_len)
    y = y.transpose(1, 0, 2).reshape(y.shape[1], -1)[:, :n_out]
    return y.reshape(-1) if (mono_in and y.shape[0] == 1) else y


def separate(session, audio: np.ndarray, pad_head: int | None = None, max_batch: int = 4096, tail: str = "zeros",
             rng=None) -> list[np.ndarray]:
    """Whole-file drop-in for the run section of `MossFormer2_SS_16K/Inference_MossFormer_SS_ONNX.py:269-340`:
    `pad_head` zeros are prepended (`:273`; default: the model file's `pad_head` metadata key), the padded signal is cut
    into fixed windows (stride = window, `:286-305`), ALL windows run as one batch with every output bound
    (`:312-317`), and each output is concatenated and trimmed to `[round(pad_head * s) : round(len(padded audio) * s)]`,
    s = the model file's input_to_output_scale (`:308-309`, `:339-340`).
    Works for any number of outputs (one list entry per `session.get_outputs()` element).  tail: 'zeros' (fold mode)
    or 'noise' (the un-folded script's RMS-matched gaussian tail)."""
    from .ort_shim import OrtValue

    i = session.get_inputs()[0]
    outs_meta = session.get_outputs()
    in_len, out_len = i.shape[-1], outs_meta[0].shape[-1]
    if pad_head is None:
        pad_head = int(session.get_modelmeta().custom_metadata_map.get("pad_head", "0"))
    a = np.asarray(audio).reshape(-1)
    a = np.concatenate([np.zeros(pad_head, dtype=a.dtype), a])
    n = a.shape[0]
    stride, num, total = plan_windows(n, in_len, in_len)            # SS strides by the input window (:285)
    a = tail_pad(a.reshape(1, -1), total - n, tail, rng)
    idx = np.arange(num)[:, None] * stride + np.arange(in_len)[None, :]
    windows = np.ascontiguousarray(a[:, idx].transpose(1, 0, 2))     # (num, 1, in_len)
    results = [[] for _ in outs_meta]
    for s in range(0, num, max_batch):
        w = np.ascontiguousarray(windows[s:s + max_batch])
        b = session.io_binding()
        b.bind_ortvalue_input(i.name, OrtValue.ortvalue_from_numpy(w))
        vouts = []
        for o in outs_meta:
            v = OrtValue.ortvalue_from_numpy(np.zeros((w.shape[0], o.shape[-2], out_len), dtype=_np_dtype(o.type)))
            b.bind_ortvalue_output(o.name, v)
            vouts.append(v)
        session.run_with_iobinding(b)
        for r, v in zip(results, vouts):
            r.append(v.numpy())
    # windows whose output is shorter than their input (length not 16 + 8k) are concatenated as produced, like the script
    scale = _rates(session)[2]
    lo, hi = int(round(pad_head * scale)), int(round(n * scale))
    return [np.concatenate(r, axis=0).reshape(-1)[lo:hi] for r in results]


def _np_dtype(ort_type: str):
    if "int16" in ort_type:
        return np.int16
    if "float16" in ort_type:
        return np.float16
    return np.float32

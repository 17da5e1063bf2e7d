"""Front and back ends of the wrappers whose backbone is not part of this library (ZipEnhancer,
MossFormerGAN-SE-16K, MossFormer2-SS-16K) and the linear resampler, on the GPU through the C ABI.

Each class mirrors the head / tail of the reference wrapper's `forward`:
`ZipEnds`  <- ZipEnhancer/Export_ZipEnhancer.py:818-850 and :880-926,
`GanEnds`  <- MossFormerGAN_SE_16K/Export_MossFormer_SE.py:539-590 and :863-897,
`SsEnds`   <- MossFormer2_SS_16K/Export_MossFormer2_SS_16K.py:403-423 and :625-660.
torch tensors are only containers for device memory.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib, stft_tables
from .stft_op import StftOp

ZIPENHANCER, MOSSFORMERGAN, MOSSFORMER2_SS = 1, 2, 3
_DT = {torch.float32: 0, torch.int16: 1, torch.float16: 2}
_OUT = {"F32": (0, torch.float32), "INT16": (1, torch.int16), "F16": (2, torch.float16)}


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _st(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _rows(x):
    assert x.is_cuda and x.is_contiguous() and x.dtype in _DT, "contiguous CUDA tensor of f32 / int16 / f16 expected"
    return x.numel() // x.shape[-1], x.shape[-1]


def resample_linear(x: torch.Tensor, size: int | None = None, scale_factor: float | None = None) -> torch.Tensor:
    """`F.interpolate(x.float(), size= | scale_factor=, mode='linear', align_corners=False)`."""
    rows, L = _rows(x)
    if (size is None) == (scale_factor is None):
        raise ValueError("exactly one of size / scale_factor")
    n = int(size) if size is not None else int(math.floor(float(L) * float(scale_factor)))
    out = torch.empty(x.shape[:-1] + (n,), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().adn_resample_linear(_p(x), _DT[x.dtype], _p(out), rows, L, n,
                                              float(scale_factor) if scale_factor is not None else 0.0, _st(x)),
               None, "adn_resample_linear")
    return out


def rms_normalize(x: torch.Tensor, pre_scale: float = 1.0, eps: float = 1e-6, pad_to: int | None = None):
    rows, L = _rows(x)
    n = pad_to or L
    out = torch.empty(x.shape[:-1] + (n,), dtype=torch.float32, device=x.device)
    nf = torch.empty(x.shape[:-1] + (1,), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().adn_rms_normalize(_p(x), _DT[x.dtype], pre_scale, eps, _p(out), _p(nf), rows, L, n, _st(x)),
               None, "adn_rms_normalize")
    return out, nf


class _Ends:
    family = 0
    geometry = ""

    def __init__(self, length: int, in_dtype: str = "INT16", out_dtype: str = "INT16", device_id: int = 0):
        g = stft_tables.GEOMETRY[self.geometry]
        self.length, self.in_dtype, self.out_dtype = length, in_dtype, out_dtype
        self.padded = length + (g.hop - length % g.hop) % g.hop if self.family == MOSSFORMERGAN else length
        self.stft = StftOp(g, self.padded, device_id)
        self.fbins, self.frames = g.fbins, self.stft.n_frames

    def _condition(self, wave, gain, group, rows_shape):
        code, tdt = _OUT[self.out_dtype]
        rows, src = _rows(wave)
        out = torch.empty(rows_shape + (self.length,), dtype=tdt, device=wave.device)
        _lib.check(_lib.lib().adn_condition_output(self.family, _p(wave), src, _p(gain), group, _p(out), code, rows,
                                                   self.length, _st(wave)), None, "adn_condition_output")
        return out

    def close(self):
        self.stft.close()


class ZipEnds(_Ends):
    family, geometry = ZIPENHANCER, "zipenhancer"

    def analyse(self, audio: torch.Tensor):
        """audio (B,1,L) -> (x (B,2,T,F) [compressed magnitude, phase], norm_factor (B,1,1))."""
        a, nf = rms_normalize(audio, 1.0 if self.in_dtype == "INT16" else 32768.0)
        spec = self.stft.forward(a)
        B = audio.shape[0]
        feat = torch.empty((B, 2, self.frames, self.fbins), dtype=torch.float32, device=audio.device)
        _lib.check(_lib.lib().adn_spec_features(self.family, _p(spec), _p(feat), None, B, self.fbins, self.frames, _st(spec)),
                   None, "adn_spec_features")
        return feat, nf

    def synthesise(self, mx: torch.Tensor, phase_ri: torch.Tensor, nf: torch.Tensor):
        """mx (B,1,T,F), phase_ri (B,2,T,F) -> audio (B,1,L) in out_dtype."""
        B = mx.shape[0]
        spec = torch.empty((B, 2 * self.fbins, self.frames), dtype=torch.float32, device=mx.device)
        _lib.check(_lib.lib().adn_spec_recombine(self.family, _p(mx.contiguous()), _p(phase_ri.contiguous()), None, _p(spec),
                                                 B, self.fbins, self.frames, _st(mx)), None, "adn_spec_recombine")
        return self._condition(self.stft.inverse(spec), nf.contiguous(), 1, (B, 1))


class GanEnds(_Ends):
    family, geometry = MOSSFORMERGAN, "mossformergan_se_16k"

    def analyse(self, audio: torch.Tensor):
        """audio (B,1,L) -> (x (B,3,T,F), compressed complex spectrum (B,2,F,T), norm_factor)."""
        a, nf = rms_normalize(audio, 1.0 if self.in_dtype == "INT16" else 32768.0, pad_to=self.padded)
        spec = self.stft.forward(a)
        B = audio.shape[0]
        feat = torch.empty((B, 3, self.frames, self.fbins), dtype=torch.float32, device=audio.device)
        keep = torch.empty((B, 2, self.fbins, self.frames), dtype=torch.float32, device=audio.device)
        _lib.check(_lib.lib().adn_spec_features(self.family, _p(spec), _p(feat), _p(keep), B, self.fbins, self.frames,
                                                _st(spec)), None, "adn_spec_features")
        return feat, keep, nf

    def synthesise(self, mask: torch.Tensor, complex_out: torch.Tensor, keep: torch.Tensor, nf: torch.Tensor):
        """mask (B,F,T), complex_out (B,2,F,T) -> audio (B,1,L) in out_dtype."""
        B = mask.shape[0]
        spec = torch.empty((B, 2 * self.fbins, self.frames), dtype=torch.float32, device=mask.device)
        _lib.check(_lib.lib().adn_spec_recombine(self.family, _p(mask.contiguous()), _p(complex_out.contiguous()), _p(keep),
                                                 _p(spec), B, self.fbins, self.frames, _st(mask)), None, "adn_spec_recombine")
        return self._condition(self.stft.inverse(spec), nf.contiguous(), 1, (B, 1))


class SsEnds:
    """MossFormer2-SS has learned encoder / decoder convolutions instead of an STFT; its ends are the two-stage RMS
    normalisation and the per-speaker gain restore."""
    family = MOSSFORMER2_SS
    TARGET = float(10.0 ** (-25.0 / 20.0))

    def __init__(self, length: int, out_dtype: str = "INT16", num_spks: int = 2):
        self.length, self.out_dtype, self.num_spks = length, out_dtype, num_spks

    def analyse(self, audio: torch.Tensor, eps: float = 1e-6):
        rows, L = _rows(audio)
        out = torch.empty(audio.shape, dtype=torch.float32, device=audio.device)
        rms_in = torch.empty((audio.shape[0], 1, 1), dtype=torch.float32, device=audio.device)
        _lib.check(_lib.lib().adn_two_stage_rms(_p(audio), _DT[audio.dtype], self.TARGET, eps, _p(out), _p(rms_in), rows, L,
                                                _st(audio)), None, "adn_two_stage_rms")
        return out, rms_in

    def synthesise(self, wav: torch.Tensor, rms_in: torch.Tensor):
        """wav (B,spks,L) fp32 decoder output -> (B,spks,L) in out_dtype."""
        code, tdt = _OUT[self.out_dtype]
        rows, src = _rows(wav)
        out = torch.empty(wav.shape[:-1] + (self.length,), dtype=tdt, device=wav.device)
        _lib.check(_lib.lib().adn_condition_output(self.family, _p(wav), src, _p(rms_in.contiguous()), self.num_spks, _p(out),
                                                   code, rows, self.length, _st(wav)), None, "adn_condition_output")
        return out

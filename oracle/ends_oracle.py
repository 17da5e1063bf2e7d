"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the waveform conditioning / spectral feature /
recombine / output-conditioning steps that sit either side of the backbones that are NOT built
(ZipEnhancer, MossFormerGAN-SE-16K, MossFormer2-SS-16K), plus the linear resampler every wrapper
shares.  SURVEY.md 8 rows a2, a4, a10, a12 and f-2.

PARITY UNPINNED BY EXECUTION: these statements live inside wrapper `forward`s whose constructors
need the un-vendored `modelscope` / `clearvoice` packages, so the reference modules cannot be
instantiated here.  Each function restates the cited lines one to one; the STFT/ISTFT they call
(`stft_oracle`) and `torch.nn.functional.interpolate` (the reference's own resampler call) are
pinned / are the reference call itself.

Only `tests/` may import this file.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from stft_oracle import SPECS, istft_packed, stft_packed

INV_INT16 = float(1.0 / 32768.0)


def resample_linear(x: torch.Tensor, size: int | None = None, scale_factor: float | None = None) -> torch.Tensor:
    """`F.interpolate(mode='linear', align_corners=False)` as every wrapper calls it
    (GTCRN/Export_GTCRN.py:638-654 with scale_factor; Export_ZipEnhancer.py:826-832 with size)."""
    return F.interpolate(x.float(), size=size, scale_factor=scale_factor, mode="linear", align_corners=False)


def resample_linear_explicit(x: torch.Tensor, size: int | None = None, scale_factor: float | None = None) -> torch.Tensor:
    """The same operator written out (ATen upsample_linear1d, align_corners=False): source index
    max(r*(i+0.5)-0.5, 0), r = 1/scale_factor when a scale factor is given else L_in/L_out; used to
    document the arithmetic the CUDA kernel follows."""
    L = x.shape[-1]
    if size is None:
        size = int(torch.floor(torch.tensor(float(L) * scale_factor, dtype=torch.float64)).item())
        r = torch.tensor(1.0 / scale_factor, dtype=torch.float64).float()
    else:
        r = torch.tensor(float(L), dtype=torch.float32) / float(size)
    if size == L and scale_factor is None:
        return x.float().clone()
    i = torch.arange(size, dtype=torch.float32)
    # ATen is built with FMA contraction: r*(i+0.5)-0.5 is ONE rounding (matters: ulp(16000) = 1e-3)
    src = torch.clamp((r.double() * (i + 0.5).double() - 0.5).float(), min=0.0)
    i0 = torch.clamp(src.floor().long(), max=L - 1)
    lam = torch.clamp(src - i0.float(), 0.0, 1.0)
    i1 = i0 + (i0 < L - 1).long()
    xf = x.float()
    return (1.0 - lam) * xf[..., i0] + lam * xf[..., i1]


# ----------------------------------------------------------------------------- ZipEnhancer
def zip_front(audio: torch.Tensor, in_dtype: str = "INT16"):
    """Export_ZipEnhancer.py:819-821, 839-850: lift float input to int16 amplitude, per-window RMS
    normalisation, STFT (400/100 hann, reflect), compressed magnitude + phase.
    Returns (x (B,2,T,F), norm_factor (B,1,1))."""
    a = audio.float()
    if "int" not in in_dtype.lower():
        a = a * 32768.0
    nf = torch.sqrt(torch.mean(a * a, dim=-1, keepdim=True) + 1e-6)
    a = a / nf
    spec = stft_packed(SPECS["zipenhancer"], a)
    fb = spec.shape[1] // 2
    re, im = spec[:, :fb], spec[:, fb:]
    mag = torch.pow(re * re + im * im + 1e-9, 0.3 * 0.5)
    pha = torch.atan2(im, re + 1e-5)
    return torch.stack((mag, pha), dim=1).transpose(2, 3).contiguous(), nf


def zip_back(mx: torch.Tensor, phase_ri: torch.Tensor, nf: torch.Tensor, length: int, out_dtype: str = "INT16"):
    """Export_ZipEnhancer.py:882-926: mx (B,1,T,F) mask-decoder output, phase_ri (B,2,T,F) rectangular
    phase; magnitude decompress, unit phase vector with the zero-phase guard, ISTFT (multiply by the
    reciprocal window sum), trim, x norm_factor, NaN/Inf handling, output dtype."""
    mag = torch.pow(F.relu(mx), 1.0 / 0.3).transpose(2, 3)
    ri = phase_ri.transpose(2, 3)
    nrm = torch.linalg.vector_norm(ri, ord=2, dim=1, keepdim=True)
    has = nrm > 0.0
    unit = torch.tensor([1.0, 0.0]).view(1, 2, 1, 1)
    ri = torch.where(has, ri, unit)
    nrm = torch.where(has, nrm, torch.ones_like(nrm))
    ri = ri * (mag / nrm)
    b, _, f, t = ri.shape
    y = istft_packed(SPECS["zipenhancer"], ri.reshape(b, 2 * f, t))[..., :length]
    y = y * nf
    if "int" in out_dtype.lower():
        y = torch.where(torch.isnan(y), torch.zeros_like(y), y)
        return y.clamp(min=-32768.0, max=32767.0).to(torch.int16)
    y = torch.nan_to_num(y, nan=0.0, posinf=32767.0, neginf=-32768.0) * INV_INT16
    return y if "32" in out_dtype else y.to(torch.float16)


# ----------------------------------------------------------------------------- MossFormerGAN-SE-16K
def gan_front(audio: torch.Tensor, in_dtype: str = "INT16"):
    """MossFormerGAN_SE_16K/Export_MossFormer_SE.py:539-541, 564-586: RMS normalisation, wrap-around
    tail pad to a hop multiple, STFT (400/100 hamming), power-law compression.
    Returns (x (B,3,T,F), complex_compress (B,2,F,T), norm_factor)."""
    a = audio.float()
    if "int" not in in_dtype.lower():
        a = a * 32768.0
    L = a.shape[-1]
    nf = torch.sqrt(torch.mean(a * a, dim=-1, keepdim=True) + 1e-6)
    a = a / nf
    pad = (100 - L % 100) % 100
    if pad:
        a = torch.cat([a, a[..., :pad]], dim=-1)
    spec = stft_packed(SPECS["mossformergan_se_16k"], a)
    b, f2, t = spec.shape
    cplx = spec.reshape(b, 2, f2 // 2, t)
    power = (cplx * cplx).sum(dim=1)
    mag_c = torch.pow(power, 0.15)
    scale = torch.pow(power.clamp_min(torch.finfo(torch.float32).tiny), 0.15 - 0.5)
    cc = cplx * scale.unsqueeze(1)
    x = torch.cat((mag_c.unsqueeze(1), cc), dim=1).transpose(-1, -2).contiguous()
    return x, cc, nf


def gan_back(mask: torch.Tensor, complex_out: torch.Tensor, cc: torch.Tensor, nf: torch.Tensor, length: int,
             out_dtype: str = "INT16"):
    """:863-897: mask (B,F,T) x compressed spectrum + complex branch (B,2,F,T), power-law
    decompression, ISTFT (divide by the window sum), trim, x norm_factor, output dtype."""
    fin = mask.unsqueeze(1) * cc + complex_out
    factor = torch.pow((fin * fin).sum(dim=1), float(0.5 / 0.3) - 0.5)
    fin = fin * factor.unsqueeze(1)
    b, _, f, t = fin.shape
    y = istft_packed(SPECS["mossformergan_se_16k"], fin.reshape(b, 2 * f, t))[..., :length]
    y = y * nf
    if "int" in out_dtype.lower():
        return y.clamp(min=-32768.0, max=32767.0).to(torch.int16)
    y = y * INV_INT16
    return y if "32" in out_dtype else y.to(torch.float16)


# ----------------------------------------------------------------------------- MossFormer2-SS-16K
SS_NORM = float(10.0 ** (-25.0 / 20.0))


def ss_front(audio: torch.Tensor, eps: float = 1e-6):
    """MossFormer2_SS_16K/Export_MossFormer2_SS_16K.py:403-423 `norm_audio`: two-stage RMS
    normalisation per window (whole-window RMS, then RMS of the above-average-power samples).
    audio (B,1,L) raw PCM amplitude -> (x (B,1,L), rms_in (B,1,1))."""
    x = audio.float() * INV_INT16
    p = x * x
    avg = p.mean(dim=(1, 2), keepdim=True)
    rms = torch.sqrt(avg)
    s1 = SS_NORM / (rms + eps)
    hot = (p > avg).to(p.dtype)
    high = torch.sqrt((p * hot).sum(dim=(1, 2), keepdim=True) / hot.sum(dim=(1, 2), keepdim=True).clamp(min=1.0))
    s2 = SS_NORM / (high * s1 + eps)
    y = (x * s1) * s2
    g = s1 * s2
    undo = 1.0 / (g + eps)
    return y, rms * g * undo * 32767.0


def ss_back(wav: torch.Tensor, rms_in: torch.Tensor, out_dtype: str = "INT16"):
    """:625-660: wav (B,spks,L) decoder output; per-speaker RMS gain restore with the silent-window
    guard, output dtype (int32 staging for int16)."""
    rms_out = torch.sqrt((wav * wav).mean(dim=2, keepdim=True))
    gain = torch.where(rms_out > 0.0, rms_in / rms_out, torch.zeros_like(rms_out))
    y = wav * gain
    if "int" in out_dtype.lower():
        return y.to(torch.int32).clamp(min=-32768, max=32767).to(torch.int16)
    y = y * INV_INT16
    return y.to(torch.float16) if "16" in out_dtype else y

"""Host-side STFT / ISTFT tables.

The reference's DFT bases are *data*, not math: they
are produced by evaluating cos/sin of the unreduced fp32 argument `(2*pi/N)*f*t`
(`GTCRN/STFT_Process.py:213-251`).  The ONNX graph carries them as initializers; here
they are generated once on the host with the very same torch expressions, stored in the
model blob and uploaded.  They must never be recomputed on the device or with numpy trig.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class StftGeometry:
    nfft: int
    win_length: int
    hop: int
    window_type: str            # hann_sqrt | hann | hamming | hamming_sym
    center: bool = True
    pad_mode: str = "reflect"   # reflect | constant
    norm: str = "divide"        # divide | multiply (ZipEnhancer stores the reciprocal)

    @property
    def fbins(self) -> int:
        return self.nfft // 2 + 1

    def n_frames(self, length: int) -> int:
        return length // self.hop + 1 if self.center else (length - self.nfft) // self.hop + 1

    def out_length(self, n_frames: int) -> int:
        raw = self.nfft + self.hop * (n_frames - 1)
        return raw - 2 * (self.nfft // 2) if self.center else raw


# per-model geometry (reference constructor calls)
GEOMETRY = {
    "gtcrn": StftGeometry(512, 512, 256, "hann_sqrt", True, "reflect", "divide"),
    "zipenhancer": StftGeometry(400, 400, 100, "hann", True, "reflect", "multiply"),
    "mossformer2_se_48k": StftGeometry(1920, 1920, 384, "hamming_sym", False, "constant", "divide"),
    "mel_band_roformer": StftGeometry(2048, 2048, 441, "hann", True, "reflect", "divide"),
    "mossformergan_se_16k": StftGeometry(400, 400, 100, "hamming", True, "reflect", "divide"),
    "h_gtcrn": StftGeometry(512, 512, 256, "hann", True, "reflect", "divide"),          # H-GTCRN/Export_H_GTCRN.py:34-39
}


def window(g: StftGeometry) -> torch.Tensor:
    L = g.win_length
    if g.window_type == "hann_sqrt":
        w = torch.hann_window(L, periodic=True).pow(0.5)
    elif g.window_type == "hann":
        w = torch.hann_window(L, periodic=True)
    elif g.window_type == "hamming":
        w = torch.hamming_window(L, periodic=True)
    elif g.window_type == "hamming_sym":
        w = torch.hamming_window(L, periodic=False)
    else:
        raise ValueError(f"unknown window_type {g.window_type!r}")
    w = w.float()
    if L < g.nfft:
        left = (g.nfft - L) // 2
        w = torch.cat([torch.zeros(left), w, torch.zeros(g.nfft - L - left)])
    elif L > g.nfft:
        s = (L - g.nfft) // 2
        w = w[s:s + g.nfft]
    return w


def _omega(g: StftGeometry) -> torch.Tensor:
    factor = 2.0 * torch.pi / g.nfft
    t = torch.arange(g.nfft, dtype=torch.float32).unsqueeze(0)
    f = torch.arange(g.fbins, dtype=torch.float32).unsqueeze(1)
    return factor * f * t


def forward_basis(g: StftGeometry, input_scale: float = 1.0) -> torch.Tensor:
    """(2F, nfft): rows [cos*w ; -sin*w] == STFT_Process.stft_kernel.squeeze(1)."""
    om = _omega(g)
    w = (window(g) * input_scale).unsqueeze(0)
    return torch.cat([torch.cos(om) * w, -torch.sin(om) * w], dim=0).contiguous()


def inverse_basis(g: StftGeometry) -> torch.Tensor:
    """(2F, nfft) == STFT_Process.inverse_kernel.squeeze(1)."""
    om = _omega(g)
    scale = 2.0 * torch.ones(g.fbins, 1)
    scale[0] = 1.0
    if g.nfft % 2 == 0:
        scale[g.fbins - 1] = 1.0
    inv_n = 1.0 / g.nfft
    w = window(g).unsqueeze(0)
    re = (scale * torch.cos(om) * inv_n) * w
    im = (scale * -torch.sin(om) * inv_n) * w
    return torch.cat([re, im], dim=0).contiguous()


def norm_table(g: StftGeometry, n_frames: int) -> torch.Tensor:
    """(L_out,) overlap-added w^2 over the kept range; its fp32 reciprocal when the model
    multiplies (ZipEnhancer/STFT_Process.py:243-248)."""
    w2 = window(g).square().reshape(1, 1, -1)
    ws = F.conv_transpose1d(torch.ones(1, 1, n_frames), w2, stride=g.hop).reshape(-1)
    if g.center:
        half = g.nfft // 2
        ws = ws[half:ws.numel() - half]
    ws = ws.contiguous()
    return (1.0 / ws) if g.norm == "multiply" else ws

"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's STFT / ISTFT.

Restates `STFT_Process` (reference `GTCRN/STFT_Process.py:129-361` and the per-model
variants listed in SURVEY.md A.1) as plain functions over torch CPU tensors.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline leg may import it.

Pinned (tests/test_oracle_pinning.py) against
  * the reference's own modules executed from /root/reference (container only), and
  * `torch.stft` / `torch.istft`, which is the reference's own known-answer check
    (`GTCRN/STFT_Process.py:384-455`, seed 1234), and
  * the committed fixtures in tests/golden/.

Key fact: the reference evaluates cos/sin of the *unreduced* fp32
argument `(2*pi/N) * f * t`, so its DFT basis is a fixed, slightly inexact matrix.  The
oracle reproduces exactly those expressions; the basis is data, not math.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class StftSpec:
    """Geometry of one STFT/ISTFT pair (metadata keys nfft/window_length/hop_length/
    window_type/center_pad/pad_mode, reference audio_onnx_metadata.py:191-197)."""
    nfft: int
    win_length: int
    hop: int
    window_type: str          # hann_sqrt | hann | hamming | hamming_sym
    center: bool = True
    pad_mode: str = "reflect"  # reflect | constant
    norm: str = "divide"       # divide (GTCRN/SE/MBR/GAN) | multiply (ZipEnhancer reciprocal)

    @property
    def fbins(self) -> int:
        return self.nfft // 2 + 1

    def n_frames(self, length: int) -> int:
        # GTCRN/STFT_Process.py:73-76
        if self.center:
            return length // self.hop + 1
        return (length - self.nfft) // self.hop + 1

    def out_length(self, n_frames: int) -> int:
        # GTCRN/STFT_Process.py:170-176
        raw = self.nfft + self.hop * (n_frames - 1)
        return raw - 2 * (self.nfft // 2) if self.center else raw


SPECS = {
    # SURVEY.md A.1
    "gtcrn": StftSpec(512, 512, 256, "hann_sqrt", True, "reflect", "divide"),
    "zipenhancer": StftSpec(400, 400, 100, "hann", True, "reflect", "multiply"),
    "mossformer2_se_48k": StftSpec(1920, 1920, 384, "hamming_sym", False, "constant", "divide"),
    "mel_band_roformer": StftSpec(2048, 2048, 441, "hann", True, "reflect", "divide"),
    "mossformergan_se_16k": StftSpec(400, 400, 100, "hamming", True, "reflect", "divide"),
}


def make_window(spec: StftSpec) -> torch.Tensor:
    """`create_padded_window` (GTCRN/STFT_Process.py:100-113) with the per-model window
    registries (`:88-96`; MossFormer2_SE_48K/STFT_Process.py:92 uses periodic=False)."""
    L = spec.win_length
    if spec.window_type == "hann_sqrt":
        w = torch.hann_window(L, periodic=True).pow(0.5)
    elif spec.window_type == "hann":
        w = torch.hann_window(L, periodic=True)
    elif spec.window_type == "hamming":
        w = torch.hamming_window(L, periodic=True)
    elif spec.window_type == "hamming_sym":
        w = torch.hamming_window(L, periodic=False)
    else:
        raise ValueError(spec.window_type)
    w = w.float()
    n = spec.nfft
    if L == n:
        return w
    if L < n:
        left = (n - L) // 2
        return torch.cat([torch.zeros(left), w, torch.zeros(n - L - left)])
    s = (L - n) // 2
    return w[s:s + n]


def forward_basis(spec: StftSpec, input_scale: float = 1.0) -> torch.Tensor:
    """(2F, nfft) rows [cos*w ; -sin*w] -- GTCRN/STFT_Process.py:213-227."""
    n = spec.nfft
    omega_factor = 2.0 * torch.pi / n
    t = torch.arange(n, dtype=torch.float32).unsqueeze(0)
    f = torch.arange(spec.fbins, dtype=torch.float32).unsqueeze(1)
    omega = omega_factor * f * t
    w = make_window(spec) * input_scale
    c = torch.cos(omega) * w.unsqueeze(0)
    s = -torch.sin(omega) * w.unsqueeze(0)
    return torch.cat([c, s], dim=0)


def inverse_basis(spec: StftSpec) -> torch.Tensor:
    """(2F, nfft) rows [s_k cos w / N ; -s_k sin w / N] -- GTCRN/STFT_Process.py:229-251."""
    n = spec.nfft
    omega_factor = 2.0 * torch.pi / n
    k = torch.arange(spec.fbins, dtype=torch.float32).unsqueeze(1)
    m = torch.arange(n, dtype=torch.float32).unsqueeze(0)
    omega = omega_factor * k * m
    cos_b = torch.cos(omega)
    sin_b = torch.sin(omega)
    scale = 2.0 * torch.ones(spec.fbins, 1)
    scale[0] = 1.0
    if n % 2 == 0:
        scale[spec.fbins - 1] = 1.0
    inv_n = 1.0 / n
    w = make_window(spec)
    re = (scale * cos_b * inv_n) * w.unsqueeze(0)
    im = (scale * -sin_b * inv_n) * w.unsqueeze(0)
    return torch.cat([re, im], dim=0)


def window_sum(spec: StftSpec, n_frames: int) -> torch.Tensor:
    """Overlap-added w^2 over the trimmed output range, (L_out,) --
    GTCRN/STFT_Process.py:254-262."""
    w2 = make_window(spec).square().reshape(1, 1, -1)
    ws = F.conv_transpose1d(torch.ones(1, 1, n_frames), w2, stride=spec.hop).reshape(-1)
    half = spec.nfft // 2
    if spec.center:
        ws = ws[half:ws.numel() - half]
    return ws.contiguous()


def pad_signal(spec: StftSpec, x: torch.Tensor) -> torch.Tensor:
    """Centre padding; reflect excludes the edge sample (GTCRN/STFT_Process.py:305-315)."""
    if not spec.center:
        return x
    half = spec.nfft // 2
    if spec.pad_mode == "reflect":
        left = x[..., 1:half + 1].flip(-1)
        right = x[..., -(half + 1):-1].flip(-1)
        return torch.cat([left, x, right], dim=-1)
    z = torch.zeros(*x.shape[:-1], half, dtype=x.dtype)
    return torch.cat([z, x, z], dim=-1)


def stft_packed(spec: StftSpec, x: torch.Tensor, input_scale: float = 1.0) -> torch.Tensor:
    """x (B,1,L) fp32 -> (B,2F,T) packed [Re;Im] -- `_stft_B_packed_forward`
    (GTCRN/STFT_Process.py:303-316): strided Conv1d with the windowed DFT rows."""
    k = forward_basis(spec, input_scale).unsqueeze(1)
    return F.conv1d(pad_signal(spec, x), k, stride=spec.hop)


def istft_packed(spec: StftSpec, s: torch.Tensor) -> torch.Tensor:
    """(B,2F,T) -> (B,1,L_out) -- `_istft_B_packed_forward` (GTCRN/STFT_Process.py:326-336),
    `inverse_packed` (ZipEnhancer/STFT_Process.py:291-300, multiplies by a precomputed
    reciprocal), `_istft_packed_forward` (MossFormer2_SE_48K/STFT_Process.py:301-309)."""
    k = inverse_basis(spec).unsqueeze(1)
    inv = F.conv_transpose1d(s, k, stride=spec.hop)
    half = spec.nfft // 2
    if spec.center:
        inv = inv[..., half:inv.shape[-1] - half]
    ws = window_sum(spec, s.shape[-1])
    if spec.norm == "multiply":
        return inv * (1.0 / ws)
    return inv / ws

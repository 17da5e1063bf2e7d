"""DFSMN (48 kHz, causal) weight packing: checkpoint-shaped `state_dict` -> flat fp32 blob.

Host-side equivalent of `DFSMN.__init__` / `_build_dfsmn_buffers` (reference `DFSMN/Export_DFSMN.py:86-189`):

  * Kaldi log-mel-fbank analysis (per-frame DC removal -> 0.97 pre-emphasis -> symmetric hamming -> 2048-point DFT) folded
    into one (2 x 1025, 1920) kernel and concatenated with the mask-STFT rows (2 x 961) -> ONE analysis matrix (:105-140),
  * Kaldi triangular mel filterbank, 120 x 1025 (:142-146),
  * the DfsmnAns affines as (K, N) matrices, every causal FSMN memory kernel with its inner residual folded into the
    current-frame tap (:176-178),
  * periodic-hamming synthesis basis and its overlap-added w^2 (`ISTFT_WINDOW_TYPE = 'hamming_periodic'`, :32).

State-dict keys are the attribute paths the wrapper dereferences on the modelscope `DfsmnAns` model
(`linear1.linear.weight`, `deepfsmn.3.conv1.weight`, `linear2.linear.bias`, ...).
"""
from __future__ import annotations

from dataclasses import dataclass

import math

import numpy as np
import torch

from . import stft_tables

FAMILY = "dfsmn"
ANALYSIS = stft_tables.StftGeometry(1920, 1920, 960, "hamming_sym", False, "constant", "divide")
SYNTHESIS = stft_tables.StftGeometry(1920, 1920, 960, "hamming", False, "constant", "divide")
KALDI_NFFT, PREEMPH = 2048, 0.97


@dataclass(frozen=True)
class DfsmnHyper:
    layers: int = 9          # upstream DfsmnAns depth (not in the reference; read off the live model there)
    lorder: int = 20
    n_mels: int = 120
    hidden: int = 256
    n_bins: int = 961
    sample_rate: int = 48000

    def n_frames(self, length: int) -> int:
        return (length - ANALYSIS.nfft) // ANALYSIS.hop + 1


def _f(t) -> np.ndarray:
    return np.ascontiguousarray(t.detach().cpu().numpy().astype(np.float32, copy=False))


def mel_matrix(n_mels: int = 120, sample_rate: float = 48000.0, f_lo: float = 20.0) -> torch.Tensor:
    """(n_mels, 1025): `torchaudio.compliance.kaldi.get_mel_banks(n_mels, 2048, sr, 20, 0, 100, -500, 1.0)` zero-padded right
    (:142-146) -- Kaldi mel scale 1127 ln(1 + f / 700), fp32 arithmetic in the reference's order, no VTLN warp at 1.0."""
    mel_lo = 1127.0 * math.log(1.0 + f_lo / 700.0)
    mel_hi = 1127.0 * math.log(1.0 + 0.5 * sample_rate / 700.0)
    delta = (mel_hi - mel_lo) / (n_mels + 1)
    b = torch.arange(n_mels).unsqueeze(1)
    left, center, right = mel_lo + b * delta, mel_lo + (b + 1.0) * delta, mel_lo + (b + 2.0) * delta
    m = (1127.0 * torch.log(1.0 + (sample_rate / KALDI_NFFT) * torch.arange(KALDI_NFFT // 2) / 700.0)).unsqueeze(0)
    up, down = (m - left) / (center - left), (right - m) / (right - center)
    banks = torch.max(torch.zeros(1), torch.min(up, down))
    return torch.nn.functional.pad(banks, (0, 1)).float()


def fbank_kernel() -> torch.Tensor:
    """(2 * 1025, 1920) fp32, `fold_preemphasis_and_dc` of the windowed 2048-point DFT rows (:105-130)."""
    n = ANALYSIS.nfft
    win = torch.hamming_window(n, periodic=False, alpha=0.54, beta=0.46, dtype=torch.float64)
    t = torch.arange(n, dtype=torch.float64).unsqueeze(0)
    f = torch.arange(KALDI_NFFT // 2 + 1, dtype=torch.float64).unsqueeze(1)
    omega = (2.0 * torch.pi / KALDI_NFFT) * f * t
    out = []
    for basis in (torch.cos(omega) * win.unsqueeze(0), -torch.sin(omega) * win.unsqueeze(0)):
        flt = torch.cat(((1.0 - PREEMPH) * basis[:, :1] - PREEMPH * basis[:, 1:2], basis[:, 1:-1] - PREEMPH * basis[:, 2:],
                         basis[:, -1:]), dim=1)
        out.append(flt - flt.mean(dim=1, keepdim=True))
    return torch.cat(out, dim=0).float()


def pack(sd: dict, h: DfsmnHyper, input_audio_length: int) -> dict[str, np.ndarray]:
    L = int(input_audio_length)
    if L < ANALYSIS.nfft or (L - ANALYSIS.nfft) % ANALYSIS.hop:
        raise ValueError("input_audio_length must be 1920 + k*960: snip-edges framing, no centre padding")
    T = h.n_frames(L)
    blob: dict[str, np.ndarray] = {}
    blob["analysis_w"] = _f(torch.cat([fbank_kernel(), stft_tables.forward_basis(ANALYSIS)], dim=0))      # (3972, 1920)
    blob["mel_t"] = _f(mel_matrix(h.n_mels, float(h.sample_rate)).t())                        # (1025, 120)
    blob["lin1_w"], blob["lin1_b"] = _f(sd["linear1.linear.weight"].t()), _f(sd["linear1.linear.bias"])
    blob["lin2_w"], blob["lin2_b"] = _f(sd["linear2.linear.weight"].t()), _f(sd["linear2.linear.bias"])
    for i in range(h.layers):
        u = f"deepfsmn.{i}"
        cw = sd[f"{u}.conv1.weight"].squeeze(-1).clone().float()[:, 0, :]            # (hidden, lorder)
        if cw.shape != (h.hidden, h.lorder):
            raise ValueError(f"{u}.conv1.weight does not match hidden={h.hidden}, lorder={h.lorder}")
        cw[:, -1] += 1.0                                                             # inner residual p1 + conv(p1)
        blob[f"uf{i}.lin_w"], blob[f"uf{i}.lin_b"] = _f(sd[f"{u}.linear.weight"].t()), _f(sd[f"{u}.linear.bias"])
        blob[f"uf{i}.proj_w"] = _f(sd[f"{u}.project.weight"].t())
        blob[f"uf{i}.conv_w"] = _f(cw.t())                                           # (lorder, hidden) tap-major
    blob["stft.inv"] = _f(stft_tables.inverse_basis(SYNTHESIS))
    blob["stft.norm"] = _f(stft_tables.norm_table(SYNTHESIS, T))
    return blob


def metadata(h: DfsmnHyper, input_audio_length: int, in_dtype: str = "INT16", out_dtype: str = "INT16") -> dict[str, str]:
    """Metadata keys of `Export_DFSMN.py:310-320` + the depth / memory order the reference reads off the live model."""
    md = {
        "audio_metadata_version": 1, "producer": "adn.dfsmn_params", "model_name": "DFSMN",
        "task": "denoise", "model_family": FAMILY, "dynamic_axes": "0", "opset": 20,
        "input_audio_dtype": in_dtype, "output_audio_dtype": out_dtype,
        "in_sample_rate": h.sample_rate, "out_sample_rate": h.sample_rate, "model_sample_rate": h.sample_rate,
        "input_audio_length": input_audio_length, "export_audio_length": input_audio_length,
        "model_audio_length": input_audio_length, "output_audio_length": input_audio_length,
        "input_to_output_scale": 1.0, "batch_window_seconds": 1.5, "use_batch_fold": "0",
        "batch_fold_inference_default": "0", "fold_window_length": 72000, "fold_input_length": 72000,
        "max_dynamic_audio_seconds": 6, "normalize_audio_default": "0", "normalize_target_rms": 4096.0,
        "window_type": "hamming", "nfft": 1920, "window_length": 1920, "hop_length": 960,
        "max_signal_length": h.n_frames(input_audio_length), "center_pad": "0", "pad_mode": "constant",
        "feature_kind": "kaldi_fbank_stft", "input_channels": 1, "output_channels": 1, "num_audio_inputs": 1,
        "n_mels": h.n_mels, "nfft_stft": 1920, "kaldi_nfft": KALDI_NFFT, "kaldi_frame_length": 1920, "kaldi_hop_length": 960,
        "preemph_coeff": PREEMPH, "istft_window_type": "hamming_periodic",
        "dfsmn_layers": h.layers, "dfsmn_lorder": h.lorder,
    }
    return {k: str(v) for k, v in md.items()}

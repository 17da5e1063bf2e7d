"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference DFSMN (48 kHz, causal) path.

Restates `DFSMN.__init__` (the Kaldi log-mel-fbank front end folded into one Conv1d together with the mask STFT, the
channels-first DfsmnAns buffers with the inner FSMN residual folded into the last memory tap) and `forward`
(reference `DFSMN/Export_DFSMN.py:71-250`) as plain functions over a flat `state_dict`.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import it.

The reference takes its weights from the un-vendored `modelscope` pipeline (`speech_dfsmn_ans_psm_48k_causal`, no pinned
version; `:273-277`).  The wrapper's forward is made of leaf torch ops; only attribute *paths and shapes* are needed:
`skeleton()` builds a holder with the attribute paths `_build_dfsmn_buffers` dereferences (`linear1.linear`,
`deepfsmn[i].{linear, project, conv1, lorder, output_dim}`, `linear2.linear`).  The layer sizes are the wrapper's in-file
comments (120 mel -> 256 -> 961 bins, `:171-175`); depth 9 and lorder 20 are the upstream DfsmnAns defaults and NOT in the
reference: parity is self-referential in those two numbers.

Pinned (tests/test_oracle_pinning.py): against the reference wrapper executed from /root/reference around the skeleton on
identical seeded weights (container only), and against the committed fixtures tests/golden/dfsmn_*.npz generated from
that execution.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F

from stft_oracle import StftSpec, forward_basis, istft_packed

ANALYSIS = StftSpec(1920, 1920, 960, "hamming_sym", False, "constant", "divide")     # WINDOW_TYPE 'hamming' of DFSMN/STFT_Process.py:92
SYNTHESIS = StftSpec(1920, 1920, 960, "hamming", False, "constant", "divide")       # ISTFT_WINDOW_TYPE 'hamming_periodic' (:93)
INT16_SCALE = 32768.0
INV_INT16 = float(1.0 / INT16_SCALE)
KALDI_NFFT, KALDI_FRAME, KALDI_HOP, PREEMPH = 2048, 1920, 960, 0.97


@dataclass(frozen=True)
class DfsmnConfig:
    layers: int = 9
    n_mels: int = 120
    hidden: int = 256
    lorder: int = 20
    n_bins: int = 961

    def n_frames(self, length: int) -> int:
        return (length - KALDI_FRAME) // KALDI_HOP + 1


class _Affine(nn.Module):
    def __init__(self, i, o):
        super().__init__()
        self.linear = nn.Linear(i, o)


class _UniDeepFsmn(nn.Module):
    def __init__(self, d, lorder):
        super().__init__()
        self.lorder, self.output_dim = lorder, d
        self.linear = nn.Linear(d, d)
        self.project = nn.Linear(d, d, bias=False)
        self.conv1 = nn.Conv2d(d, d, [lorder, 1], [1, 1], groups=d, bias=False)


def skeleton(c: DfsmnConfig = DfsmnConfig()) -> nn.Module:
    m = nn.Module()
    m.linear1 = _Affine(c.n_mels, c.hidden)
    m.deepfsmn = nn.ModuleList([_UniDeepFsmn(c.hidden, c.lorder) for _ in range(c.layers)])
    m.linear2 = _Affine(c.hidden, c.n_bins)
    return m


def random_state_dict(c: DfsmnConfig = DfsmnConfig(), seed: int = 0) -> dict[str, torch.Tensor]:
    torch.manual_seed(seed)
    g = torch.Generator().manual_seed(seed)
    sd = {k: v.clone().float() for k, v in skeleton(c).state_dict().items()}
    for k, v in sd.items():
        if k.endswith("bias"):
            sd[k] = v + 0.05 * torch.randn(v.shape, generator=g)
    return sd


def mel_banks(n_mels: int = 120) -> torch.Tensor:
    """`torchaudio.compliance.kaldi.get_mel_banks(n_mels, 2048, 48000, 20, 0, 100, -500, 1.0)` zero-padded right (:142-146),
    restated (Kaldi mel scale 1127 ln(1 + f / 700); no VTLN warp at factor 1.0)."""
    nfft, sr, lo, hi = KALDI_NFFT, 48000.0, 20.0, 0.0
    nyq = 0.5 * sr
    hi = hi + nyq if hi <= 0.0 else hi
    fft_bin_width = sr / nfft

    def mel(f):
        return 1127.0 * torch.log(1.0 + f / 700.0)

    mel_lo, mel_hi = 1127.0 * __import__("math").log(1.0 + lo / 700.0), 1127.0 * __import__("math").log(1.0 + hi / 700.0)
    delta = (mel_hi - mel_lo) / (n_mels + 1)
    b = torch.arange(n_mels).unsqueeze(1)
    left, center, right = mel_lo + b * delta, mel_lo + (b + 1.0) * delta, mel_lo + (b + 2.0) * delta
    m = mel(fft_bin_width * torch.arange(nfft // 2)).unsqueeze(0)
    up, down = (m - left) / (center - left), (right - m) / (right - center)
    banks = torch.max(torch.zeros(1), torch.min(up, down))
    return F.pad(banks, (0, 1)).float()                                                 # (n_mels, 1025)


def fbank_kernel() -> torch.Tensor:
    """(2 * 1025, 1920): per-frame DC removal -> 0.97 pre-emphasis -> symmetric hamming -> 2048-point DFT, folded (:105-130)."""
    n = KALDI_FRAME
    win = torch.hamming_window(n, periodic=False, alpha=0.54, beta=0.46, dtype=torch.float64)
    t = torch.arange(n, dtype=torch.float64).unsqueeze(0)
    f = torch.arange(KALDI_NFFT // 2 + 1, dtype=torch.float64).unsqueeze(1)
    omega = (2.0 * torch.pi / KALDI_NFFT) * f * t
    cw, sw = torch.cos(omega) * win.unsqueeze(0), -torch.sin(omega) * win.unsqueeze(0)

    def fold(b):
        flt = torch.cat(((1.0 - PREEMPH) * b[:, :1] - PREEMPH * b[:, 1:2], b[:, 1:-1] - PREEMPH * b[:, 2:], b[:, -1:]), dim=1)
        return flt - flt.mean(dim=1, keepdim=True)

    return torch.cat([fold(cw), fold(sw)], dim=0).float()


def fold(sd: dict, c: DfsmnConfig) -> dict[str, torch.Tensor]:
    """Raw state_dict -> the buffers the forward uses (`__init__` :86-147, `_build_dfsmn_buffers` :151-189)."""
    P = {"analysis_w": torch.cat([fbank_kernel(), forward_basis(ANALYSIS)], dim=0),      # (2050 + 1922, 1920)
         "mel_banks": mel_banks(c.n_mels),
         "lin1_w": sd["linear1.linear.weight"].float(), "lin1_b": sd["linear1.linear.bias"].float(),
         "lin2_w": sd["linear2.linear.weight"].float(), "lin2_b": sd["linear2.linear.bias"].float()}
    for i in range(c.layers):
        u = f"deepfsmn.{i}"
        cw = sd[f"{u}.conv1.weight"].squeeze(-1).clone().float()                          # (256, 1, lorder)
        cw[:, 0, -1] += 1.0                                                                # inner residual p1 + conv(p1)
        P[f"uf{i}.lin_w"], P[f"uf{i}.lin_b"] = sd[f"{u}.linear.weight"].float(), sd[f"{u}.linear.bias"].float()
        P[f"uf{i}.proj_w"], P[f"uf{i}.conv_w"] = sd[f"{u}.project.weight"].float(), cw[:, 0, :].contiguous()
    return P


def dfsmn_forward(sd: dict, audio: torch.Tensor, c: DfsmnConfig = DfsmnConfig(), in_dtype: str = "F32", out_dtype: str = "F32",
                  dbg=None, folded: dict | None = None) -> torch.Tensor:
    """audio (B,1,L) at 48 kHz in `in_dtype`, (L - 1920) % 960 == 0 -> (B,1,L) in `out_dtype`; windows independent (:191-250)."""
    P = folded if folded is not None else fold(sd, c)
    x = audio.float()
    if "int" in in_dtype.lower():
        x = x * INV_INT16
    an = F.conv1d(x, P["analysis_w"].unsqueeze(1), stride=KALDI_HOP)                      # (B, 3972, T)
    kb = KALDI_NFFT // 2 + 1
    re, im, spec = an[:, :kb], an[:, kb:2 * kb], an[:, 2 * kb:]
    power = (re * re + im * im) * (INT16_SCALE * INT16_SCALE)
    feat = torch.matmul(P["mel_banks"].unsqueeze(0), power).clamp(min=torch.finfo(torch.float32).eps).log()
    h = F.relu(F.conv1d(feat, P["lin1_w"].unsqueeze(-1), P["lin1_b"]))
    if dbg is not None:
        dbg["feat"], dbg["lin1"] = feat, h
    pad = torch.zeros(h.shape[0], c.hidden, c.lorder - 1)
    for i in range(c.layers):
        f1 = F.relu(F.conv1d(h, P[f"uf{i}.lin_w"].unsqueeze(-1), P[f"uf{i}.lin_b"]))
        p1 = F.conv1d(f1, P[f"uf{i}.proj_w"].unsqueeze(-1), None)
        h = h + F.conv1d(torch.cat((pad, p1), dim=2), P[f"uf{i}.conv_w"].unsqueeze(1), None, groups=c.hidden)
        if dbg is not None:
            dbg[f"uf{i}"] = h
    mask = torch.sigmoid(F.conv1d(h, P["lin2_w"].unsqueeze(-1), P["lin2_b"]))             # (B, 961, T)
    if dbg is not None:
        dbg["mask"] = mask
    y = istft_packed(SYNTHESIS, spec * torch.cat((mask, mask), dim=1))
    if "int" in out_dtype.lower():
        return (y * INT16_SCALE).clamp(min=-32768.0, max=32767.0).to(torch.int16)
    return y if "32" in out_dtype else y.to(torch.float16)


def dfsmn_forward_batch(sd, audio, c: DfsmnConfig = DfsmnConfig(), in_dtype="F32", out_dtype="F32"):
    return torch.cat([dfsmn_forward(sd, audio[i:i + 1], c, in_dtype, out_dtype) for i in range(audio.shape[0])], dim=0)

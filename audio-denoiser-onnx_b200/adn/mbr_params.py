"""Mel-Band-Roformer (stereo) weight packing: checkpoint-shaped `state_dict` -> flat fp32 blob.

Host-side equivalent of the fusions in `MelBandRoformer.__init__` (reference
`Mel_Band_Roformer/Stereo/Export_MelBandRoformer.py:340-531`):

  * mel band layout (`create_mel_filter_bank` :119-142 -> `freq_indices`, `dim_inputs` :350-368),
  * RMSNorm gains folded into the consuming Linear in float64 (:455-463, 504-531),
  * attention scale folded into the Q rows (:514-516),
  * scatter-average denominator folded into the GLU value rows (:472-498),
  * rotary tables with the GPT-J sign folded into sin; the time tables take the reference's
    fp16 round trip, the frequency tables stay fp32 (:371-378, 438-452).

Weights are stored as (N, K) row-major (= nn.Linear.weight), which is the K-major operand layout
the tcgen05 GEMM wants.  Tensor names are the keys csrc/mbr.cu looks up.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import stft_tables

FAMILY = "mel_band_roformer"


@dataclass(frozen=True)
class MbrHyper:
    dim: int = 384
    depth: int = 6
    heads: int = 8
    dim_head: int = 64
    num_bands: int = 60
    sample_rate: int = 44100
    nfft: int = 2048
    hop: int = 441
    mlp_expansion_factor: int = 4


def _f(t) -> np.ndarray:
    return np.ascontiguousarray(t.detach().cpu().numpy().astype(np.float32, copy=False))


def _mel_points(n, fmax):
    """Slaney mel <-> Hz (reference :68-116), n+2 band edges in Hz."""
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0

    def to_mel(f):
        return min_log_mel + np.log(f / min_log_hz) / logstep if f >= min_log_hz else f / f_sp

    mels = np.linspace(to_mel(0.0), to_mel(fmax), n + 2)
    hz = f_sp * mels
    hi = mels >= min_log_mel
    hz[hi] = min_log_hz * np.exp(logstep * (mels[hi] - min_log_mel))
    return hz


def band_layout(h: MbrHyper):
    """Returns (freq_indices over the (freq,chan)-interleaved axis, per-band input widths,
    per-(source, re/im) averaging scale)."""
    nf = h.nfft // 2 + 1
    fmax = h.sample_rate / 2.0
    edges = _mel_points(h.num_bands, fmax)
    fftfreqs = np.linspace(0, fmax, nf)
    fdiff = np.diff(edges)
    ramps = np.subtract.outer(edges, fftfreqs)
    w = np.zeros((h.num_bands, nf), dtype=np.float32)
    for i in range(h.num_bands):
        w[i] = np.maximum(0, np.minimum(-ramps[i] / fdiff[i], ramps[i + 2] / fdiff[i + 1]))
    w *= (2.0 / (edges[2:h.num_bands + 2] - edges[:h.num_bands]))[:, None]
    w = w.astype(np.float32, copy=False)
    w[0, 0] = 1.0
    w[-1, -1] = 1.0
    member = torch.from_numpy(w > 0)
    idx = torch.arange(nf).expand(h.num_bands, -1)[member]
    idx = (idx.unsqueeze(1).expand(-1, 2) * 2 + torch.arange(2)).flatten()          # stereo interleave
    widths = tuple(int(4 * c) for c in member.sum(dim=1).tolist())                  # 2 (re/im) * bins * 2 (chan)
    per_freq = member.sum(dim=0).repeat_interleave(2).clamp(min=1e-8).double()
    denom = (1.0 / per_freq)[idx.long()].repeat_interleave(2)
    return idx.to(torch.int64), widths, denom


def model_length(input_audio_length: int, in_rate: int = 44100) -> int:
    """Window length at the 44.1 kHz model rate: F.interpolate(scale_factor=44100/in_rate) yields floor(L * scale)."""
    if in_rate == 44100:
        return int(input_audio_length)
    return int(np.floor(float(input_audio_length) * float(44100 / in_rate)))


def pack(sd: dict, h: MbrHyper, input_audio_length: int, in_rate: int = 44100) -> dict[str, np.ndarray]:
    input_audio_length = model_length(input_audio_length, in_rate)
    if input_audio_length % h.hop:
        raise ValueError("input_audio_length must be a multiple of hop_length (441)")
    idx, widths, denom = band_layout(h)
    d, di = h.dim, h.heads * h.dim_head
    blob: dict[str, np.ndarray] = {}
    blob["freq_indices"] = idx.numpy().astype(np.float32)
    blob["band_din"] = np.asarray(widths, np.float32)

    off = 0
    w1, b1, w2, b2 = [], [], [], []
    for b, din in enumerate(widths):
        g = (din ** 0.5) * sd[f"band_split.to_features.{b}.0.gamma"].double()
        blob[f"bs_w.{b}"] = _f((sd[f"band_split.to_features.{b}.1.weight"].double() * g.unsqueeze(0)).float())
        blob[f"bs_b.{b}"] = _f(sd[f"band_split.to_features.{b}.1.bias"].float())
        q = f"mask_estimators.0.to_freqs.{b}.0"
        w1.append(sd[f"{q}.0.weight"].float()); b1.append(sd[f"{q}.0.bias"].float())
        w2.append(sd[f"{q}.2.weight"].float()); b2.append(sd[f"{q}.2.bias"].float())
        dv = denom[off:off + din]
        off += din
        w3 = sd[f"{q}.4.weight"].double().clone()
        b3 = sd[f"{q}.4.bias"].double().clone()
        w3[:din] *= dv.unsqueeze(1)
        b3[:din] *= dv
        blob[f"me_w3.{b}"] = _f(w3.float())
        blob[f"me_b3.{b}"] = _f(b3.float())
    blob["me_w1"] = _f(torch.stack(w1, 0))          # (bands, 4*dim, dim)   == (N, K) per band
    blob["me_b1"] = _f(torch.stack(b1, 0))
    blob["me_w2"] = _f(torch.stack(w2, 0))
    blob["me_b2"] = _f(torch.stack(b2, 0))

    scale = h.dim_head ** -0.5
    for i in range(h.depth):
        for j in (0, 1):                            # 0 = time transformer, 1 = frequency transformer
            p = f"layers.{i}.{j}"
            a, f = f"{p}.layers.0.0", f"{p}.layers.0.1"
            n = f"tf.{2 * i + j}"
            g_in = (d ** 0.5) * sd[f"{a}.norm.gamma"].double()
            wqkv = sd[f"{a}.to_qkv.weight"].double()
            fused = torch.cat([wqkv[:di] * scale, wqkv[di:2 * di], wqkv[2 * di:3 * di],
                               sd[f"{a}.to_gates.weight"].double()], dim=0) * g_in.unsqueeze(0)
            blob[f"{n}.in_w"] = _f(fused.float())
            blob[f"{n}.in_b"] = _f(torch.cat([torch.zeros(3 * di, dtype=torch.float64),
                                              sd[f"{a}.to_gates.bias"].double()]).float())
            blob[f"{n}.out_w"] = _f(sd[f"{a}.to_out.0.weight"].float())
            g_ff = (d ** 0.5) * sd[f"{f}.net.0.gamma"].double()
            blob[f"{n}.ff1_w"] = _f((sd[f"{f}.net.1.weight"].double() * g_ff.unsqueeze(0)).float())
            blob[f"{n}.ff1_b"] = _f(sd[f"{f}.net.1.bias"].float())
            blob[f"{n}.ff2_w"] = _f(sd[f"{f}.net.4.weight"].float())
            blob[f"{n}.ff2_b"] = _f(sd[f"{f}.net.4.bias"].float())
            blob[f"{n}.out_g"] = _f(((d ** 0.5) * sd[f"{p}.norm.gamma"].double()).float())

    geom = stft_tables.GEOMETRY["mel_band_roformer"]
    t = geom.n_frames(input_audio_length)
    table_len = max(t, h.num_bands)
    pos = torch.arange(table_len, dtype=torch.float32).unsqueeze(-1)
    inv_freq = 10000.0 ** -(torch.arange(0, h.dim_head, 2, dtype=torch.float32) / h.dim_head)
    rot = torch.repeat_interleave(pos * inv_freq, repeats=2, dim=-1)
    cos, sin = torch.cos(rot), torch.sin(rot)
    sign = torch.ones(h.dim_head)
    sign[0::2] = -1.0
    blob["rope.tcos"] = _f(cos.half()[:t].float())
    blob["rope.tsin"] = _f((sin.half().float() * sign)[:t])
    blob["rope.fcos"] = _f(cos[:h.num_bands])
    blob["rope.fsin"] = _f(sin[:h.num_bands] * sign)

    blob["stft.fwd"] = _f(stft_tables.forward_basis(geom))
    blob["istft.inv"] = _f(stft_tables.inverse_basis(geom))
    blob["istft.norm"] = _f(stft_tables.norm_table(geom, t))
    return blob


def metadata(h: MbrHyper, input_audio_length: int, in_dtype: str = "INT16", out_dtype: str = "INT16",
             in_rate: int = 44100, out_rate: int = 44100) -> dict[str, str]:
    """Metadata keys of `Export_MelBandRoformer.py:728-733` + the model hyper-parameters the
    reference reads from the (absent) YAML."""
    g = stft_tables.GEOMETRY["mel_band_roformer"]
    mlen = model_length(input_audio_length, in_rate)
    olen = mlen if out_rate == 44100 else int(np.floor(float(mlen) * float(out_rate / 44100)))
    md = {
        "audio_metadata_version": 1, "producer": "adn.mbr_params", "model_name": "MelBandRoformer_Stereo",
        "task": "denoise", "model_family": FAMILY, "dynamic_axes": "0", "opset": 20,
        "input_audio_dtype": in_dtype, "output_audio_dtype": out_dtype,
        "in_sample_rate": in_rate, "out_sample_rate": out_rate, "model_sample_rate": h.sample_rate,
        "input_audio_length": input_audio_length, "export_audio_length": input_audio_length,
        "model_audio_length": mlen, "output_audio_length": olen,
        "input_to_output_scale": float(out_rate / in_rate), "batch_window_seconds": 1.5, "use_batch_fold": "0",
        "batch_fold_inference_default": "0", "fold_window_length": 66150, "fold_input_length": 66150,
        "max_dynamic_audio_seconds": 6, "normalize_audio_default": "0", "normalize_target_rms": 4096.0,
        "window_type": g.window_type, "nfft": g.nfft, "window_length": g.win_length, "hop_length": g.hop,
        "max_signal_length": g.n_frames(mlen), "center_pad": "1", "pad_mode": "reflect",
        "feature_kind": "stft_mel_band", "input_channels": 2, "output_channels": 2, "num_audio_inputs": 1,
        "mbr_dim": h.dim, "mbr_depth": h.depth, "mbr_heads": h.heads, "mbr_dim_head": h.dim_head,
        "mbr_num_bands": h.num_bands,
    }
    return {k: str(v) for k, v in md.items()}

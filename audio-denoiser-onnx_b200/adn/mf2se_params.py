"""MossFormer2-SE-48K weight packing: checkpoint-shaped `state_dict` -> flat fp32 blob.

Host-side equivalent of `MOSSFORMER_SE.__init__` (reference
`MossFormer2_SE_48K/Export_MossFormer_SE.py:74-283`):

  * Kaldi fbank frontend (DC removal, 0.97 pre-emphasis, symmetric Hamming, 2048-point real
    DFT) folded into one matrix and concatenated with the analysis-STFT rows (:228-252, :96-102),
  * Kaldi mel filter matrix (:254-275),
  * ScaleNorm gains folded into to_hidden||to_qk and to_out (:158-176), quadratic 1/group and
    linear 1/n folded into the OffsetScale rows (:177-182),
  * LayerNorm affines folded into to_u||to_v (:200-218),
  * speaker-0 rows of conv1d_out folded into output||output_gate (:209-226),
  * sinusoidal position table and rotary tables with their fp16 storage round trip (:109-117,
    :139-147).

The state_dict keys are the attribute paths the reference wrapper dereferences on the upstream
`clearvoice` model object (`mossformer.*`), e.g.
`mdl.intra_mdl.mossformerM.layers.3.to_hidden.mdl.1.weight`.  Linear weights are stored
(N, K) row-major; depthwise taps are stored tap-major (k, C) for coalesced channel access.
Tensor names are the keys csrc/mf2se.cu looks up.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import stft_tables

FAMILY = "mossformer2_se"
GEOM_KEY = "mossformer2_se_48k"


@dataclass(frozen=True)
class Mf2Hyper:
    layers: int = 24
    dim: int = 512
    vu: int = 1024
    qk: int = 128
    group: int = 256
    dw_kernel: int = 17
    fsmn_inner: int = 256
    lorder: int = 20
    rot_dim: int = 32
    n_mels: int = 60
    out_bins: int = 961
    sample_rate: int = 48000


def _f(t) -> np.ndarray:
    return np.ascontiguousarray(t.detach().cpu().numpy().astype(np.float32, copy=False))


def kaldi_basis(win_len: int, preemph: float = 0.97) -> torch.Tensor:
    """(2*(P/2+1), win_len): rows [Re ; Im] of  DFT_P . diag(hamming_sym) . preemphasis . (I - 1/N)."""
    P = 1 << (win_len - 1).bit_length()
    k = torch.arange(P // 2 + 1, dtype=torch.float64).unsqueeze(1)
    n = torch.arange(win_len, dtype=torch.float64).unsqueeze(0)
    w = torch.hamming_window(win_len, periodic=False, alpha=0.54, beta=0.46, dtype=torch.float64).unsqueeze(0)
    phase = (2.0 * torch.pi / P) * k * n
    dft = torch.cat([torch.cos(phase) * w, -torch.sin(phase) * w], dim=0)
    emph = torch.eye(win_len, dtype=torch.float64)
    emph -= preemph * torch.diag(torch.ones(win_len - 1, dtype=torch.float64), -1)
    emph[0, 0] -= preemph                                   # Kaldi replicates the first sample
    center = torch.eye(win_len, dtype=torch.float64) - 1.0 / win_len
    return (dft @ (emph @ center)).float()


def mel_matrix(n_mels: int, P: int, fs: float, f_lo: float = 20.0) -> torch.Tensor:
    """(n_mels, P/2+1) Kaldi triangular filters on the mel axis, Nyquist column zero."""
    to_mel = lambda hz: 1127.0 * float(np.log(1.0 + hz / 700.0))
    lo, hi = to_mel(f_lo), to_mel(0.5 * fs)
    d = (hi - lo) / (n_mels + 1)
    j = torch.arange(n_mels, dtype=torch.float64).unsqueeze(1)
    bins = 1127.0 * torch.log(1.0 + (fs / P) * torch.arange(P // 2, dtype=torch.float64) / 700.0)
    bins = bins.unsqueeze(0)
    # quotients over the explicit edge differences (not over d), as the reference evaluates them
    rise = (bins - (lo + j * d)) / ((lo + (j + 1.0) * d) - (lo + j * d))
    fall = ((lo + (j + 2.0) * d) - bins) / ((lo + (j + 2.0) * d) - (lo + (j + 1.0) * d))
    tri = torch.clamp(torch.minimum(rise, fall), min=0.0)
    return torch.cat([tri, torch.zeros(n_mels, 1, dtype=torch.float64)], dim=1).float()


def model_length(h: Mf2Hyper, input_audio_length: int, in_rate: int | None = None) -> int:
    """MODEL_AUDIO_LENGTH (Export_MossFormer_SE.py:48): the window length at the 48 kHz model rate."""
    in_rate = in_rate or h.sample_rate
    return int(round(input_audio_length * h.sample_rate / in_rate))


def pack(sd: dict, h: Mf2Hyper, input_audio_length: int, in_rate: int | None = None) -> dict[str, np.ndarray]:
    """input_audio_length is at `in_rate` (default: the model rate); tables are sized for the model-rate window."""
    input_audio_length = model_length(h, input_audio_length, in_rate)
    geom = stft_tables.GEOMETRY[GEOM_KEY]
    if input_audio_length < geom.nfft or (input_audio_length - geom.nfft) % geom.hop:
        raise ValueError("input_audio_length must be nfft + k*hop (1920 + k*384): snip-edges framing, no centre padding")
    T = geom.n_frames(input_audio_length)
    if T > h.group:
        raise ValueError(f"{T} frames exceed one FLASH group ({h.group}); fold long audio into shorter windows")
    blob: dict[str, np.ndarray] = {}
    P = 1 << (geom.nfft - 1).bit_length()
    blob["frontend"] = _f(torch.cat([kaldi_basis(geom.nfft), stft_tables.forward_basis(geom)], dim=0))
    blob["mel_banks"] = _f(mel_matrix(h.n_mels, P, float(h.sample_rate)))
    blob["norm.w"], blob["norm.b"] = _f(sd["norm.weight"]), _f(sd["norm.bias"])
    blob["enc.w"] = _f(sd["conv1d_encoder.weight"][:, :, 0])

    pos = torch.arange(T, dtype=torch.float32).unsqueeze(-1)
    sinu = pos * sd["pos_enc.inv_freq"].float()
    table = torch.cat((sinu.sin(), sinu.cos()), dim=-1) * sd["pos_enc.scale"].float()
    blob["emb_pos"] = _f(table.half().float())                                   # (T, dim)
    freqs = sd["mdl.intra_mdl.mossformerM.layers.0.rotary_pos_emb.freqs"]
    ang = torch.repeat_interleave(torch.arange(T, dtype=freqs.dtype).unsqueeze(-1) * freqs, 2, dim=-1)
    if ang.shape[-1] != h.rot_dim:
        raise ValueError("rotary width mismatch")
    blob["rot_cos"], blob["rot_sin"] = _f(ang.cos().half().float()), _f(ang.sin().half().float())

    inv_scale_in, inv_scale_out = float(h.dim) ** 0.5, float(h.vu) ** 0.5        # 1 / ScaleNorm.scale
    hs = torch.ones(4, 1, dtype=torch.float64)
    hs[0, 0] = 1.0 / h.group
    hs[3, 0] = 1.0 / float(T)
    for i in range(h.layers):
        f, b, n = f"mdl.intra_mdl.mossformerM.layers.{i}", f"mdl.intra_mdl.mossformerM.fsmn.{i}", f"L{i}"
        rows = [sd[f"{f}.{br}.mdl.1.weight"].double() * sd[f"{f}.{br}.mdl.0.g"].double() * inv_scale_in
                for br in ("to_hidden", "to_qk")]
        blob[f"{n}.in_w"] = _f(torch.cat(rows, 0).float())
        blob[f"{n}.in_b"] = _f(torch.cat([sd[f"{f}.to_hidden.mdl.1.bias"], sd[f"{f}.to_qk.mdl.1.bias"]], 0))
        taps = torch.cat([sd[f"{f}.{br}.mdl.3.sequential.1.conv.weight"][:, 0, :] for br in ("to_hidden", "to_qk")], 0)
        blob[f"{n}.in_c"] = _f(taps.t())                                         # (k, 2*vu+qk)
        blob[f"{n}.qk_gamma"] = _f((sd[f"{f}.qk_offset_scale.gamma"].double() * hs).float())
        blob[f"{n}.qk_beta"] = _f((sd[f"{f}.qk_offset_scale.beta"].double() * hs).float())
        blob[f"{n}.out_w"] = _f((sd[f"{f}.to_out.mdl.1.weight"].double() * sd[f"{f}.to_out.mdl.0.g"].double()
                                 * inv_scale_out).float())
        blob[f"{n}.out_b"] = _f(sd[f"{f}.to_out.mdl.1.bias"])
        blob[f"{n}.out_c"] = _f(sd[f"{f}.to_out.mdl.3.sequential.1.conv.weight"][:, 0, :].t())

        blob[f"{n}.c1_w"] = _f(sd[f"{b}.conv1.0.weight"][:, :, 0])
        blob[f"{n}.c1_b"] = _f(sd[f"{b}.conv1.0.bias"])
        blob[f"{n}.c1_a"] = _f(sd[f"{b}.conv1.1.weight"].reshape(-1)[:1])
        blob[f"{n}.n1_w"], blob[f"{n}.n1_b"] = _f(sd[f"{b}.norm1.weight"]), _f(sd[f"{b}.norm1.bias"])
        uw, ub, uc = [], [], []
        for br in ("to_u", "to_v"):
            q = f"{b}.gated_fsmn.{br}.mdl"
            w = sd[f"{q}.1.weight"].double()
            uw.append(w * sd[f"{q}.0.weight"].double().unsqueeze(0))
            ub.append(w @ sd[f"{q}.0.bias"].double() + sd[f"{q}.1.bias"].double())
            uc.append(sd[f"{q}.3.sequential.1.conv.weight"][:, 0, :])
        blob[f"{n}.uv_w"] = _f(torch.cat(uw, 0).float())
        blob[f"{n}.uv_b"] = _f(torch.cat(ub, 0).float())
        blob[f"{n}.uv_c"] = _f(torch.cat(uc, 0).t())
        m = f"{b}.gated_fsmn.fsmn"
        blob[f"{n}.ul_w"], blob[f"{n}.ul_b"] = _f(sd[f"{m}.linear.weight"]), _f(sd[f"{m}.linear.bias"])
        blob[f"{n}.up_w"] = _f(sd[f"{m}.project.weight"])
        blob[f"{n}.mem_c"] = _f(sd[f"{m}.conv1.weight"][:, 0, :, 0].t())          # (2*lorder-1, inner)
        blob[f"{n}.n2_w"], blob[f"{n}.n2_b"] = _f(sd[f"{b}.norm2.weight"]), _f(sd[f"{b}.norm2.bias"])
        blob[f"{n}.c2_w"] = _f(sd[f"{b}.conv2.weight"][:, :, 0])
        blob[f"{n}.c2_b"] = _f(sd[f"{b}.conv2.bias"])

    blob["mm_norm.w"], blob["mm_norm.b"] = _f(sd["mdl.intra_mdl.norm.weight"]), _f(sd["mdl.intra_mdl.norm.bias"])
    blob["intra_norm.w"], blob["intra_norm.b"] = _f(sd["mdl.intra_norm.weight"]), _f(sd["mdl.intra_norm.bias"])
    blob["prelu_a"] = _f(sd["prelu.weight"].reshape(-1)[:1])
    d = h.dim
    first_w = sd["conv1d_out.weight"][:d, :, 0].double()
    first_b = sd["conv1d_out.bias"][:d].double()
    pair_w = torch.cat([sd["output.0.weight"], sd["output_gate.0.weight"]], 0)[:, :, 0].double()
    pair_b = torch.cat([sd["output.0.bias"], sd["output_gate.0.bias"]], 0).double()
    blob["gate_w"] = _f((pair_w @ first_w).float())
    blob["gate_b"] = _f((pair_w @ first_b + pair_b).float())
    blob["dec_w"] = _f(sd["conv1_decoder.weight"][:, :, 0])
    blob["istft.inv"] = _f(stft_tables.inverse_basis(geom))
    blob["istft.norm"] = _f(stft_tables.norm_table(geom, T))
    return blob


def metadata(h: Mf2Hyper, input_audio_length: int, in_dtype: str = "INT16", out_dtype: str = "INT16",
             matmul_dtype: str = "F32", in_rate: int | None = None, out_rate: int | None = None) -> dict[str, str]:
    """Metadata keys of `Export_MossFormer_SE.py:557-561` + the hyper-parameters the reference
    reads off the live upstream modules."""
    g = stft_tables.GEOMETRY[GEOM_KEY]
    in_rate, out_rate = in_rate or h.sample_rate, out_rate or h.sample_rate
    mlen = model_length(h, input_audio_length, in_rate)
    olen = mlen if out_rate == h.sample_rate else int(round(input_audio_length * out_rate / in_rate))
    md = {
        "audio_metadata_version": 1, "producer": "adn.mf2se_params", "model_name": "MossFormer2_SE_48K",
        "task": "denoise", "model_family": FAMILY, "dynamic_axes": "0", "opset": 20,
        "input_audio_dtype": in_dtype, "output_audio_dtype": out_dtype,
        "in_sample_rate": in_rate, "out_sample_rate": out_rate, "model_sample_rate": h.sample_rate,
        "input_audio_length": input_audio_length, "export_audio_length": input_audio_length,
        "model_audio_length": mlen, "output_audio_length": olen,
        "input_to_output_scale": float(out_rate / in_rate), "batch_window_seconds": 1.5, "use_batch_fold": "0",
        "batch_fold_inference_default": "0", "fold_window_length": 72192, "fold_input_length": 72192,
        "max_dynamic_audio_seconds": 6, "normalize_audio_default": "0", "normalize_target_rms": 4096.0,
        "window_type": "hamming", "nfft": g.nfft, "window_length": g.win_length, "hop_length": g.hop,
        "max_signal_length": g.n_frames(mlen), "center_pad": "0", "pad_mode": "constant",
        "feature_kind": "kaldi_fbank_stft", "input_channels": 1, "output_channels": 1, "num_audio_inputs": 1,
        "n_mels": h.n_mels, "mf2_layers": h.layers,
        # F32 = 3xTF32 tensor-core GEMMs with fp32-class accuracy (default, the 1e-4 parity path); BF16 = the 24 layers'
        # GEMMs on bf16 operands with fp32 accumulation (BASELINE.json configs[2] "bf16 matmuls"; frontend, log-mel,
        # norms, gates and the ISTFT stay fp32, cf. MossFormer2_SE_48K/Optimize_ONNX.py:27-108)
        "matmul_dtype": matmul_dtype,
    }
    return {k: str(v) for k, v in md.items()}

"""MossFormer2-SS-16K weight packing: checkpoint-shaped `state_dict` -> flat fp32 blob.

Host-side equivalent of `MOSSFORMER_SS.__init__` (reference
`MossFormer2_SS_16K/Export_MossFormer2_SS_16K.py:84-396`):

  * front GroupNorm affine folded into the 1x1 `conv1d_encoder` (:219-228),
  * ScaleNorm gains folded into to_hidden||to_qk and to_out (:238-247), quadratic 1/group and linear
    1/n folded into the OffsetScale rows (:248-253),
  * LayerNorm affines folded into to_u||to_v (:309-319),
  * `conv1d_out` folded into output||output_gate per speaker (:370-389),
  * sinusoidal position table and rotary tables, fp32 (:156-162, :195-206),
  * width-one Conv2d memory kernels of the dilated dense FSMN as Conv1d taps (:320-327).

The state_dict keys are the attribute paths the reference wrapper dereferences on the upstream `clearvoice`
model object (`mossformer_ss.*`), e.g. `mask_net.mdl.intra_mdl.mossformerM.fsmn.3.gated_fsmn.fsmn.conv.conv2.weight`.
Linear weights are stored (N, K) row-major; depthwise taps tap-major (k, C) for coalesced channel access.
Tensor names are the keys csrc/mf2ss.cu looks up.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

FAMILY = "mossformer2_ss"


@dataclass(frozen=True)
class SsHyper:
    layers: int = 24
    dim: int = 512
    vu: int = 1024
    qk: int = 128
    group: int = 256
    dw_kernel: int = 17
    fsmn_inner: int = 256
    lorder: int = 20
    mem_depth: int = 2
    rot_dim: int = 32
    num_spks: int = 2
    enc_kernel: int = 16
    enc_stride: int = 8
    sample_rate: int = 16000
    pad_head: int = 8000

    def n_frames(self, length: int) -> int:
        return (length - self.enc_kernel) // self.enc_stride + 1

    def out_len(self, length: int) -> int:
        return (self.n_frames(length) - 1) * self.enc_stride + self.enc_kernel


def _f(t) -> np.ndarray:
    return np.ascontiguousarray(t.detach().cpu().numpy().astype(np.float32, copy=False))


def model_length(h: SsHyper, input_audio_length: int, in_rate: int | None = None) -> int:
    """MODEL_AUDIO_LENGTH (Export_MossFormer2_SS_16K.py:36): the window length at the 16 kHz model rate."""
    in_rate = in_rate or h.sample_rate
    return int(round(input_audio_length * h.sample_rate / in_rate))


def pack(sd: dict, h: SsHyper, input_audio_length: int, in_rate: int | None = None) -> dict[str, np.ndarray]:
    """input_audio_length is at `in_rate` (default: the model rate); tables are sized for the model-rate window."""
    input_audio_length = model_length(h, input_audio_length, in_rate)
    if input_audio_length < h.enc_kernel:
        raise ValueError("input_audio_length must cover one encoder kernel (16 samples)")
    if h.mem_depth != 2:
        raise ValueError("the dilated FSMN memory is built for depth 2 (conv1 dil 1, conv2 dil 2)")
    n = h.n_frames(input_audio_length)
    mn, core = "mask_net", "mask_net.mdl.intra_mdl.mossformerM"
    blob: dict[str, np.ndarray] = {}
    blob["enc_w"] = _f(sd["enc.conv1d.weight"][:, 0, :])
    blob["enc_b"] = _f(sd["enc.conv1d.bias"]) if "enc.conv1d.bias" in sd else np.zeros(h.dim, np.float32)
    blob["dec_w"] = _f(sd["dec.weight"][:, 0, :])
    blob["dec_b"] = _f(sd["dec.bias"].reshape(-1)[:1]) if "dec.bias" in sd else np.zeros(1, np.float32)
    conv = sd[f"{mn}.conv1d_encoder.weight"].double()
    blob["front_w"] = _f((conv * sd[f"{mn}.norm.weight"].double().reshape(1, -1, 1)).float()[:, :, 0])
    shift = conv.squeeze(-1) @ sd[f"{mn}.norm.bias"].double()
    if f"{mn}.conv1d_encoder.bias" in sd:
        shift = shift + sd[f"{mn}.conv1d_encoder.bias"].double()
    blob["front_b"] = _f(shift.float())

    pos = torch.arange(n, dtype=torch.float32).unsqueeze(-1)
    sinu = pos * sd[f"{mn}.pos_enc.inv_freq"].float()
    blob["emb_pos"] = _f(torch.cat((sinu.sin(), sinu.cos()), dim=-1) * sd[f"{mn}.pos_enc.scale"].float())   # (n, dim) fp32
    freqs = sd[f"{core}.layers.0.rotary_pos_emb.freqs"]
    ang = torch.repeat_interleave(torch.arange(n, dtype=freqs.dtype).unsqueeze(-1) * freqs, 2, dim=-1)
    if ang.shape[-1] != h.rot_dim:
        raise ValueError("rotary width mismatch")
    blob["rot_cos"], blob["rot_sin"] = _f(ang.cos()), _f(ang.sin())

    inv_scale_in, inv_scale_out = float(1.0 / (h.dim ** -0.5)), float(1.0 / (h.vu ** -0.5))
    hs = torch.ones(4, 1, dtype=torch.float64)
    hs[0, 0] = float(1.0 / h.group)
    hs[3, 0] = float(1.0 / n)
    for i in range(h.layers):
        f, b, k = f"{core}.layers.{i}", f"{core}.fsmn.{i}", f"L{i}"
        rows = [sd[f"{f}.{br}.mdl.1.weight"].double() * sd[f"{f}.{br}.mdl.0.g"].double() * inv_scale_in
                for br in ("to_hidden", "to_qk")]
        blob[f"{k}.in_w"] = _f(torch.cat(rows, 0).float())
        blob[f"{k}.in_b"] = _f(torch.cat([sd[f"{f}.to_hidden.mdl.1.bias"], sd[f"{f}.to_qk.mdl.1.bias"]], 0))
        taps = torch.cat([sd[f"{f}.{br}.mdl.3.sequential.1.conv.weight"][:, 0, :] for br in ("to_hidden", "to_qk")], 0)
        blob[f"{k}.in_c"] = _f(taps.t())
        blob[f"{k}.qk_gamma"] = _f((sd[f"{f}.qk_offset_scale.gamma"].double() * hs).float())
        blob[f"{k}.qk_beta"] = _f((sd[f"{f}.qk_offset_scale.beta"].double() * hs).float())
        blob[f"{k}.out_w"] = _f((sd[f"{f}.to_out.mdl.1.weight"].double() * sd[f"{f}.to_out.mdl.0.g"].double()
                                 * inv_scale_out).float())
        blob[f"{k}.out_b"] = _f(sd[f"{f}.to_out.mdl.1.bias"])
        blob[f"{k}.out_c"] = _f(sd[f"{f}.to_out.mdl.3.sequential.1.conv.weight"][:, 0, :].t())

        blob[f"{k}.c1_w"] = _f(sd[f"{b}.conv1.0.weight"][:, :, 0])
        blob[f"{k}.c1_b"] = _f(sd[f"{b}.conv1.0.bias"])
        blob[f"{k}.c1_a"] = _f(sd[f"{b}.conv1.1.weight"].reshape(-1)[:1])
        blob[f"{k}.n1_w"], blob[f"{k}.n1_b"] = _f(sd[f"{b}.norm1.weight"]), _f(sd[f"{b}.norm1.bias"])
        uw, ub, uc = [], [], []
        for br in ("to_u", "to_v"):
            q = f"{b}.gated_fsmn.{br}.mdl"
            w = sd[f"{q}.1.weight"].double()
            uw.append(w * sd[f"{q}.0.weight"].double().unsqueeze(0))
            ub.append(w @ sd[f"{q}.0.bias"].double() + sd[f"{q}.1.bias"].double())
            uc.append(sd[f"{q}.3.sequential.1.conv.weight"][:, 0, :])
        blob[f"{k}.uv_w"] = _f(torch.cat(uw, 0).float())
        blob[f"{k}.uv_b"] = _f(torch.cat(ub, 0).float())
        blob[f"{k}.uv_c"] = _f(torch.cat(uc, 0).t())
        m = f"{b}.gated_fsmn.fsmn"
        blob[f"{k}.ul_w"], blob[f"{k}.ul_b"] = _f(sd[f"{m}.linear.weight"]), _f(sd[f"{m}.linear.bias"])
        blob[f"{k}.up_w"] = _f(sd[f"{m}.project.weight"])
        for j in range(h.mem_depth):
            w = sd[f"{m}.conv.conv{j + 1}.weight"][:, :, :, 0]                   # (inner, j+1, 2*lorder-1)
            if tuple(w.shape) != (h.fsmn_inner, j + 1, 2 * h.lorder - 1):
                raise ValueError("dilated FSMN memory kernel: unexpected geometry")
            blob[f"{k}.mem{j}_c"] = _f(w.permute(1, 2, 0))                        # (input slot, tap, channel)
            blob[f"{k}.mem{j}_nw"] = _f(sd[f"{m}.conv.norm{j + 1}.weight"])
            blob[f"{k}.mem{j}_nb"] = _f(sd[f"{m}.conv.norm{j + 1}.bias"])
            blob[f"{k}.mem{j}_a"] = _f(sd[f"{m}.conv.prelu{j + 1}.weight"])
        blob[f"{k}.n2_w"], blob[f"{k}.n2_b"] = _f(sd[f"{b}.norm2.weight"]), _f(sd[f"{b}.norm2.bias"])
        blob[f"{k}.c2_w"] = _f(sd[f"{b}.conv2.weight"][:, :, 0])
        blob[f"{k}.c2_b"] = _f(sd[f"{b}.conv2.bias"])

    blob["mm_norm.w"], blob["mm_norm.b"] = _f(sd[f"{mn}.mdl.intra_mdl.norm.weight"]), _f(sd[f"{mn}.mdl.intra_mdl.norm.bias"])
    blob["intra_norm.w"], blob["intra_norm.b"] = _f(sd[f"{mn}.mdl.intra_norm.weight"]), _f(sd[f"{mn}.mdl.intra_norm.bias"])
    blob["prelu_a"] = _f(sd[f"{mn}.prelu.weight"].reshape(-1)[:1])
    d = h.dim
    pair_w = torch.cat([sd[f"{mn}.output.0.weight"], sd[f"{mn}.output_gate.0.weight"]], 0)[:, :, 0].double()
    pair_b = torch.cat([sd[f"{mn}.output.0.bias"], sd[f"{mn}.output_gate.0.bias"]], 0).double()
    gw, gb = [], []
    for s in range(h.num_spks):
        sw = sd[f"{mn}.conv1d_out.weight"][s * d:(s + 1) * d, :, 0].double()
        sb = sd[f"{mn}.conv1d_out.bias"][s * d:(s + 1) * d].double()
        gw.append((pair_w @ sw).float())
        gb.append((pair_w @ sb + pair_b).float())
    blob["gate_w"] = _f(torch.cat(gw, 0))                                          # (spks * 2 * dim, dim)
    blob["gate_b"] = _f(torch.cat(gb, 0))
    blob["mask_w"] = _f(sd[f"{mn}.conv1_decoder.weight"][:, :, 0])
    return blob


def metadata(h: SsHyper, input_audio_length: int, in_dtype: str = "INT16", out_dtype: str = "INT16",
             in_rate: int | None = None, out_rate: int | None = None) -> dict[str, str]:
    """Metadata keys of `Export_MossFormer2_SS_16K.py:703-708` (no STFT keys: learned encoder / decoder)
    + the layer count the reference reads off the live upstream module.  in_rate / out_rate != 16000: the model
    resamples linearly either side (`:564-579`, `:633-648`); input_audio_length is at in_rate."""
    in_rate, out_rate = in_rate or h.sample_rate, out_rate or h.sample_rate
    mlen = model_length(h, input_audio_length, in_rate)
    out_len = h.out_len(mlen) if out_rate == h.sample_rate else int(round(input_audio_length * out_rate / in_rate))
    md = {
        "audio_metadata_version": 1, "producer": "adn.mf2ss_params", "model_name": "MossFormer2_SS_16K",
        "task": "source_separation", "model_family": FAMILY, "dynamic_axes": "0", "opset": 20,
        "input_audio_dtype": in_dtype, "output_audio_dtype": out_dtype,
        "in_sample_rate": in_rate, "out_sample_rate": out_rate, "model_sample_rate": h.sample_rate,
        "input_audio_length": input_audio_length, "export_audio_length": input_audio_length,
        "model_audio_length": mlen, "output_audio_length": out_len,
        "input_to_output_scale": float(out_rate / in_rate), "batch_window_seconds": 1.5, "use_batch_fold": "0",
        "batch_fold_inference_default": "0", "fold_window_length": 24000, "fold_input_length": 24000,
        "max_dynamic_audio_seconds": 6, "normalize_audio_default": "0", "normalize_target_rms": 4096.0,
        "feature_kind": "conv_encoder_decoder", "center_pad": "0", "input_channels": 1, "output_channels": 1,
        "num_audio_inputs": 1, "pad_head": h.pad_head, "enc_stride": h.enc_stride, "output_sources": h.num_spks,
        "mf2_layers": h.layers,
    }
    return {k: str(v) for k, v in md.items()}

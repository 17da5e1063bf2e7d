"""MossFormerGAN-SE-16K weight packing: checkpoint-shaped `state_dict` -> flat fp32 blob.

Host-side equivalent of `MOSSFORMER_SE.__init__` (reference
`MossFormerGAN_SE_16K/Export_MossFormer_SE.py:83-131, :263-530`):

  * FFConvM LayerNorm affines folded into the Linear that follows (`_fold_ln_linear`, :83-92); FFConvM pairs that
    read the same tensor fused into one Linear + one depthwise conv (`_fuse_pair`, :440-449),
  * LayerNormalization4D affine folded into the grouped (1, ks) conv of the intra path (:95-111) and into the
    unfold of the inter path (:114-131),
  * 1/Q folded into the lin_k and quad_k OffsetScale rows (:451-486), signed rotary tables,
  * triple attention: all heads' Q | K | V 1x1 convs stacked, 1/sqrt(D) folded as D^-1/4 into both the Q and K
    (channel, sub-band) affines (:488-529).

The state_dict keys are the attribute paths the reference wrapper dereferences on the upstream `clearvoice`
generator (`dense_encoder.conv_1.0.weight`, `blocks.3.intra_mossformer.to_hidden.mdl.1.weight`, ...).
Layouts for csrc/mfgan_ops.cuh: Linear / 1x1 weights transposed to (K, N), conv kernels (kt, kf, Cin, Cout),
depthwise taps tap-major (k, C), ConvTranspose1d (k, Cin, Cout) -- adjacent threads read adjacent weights.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import stft_tables

FAMILY = "mossformergan_se"
GEOM_KEY = "mossformergan_se_16k"


@dataclass(frozen=True)
class GanHyper:
    """Dimensions compiled into csrc/mfgan_ops.cuh (the upstream generator's); only the depth is free."""
    layers: int = 6
    emb: int = 64
    n_bins: int = 201
    n_freqs: int = 101
    emb_ks: int = 2
    uv: int = 128
    mf_hidden: int = 256
    mf_qk: int = 128
    rot_freqs: int = 16
    dw_kernel: int = 31
    lorder: int = 20
    heads: int = 4
    attn_e: int = 6
    dense_depth: int = 4
    dense_lorder: int = 5
    sample_rate: int = 16000
    hop: int = 100

    def padded(self, length: int) -> int:
        return length + (self.hop - length % self.hop) % self.hop

    def n_frames(self, length: int) -> int:
        return self.padded(length) // self.hop + 1


def _f(t) -> np.ndarray:
    return np.ascontiguousarray(t.detach().cpu().numpy().astype(np.float32, copy=False))


def _fold_ln(sd, ln: str, lin: str):
    w, b = sd[f"{lin}.weight"].double(), sd[f"{lin}.bias"].double()
    return (w * sd[f"{ln}.weight"].double()[None, :]).float(), (w @ sd[f"{ln}.bias"].double() + b).float()


def _pair(sd, a: str, b: str):
    wa, ba = _fold_ln(sd, f"{a}.mdl.0", f"{a}.mdl.1")
    wb, bb = _fold_ln(sd, f"{b}.mdl.0", f"{b}.mdl.1")
    taps = torch.cat((sd[f"{a}.mdl.3.sequential.1.conv.weight"], sd[f"{b}.mdl.3.sequential.1.conv.weight"]), 0)[:, 0, :]
    return torch.cat((wa, wb), 0), torch.cat((ba, bb), 0), taps.float()


def _dense(sd, pre: str, h: GanHyper, blob: dict, out: str):
    for i in range(h.dense_depth):
        blob[f"{out}{i}.conv_w"] = _f(sd[f"{pre}.conv{i + 1}.weight"].float().permute(2, 3, 1, 0))       # (kt, kf, Cin, Cout)
        blob[f"{out}{i}.conv_b"] = _f(sd[f"{pre}.conv{i + 1}.bias"])
        blob[f"{out}{i}.nw"], blob[f"{out}{i}.nb"] = _f(sd[f"{pre}.norm{i + 1}.weight"]), _f(sd[f"{pre}.norm{i + 1}.bias"])
        blob[f"{out}{i}.pa"] = _f(sd[f"{pre}.prelu{i + 1}.weight"])
        f = f"{pre}.fsmn{i + 1}.fsmn"
        blob[f"{out}{i}.fl_w"], blob[f"{out}{i}.fl_b"] = _f(sd[f"{f}.linear.weight"].t()), _f(sd[f"{f}.linear.bias"])
        blob[f"{out}{i}.fp_w"] = _f(sd[f"{f}.project.weight"].t())
        blob[f"{out}{i}.fm_w"] = _f(sd[f"{f}.conv1.weight"][:, 0, :, 0].t())                               # (2*lorder-1, C)


def model_length(h: GanHyper, input_audio_length: int, in_rate: int | None = None) -> int:
    """MODEL_AUDIO_LENGTH (Export_MossFormer_SE.py:37): the window length at the 16 kHz model rate."""
    return int(round(input_audio_length * h.sample_rate / (in_rate or h.sample_rate)))


def pack(sd: dict, h: GanHyper, input_audio_length: int, in_rate: int | None = None) -> dict[str, np.ndarray]:
    """input_audio_length is at `in_rate` (default: the model rate); tables are sized for the model-rate window."""
    input_audio_length = model_length(h, input_audio_length, in_rate)
    if input_audio_length < 400:
        raise ValueError("input_audio_length must cover one STFT frame (400 samples)")
    T = h.n_frames(input_audio_length)
    geom = stft_tables.GEOMETRY[GEOM_KEY]
    blob: dict[str, np.ndarray] = {}
    e = "dense_encoder"
    blob["enc.c1_w"], blob["enc.c1_b"] = _f(sd[f"{e}.conv_1.0.weight"][:, :, 0, 0]), _f(sd[f"{e}.conv_1.0.bias"])
    blob["enc.n1_w"], blob["enc.n1_b"], blob["enc.p1"] = _f(sd[f"{e}.conv_1.1.weight"]), _f(sd[f"{e}.conv_1.1.bias"]), _f(sd[f"{e}.conv_1.2.weight"])
    _dense(sd, f"{e}.dilated_dense", h, blob, "enc.dd")
    blob["enc.c2_w"], blob["enc.c2_b"] = _f(sd[f"{e}.conv_2.0.weight"].float().permute(2, 3, 1, 0)), _f(sd[f"{e}.conv_2.0.bias"])
    blob["enc.n2_w"], blob["enc.n2_b"], blob["enc.p2"] = _f(sd[f"{e}.conv_2.1.weight"]), _f(sd[f"{e}.conv_2.1.bias"]), _f(sd[f"{e}.conv_2.2.weight"])

    fr = sd["blocks.0.intra_mossformer.rotary_pos_emb.freqs"]
    pos = torch.arange(max(h.n_freqs, T) + 2, dtype=fr.dtype)
    ang = pos.unsqueeze(-1) * fr
    blob["rot_cos"] = _f(torch.stack((ang.cos(), ang.cos()), dim=-1).flatten(-2))
    blob["rot_sin"] = _f(torch.stack((-ang.sin(), ang.sin()), dim=-1).flatten(-2))       # rotate-half sign folded in

    C, ks = h.emb, h.emb_ks
    for i in range(h.layers):
        b, o = f"blocks.{i}", f"B{i}"
        w = sd[f"{b}.Fconv.weight"].double()
        g, bt = sd[f"{b}.intra_norm.gamma"].reshape(-1).double(), sd[f"{b}.intra_norm.beta"].reshape(-1).double()
        wg = w.view(C, ks, 1, 1, ks)
        bias = sd[f"{b}.Fconv.bias"].double().view(C, ks) + (wg * bt.view(C, 1, 1, 1, 1)).sum(dim=(2, 3, 4))
        blob[f"{o}.intra.gw"] = _f((wg * g.view(C, 1, 1, 1, 1)).float().reshape(C * ks, ks))
        blob[f"{o}.intra.gb"] = _f(bias.reshape(-1).float())
        g, bt = sd[f"{b}.inter_norm.gamma"].reshape(-1).float(), sd[f"{b}.inter_norm.beta"].reshape(-1).float()
        uw = torch.zeros(C * ks, ks)
        for k in range(ks):
            uw[k::ks, k] = g
        blob[f"{o}.inter.gw"], blob[f"{o}.inter.gb"] = _f(uw), _f(bt.repeat_interleave(ks))
        for p, q_len in (("intra", h.n_freqs), ("inter", T)):
            uw_, ub_, uc_ = _pair(sd, f"{b}.{p}_to_u", f"{b}.{p}_to_v")
            blob[f"{o}.{p}.uv_w"], blob[f"{o}.{p}.uv_b"], blob[f"{o}.{p}.uv_c"] = _f(uw_.t()), _f(ub_), _f(uc_.t())
            r = f"{b}.{p}_rnn.0"
            blob[f"{o}.{p}.rl_w"], blob[f"{o}.{p}.rl_b"] = _f(sd[f"{r}.linear.weight"].t()), _f(sd[f"{r}.linear.bias"])
            blob[f"{o}.{p}.rp_w"] = _f(sd[f"{r}.project.weight"].t())
            blob[f"{o}.{p}.rm_w"] = _f(sd[f"{r}.conv1.weight"][:, 0, :, 0].t())
            blob[f"{o}.{p}.lin_w"] = _f(sd[f"{b}.{p}_linear.weight"].float().permute(2, 0, 1))             # (k, Cin, Cout)
            blob[f"{o}.{p}.lin_b"] = _f(sd[f"{b}.{p}_linear.bias"])
            m = f"{b}.{p}_mossformer"
            iw, ib, ic = _pair(sd, f"{m}.to_hidden", f"{m}.to_qk")
            blob[f"{o}.{p}.mf.in_w"], blob[f"{o}.{p}.mf.in_b"], blob[f"{o}.{p}.mf.in_c"] = _f(iw.t()), _f(ib), _f(ic.t())
            ow, ob = _fold_ln(sd, f"{m}.to_out.mdl.0", f"{m}.to_out.mdl.1")
            blob[f"{o}.{p}.mf.out_w"], blob[f"{o}.{p}.mf.out_b"] = _f(ow.t()), _f(ob)
            blob[f"{o}.{p}.mf.out_c"] = _f(sd[f"{m}.to_out.mdl.3.sequential.1.conv.weight"][:, 0, :].float().t())
            gamma, beta = sd[f"{m}.qk_offset_scale.gamma"].clone().float(), sd[f"{m}.qk_offset_scale.beta"].clone().float()
            inv_q = 1.0 / float(q_len)
            for head in (3, 2):                              # lin_k, quad_k (fp32 in-place scaling, as the reference)
                gamma[head].mul_(inv_q)
                beta[head].mul_(inv_q)
            blob[f"{o}.{p}.mf.gamma"], blob[f"{o}.{p}.mf.beta"] = _f(gamma), _f(beta)
            for kind in ("avg", "max"):
                for j in (0, 2):
                    blob[f"{o}.{p}.se_{kind}{j}_w"] = _f(sd[f"{b}.{p}_se.{kind}_pool_layer.{j}.weight"])
                    blob[f"{o}.{p}.se_{kind}{j}_b"] = _f(sd[f"{b}.{p}_se.{kind}_pool_layer.{j}.bias"])
        names = [f"{b}.attn_conv_{t}_{j}" for t in "QKV" for j in range(h.heads)]
        blob[f"{o}.att.w"] = _f(torch.cat([sd[f"{n}.0.weight"] for n in names], 0)[:, :, 0, 0].float().t())
        blob[f"{o}.att.b"] = _f(torch.cat([sd[f"{n}.0.bias"] for n in names], 0))
        blob[f"{o}.att.a"] = _f(torch.cat([sd[f"{n}.1.weight"].expand(sd[f"{n}.0.weight"].shape[0]) for n in names], 0))
        s = float((h.attn_e * h.n_freqs) ** -0.25)
        gs, bs = [], []
        for t in "QKV":
            sc = s if t in "QK" else 1.0
            for j in range(h.heads):
                gs.append(sd[f"{b}.attn_conv_{t}_{j}.2.gamma"][0, :, 0, :].float() * sc)
                bs.append(sd[f"{b}.attn_conv_{t}_{j}.2.beta"][0, :, 0, :].float() * sc)
        blob[f"{o}.att.g"], blob[f"{o}.att.beta"] = _f(torch.cat(gs, 0)), _f(torch.cat(bs, 0))             # (112, n_freqs)
        blob[f"{o}.att.p_w"] = _f(sd[f"{b}.attn_concat_proj.0.weight"][:, :, 0, 0].float().t())
        blob[f"{o}.att.p_b"] = _f(sd[f"{b}.attn_concat_proj.0.bias"])
        blob[f"{o}.att.p_a"] = _f(sd[f"{b}.attn_concat_proj.1.weight"].expand(C))
        blob[f"{o}.att.p_g"] = _f(sd[f"{b}.attn_concat_proj.2.gamma"][0, :, 0, :])
        blob[f"{o}.att.p_beta"] = _f(sd[f"{b}.attn_concat_proj.2.beta"][0, :, 0, :])

    for dec, o in (("mask_decoder", "md"), ("complex_decoder", "cd")):
        _dense(sd, f"{dec}.dense_block", h, blob, f"{o}.dd")
        blob[f"{o}.sp_w"] = _f(sd[f"{dec}.sub_pixel.conv.weight"].float().permute(2, 3, 1, 0))
        blob[f"{o}.sp_b"] = _f(sd[f"{dec}.sub_pixel.conv.bias"])
        blob[f"{o}.nw"], blob[f"{o}.nb"], blob[f"{o}.pa"] = _f(sd[f"{dec}.norm.weight"]), _f(sd[f"{dec}.norm.bias"]), _f(sd[f"{dec}.prelu.weight"])
    blob["md.c1_w"] = _f(sd["mask_decoder.conv_1.weight"].float().permute(2, 3, 1, 0))
    blob["md.c1_b"] = _f(sd["mask_decoder.conv_1.bias"])
    blob["md.fin_w"], blob["md.fin_b"] = _f(sd["mask_decoder.final_conv.weight"].reshape(-1)), _f(sd["mask_decoder.final_conv.bias"])
    blob["md.pout"] = _f(sd["mask_decoder.prelu_out.weight"])
    blob["cd.c_w"] = _f(sd["complex_decoder.conv.weight"].float().permute(2, 3, 1, 0))
    blob["cd.c_b"] = _f(sd["complex_decoder.conv.bias"])
    blob["stft.fwd"] = _f(stft_tables.forward_basis(geom))
    blob["stft.inv"] = _f(stft_tables.inverse_basis(geom))
    blob["stft.norm"] = _f(stft_tables.norm_table(geom, T))
    return blob


def metadata(h: GanHyper, input_audio_length: int, in_dtype: str = "INT16", out_dtype: str = "INT16",
             in_rate: int | None = None, out_rate: int | None = None) -> dict[str, str]:
    """Metadata keys of `MossFormerGAN_SE_16K/Export_MossFormer_SE.py:941-947` + the block count the reference
    reads off the live upstream module."""
    g = stft_tables.GEOMETRY[GEOM_KEY]
    in_rate, out_rate = in_rate or h.sample_rate, out_rate or h.sample_rate
    mlen = model_length(h, input_audio_length, in_rate)
    # OUTPUT_AUDIO_LENGTH (:38): the reference scales the INPUT length by out / model rate
    olen = mlen if out_rate == h.sample_rate else int(round(input_audio_length * out_rate / h.sample_rate))
    md = {
        "audio_metadata_version": 1, "producer": "adn.mfgan_params", "model_name": "MossFormerGAN_SE_16K",
        "task": "denoise", "model_family": FAMILY, "dynamic_axes": "0", "opset": 20,
        "input_audio_dtype": in_dtype, "output_audio_dtype": out_dtype,
        "in_sample_rate": in_rate, "out_sample_rate": out_rate, "model_sample_rate": h.sample_rate,
        "input_audio_length": input_audio_length, "export_audio_length": input_audio_length,
        "model_audio_length": mlen, "output_audio_length": olen,
        "input_to_output_scale": float(out_rate / in_rate), "batch_window_seconds": 1.0, "use_batch_fold": "0",
        "batch_fold_inference_default": "0", "fold_window_length": 16000, "fold_input_length": 16000,
        "max_dynamic_audio_seconds": 6, "normalize_audio_default": "0", "normalize_target_rms": 4096.0,
        "window_type": "hamming", "nfft": g.nfft, "window_length": g.win_length, "hop_length": g.hop,
        "max_signal_length": h.n_frames(mlen), "center_pad": "1", "pad_mode": "reflect",
        "feature_kind": "stft_power_compressed", "input_channels": 1, "output_channels": 1, "num_audio_inputs": 1,
        "gan_layers": h.layers,
    }
    return {k: str(v) for k, v in md.items()}

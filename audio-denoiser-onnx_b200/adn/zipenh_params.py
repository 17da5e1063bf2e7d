"""ZipEnhancer weight packing: checkpoint-shaped `state_dict` -> flat fp32 blob for csrc/zipenh_ops.cuh.

Host-side equivalent of `ZipEnhancer.__init__` (reference `ZipEnhancer/Export_ZipEnhancer.py:357-699`):

  * SwooshL / SwooshR constant offsets folded into the bias of the linear that follows, in double (:446-455),
  * BiasNorm scale exp(log_scale) * sqrt(C), the layer bypass and the enclosing dual-path bypass folded into one
    per-channel (norm scale, residual scale) pair, in double (:659-676),
  * SimpleDownsample weights softmax(bias) (:456-460); out-combiner residual scale 1 - scale in double (:592-597),
  * the relative-position table `linear_pos(pe rows -(S-1) .. S-1)` per layer and sequence length (:606-613, :690-699),
    laid out (head, pos_head_dim, 2S-1),
  * attention in_proj rows regrouped as per-head [q | k | p] blocks (:615-637).

The state_dict keys are the attribute paths the reference wrapper dereferences on the upstream modelscope model
(`dense_encoder.dense_block.dense_block.2.1.weight`, `TSConformer.encoders.1.encoder.t_layers.0.feed_forward3.out_proj.weight`, ...).
Layouts: every dense contraction is a zero-padded (n_pad, k_pad) row-major matrix W[n][k] (n_pad multiple of 64, k_pad of 32);
(2,3) conv kernels are flattened tap-major W[n][(kt*3 + kf)*Cin + cin], the (1,3) convs W[n][kf*C + c].
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

from . import stft_tables

FAMILY = "zipenhancer"
GEOM_KEY = "zipenhancer"
SWOOSH_L_OFFSET = 0.035
SWOOSH_R_OFFSET = 0.313261687


@dataclass(frozen=True)
class ZipHyper:
    """Dimensions compiled into csrc/zipenh_ops.cuh (speech_zipenhancer_ans_multiloss_16k_base as far as the reference
    constrains it, upstream defaults otherwise; see oracle/zipenh_oracle.py)."""
    channels: int = 64
    heads: int = 4
    query_head_dim: int = 12
    pos_head_dim: int = 4
    value_head_dim: int = 12
    pos_dim: int = 24
    ff_dim: int = 256
    conv_kernel: int = 15
    downsample: tuple = (1, 2, 2, 1)
    dense_depth: int = 4
    up_factor: int = 2
    n_bins: int = 201
    n_sub: int = 101
    sample_rate: int = 16000
    hop: int = 100
    pe_max_len: int = 1000

    def n_frames(self, length: int) -> int:
        return length // self.hop + 1


def _f(t) -> np.ndarray:
    return np.ascontiguousarray(t.detach().cpu().numpy().astype(np.float32, copy=False))


def _pad_to(a: int, b: int) -> int:
    return (a + b - 1) // b * b


def _lin(blob: dict, name: str, w: torch.Tensor, b: torch.Tensor | None):
    n, k = w.shape
    wp = torch.zeros(_pad_to(n, 64), _pad_to(k, 32))
    wp[:n, :k] = w.float()
    blob[f"{name}.w"] = _f(wp)
    if b is not None:
        bp = torch.zeros(wp.shape[0])
        bp[:n] = b.float()
        blob[f"{name}.b"] = _f(bp)


def _norm_act(blob: dict, name: str, sd, pre: str, ni: int, ai: int):
    blob[f"{name}.in_w"], blob[f"{name}.in_b"] = _f(sd[f"{pre}.{ni}.weight"]), _f(sd[f"{pre}.{ni}.bias"])
    blob[f"{name}.prelu"] = _f(sd[f"{pre}.{ai}.weight"])


def _dense(blob: dict, out: str, sd, pre: str, h: ZipHyper):
    for i in range(h.dense_depth):
        p = f"{pre}.dense_block.{i}"
        w = sd[f"{p}.1.weight"].float()                                    # (C, Cin, 2, 3)
        _lin(blob, f"{out}.d{i}", w.permute(0, 2, 3, 1).reshape(w.shape[0], -1), sd[f"{p}.1.bias"])
        _norm_act(blob, f"{out}.d{i}", sd, p, 2, 3)


def compact_rel_pe(embed_dim: int, max_len: int) -> torch.Tensor:
    """CompactRelPositionalEncoding.pe of the upstream Zipformer2 (length_factor 1): rows = offsets -(max_len-1) .. max_len-1."""
    x = torch.arange(-(max_len - 1), max_len, dtype=torch.float32).unsqueeze(1)
    freqs = 1 + torch.arange(embed_dim // 2)
    clen = embed_dim ** 0.5
    xc = clen * x.sign() * ((x.abs() + clen).log() - math.log(clen))
    xa = (xc / (embed_dim / (2.0 * math.pi))).atan()
    pe = torch.zeros(x.shape[0], embed_dim)
    pe[:, 0::2] = (xa * freqs).cos()
    pe[:, 1::2] = (xa * freqs).sin()
    pe[:, -1] = 1.0
    return pe


def _swoosh_out(blob, name, sd, pre, offset):
    w, b = sd[f"{pre}.weight"], sd[f"{pre}.bias"]
    _lin(blob, name, w, (b.double() - offset * w.double().sum(dim=1)).to(b.dtype))


def _layer(blob: dict, out: str, sd, pre: str, outer_scale: torch.Tensor, seq_len: int, h: ZipHyper, pe: torch.Tensor):
    H, q, pd = h.heads, h.query_head_dim, h.pos_head_dim
    aw = f"{pre}.self_attn_weights"
    w, b = sd[f"{aw}.in_proj.weight"].float(), sd[f"{aw}.in_proj.bias"].float()
    kin = w.shape[1]
    wq, wk, wp = w[:H * q].reshape(H, q, kin), w[H * q:2 * H * q].reshape(H, q, kin), w[2 * H * q:].reshape(H, pd, kin)
    bq, bk, bp = b[:H * q].reshape(H, q), b[H * q:2 * H * q].reshape(H, q), b[2 * H * q:].reshape(H, pd)
    _lin(blob, f"{out}.attn_in", torch.cat((wq, wk, wp), dim=1).reshape(-1, kin), torch.cat((bq, bk, bp), dim=1).reshape(-1))
    half = pe.shape[0] // 2
    rel = F.linear(pe[half - seq_len + 1:half + seq_len], sd[f"{aw}.linear_pos.weight"].float())       # (2S-1, H*pd)
    blob[f"{out}.pos"] = _f(rel.reshape(2 * seq_len - 1, H, pd).permute(1, 2, 0))
    for j, ff in ((1, "feed_forward1"), (2, "feed_forward2"), (3, "feed_forward3")):
        _lin(blob, f"{out}.ff{j}_in", sd[f"{pre}.{ff}.in_proj.weight"], sd[f"{pre}.{ff}.in_proj.bias"])
        _swoosh_out(blob, f"{out}.ff{j}_out", sd, f"{pre}.{ff}.out_proj", SWOOSH_L_OFFSET)
    na = f"{pre}.nonlin_attention"
    _lin(blob, f"{out}.nl_in", sd[f"{na}.in_proj.weight"], sd[f"{na}.in_proj.bias"])
    _lin(blob, f"{out}.nl_out", sd[f"{na}.out_proj.weight"], sd[f"{na}.out_proj.bias"])
    for j in (1, 2):
        sa, cv = f"{pre}.self_attn{j}", f"{pre}.conv_module{j}"
        _lin(blob, f"{out}.sa{j}_in", sd[f"{sa}.in_proj.weight"], sd[f"{sa}.in_proj.bias"])
        _lin(blob, f"{out}.sa{j}_out", sd[f"{sa}.out_proj.weight"], sd[f"{sa}.out_proj.bias"])
        _lin(blob, f"{out}.cv{j}_in", sd[f"{cv}.in_proj.weight"], sd[f"{cv}.in_proj.bias"])
        _swoosh_out(blob, f"{out}.cv{j}_out", sd, f"{cv}.out_proj", SWOOSH_R_OFFSET)
        blob[f"{out}.dw{j}.w"] = _f(sd[f"{cv}.depthwise_conv.weight"][:, 0, :])
        blob[f"{out}.dw{j}.b"] = _f(sd[f"{cv}.depthwise_conv.bias"])
    blob[f"{out}.mid_scale"] = _f(sd[f"{pre}.bypass_mid.bypass_scale"])
    blob[f"{out}.norm_bias"] = _f(sd[f"{pre}.norm.bias"])
    cs = sd[f"{pre}.bypass.bypass_scale"].double() * outer_scale.double()
    l2 = sd[f"{pre}.norm.log_scale"].double().exp() * math.sqrt(h.channels)
    blob[f"{out}.norm_scale"] = _f((cs * l2).float())
    blob[f"{out}.res_scale"] = _f((1.0 - cs).float())


def pack(sd: dict, h: ZipHyper, input_audio_length: int) -> dict[str, np.ndarray]:
    geom = stft_tables.GEOMETRY[GEOM_KEY]
    T = h.n_frames(input_audio_length)
    blob: dict[str, np.ndarray] = {}
    de = "dense_encoder"
    blob["enc.c1.w"] = _f(sd[f"{de}.dense_conv_1.0.weight"].reshape(h.channels, 2))
    blob["enc.c1.b"] = _f(sd[f"{de}.dense_conv_1.0.bias"])
    _norm_act(blob, "enc.c1", sd, f"{de}.dense_conv_1", 1, 2)
    _dense(blob, "enc", sd, f"{de}.dense_block", h)
    w = sd[f"{de}.dense_conv_2.0.weight"].float()                          # (C, C, 1, 3)
    _lin(blob, "enc.c2", w[:, :, 0, :].permute(0, 2, 1).reshape(w.shape[0], -1), sd[f"{de}.dense_conv_2.0.bias"])
    _norm_act(blob, "enc.c2", sd, f"{de}.dense_conv_2", 1, 2)
    pe = compact_rel_pe(h.pos_dim, h.pe_max_len)
    for k, ds in enumerate(h.downsample):
        pre = f"TSConformer.encoders.{k}"
        inner = pre if ds == 1 else f"{pre}.encoder"
        Tk, Fk = -(-T // ds), -(-h.n_sub // ds)
        if ds > 1:
            blob[f"ts{k}.down_t"] = _f(sd[f"{pre}.downsample_t.bias"].float().softmax(dim=0))
            blob[f"ts{k}.down_f"] = _f(sd[f"{pre}.downsample_f.bias"].float().softmax(dim=0))
            sc = sd[f"{pre}.out_combiner.bypass_scale"]
            blob[f"ts{k}.comb_scale"] = _f(sc)
            blob[f"ts{k}.comb_rscale"] = _f((1.0 - sc.double()).to(sc.dtype))
        _layer(blob, f"ts{k}.f", sd, f"{inner}.f_layers.0", sd[f"{inner}.bypass_layers.0.bypass_scale"], Fk, h, pe)
        _layer(blob, f"ts{k}.t", sd, f"{inner}.t_layers.0", sd[f"{inner}.bypass_layers.1.bypass_scale"], Tk, h, pe)
    for dec, o, seq in (("mask_decoder", "mask", "mask_conv"), ("phase_decoder", "phase", "phase_conv")):
        _dense(blob, o, sd, f"{dec}.dense_block", h)
        w = sd[f"{dec}.{seq}.0.conv1.weight"].float()                       # (UPF*C, C, 1, 3)
        _lin(blob, f"{o}.up", w[:, :, 0, :].permute(0, 2, 1).reshape(w.shape[0], -1), sd[f"{dec}.{seq}.0.conv1.bias"])
        _norm_act(blob, f"{o}.up", sd, f"{dec}.{seq}", 1, 2)
    w = sd["mask_decoder.mask_conv.3.weight"].float()                      # (1, C, 1, 2)
    blob["mask.out.w"] = _f(w[:, :, 0, :].permute(0, 2, 1).reshape(-1))
    blob["mask.out.b"] = _f(sd["mask_decoder.mask_conv.3.bias"])
    wr, wi = sd["phase_decoder.phase_conv_r.weight"].float(), sd["phase_decoder.phase_conv_i.weight"].float()
    blob["phase.out.w"] = _f(torch.cat((wr, wi), dim=0)[:, :, 0, :].permute(0, 2, 1).reshape(-1))
    blob["phase.out.b"] = _f(torch.cat((sd["phase_decoder.phase_conv_r.bias"], sd["phase_decoder.phase_conv_i.bias"])))
    blob["stft.fwd"] = _f(stft_tables.forward_basis(geom))
    blob["stft.inv"] = _f(stft_tables.inverse_basis(geom))
    blob["stft.norm"] = _f(stft_tables.norm_table(geom, T))
    return blob


def metadata(h: ZipHyper, input_audio_length: int, in_dtype: str = "INT16", out_dtype: str = "INT16") -> dict[str, str]:
    """Metadata keys of `ZipEnhancer/Export_ZipEnhancer.py:986-991` (un-folded static export at the model rate) + the
    hyper-parameters the reference reads off the live upstream module."""
    g = stft_tables.GEOMETRY[GEOM_KEY]
    olen = g.hop * (input_audio_length // g.hop)
    md = {
        "audio_metadata_version": 1, "producer": "adn.zipenh_params", "model_name": "ZipEnhancer",
        "task": "denoise", "model_family": FAMILY, "dynamic_axes": "0", "opset": 20,
        "input_audio_dtype": in_dtype, "output_audio_dtype": out_dtype,
        "in_sample_rate": h.sample_rate, "out_sample_rate": h.sample_rate, "model_sample_rate": h.sample_rate,
        "input_audio_length": input_audio_length, "export_audio_length": input_audio_length,
        "model_audio_length": input_audio_length, "output_audio_length": olen,
        "input_to_output_scale": 1.0, "batch_window_seconds": 1.5, "use_batch_fold": "0",
        "batch_fold_inference_default": "0", "fold_window_length": 24000, "fold_input_length": 24000,
        "max_dynamic_audio_seconds": 2, "normalize_audio_default": "0", "normalize_target_rms": 4096.0,
        "window_type": "hann", "nfft": g.nfft, "window_length": g.win_length, "hop_length": g.hop,
        "max_signal_length": h.n_frames(input_audio_length), "center_pad": "1", "pad_mode": "reflect",
        "feature_kind": "stft_zipformer", "input_channels": 1, "output_channels": 1, "num_audio_inputs": 1, "n_mels": 100,
        "zip_channels": h.channels, "zip_heads": h.heads, "zip_query_head_dim": h.query_head_dim, "zip_pos_head_dim": h.pos_head_dim,
        "zip_value_head_dim": h.value_head_dim, "zip_ff_dim": h.ff_dim, "zip_conv_kernel": h.conv_kernel,
        "zip_downsample": ",".join(str(d) for d in h.downsample),
    }
    return {k: str(v) for k, v in md.items()}

"""H-GTCRN weight packing: reference `GTCRN_IVA` `state_dict` (raw, BatchNorm not folded) -> flat fp32 blob for libadn.

Host-side equivalent of `GTCRN_IVA.fuse_bn_` (reference `H-GTCRN/Export_H_GTCRN.py:207-230`, `:269-273`, `:459-462`).  The
network between en_convs.1 and the band synthesis is GTCRN's (same tensor names, csrc/gtcrn.cuh structs) except that the
decoder's GTConv blocks are plain causal convolutions (`:405-413`), so `gtcrn_params.pack_backbone` packs it after the
`ConvBlock`-wrapper key names (`point_conv1.conv / .bn / .act`, `:253-260`) are mapped onto GTCRN's.  en_convs.0 takes the
18 = 6 x 3 SFE channels of [mic0 re, mic0 im, mic1 re, mic1 im, selected log|Y|, other log|Y|] (`:383`, `:1018-1024`).
The WPE / AuxIVA front end has no weights; its constants travel as metadata (`:1182`).
"""
from __future__ import annotations

import numpy as np
import torch

from . import gtcrn_params as gp
from . import stft_tables

FAMILY = "h_gtcrn"
WPE_RT60, WPE_DELAY, WPE_ITER, IVA_ITER, CG_SOLVE_ITER = 0.3, 2, 1, 10, 6      # Export_H_GTCRN.py:46-50


def _gtcrn_names(sd: dict) -> dict:
    """`point_conv1.conv.weight` -> `point_conv1.weight`, `point_conv1.bn.*` -> `point_bn1.*`, `point_conv1.act.weight` ->
    `point_act.weight` (and the same for depth_conv / point_conv2) inside the six GTConv blocks."""
    out = {}
    ren = {"point_conv1": ("point_conv1", "point_bn1", "point_act"), "depth_conv": ("depth_conv", "depth_bn", "depth_act"),
           "point_conv2": ("point_conv2", "point_bn2", None)}
    for k, v in sd.items():
        parts = k.split(".")
        if len(parts) >= 6 and parts[3] in ren and parts[4] in ("conv", "bn", "act"):
            conv, bn, act = ren[parts[3]]
            tgt = {"conv": conv, "bn": bn, "act": act}[parts[4]]
            if tgt is None:
                continue
            out[".".join(parts[:3] + [tgt] + parts[5:])] = v
        else:
            out[k] = v
    return out


def pack(state_dict: dict, input_audio_length: int) -> dict[str, np.ndarray]:
    sd = _gtcrn_names({k: v for k, v in state_dict.items()})
    geom = stft_tables.GEOMETRY["h_gtcrn"]
    blob: dict[str, np.ndarray] = {}
    w0, b0 = gp._fold(sd, "encoder.en_convs.0.conv", "encoder.en_convs.0.bn")       # (16,18,1,5)
    w1, b1 = gp._fold(sd, "encoder.en_convs.1.conv", "encoder.en_convs.1.bn")
    w0p = w0[:, :, 0, :].permute(2, 1, 0).contiguous()                               # (o,ci,k) -> [k][ci][o]
    w1p = w1[:, :, 0, :].reshape(2, 8, 8, 5).permute(0, 2, 3, 1).contiguous()        # (grp,ol,ci,k) -> [grp][ci][k][ol]
    blob["enc_front_h"] = gp._f(torch.cat([
        w0p.reshape(-1), b0, w1p.reshape(-1), b1,
        sd["encoder.en_convs.0.act.weight"].reshape(-1), sd["encoder.en_convs.1.act.weight"].reshape(-1)]))
    assert blob["enc_front_h"].size == 1440 + 16 + 640 + 16 + 2
    gp.pack_backbone(sd, blob, False)
    n_frames = geom.n_frames(input_audio_length)
    blob["stft.fwd"] = gp._f(stft_tables.forward_basis(geom))
    blob["istft.inv"] = gp._f(stft_tables.inverse_basis(geom))
    blob["istft.norm"] = gp._f(stft_tables.norm_table(geom, n_frames))
    return blob


def metadata(input_audio_length: int, in_dtype: str = "INT16", out_dtype: str = "INT16") -> dict[str, str]:
    """The keys `Export_H_GTCRN.py:1177-1184` stamps, as strings.  16 kHz I/O only."""
    g = stft_tables.GEOMETRY["h_gtcrn"]
    if input_audio_length % g.hop:
        raise ValueError("H-GTCRN windows are a whole number of hops (Export_H_GTCRN.py:31)")
    t = g.n_frames(input_audio_length)
    md = {
        "audio_metadata_version": 1, "producer": "adn.hgtcrn_params", "model_name": "H_GTCRN", "task": "denoise",
        "model_family": FAMILY, "dynamic_axes": "0", "opset": 20,
        "input_audio_dtype": in_dtype, "output_audio_dtype": out_dtype,
        "in_sample_rate": 16000, "out_sample_rate": 16000, "model_sample_rate": 16000,
        "input_audio_length": input_audio_length, "export_audio_length": input_audio_length,
        "model_audio_length": input_audio_length, "output_audio_length": g.out_length(t),
        "input_to_output_scale": 1.0, "batch_window_seconds": 1.5, "use_batch_fold": "0",
        "batch_fold_inference_default": "0", "fold_window_length": 24064, "fold_input_length": 24064,
        "max_dynamic_audio_seconds": 30, "normalize_audio_default": "0", "normalize_target_rms": 4096.0,
        "window_type": g.window_type, "nfft": g.nfft, "window_length": g.win_length, "hop_length": g.hop,
        "max_signal_length": t, "center_pad": "1", "pad_mode": g.pad_mode, "feature_kind": "stft_wpe_auxiva",
        "input_channels": 2, "output_channels": 1, "num_audio_inputs": 1, "n_mels": 100,
        "wpe_rt60": WPE_RT60, "wpe_delay": WPE_DELAY, "wpe_iter": WPE_ITER, "iva_iter": IVA_ITER, "cg_solve_iter": CG_SOLVE_ITER,
    }
    return {k: str(v) for k, v in md.items()}

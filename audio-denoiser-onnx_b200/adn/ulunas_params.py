"""UL-UNAS weight packing: RAW `ULUNAS().state_dict()` (reference `UL-UNAS/Export_UL_UNAS.py:654-707`, the checkpoint after its
own `convert_state_dict`) -> flat fp32 blob.

Host-side equivalent of `prepare_for_export_` (:692-707): BatchNorm folded into every (de)conv (`fuse_bn_`, :240-262, fp32 in the
reference's expression order), AffinePReLU as positive / negative slopes (:122-129), the 0.5 / ln 10 of the log10 magnitude
folded into the first conv (:697-700).  The grouped GRUs stay in the un-fused two-GRU form (`fuse_for_export_` builds a
block-diagonal GRU that is algebraically the same), packed per (group, direction).  DFT bases / reciprocal window sum carry the
int16 scales as in `UL-UNAS/STFT_Process.py:221, :264`.
"""
from __future__ import annotations

import numpy as np
import torch

from . import stft_tables

FAMILY = "ulunas"
GEOM = stft_tables.StftGeometry(512, 512, 256, "hann", True, "reflect", "multiply")
TYPES, STRIDES, GROUPS = [0, 2, 1, 2, 1], [2, 2, 1, 1, 1], [1, 2, 2, 2, 2]
CHANNELS, KERNELS = [12, 24, 24, 32, 16], [(3, 3), (2, 3), (2, 3), (1, 5), (1, 5)]


def _f(t) -> np.ndarray:
    return np.ascontiguousarray(t.detach().cpu().numpy().astype(np.float32, copy=False))


def _fold_bn(sd, conv: str, bn: str, transposed: bool, groups: int):
    w, b = sd[f"{conv}.weight"].float(), sd.get(f"{conv}.bias")
    std = torch.sqrt(sd[f"{bn}.running_var"] + 1e-5)
    scale = sd[f"{bn}.weight"] / std
    if transposed:
        opg, ipg = w.shape[1], w.shape[0] // groups
        fw = (w.view(groups, ipg, opg, w.shape[2], w.shape[3]) * scale.view(groups, 1, opg, 1, 1)).view_as(w)
    else:
        fw = w * scale.view(-1, 1, 1, 1)
    fb = sd[f"{bn}.bias"] - sd[f"{bn}.running_mean"] * scale if b is None else (b - sd[f"{bn}.running_mean"]) * scale + sd[f"{bn}.bias"]
    return fw, fb


def _act(sd, pre: str, blob: dict, out: str):
    aw = sd[f"{pre}.affine_weight"][0, :, 0, :]
    blob[f"{out}_pos"], blob[f"{out}_neg"] = _f(aw + 1.0), _f(aw + sd[f"{pre}.slope_weight"][0, :, 0, :])
    blob[f"{out}_bias"] = _f(sd[f"{pre}.affine_bias"][0, :, 0, :])


def _gru(sd, pre: str, bidirectional: bool):
    sfx = ("", "_reverse") if bidirectional else ("",)
    return [torch.stack([sd[f"{pre}.{k}_l0{s}"] for s in sfx]) for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]


def _ctfa(sd, pre: str, blob: dict, out: str):
    for k, v in zip(("ta_wih", "ta_whh", "ta_bih", "ta_bhh"), _gru(sd, f"{pre}.ta_gru", False)):
        blob[f"{out}.{k}"] = _f(v)
    blob[f"{out}.ta_fc_w"], blob[f"{out}.ta_fc_b"] = _f(sd[f"{pre}.ta_fc.weight"].t()), _f(sd[f"{pre}.ta_fc.bias"])
    for k, v in zip(("fa_wih", "fa_whh", "fa_bih", "fa_bhh"), _gru(sd, f"{pre}.fa.gru", True)):
        blob[f"{out}.{k}"] = _f(v)
    blob[f"{out}.fa_fc_w"], blob[f"{out}.fa_fc_b"] = _f(sd[f"{pre}.fa.fc.weight"].t()), _f(sd[f"{pre}.fa.fc.bias"])


def _block(sd, pre: str, typ: int, groups: int, cout: int, deconv: bool, last: bool, blob: dict, out: str, first_scale: float = 1.0):
    if typ == 0:
        w, b = _fold_bn(sd, f"{pre}.conv", f"{pre}.bn", deconv, groups)
        blob[f"{out}.c0_w"], blob[f"{out}.c0_b"] = _f(w * first_scale if first_scale != 1.0 else w), _f(b)
        if not last:
            _act(sd, f"{pre}.act", blob, f"{out}.a0")
        _ctfa(sd, f"{pre}.ctfa", blob, out)
        return
    p0 = "pconv" if typ == 1 else "pconv1"
    w, b = _fold_bn(sd, f"{pre}.{p0}_conv", f"{pre}.{p0}_bn", False, groups)
    blob[f"{out}.c0_w"], blob[f"{out}.c0_b"] = _f(w), _f(b)
    _act(sd, f"{pre}.{p0}_act", blob, f"{out}.a0")
    w, b = _fold_bn(sd, f"{pre}.dconv_conv", f"{pre}.dconv_bn", deconv, cout)
    blob[f"{out}.c1_w"], blob[f"{out}.c1_b"] = _f(w), _f(b)
    if not (typ == 1 and last):
        _act(sd, f"{pre}.dconv_act", blob, f"{out}.a1")
    if typ == 2:
        w, b = _fold_bn(sd, f"{pre}.pconv2_conv", f"{pre}.pconv2_bn", False, groups)
        blob[f"{out}.c2_w"], blob[f"{out}.c2_b"] = _f(w), _f(b)
    _ctfa(sd, f"{pre}.{'dconv_ctfa' if typ == 1 else 'pconv2_ctfa'}", blob, out)


def pack(sd: dict, input_audio_length: int, in_dtype: str = "INT16", out_dtype: str = "INT16") -> dict[str, np.ndarray]:
    L = int(input_audio_length)
    if L < 512:
        raise ValueError("input_audio_length must cover the reflect padding (>= 512 samples)")
    T = GEOM.n_frames(L)
    blob: dict[str, np.ndarray] = {"erb": _f(sd["erb.erb_fc.weight"]), "ierb": _f(sd["erb.ierb_fc.weight"])}
    for i in range(5):
        _block(sd, f"encoder.en_convs.{i}", TYPES[i], GROUPS[i], CHANNELS[i], False, False, blob, f"enc{i}",
               float(0.5 / np.log(10.0)) if i == 0 else 1.0)
    for i in range(5):
        k = 4 - i
        _block(sd, f"decoder.de_convs.{i}", TYPES[k], GROUPS[k], CHANNELS[k - 1] if k > 0 else 1, True, k == 0, blob, f"dec{i}")
    for j in range(2):
        p, o = f"dpgrnn.{j}", f"dp{j}"
        for side, name, bi in (("i", "intra", True), ("e", "inter", False)):
            parts = [_gru(sd, f"{p}.{name}_rnn.rnn{g}", bi) for g in (1, 2)]
            for idx, k in enumerate(("wih", "whh", "bih", "bhh")):
                blob[f"{o}.{side}_{k}"] = _f(torch.stack([parts[0][idx], parts[1][idx]]))          # (group, direction, ...)
            blob[f"{o}.{side}_fc_w"], blob[f"{o}.{side}_fc_b"] = _f(sd[f"{p}.{name}_fc.weight"].t()), _f(sd[f"{p}.{name}_fc.bias"])
            blob[f"{o}.{side}_ln_g"], blob[f"{o}.{side}_ln_b"] = _f(sd[f"{p}.{name}_ln.weight"]), _f(sd[f"{p}.{name}_ln.bias"])
    is_in, is_out = "int" in in_dtype.lower(), "int" in out_dtype.lower()
    blob["stft.fwd"] = _f(stft_tables.forward_basis(GEOM, float(1.0 / 32768.0) if is_in else 1.0))
    blob["stft.inv"] = _f(stft_tables.inverse_basis(GEOM))
    ws = 1.0 / stft_tables.norm_table(GEOM, T)                      # norm_table(multiply) = 1 / window sum -> back to the sum
    blob["stft.norm"] = _f((32767.0 if is_out else 1.0) / ws)       # output_scale / win_sum (STFT_Process.py:264)
    return blob


def metadata(input_audio_length: int, in_dtype: str = "INT16", out_dtype: str = "INT16") -> dict[str, str]:
    """Metadata keys of `Export_UL_UNAS.py:1000-1010`."""
    T = GEOM.n_frames(input_audio_length)
    md = {
        "audio_metadata_version": 1, "producer": "adn.ulunas_params", "model_name": "UL_UNAS",
        "task": "denoise", "model_family": FAMILY, "dynamic_axes": "0", "opset": 20,
        "input_audio_dtype": in_dtype, "output_audio_dtype": out_dtype,
        "in_sample_rate": 16000, "out_sample_rate": 16000, "model_sample_rate": 16000,
        "input_audio_length": input_audio_length, "export_audio_length": input_audio_length,
        "model_audio_length": input_audio_length, "output_audio_length": GEOM.out_length(T),
        "input_to_output_scale": 1.0, "batch_window_seconds": 1.5, "use_batch_fold": "0",
        "batch_fold_inference_default": "0", "fold_window_length": 24064, "fold_input_length": 24064,
        "max_dynamic_audio_seconds": 6, "normalize_audio_default": "0", "normalize_target_rms": 4096.0,
        "window_type": "hann", "nfft": 512, "window_length": 512, "hop_length": 256,
        "max_signal_length": T, "center_pad": "1", "pad_mode": "reflect",
        "feature_kind": "stft_log_power_erb", "input_channels": 1, "output_channels": 1, "num_audio_inputs": 1,
        "n_mels": 100,
    }
    return {k: str(v) for k, v in md.items()}

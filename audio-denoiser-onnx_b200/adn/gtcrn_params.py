"""GTCRN weight packing: reference `state_dict` -> flat fp32 blob for libadn.

Host-side equivalent of `GTCRN.prepare_for_export_` (reference
`GTCRN/Export_GTCRN.py:546-563`): BatchNorm folding (`:171-194`, `:244-267`), ERB matrix
transposes (`:109-114`), plus the layout changes the CUDA kernels want:

  * ConvTranspose2d blocks of the decoder are rewritten as the equivalent causal
    convolutions (1x1: transpose; depthwise (3,3) with output `[..., :-pad, :]` (`:311-312`):
    flip both kernel axes),
  * LayerNorm tables (33,16) are transposed to the frame layout [c][f],
  * ERB matrices get per-band nonzero ranges so the kernels skip exact zeros.

Tensor names are the keys csrc/api.cu looks up; struct-like tensors are packed in the
field order of csrc/gtcrn.cuh.
"""
from __future__ import annotations

import numpy as np
import torch

from . import stft_tables

BN_EPS = 1e-5
FAMILY = "gtcrn"


def _f(t) -> np.ndarray:
    return np.ascontiguousarray(t.detach().cpu().numpy().astype(np.float32, copy=False))


def _fold(sd, conv, bn, deconv=False, groups=1):
    w, b = sd[f"{conv}.weight"].float(), sd[f"{conv}.bias"].float()
    scale = sd[f"{bn}.weight"].float() / torch.sqrt(sd[f"{bn}.running_var"].float() + BN_EPS)
    if deconv:
        cin, opg = w.shape[0], w.shape[1]
        fw = (w.view(groups, cin // groups, opg, w.shape[2], w.shape[3]) * scale.view(groups, 1, opg, 1, 1)).view_as(w)
    else:
        fw = w * scale.view(-1, 1, 1, 1)
    fb = (b - sd[f"{bn}.running_mean"].float()) * scale + sd[f"{bn}.bias"].float()
    return fw, fb


def _gt_block(sd, p, deconv) -> np.ndarray:
    w1, b1 = _fold(sd, f"{p}.point_conv1", f"{p}.point_bn1", deconv)
    wd, bd = _fold(sd, f"{p}.depth_conv", f"{p}.depth_bn", deconv, groups=16)
    w2, b2 = _fold(sd, f"{p}.point_conv2", f"{p}.point_bn2", deconv)
    if deconv:
        w1 = w1[:, :, 0, 0].T            # (in=24,out=16) -> (16,24)
        w2 = w2[:, :, 0, 0].T            # (in=16,out=8)  -> (8,16)
        wd = wd[:, 0].flip(-1).flip(-2)  # transposed depthwise == correlation with flipped taps
    else:
        w1 = w1[:, :, 0, 0]
        w2 = w2[:, :, 0, 0]
        wd = wd[:, 0]
    parts = [w1.reshape(-1), b1, wd.reshape(-1), bd, w2.T.contiguous().reshape(-1), b2,   # w2 stored [c][o]
             sd[f"{p}.point_act.weight"].reshape(-1), sd[f"{p}.depth_act.weight"].reshape(-1)]
    out = torch.cat([x.float().reshape(-1) for x in parts])
    assert out.numel() == 16 * 24 + 16 + 144 + 16 + 128 + 8 + 2
    return _f(out)


def _gru(blob, name, sd, p, sfx=""):
    blob[f"{name}.w_ih"] = _f(sd[f"{p}.weight_ih_l0{sfx}"])
    blob[f"{name}.w_hh"] = _f(sd[f"{p}.weight_hh_l0{sfx}"])
    blob[f"{name}.b_ih"] = _f(sd[f"{p}.bias_ih_l0{sfx}"])
    blob[f"{name}.b_hh"] = _f(sd[f"{p}.bias_hh_l0{sfx}"])


def _nonzero_ranges(mat: torch.Tensor):
    """For each column j of `mat` (rows = summation index): [lo, hi) covering all nonzeros."""
    nz = (mat != 0)
    rows = mat.shape[0]
    lo = torch.full((mat.shape[1],), 0, dtype=torch.float32)
    hi = torch.full((mat.shape[1],), 0, dtype=torch.float32)
    idx = torch.arange(rows).unsqueeze(1)
    for j in range(mat.shape[1]):
        col = nz[:, j]
        if col.any():
            ids = idx[col]
            lo[j] = float(ids.min())
            hi[j] = float(ids.max() + 1)
    return lo, hi


def model_length(input_audio_length: int, in_rate: int = 16000) -> int:
    """Window length at the 16 kHz model rate: F.interpolate(scale_factor=16000/in_rate) yields floor(L * scale)
    samples (Export_GTCRN.py:626, :638-654)."""
    if in_rate == 16000:
        return int(input_audio_length)
    return int(np.floor(float(input_audio_length) * (1.0 / (in_rate / 16000.0))))


def output_length(model_out_length: int, out_rate: int = 16000) -> int:
    if out_rate == 16000:
        return int(model_out_length)
    return int(np.floor(float(model_out_length) * (out_rate / 16000.0)))


def pack_backbone(sd: dict, blob: dict, dec_deconv: bool) -> None:
    """Everything between en_convs.1 and the band synthesis: the six GTConv blocks with their TRA GRUs, the two DPGRNNs,
    de_convs.3 / .4 and the ERB matrices.  dec_deconv: GTCRN's decoder blocks are transposed convolutions, H-GTCRN's plain."""
    for i in range(3):
        pe = f"encoder.en_convs.{i + 2}"
        blob[f"enc_gt.{i}"] = _gt_block(sd, pe, False)
        _gru(blob, f"enc_tra.{i}", sd, f"{pe}.tra.att_gru")
        blob[f"enc_tra.{i}.fc_w"] = _f(sd[f"{pe}.tra.att_fc.weight"])
        blob[f"enc_tra.{i}.fc_b"] = _f(sd[f"{pe}.tra.att_fc.bias"])
        pd = f"decoder.de_convs.{i}"
        blob[f"dec_gt.{i}"] = _gt_block(sd, pd, dec_deconv)
        _gru(blob, f"dec_tra.{i}", sd, f"{pd}.tra.att_gru")
        blob[f"dec_tra.{i}.fc_w"] = _f(sd[f"{pd}.tra.att_fc.weight"])
        blob[f"dec_tra.{i}.fc_b"] = _f(sd[f"{pd}.tra.att_fc.bias"])

    for i, n in enumerate(("dpgrnn1", "dpgrnn2")):
        for g, r in enumerate(("rnn1", "rnn2")):
            _gru(blob, f"dp.{i}.intra.{g}.0", sd, f"{n}.intra_rnn.{r}")
            _gru(blob, f"dp.{i}.intra.{g}.1", sd, f"{n}.intra_rnn.{r}", "_reverse")
            _gru(blob, f"dp.{i}.inter.{g}", sd, f"{n}.inter_rnn.{r}")
        for path in ("intra", "inter"):
            blob[f"dp.{i}.{path}_fc_w"] = _f(sd[f"{n}.{path}_fc.weight"])
            blob[f"dp.{i}.{path}_fc_b"] = _f(sd[f"{n}.{path}_fc.bias"])
            blob[f"dp.{i}.{path}_ln_w"] = _f(sd[f"{n}.{path}_ln.weight"].T.contiguous())   # (33,16)->(16,33)
            blob[f"dp.{i}.{path}_ln_b"] = _f(sd[f"{n}.{path}_ln.bias"].T.contiguous())

    # decoder tail: de_convs.3 ConvT(16->16, groups 2) weight (16, 8, 1, 5); de_convs.4 (16, 2, 1, 5)
    w3, b3 = _fold(sd, "decoder.de_convs.3.conv", "decoder.de_convs.3.bn", True, groups=2)
    w4, b4 = _fold(sd, "decoder.de_convs.4.conv", "decoder.de_convs.4.bn", True)
    w3p = w3[:, :, 0, :].permute(0, 2, 1).contiguous()   # (ci,ol,k) -> [ci][k][ol]
    w4p = w4[:, :, 0, :].permute(0, 2, 1).contiguous()   # (ci,o,k)  -> [ci][k][o]
    blob["dec_tail"] = _f(torch.cat([
        w3p.reshape(-1), b3, w4p.reshape(-1), b4,
        sd["decoder.de_convs.3.act.weight"].reshape(-1)]))
    assert blob["dec_tail"].size == 640 + 16 + 160 + 2 + 1

    # ERB: bm uses erb_fc.weight.T (192,64); bs uses ierb_fc.weight.T (64,192)  (:109-114)
    bm = sd["erb.erb_fc.weight"].float().T.contiguous()
    bs = sd["erb.ierb_fc.weight"].float().T.contiguous()
    blob["erb.bm"] = _f(bm)
    lo, hi = _nonzero_ranges(bm)
    blob["erb.bm_lo"], blob["erb.bm_hi"] = _f(lo), _f(hi)
    blob["erb.bs"] = _f(bs)
    lo, hi = _nonzero_ranges(bs)
    blob["erb.bs_lo"], blob["erb.bs_hi"] = _f(lo), _f(hi)



def pack(state_dict: dict, input_audio_length: int, in_rate: int = 16000) -> dict[str, np.ndarray]:
    """Returns {tensor name: fp32 array} for one static chunk length (given at `in_rate`)."""
    input_audio_length = model_length(input_audio_length, in_rate)
    sd = {k: v for k, v in state_dict.items()}
    geom = stft_tables.GEOMETRY["gtcrn"]
    blob: dict[str, np.ndarray] = {}

    # encoder front: en_convs.0 (16,9,1,5) and en_convs.1 (16,8,1,5, groups 2)
    w0, b0 = _fold(sd, "encoder.en_convs.0.conv", "encoder.en_convs.0.bn")
    w1, b1 = _fold(sd, "encoder.en_convs.1.conv", "encoder.en_convs.1.bn")
    w0p = w0[:, :, 0, :].permute(2, 1, 0).contiguous()                       # (o,ci,k) -> [k][ci][o]
    w1p = w1[:, :, 0, :].reshape(2, 8, 8, 5).permute(0, 2, 3, 1).contiguous()  # (grp,ol,ci,k) -> [grp][ci][k][ol]
    blob["enc_front"] = _f(torch.cat([
        w0p.reshape(-1), b0, w1p.reshape(-1), b1,
        sd["encoder.en_convs.0.act.weight"].reshape(-1), sd["encoder.en_convs.1.act.weight"].reshape(-1)]))
    assert blob["enc_front"].size == 720 + 16 + 640 + 16 + 2

    pack_backbone(sd, blob, True)

    n_frames = geom.n_frames(input_audio_length)
    blob["stft.fwd"] = _f(stft_tables.forward_basis(geom))
    blob["istft.inv"] = _f(stft_tables.inverse_basis(geom))
    blob["istft.norm"] = _f(stft_tables.norm_table(geom, n_frames))
    return blob


def metadata(input_audio_length: int, in_dtype: str = "INT16", out_dtype: str = "INT16", in_rate: int = 16000,
             out_rate: int = 16000) -> dict[str, str]:
    """The metadata keys `Export_GTCRN.py:784-788` stamps (via
    audio_onnx_metadata.build_audio_metadata_from_globals), as strings.  in_rate / out_rate != 16000: the model resamples
    linearly either side; input_audio_length is at in_rate."""
    g = stft_tables.GEOMETRY["gtcrn"]
    mlen = model_length(input_audio_length, in_rate)
    t = g.n_frames(mlen)
    md = {
        "audio_metadata_version": 1, "producer": "adn.gtcrn_params", "model_name": "GTCRN", "task": "denoise",
        "model_family": FAMILY, "dynamic_axes": "0", "opset": 20,
        "input_audio_dtype": in_dtype, "output_audio_dtype": out_dtype,
        "in_sample_rate": in_rate, "out_sample_rate": out_rate, "model_sample_rate": 16000,
        "input_audio_length": input_audio_length, "export_audio_length": input_audio_length,
        "model_audio_length": mlen, "output_audio_length": output_length(g.out_length(t), out_rate),
        "input_to_output_scale": float(out_rate / in_rate), "batch_window_seconds": 1.5, "use_batch_fold": "0",
        "batch_fold_inference_default": "0", "fold_window_length": 24064, "fold_input_length": 24064,
        "max_dynamic_audio_seconds": 30, "normalize_audio_default": "0", "normalize_target_rms": 4096.0,
        "window_type": g.window_type, "nfft": g.nfft, "window_length": g.win_length, "hop_length": g.hop,
        "max_signal_length": t, "center_pad": "1", "pad_mode": g.pad_mode, "feature_kind": "stft",
        "input_channels": 1, "output_channels": 1, "num_audio_inputs": 1, "n_mels": 100,
    }
    return {k: str(v) for k, v in md.items()}

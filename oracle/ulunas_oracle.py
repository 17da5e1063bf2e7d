"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference UL-UNAS path; the oracle of the CUDA path
csrc/ulunas_ops.cuh + csrc/ulunas.cu (model family `ulunas`), whose launch sequence tests/test_ulunas_host.py checks on the CPU.

Restates `ULUNAS` + `ULUNAS_CUSTOM.forward` (reference `UL-UNAS/Export_UL_UNAS.py:51-912`) as plain functions over the RAW
(pre-fold) `state_dict` of `ULUNAS()`: BatchNorm folds (`fuse_bn_`, :240-262), AffinePReLU slopes (:122-129), the
0.5 / ln 10 scale of the first conv (:697-700), the ERB split / merge (:92-101), cTFA (causal time attention GRU + frequency
attention bi-GRU, :132-195), the three block types (XConvBlock :211-274, XDWSBlock :277-357, XMBBlocks :360-453), grouped
dual-path GRUs with LayerNorm (:456-574; restated in the un-fused two-GRU form, which is what `fuse_for_export_` /
`fold_fused_intra_order_` are algebraically equal to), encoder / decoder with skip additions (:577-651) and the wrapper's
STFT -> power -> mask -> ISTFT -> output rule (:849-912; int16 scales folded into the STFT window and the ISTFT reciprocal
window sum, `UL-UNAS/STFT_Process.py:221, :264`).  GRUs run on torch's `nn.GRU`, like the reference's.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import it.
Pinned (tests/test_oracle_pinning.py): against the reference executed from /root/reference on the same raw state_dict
(container only) and against the committed fixtures tests/golden/ulunas_*.npz (outputs of that execution).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from stft_oracle import StftSpec, forward_basis, inverse_basis, pad_signal, window_sum

SPEC = StftSpec(512, 512, 256, "hann", True, "reflect", "multiply")
TYPES, STRIDES, GROUPS = [0, 2, 1, 2, 1], [2, 2, 1, 1, 1], [1, 2, 2, 2, 2]
CHANNELS, KERNELS, WIDTHS = [12, 24, 24, 32, 16], [(3, 3), (2, 3), (2, 3), (1, 5), (1, 5)], [65, 33, 33, 33, 33]
ERB_LOW, ERB_HIGH = 65, 64
INV_INT16 = float(1.0 / 32768.0)


def _gru(sd, pre: str, x: torch.Tensor, bidirectional: bool) -> torch.Tensor:
    """nn.GRU (sequence-first, zero initial state) with the weights under `pre`."""
    w_ih = sd[f"{pre}.weight_ih_l0"]
    g = nn.GRU(w_ih.shape[1], w_ih.shape[0] // 3, batch_first=False, bidirectional=bidirectional)
    g.load_state_dict({k[len(pre) + 1:]: v for k, v in sd.items() if k.startswith(pre + ".")})
    return g.eval()(x)[0]


def _fold_bn(sd, conv: str, bn: str, transposed: bool, groups: int):
    """`fuse_bn_` (:240-262): fp32, the reference's expression order."""
    w, b = sd[f"{conv}.weight"], sd.get(f"{conv}.bias")
    std = torch.sqrt(sd[f"{bn}.running_var"] + 1e-5)
    scale = sd[f"{bn}.weight"] / std
    if transposed:
        opg, ipg = w.shape[1], w.shape[0] // groups
        fw = (w.view(groups, ipg, opg, w.shape[2], w.shape[3]) * scale.view(groups, 1, opg, 1, 1)).view_as(w)
    else:
        fw = w * scale.view(-1, 1, 1, 1)
    fb = sd[f"{bn}.bias"] - sd[f"{bn}.running_mean"] * scale if b is None else (b - sd[f"{bn}.running_mean"]) * scale + sd[f"{bn}.bias"]
    return fw, fb


def _act(sd, pre: str, x: torch.Tensor) -> torch.Tensor:
    """AffinePReLU after `fuse_for_export_` (:122-129)."""
    pos, neg = sd[f"{pre}.affine_weight"] + 1.0, sd[f"{pre}.affine_weight"] + sd[f"{pre}.slope_weight"]
    return torch.where(x > 0, pos, neg) * x + sd[f"{pre}.affine_bias"]


def _shuffle(x: torch.Tensor) -> torch.Tensor:
    half = x.shape[1] // 2
    idx = torch.stack((torch.arange(half), torch.arange(half) + half), dim=1).reshape(-1)
    return torch.index_select(x, 1, idx)


def _ctfa(sd, pre: str, x: torch.Tensor) -> torch.Tensor:
    """cTFA (:173-195) with FA.forward_power (:152-167); x (B,C,T,F)."""
    B, C, T, Fw = x.shape
    power = x * x
    at = _gru(sd, f"{pre}.ta_gru", torch.mean(power, dim=-1).permute(2, 0, 1), False)
    at = F.linear(at, sd[f"{pre}.ta_fc.weight"], sd[f"{pre}.ta_fc.bias"]).permute(1, 2, 0)
    at = torch.sigmoid(at).unsqueeze(-1)
    r = 4
    pad = (r - Fw % r) % r
    z = F.pad(torch.mean(power, dim=1), (0, pad))
    z = z.reshape(-1, (Fw + pad) // r, r).transpose(0, 1)
    z = F.linear(_gru(sd, f"{pre}.fa.gru", z, True), sd[f"{pre}.fa.fc.weight"], sd[f"{pre}.fa.fc.bias"])
    af = torch.sigmoid(z.transpose(0, 1).reshape(B, 1, T, Fw + pad)[..., :Fw])
    return at * x * af


def _conv(sd, conv: str, bn: str, x, kernel, stride: int, groups: int, deconv: bool, w_scale: float = 1.0):
    """(De)conv + folded BN + causal trim of the last kt - 1 frames."""
    kt, kf = kernel
    w, b = _fold_bn(sd, conv, bn, deconv, groups)
    if w_scale != 1.0:
        w = w * w_scale
    if deconv:
        y = F.conv_transpose2d(x, w, b, stride=(1, stride), padding=(0, kf // 2), groups=groups)
    else:
        y = F.conv2d(x, w, b, stride=(1, stride), padding=(kt - 1, kf // 2), groups=groups)
    return y[..., :-(kt - 1), :] if kt > 1 else y


def _block(sd, pre: str, typ: int, x, cin: int, cout: int, kernel, stride: int, groups: int, deconv=False, last=False, first_scale=1.0):
    if typ == 0:                                                               # XConvBlock
        y = _conv(sd, f"{pre}.conv", f"{pre}.bn", x, kernel, stride, groups, deconv, first_scale)
        if not last:
            y = _act(sd, f"{pre}.act", y)
        y = _ctfa(sd, f"{pre}.ctfa", y)
        return _shuffle(y) if (not last and groups == 2) else y
    if typ == 1:                                                               # XDWSBlock
        h = _conv(sd, f"{pre}.pconv_conv", f"{pre}.pconv_bn", x, (1, 1), 1, groups, False)
        h = _act(sd, f"{pre}.pconv_act", h)
        if groups == 2:
            h = _shuffle(h)
        h = _conv(sd, f"{pre}.dconv_conv", f"{pre}.dconv_bn", h, kernel, stride, cout, deconv)
        if not last:
            h = _act(sd, f"{pre}.dconv_act", h)
        return _ctfa(sd, f"{pre}.dconv_ctfa", h)
    y = _conv(sd, f"{pre}.pconv1_conv", f"{pre}.pconv1_bn", x, (1, 1), 1, groups, False)        # XMBBlocks
    y = _act(sd, f"{pre}.pconv1_act", y)
    if groups == 2:
        y = _shuffle(y)
    y = _conv(sd, f"{pre}.dconv_conv", f"{pre}.dconv_bn", y, kernel, stride, cout, deconv)
    y = _act(sd, f"{pre}.dconv_act", y)
    y = _conv(sd, f"{pre}.pconv2_conv", f"{pre}.pconv2_bn", y, (1, 1), 1, groups, False)
    y = _ctfa(sd, f"{pre}.pconv2_ctfa", y)
    if cin == cout and stride == 1:
        y = y + x
    return _shuffle(y) if (not last and groups == 2) else y


def _grnn(sd, pre: str, x: torch.Tensor, bidirectional: bool) -> torch.Tensor:
    """Grouped GRU, un-fused form (:517-524)."""
    x1, x2 = x.split(x.shape[-1] // 2, dim=-1)
    return torch.cat([_gru(sd, f"{pre}.rnn1", x1, bidirectional), _gru(sd, f"{pre}.rnn2", x2, bidirectional)], dim=-1)


def _dpgrnn(sd, pre: str, x: torch.Tensor) -> torch.Tensor:
    """DPGRNN (:557-574); x (B,T,F,C)."""
    B, T, W, C = x.shape
    a = _grnn(sd, f"{pre}.intra_rnn", x.permute(2, 0, 1, 3).reshape(W, -1, C), True)
    a = F.linear(a, sd[f"{pre}.intra_fc.weight"], sd[f"{pre}.intra_fc.bias"]).reshape(W, B, -1, C).permute(1, 2, 0, 3)
    a = F.layer_norm(a, (W, C), sd[f"{pre}.intra_ln.weight"], sd[f"{pre}.intra_ln.bias"], 1e-8)
    intra = x + a
    e = _grnn(sd, f"{pre}.inter_rnn", intra.permute(1, 0, 2, 3).reshape(-1, B * W, C), False)
    e = F.linear(e, sd[f"{pre}.inter_fc.weight"], sd[f"{pre}.inter_fc.bias"]).reshape(-1, B, W, C).permute(1, 0, 2, 3)
    e = F.layer_norm(e, (W, C), sd[f"{pre}.inter_ln.weight"], sd[f"{pre}.inter_ln.bias"], 1e-8)
    return intra + e


def ulunas_mask(sd: dict, power: torch.Tensor, dbg=None) -> torch.Tensor:
    """`ULUNAS.forward` (:709-739): power (B,F,T) -> sigmoid mask (B,1,F,T)."""
    erb = sd["erb.erb_fc.weight"]                                               # (64, 192)
    feat = torch.log(power.clamp_min(1e-24).unsqueeze(1).transpose(-1, -2))     # (B,1,T,257)
    lo, hi = feat.split([ERB_LOW, feat.shape[-1] - ERB_LOW], dim=-1)
    x = torch.cat([lo, torch.matmul(hi, erb.transpose(0, 1).contiguous())], dim=-1)
    outs, cin = [], 1
    for i in range(5):
        x = _block(sd, f"encoder.en_convs.{i}", TYPES[i], x, cin, CHANNELS[i], KERNELS[i], STRIDES[i], GROUPS[i],
                   first_scale=float(0.5 / np.log(10.0)) if i == 0 else 1.0)
        outs.append(x)
        cin = CHANNELS[i]
        if dbg is not None:
            dbg[f"enc{i}"] = x
    x = x.permute(0, 2, 3, 1)
    for j in range(2):
        x = _dpgrnn(sd, f"dpgrnn.{j}", x)
        if dbg is not None:
            dbg[f"dp{j}"] = x
    x = x.permute(0, 3, 1, 2)
    for i in range(5):
        k = 4 - i                                                                # decoder block i mirrors encoder block k
        cout = CHANNELS[k - 1] if k > 0 else 1
        x = _block(sd, f"decoder.de_convs.{i}", TYPES[k], x + outs[4 - i], cin, cout, KERNELS[k], STRIDES[k], GROUPS[k], deconv=True,
                   last=(k == 0))
        cin = cout
        if dbg is not None:
            dbg[f"dec{i}"] = x
    m = torch.sigmoid(x)
    lo, hi = m.split([ERB_LOW, ERB_HIGH], dim=-1)
    m = torch.cat([lo, torch.matmul(hi, sd["erb.ierb_fc.weight"].transpose(0, 1).contiguous())], dim=-1)
    return m.transpose(-1, -2)


def ulunas_forward(sd: dict, audio: torch.Tensor, in_dtype: str = "F32", out_dtype: str = "F32", dbg=None) -> torch.Tensor:
    """`ULUNAS_CUSTOM.forward` at 16 kHz, no fold (:849-912): audio (B,1,L) -> (B,1,hop * (L // hop)); windows independent."""
    x = audio.float()
    k = forward_basis(SPEC, INV_INT16 if "int" in in_dtype.lower() else 1.0).unsqueeze(1)
    packed = F.conv1d(pad_signal(SPEC, x), k, stride=SPEC.hop)                    # (B, 514, T)
    B, _, T = packed.shape
    spec = packed.reshape(B, 2, SPEC.fbins, T)
    mask = torch.cat([ulunas_mask(sd, (spec[i:i + 1] * spec[i:i + 1]).sum(dim=1), dbg if i == 0 else None) for i in range(B)], dim=0)
    spec = (spec * mask).reshape(B, 2 * SPEC.fbins, T)
    inv = F.conv_transpose1d(spec, inverse_basis(SPEC).unsqueeze(1), stride=SPEC.hop)
    half = SPEC.nfft // 2
    scale = 32767.0 if "int" in out_dtype.lower() else 1.0
    y = inv[..., half:inv.shape[-1] - half] * (scale / window_sum(SPEC, T))      # reciprocal window sum carries the PCM scale
    if "int" not in in_dtype.lower():
        y = torch.nan_to_num(y, nan=0.0, posinf=32767.0, neginf=-32768.0)
    if "int" in out_dtype.lower():
        return y.clamp(min=-32768.0, max=32767.0).to(torch.int16)
    return y if "32" in out_dtype else y.to(torch.float16)

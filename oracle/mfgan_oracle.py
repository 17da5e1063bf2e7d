"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference MossFormerGAN-SE-16K path
(the enhancement half of BASELINE.json configs[4]).

Restates `MOSSFORMER_SE.__init__` (weight folds) and `forward` / `_mossformer_block` (reference
`MossFormerGAN_SE_16K/Export_MossFormer_SE.py:83-897`) as plain functions over a flat `state_dict`.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import it.  The CUDA
path it checks is csrc/mfgan_ops.cuh + csrc/mfgan.cu (model family `mossformergan_se`); `dbg` collects the
stage dumps the host-harness test and the GPU stage test compare against.

The reference wrapper reads its parameters from the un-vendored `clearvoice` package
(`clearvoice.models.mossformer_gan_se.generator`, no pinned version).  The wrapper's
forward is made of leaf torch ops; only attribute *paths and shapes* are needed.  `skeleton()` builds a
parameter holder with exactly the attribute paths `MOSSFORMER_SE.__init__` / `forward` dereference.
Dimensions the wrapper infers from the live module (`:263-282`) are NOT in the reference; the ones used
here are the upstream generator's as far as the wrapper constrains them (emb_dim 64, n_freqs 101 = 201
bins after the stride-2 encoder conv, emb_ks 2 / emb_hs 1 so that 100 unfold steps + ConvTranspose1d(k 2)
return 101, FFConvM depthwise k 31 pad 15 `:132-133`, UniDeepFsmn lorder 20 `:659`, 4 heads `:500`,
sub-pixel r 2 + (1,2) conv -> 201 bins) and otherwise free choices (`GanConfig`): parity for this family
is self-referential in those free dimensions, exactly like the reference's own export would be for a
different checkpoint.

Pinned (tests/test_oracle_pinning.py): against the reference wrapper executed from /root/reference
around the skeleton on identical seeded weights (container only), and against the committed fixtures
tests/golden/mfgan_*.npz generated from that execution.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F

from mf2se_oracle import _FFConvM, _OffsetScale, _Rotary
from stft_oracle import SPECS, forward_basis, inverse_basis, istft_packed, stft_packed


@dataclass(frozen=True)
class GanConfig:
    layers: int = 6
    emb: int = 64             # feature-map channels
    n_bins: int = 201         # nfft/2 + 1
    n_freqs: int = 101        # sub-bands after the stride-2 encoder conv
    emb_ks: int = 2
    emb_hs: int = 1
    uv: int = 128             # FFConvM width of the intra / inter paths (to_u, to_v)
    rnn_hidden: int = 128     # UniDeepFsmn hidden width of the paths
    lorder: int = 20
    mf_hidden: int = 256      # MossFormer to_hidden width (v and u halves of 128)
    mf_qk: int = 128
    rot_freqs: int = 16
    dw_kernel: int = 31
    heads: int = 4
    attn_e: int = 6           # per-head Q/K channels (ceil(512 / n_freqs) upstream)
    dense_depth: int = 4
    dense_lorder: int = 5     # FSMN memory order inside the dilated dense blocks
    dense_hidden: int = 64
    se_reduction: int = 1
    sp_r: int = 2
    nfft: int = 400
    hop: int = 100

    @property
    def path_in(self) -> int:
        return self.emb * self.emb_ks

    def n_frames(self, length: int) -> int:
        return ((length + self.hop - 1) // self.hop * self.hop) // self.hop + 1


INV_INT16 = float(1.0 / 32768.0)


# ----------------------------------------------------------------------------- parameter holder
class _Norm4D(nn.Module):
    """LayerNormalization4D: statistics over channels, per-channel affine (1, C, 1, 1)."""

    def __init__(self, c, eps=1e-5):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(1, c, 1, 1))
        self.beta = nn.Parameter(torch.zeros(1, c, 1, 1))
        self.eps = eps


class _Norm4DCF(nn.Module):
    """LayerNormalization4DCF: statistics over (channel, freq), affine (1, C, 1, F)."""

    def __init__(self, c, f, eps=1e-5):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(1, c, 1, f))
        self.beta = nn.Parameter(torch.zeros(1, c, 1, f))
        self.eps = eps


class _Fsmn(nn.Module):
    def __init__(self, d, hidden, lorder):
        super().__init__()
        self.lorder = lorder
        self.linear = nn.Linear(d, hidden)
        self.project = nn.Linear(hidden, d, bias=False)
        self.conv1 = nn.Conv2d(d, d, [2 * lorder - 1, 1], [1, 1], groups=d, bias=False)


class _FsmnWrap(nn.Module):
    def __init__(self, d, hidden, lorder):
        super().__init__()
        self.fsmn = _Fsmn(d, hidden, lorder)


class _DenseBlock(nn.Module):
    def __init__(self, c: GanConfig):
        super().__init__()
        self.depth = c.dense_depth
        for i in range(c.dense_depth):
            setattr(self, f"conv{i + 1}", nn.Conv2d(c.emb * (i + 1), c.emb, (2, 3), dilation=(2 ** i, 1)))
            setattr(self, f"norm{i + 1}", nn.InstanceNorm2d(c.emb, affine=True))
            setattr(self, f"prelu{i + 1}", nn.PReLU(c.emb))
            setattr(self, f"fsmn{i + 1}", _FsmnWrap(c.emb, c.dense_hidden, c.dense_lorder))


class _MossFormer(nn.Module):
    def __init__(self, c: GanConfig, group_size: int):
        super().__init__()
        self.group_size = group_size
        self.to_hidden = _FFConvM(c.emb, c.mf_hidden, c.dw_kernel, nn.LayerNorm(c.emb))
        self.to_qk = _FFConvM(c.emb, c.mf_qk, c.dw_kernel, nn.LayerNorm(c.emb))
        self.to_out = _FFConvM(c.mf_hidden // 2, c.emb, c.dw_kernel, nn.LayerNorm(c.mf_hidden // 2))
        self.qk_offset_scale = _OffsetScale(c.mf_qk)
        self.rotary_pos_emb = _Rotary(c.rot_freqs)


class _SE(nn.Module):
    def __init__(self, ch, r):
        super().__init__()
        self.avg_pool_layer = nn.Sequential(nn.Linear(ch, ch // r), nn.ReLU(), nn.Linear(ch // r, ch))
        self.max_pool_layer = nn.Sequential(nn.Linear(ch, ch // r), nn.ReLU(), nn.Linear(ch // r, ch))


class _Block(nn.Module):
    def __init__(self, c: GanConfig):
        super().__init__()
        self.emb_dim, self.emb_ks, self.emb_hs, self.n_head = c.emb, c.emb_ks, c.emb_hs, c.heads
        for p in ("intra", "inter"):
            setattr(self, f"{p}_norm", _Norm4D(c.emb))
            setattr(self, f"{p}_to_u", _FFConvM(c.path_in, c.uv, c.dw_kernel, nn.LayerNorm(c.path_in)))
            setattr(self, f"{p}_to_v", _FFConvM(c.path_in, c.uv, c.dw_kernel, nn.LayerNorm(c.path_in)))
            setattr(self, f"{p}_rnn", nn.ModuleList([_Fsmn(c.uv, c.rnn_hidden, c.lorder)]))
            setattr(self, f"{p}_linear", nn.ConvTranspose1d(c.uv, c.emb, c.emb_ks, stride=c.emb_hs))
            setattr(self, f"{p}_se", _SE(c.emb, c.se_reduction))
        self.Fconv = nn.Conv2d(c.emb, c.path_in, (1, c.emb_ks), groups=c.emb)
        self.intra_mossformer = _MossFormer(c, c.n_freqs)
        self.inter_mossformer = _MossFormer(c, c.n_freqs)
        for j in range(c.heads):
            setattr(self, f"attn_conv_Q_{j}", nn.Sequential(nn.Conv2d(c.emb, c.attn_e, 1), nn.PReLU(), _Norm4DCF(c.attn_e, c.n_freqs)))
            setattr(self, f"attn_conv_K_{j}", nn.Sequential(nn.Conv2d(c.emb, c.attn_e, 1), nn.PReLU(), _Norm4DCF(c.attn_e, c.n_freqs)))
            setattr(self, f"attn_conv_V_{j}", nn.Sequential(nn.Conv2d(c.emb, c.emb // c.heads, 1), nn.PReLU(),
                                                           _Norm4DCF(c.emb // c.heads, c.n_freqs)))
        self.attn_concat_proj = nn.Sequential(nn.Conv2d(c.emb, c.emb, 1), nn.PReLU(), _Norm4DCF(c.emb, c.n_freqs))


class _SubPixel(nn.Module):
    def __init__(self, c: GanConfig):
        super().__init__()
        self.r = c.sp_r
        self.conv = nn.Conv2d(c.emb, c.emb * c.sp_r, (1, 3))


def skeleton(c: GanConfig = GanConfig()) -> nn.Module:
    """Parameter holder with the attribute paths `MOSSFORMER_SE` dereferences (`:263-530`, `:588-861`)."""
    m = nn.Module()
    m.n_layers = c.layers
    enc = nn.Module()
    enc.conv_1 = nn.Sequential(nn.Conv2d(3, c.emb, (1, 1)), nn.InstanceNorm2d(c.emb, affine=True), nn.PReLU(c.emb))
    enc.dilated_dense = _DenseBlock(c)
    enc.conv_2 = nn.Sequential(nn.Conv2d(c.emb, c.emb, (1, 3), (1, 2), padding=(0, 1)), nn.InstanceNorm2d(c.emb, affine=True),
                               nn.PReLU(c.emb))
    m.dense_encoder = enc
    m.blocks = nn.ModuleList([_Block(c) for _ in range(c.layers)])
    md = nn.Module()
    md.dense_block = _DenseBlock(c)
    md.sub_pixel = _SubPixel(c)
    md.conv_1 = nn.Conv2d(c.emb, 1, (1, 2))
    md.norm = nn.InstanceNorm2d(1, affine=True)
    md.prelu = nn.PReLU(1)
    md.final_conv = nn.Conv2d(1, 1, (1, 1))
    md.prelu_out = nn.PReLU(c.n_bins, init=-0.25)
    m.mask_decoder = md
    cd = nn.Module()
    cd.dense_block = _DenseBlock(c)
    cd.sub_pixel = _SubPixel(c)
    cd.prelu = nn.PReLU(c.emb)
    cd.norm = nn.InstanceNorm2d(c.emb, affine=True)
    cd.conv = nn.Conv2d(c.emb, 2, (1, 2))
    m.complex_decoder = cd
    return m


def random_state_dict(c: GanConfig = GanConfig(), seed: int = 0) -> dict[str, torch.Tensor]:
    """Seeded weights: default inits with every gain / bias / slope perturbed so each fold is exercised."""
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    sd = {k: v.clone().float() for k, v in skeleton(c).state_dict().items()}
    for k, v in sd.items():
        if k.endswith("freqs"):
            continue
        r = torch.randn(v.shape, generator=g)
        if k.endswith("gamma") and "offset_scale" not in k:
            sd[k] = 1.0 + 0.2 * r
        elif k.endswith("beta") and "offset_scale" not in k:
            sd[k] = 0.1 * r
        elif "qk_offset_scale.gamma" in k:
            sd[k] = 1.0 + 0.25 * r
        elif "qk_offset_scale.beta" in k:
            sd[k] = 0.1 * r
        elif ("norm" in k or ".mdl.0." in k) and k.endswith("weight") and v.ndim == 1:
            sd[k] = 1.0 + 0.2 * r
        elif k.endswith("bias"):
            sd[k] = v + 0.05 * r
        elif "prelu" in k or (v.ndim == 1 and v.numel() == 1):
            sd[k] = v + 0.05 * r
    return sd


# ----------------------------------------------------------------------------- weight folds
def _fold_ln(sd, ln: str, lin: str):
    """LayerNorm affine folded into the Linear that follows (`_fold_ln_linear`, :83-92)."""
    w = sd[f"{lin}.weight"].double()
    b = sd[f"{lin}.bias"].double()
    return (w * sd[f"{ln}.weight"].double()[None, :]).float(), (w @ sd[f"{ln}.bias"].double() + b).float()


def _pair(sd, a: str, b: str):
    """Two FFConvMs reading the same tensor -> one Linear + one depthwise conv (`_fuse_pair`, :440-449)."""
    wa, ba = _fold_ln(sd, f"{a}.mdl.0", f"{a}.mdl.1")
    wb, bb = _fold_ln(sd, f"{b}.mdl.0", f"{b}.mdl.1")
    taps = torch.cat((sd[f"{a}.mdl.3.sequential.1.conv.weight"], sd[f"{b}.mdl.3.sequential.1.conv.weight"]), 0)[:, 0, :]
    return torch.cat((wa, wb), 0).contiguous(), torch.cat((ba, bb), 0).contiguous(), taps.float().contiguous()


def _dense_params(sd, pre: str, c: GanConfig, P: dict, out: str):
    for i in range(c.dense_depth):
        P[f"{out}{i}.conv_w"], P[f"{out}{i}.conv_b"] = sd[f"{pre}.conv{i + 1}.weight"].float(), sd[f"{pre}.conv{i + 1}.bias"].float()
        P[f"{out}{i}.nw"], P[f"{out}{i}.nb"] = sd[f"{pre}.norm{i + 1}.weight"].float(), sd[f"{pre}.norm{i + 1}.bias"].float()
        P[f"{out}{i}.pa"] = sd[f"{pre}.prelu{i + 1}.weight"].float()
        f = f"{pre}.fsmn{i + 1}.fsmn"
        P[f"{out}{i}.fl_w"], P[f"{out}{i}.fl_b"] = sd[f"{f}.linear.weight"].float(), sd[f"{f}.linear.bias"].float()
        P[f"{out}{i}.fp_w"] = sd[f"{f}.project.weight"].float()
        P[f"{out}{i}.fm_w"] = sd[f"{f}.conv1.weight"][:, 0, :, 0].float().contiguous()          # (C, 2*lorder-1)


def _mossformer_params(sd, pre: str, c: GanConfig, q_len: int, bt_len: int, P: dict, out: str):
    """`_mossformer_params` (:451-486): fused in/out projections; 1/Q folded into the lin_k and quad_k OffsetScale rows."""
    P[f"{out}.in_w"], P[f"{out}.in_b"], P[f"{out}.in_c"] = _pair(sd, f"{pre}.to_hidden", f"{pre}.to_qk")
    P[f"{out}.out_w"], P[f"{out}.out_b"] = _fold_ln(sd, f"{pre}.to_out.mdl.0", f"{pre}.to_out.mdl.1")
    P[f"{out}.out_c"] = sd[f"{pre}.to_out.mdl.3.sequential.1.conv.weight"][:, 0, :].float().contiguous()
    gamma, beta = sd[f"{pre}.qk_offset_scale.gamma"].clone().float(), sd[f"{pre}.qk_offset_scale.beta"].clone().float()
    inv_q = 1.0 / float(q_len)
    for head in (3, 2):                                 # lin_k, then quad_k (fp32 in-place scaling, as the reference)
        gamma[head].mul_(inv_q)
        beta[head].mul_(inv_q)
    P[f"{out}.gamma"], P[f"{out}.beta"] = gamma, beta
    P[f"{out}.cross_scale"] = torch.tensor(float(q_len / bt_len), dtype=torch.float64)   # python scalar in the reference


def fold(sd: dict, c: GanConfig, n_frames: int) -> dict[str, torch.Tensor]:
    """Raw state_dict -> the fused tensors the forward uses (restates :263-530)."""
    P: dict[str, torch.Tensor] = {}
    e = "dense_encoder"
    P["enc.c1_w"], P["enc.c1_b"] = sd[f"{e}.conv_1.0.weight"].float(), sd[f"{e}.conv_1.0.bias"].float()
    P["enc.n1_w"], P["enc.n1_b"], P["enc.p1"] = sd[f"{e}.conv_1.1.weight"].float(), sd[f"{e}.conv_1.1.bias"].float(), sd[f"{e}.conv_1.2.weight"].float()
    _dense_params(sd, f"{e}.dilated_dense", c, P, "enc.dd")
    P["enc.c2_w"], P["enc.c2_b"] = sd[f"{e}.conv_2.0.weight"].float(), sd[f"{e}.conv_2.0.bias"].float()
    P["enc.n2_w"], P["enc.n2_b"], P["enc.p2"] = sd[f"{e}.conv_2.1.weight"].float(), sd[f"{e}.conv_2.1.bias"].float(), sd[f"{e}.conv_2.2.weight"].float()

    fr = sd["blocks.0.intra_mossformer.rotary_pos_emb.freqs"]
    pos = torch.arange(max(c.n_freqs, n_frames) + 2, dtype=fr.dtype)
    ang = pos.unsqueeze(-1) * fr
    P["rot_cos"] = torch.stack((ang.cos(), ang.cos()), dim=-1).flatten(-2)                    # (max_seq, 32)
    P["rot_sin"] = torch.stack((-ang.sin(), ang.sin()), dim=-1).flatten(-2)                   # signed: rotate-half folded in

    for i in range(c.layers):
        b, o = f"blocks.{i}", f"B{i}"
        # intra: LayerNormalization4D affine folded into the grouped (1, ks) conv (`_fold_norm4d_conv2d`, :95-111)
        w = sd[f"{b}.Fconv.weight"].double()                                                   # (emb*ks, 1, 1, ks)
        g, bt = sd[f"{b}.intra_norm.gamma"].reshape(-1).double(), sd[f"{b}.intra_norm.beta"].reshape(-1).double()
        wg = w.view(c.emb, c.emb_ks, 1, 1, c.emb_ks)
        bias = sd[f"{b}.Fconv.bias"].double().view(c.emb, c.emb_ks) + (wg * bt.view(c.emb, 1, 1, 1, 1)).sum(dim=(2, 3, 4))
        P[f"{o}.intra.fconv_w"] = (wg * g.view(c.emb, 1, 1, 1, 1)).reshape_as(w).float()
        P[f"{o}.intra.fconv_b"] = bias.reshape(-1).float()
        # inter: the unfold as a one-hot grouped Conv1d carrying gamma / beta (`_fold_norm4d_unfold1d`, :114-131)
        g, bt = sd[f"{b}.inter_norm.gamma"].reshape(-1).float(), sd[f"{b}.inter_norm.beta"].reshape(-1).float()
        uw = torch.zeros(c.emb * c.emb_ks, 1, c.emb_ks)
        for ch in range(c.emb):
            for k in range(c.emb_ks):
                uw[ch * c.emb_ks + k, 0, k] = g[ch]
        P[f"{o}.inter.unfold_w"], P[f"{o}.inter.unfold_b"] = uw, bt.repeat_interleave(c.emb_ks)
        for p, q_len, bt_len in (("intra", c.n_freqs, n_frames), ("inter", n_frames, c.n_freqs)):
            P[f"{o}.{p}.uv_w"], P[f"{o}.{p}.uv_b"], P[f"{o}.{p}.uv_c"] = _pair(sd, f"{b}.{p}_to_u", f"{b}.{p}_to_v")
            r = f"{b}.{p}_rnn.0"
            P[f"{o}.{p}.rl_w"], P[f"{o}.{p}.rl_b"] = sd[f"{r}.linear.weight"].float(), sd[f"{r}.linear.bias"].float()
            P[f"{o}.{p}.rp_w"] = sd[f"{r}.project.weight"].float()
            P[f"{o}.{p}.rm_w"] = sd[f"{r}.conv1.weight"][:, 0, :, 0].float().contiguous()
            P[f"{o}.{p}.lin_w"], P[f"{o}.{p}.lin_b"] = sd[f"{b}.{p}_linear.weight"].float(), sd[f"{b}.{p}_linear.bias"].float()
            _mossformer_params(sd, f"{b}.{p}_mossformer", c, q_len, bt_len, P, f"{o}.{p}.mf")
            for kind in ("avg", "max"):
                for j in (0, 2):
                    P[f"{o}.{p}.se_{kind}{j}_w"] = sd[f"{b}.{p}_se.{kind}_pool_layer.{j}.weight"].float()
                    P[f"{o}.{p}.se_{kind}{j}_b"] = sd[f"{b}.{p}_se.{kind}_pool_layer.{j}.bias"].float()
        # triple attention: all heads' Q | K | V 1x1 convs stacked; 1/sqrt(D) folded as D^-1/4 into both Q and K affines (:488-529)
        names = [f"{b}.attn_conv_{t}_{j}" for t in "QKV" for j in range(c.heads)]
        P[f"{o}.att.w"] = torch.cat([sd[f"{n}.0.weight"] for n in names], 0)[:, :, 0, 0].float().contiguous()
        P[f"{o}.att.b"] = torch.cat([sd[f"{n}.0.bias"] for n in names], 0).float()
        P[f"{o}.att.a"] = torch.cat([sd[f"{n}.1.weight"].expand(sd[f"{n}.0.weight"].shape[0]) for n in names], 0).float()
        s = float((c.attn_e * c.n_freqs) ** -0.25)
        for t in "QK":
            P[f"{o}.att.{t.lower()}_g"] = torch.stack([sd[f"{b}.attn_conv_{t}_{j}.2.gamma"][0, :, 0, :] for j in range(c.heads)], 0).float() * s
            P[f"{o}.att.{t.lower()}_b"] = torch.stack([sd[f"{b}.attn_conv_{t}_{j}.2.beta"][0, :, 0, :] for j in range(c.heads)], 0).float() * s
        P[f"{o}.att.v_g"] = torch.stack([sd[f"{b}.attn_conv_V_{j}.2.gamma"][0, :, 0, :] for j in range(c.heads)], 0).float()
        P[f"{o}.att.v_b"] = torch.stack([sd[f"{b}.attn_conv_V_{j}.2.beta"][0, :, 0, :] for j in range(c.heads)], 0).float()
        P[f"{o}.att.p_w"] = sd[f"{b}.attn_concat_proj.0.weight"][:, :, 0, 0].float()
        P[f"{o}.att.p_b"], P[f"{o}.att.p_a"] = sd[f"{b}.attn_concat_proj.0.bias"].float(), sd[f"{b}.attn_concat_proj.1.weight"].float()
        P[f"{o}.att.p_g"], P[f"{o}.att.p_beta"] = sd[f"{b}.attn_concat_proj.2.gamma"][0, :, 0, :].float(), sd[f"{b}.attn_concat_proj.2.beta"][0, :, 0, :].float()

    for dec, o in (("mask_decoder", "md"), ("complex_decoder", "cd")):
        _dense_params(sd, f"{dec}.dense_block", c, P, f"{o}.dd")
        P[f"{o}.sp_w"], P[f"{o}.sp_b"] = sd[f"{dec}.sub_pixel.conv.weight"].float(), sd[f"{dec}.sub_pixel.conv.bias"].float()
        P[f"{o}.nw"], P[f"{o}.nb"], P[f"{o}.pa"] = sd[f"{dec}.norm.weight"].float(), sd[f"{dec}.norm.bias"].float(), sd[f"{dec}.prelu.weight"].float()
    P["md.c1_w"], P["md.c1_b"] = sd["mask_decoder.conv_1.weight"].float(), sd["mask_decoder.conv_1.bias"].float()
    P["md.fin_w"], P["md.fin_b"] = sd["mask_decoder.final_conv.weight"].float(), sd["mask_decoder.final_conv.bias"].float()
    P["md.pout"] = sd["mask_decoder.prelu_out.weight"].float()
    P["cd.c_w"], P["cd.c_b"] = sd["complex_decoder.conv.weight"].float(), sd["complex_decoder.conv.bias"].float()
    return P


# ----------------------------------------------------------------------------- forward
def _dwconv_res(x: torch.Tensor, taps: torch.Tensor) -> torch.Tensor:
    """x (N, S, C) + depthwise 'same' conv over S, taps (C, k)."""
    k = taps.shape[-1]
    return x + F.conv1d(x.transpose(1, 2), taps.unsqueeze(1), None, padding=(k - 1) // 2, groups=taps.shape[0]).transpose(1, 2)


def _fsmn_seq(x: torch.Tensor, lw, lb, pw, mw, lorder: int) -> torch.Tensor:
    """UniDeepFsmn on (N, S, C): x + project + depthwise memory over S (:653-664)."""
    p1 = F.linear(F.relu(F.linear(x, lw, lb)), pw)
    mem = F.conv1d(p1.transpose(1, 2), mw.unsqueeze(1), padding=lorder - 1, groups=mw.shape[0]).transpose(1, 2)
    return x + (p1 + mem)


def _dense_block(x: torch.Tensor, P: dict, pre: str, c: GanConfig) -> torch.Tensor:
    """DilatedDenseNet with an FSMN along frequency after every layer (:601-623, :792-813); x (B, C, T, F)."""
    skip, out = x, x
    for i in range(c.dense_depth):
        dil = 2 ** i
        # causal over time: pad `dil` rows on top only == symmetric pad + drop the last `dil` rows
        out = F.conv2d(F.pad(skip, (1, 1, dil, 0)), P[f"{pre}{i}.conv_w"], P[f"{pre}{i}.conv_b"], dilation=(dil, 1))
        out = F.prelu(F.instance_norm(out, None, None, P[f"{pre}{i}.nw"], P[f"{pre}{i}.nb"], True, 0.1, 1e-5), P[f"{pre}{i}.pa"])
        f1 = F.relu(F.conv2d(out, P[f"{pre}{i}.fl_w"][:, :, None, None], P[f"{pre}{i}.fl_b"]))
        p1 = F.conv2d(f1, P[f"{pre}{i}.fp_w"][:, :, None, None])
        mem = F.conv2d(p1, P[f"{pre}{i}.fm_w"][:, None, None, :], padding=(0, c.dense_lorder - 1), groups=c.emb)
        out = out + (p1 + mem)
        skip = torch.cat((out, skip), dim=1)
    return out


def _mossformer(P: dict, pre: str, c: GanConfig, x0: torch.Tensor, Q: int, BT: int, b: int, dbg=None) -> torch.Tensor:
    """Inlined MossFormer block (`_mossformer_block`, :137-244): sequences of length Q, BT of them per window;
    local (within a sequence), cross-token (same position, across the BT sequences of a window, diagonal removed)
    and linear attention share the packed [v | u] value columns."""
    C, hidden, qk = c.emb, c.mf_hidden, c.mf_qk
    vdim, rot = hidden // 2, 2 * c.rot_freqs
    half = C // 2
    shifted = torch.cat((torch.zeros(x0.shape[0], 1, half), x0[:, :-1, :half]), dim=1)
    base = F.layer_norm(torch.cat((shifted, x0[..., half:]), dim=-1), (C,), None, None, 1e-5)
    huv = _dwconv_res(F.silu(F.linear(base, P[f"{pre}.in_w"], P[f"{pre}.in_b"])), P[f"{pre}.in_c"])
    hs, qkv = huv[..., :hidden], huv[..., hidden:]
    heads = qkv.unsqueeze(-2) * P[f"{pre}.gamma"] + P[f"{pre}.beta"]                 # (N, Q, 4, qk)
    cos, sin = P["rot_cos"][None, :Q, None, :], P["rot_sin"][None, :Q, None, :]
    tm = heads[..., :rot]
    swapped = tm.reshape(*tm.shape[:-1], rot // 2, 2).flip(-1).reshape(tm.shape)
    heads = torch.cat((tm * cos + swapped * sin, heads[..., rot:]), dim=-1)
    quad_q, lin_q, quad_k, lin_k = heads.unbind(dim=-2)
    # quad_k carries 1/Q; the cross-token similarity (normalised by BT upstream) corrects by Q/BT
    sim = torch.matmul(quad_q, quad_k.transpose(-1, -2))
    qq_c = quad_q.reshape(b, BT, Q, qk).transpose(1, 2)
    kk_c = quad_k.reshape(b, BT, Q, qk).transpose(1, 2)
    sim_c = torch.matmul(qq_c, kk_c.transpose(-1, -2)) * float(P[f"{pre}.cross_scale"])
    attn = F.relu(sim) ** 2
    attn_c = (F.relu(sim_c) ** 2).masked_fill(torch.eye(BT, dtype=torch.bool), 0.0)
    hs_c = hs.reshape(b, BT, Q, hidden).transpose(1, 2)
    att = torch.matmul(attn, hs) + torch.matmul(attn_c, hs_c).transpose(1, 2).reshape(b * BT, Q, hidden)
    att = att + torch.matmul(lin_q, torch.matmul(lin_k.transpose(-1, -2), hs))      # lin_k carries 1/Q
    out = (att[..., vdim:] * hs[..., :vdim]) * torch.sigmoid(att[..., :vdim] * hs[..., vdim:])
    ho = F.silu(F.linear(F.layer_norm(out, (vdim,), None, None, 1e-5), P[f"{pre}.out_w"], P[f"{pre}.out_b"]))
    if dbg is not None:
        dbg[f"{pre}.huv"], dbg[f"{pre}.att"] = huv, att
    return x0 + _dwconv_res(ho, P[f"{pre}.out_c"])


def _se(P: dict, pre: str, t: torch.Tensor) -> torch.Tensor:
    """SELayer: sigmoid MLPs of the global average and global maximum, summed (:689-696)."""
    def mlp(v, kind):
        return torch.sigmoid(F.linear(F.relu(F.linear(v, P[f"{pre}.se_{kind}0_w"], P[f"{pre}.se_{kind}0_b"])),
                                      P[f"{pre}.se_{kind}2_w"], P[f"{pre}.se_{kind}2_b"]))
    return (mlp(t.mean(dim=(2, 3)), "avg") + mlp(t.amax(dim=(2, 3)), "max"))[:, :, None, None] * t


def _norm_ch(x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    mu = x.mean(dim=1, keepdim=True)
    d = x - mu
    return d / torch.sqrt((d * d).mean(dim=1, keepdim=True) + eps)


def _path(P: dict, pre: str, c: GanConfig, seq: torch.Tensor, Q: int, BT: int, b: int, dbg=None) -> torch.Tensor:
    """Shared tail of the intra / inter paths on (N, steps, emb*ks): fused to_u||to_v, UniDeepFsmn on u, gate,
    ConvTranspose1d back to Q positions, MossFormer (:639-679)."""
    huv = F.layer_norm(seq, (c.path_in,), None, None, 1e-5)
    huv = _dwconv_res(F.silu(F.linear(huv, P[f"{pre}.uv_w"], P[f"{pre}.uv_b"])), P[f"{pre}.uv_c"])
    iu, iv = huv[..., :c.uv], huv[..., c.uv:]
    iu = _fsmn_seq(iu, P[f"{pre}.rl_w"], P[f"{pre}.rl_b"], P[f"{pre}.rp_w"], P[f"{pre}.rm_w"], c.lorder)
    t = F.conv_transpose1d((iv * iu).transpose(1, 2), P[f"{pre}.lin_w"], P[f"{pre}.lin_b"], stride=c.emb_hs).transpose(1, 2)
    if dbg is not None:
        dbg[f"{pre}.huv"], dbg[f"{pre}.lin"] = huv, t
    return _mossformer(P, f"{pre}.mf", c, t, Q, BT, b, dbg)


def _triple_attention(P: dict, pre: str, c: GanConfig, x: torch.Tensor) -> torch.Tensor:
    """Per-head softmax attention over time with (channel, freq) flattened features (:751-784); x (B, C, T, F)."""
    B, _, T, Fq = x.shape
    H, E, Vc = c.heads, c.attn_e, c.emb // c.heads
    qkv = F.prelu(F.conv2d(x, P[f"{pre}.w"][:, :, None, None], P[f"{pre}.b"]), P[f"{pre}.a"])
    qk = qkv[:, :2 * H * E].reshape(B, 2, H, E, T, Fq).permute(0, 1, 2, 4, 3, 5)                  # (B,2,H,T,E,F)
    qk = F.layer_norm(qk, (E, Fq), None, None, 1e-5)
    q = qk[:, 0] * P[f"{pre}.q_g"][None, :, None] + P[f"{pre}.q_b"][None, :, None]
    k = qk[:, 1] * P[f"{pre}.k_g"][None, :, None] + P[f"{pre}.k_b"][None, :, None]
    v = qkv[:, 2 * H * E:].reshape(B, H, Vc, T, Fq).permute(0, 1, 3, 2, 4)
    v = F.layer_norm(v, (Vc, Fq), None, None, 1e-5) * P[f"{pre}.v_g"][None, :, None] + P[f"{pre}.v_b"][None, :, None]
    a = F.softmax(torch.matmul(q.reshape(B, H, T, E * Fq), k.reshape(B, H, T, E * Fq).transpose(-1, -2)), dim=-1)
    o = torch.matmul(a, v.reshape(B, H, T, Vc * Fq)).reshape(B, H, T, Vc, Fq).permute(0, 1, 3, 2, 4).reshape(B, H * Vc, T, Fq)
    o = F.prelu(F.conv2d(o, P[f"{pre}.p_w"][:, :, None, None], P[f"{pre}.p_b"]), P[f"{pre}.p_a"])
    mu = o.mean(dim=(1, 3), keepdim=True)
    d = o - mu
    o = d / torch.sqrt((d * d).mean(dim=(1, 3), keepdim=True) + 1e-5)
    return o * P[f"{pre}.p_g"][None, :, None, :] + P[f"{pre}.p_beta"][None, :, None, :]


def model_len(length: int, in_rate: int = 16000) -> int:
    """MODEL_AUDIO_LENGTH (:37)."""
    return int(round(length * 16000 / in_rate))


def out_len(length: int, out_rate: int = 16000, in_rate: int = 16000) -> int:
    """OUTPUT_AUDIO_LENGTH (:38) -- the reference scales the INPUT length by out / model rate."""
    return model_len(length, in_rate) if out_rate == 16000 else int(round(length * out_rate / 16000))


def mfgan_forward(sd: dict, audio: torch.Tensor, c: GanConfig = GanConfig(), in_dtype: str = "F32", out_dtype: str = "F32",
                  dbg=None, folded: dict | None = None, in_rate: int = 16000, out_rate: int = 16000) -> torch.Tensor:
    """audio (B,1,L) in `in_dtype` ([-1,1] for float dtypes, :540-541) -> (B,1,L_out) in `out_dtype`; windows independent.
    in_rate / out_rate != 16 kHz: `F.interpolate(size=...)` behind the int16 lift (:542-549) and behind the x norm_factor
    (:884-891)."""
    L_in = audio.shape[-1]
    spec = SPECS["mossformergan_se_16k"]
    x = audio.float()
    if "int" not in in_dtype.lower():
        x = x * 32768.0
    if in_rate != 16000:
        x = F.interpolate(x, size=model_len(L_in, in_rate), mode="linear", align_corners=False)
    B, _, L = x.shape
    nf = torch.sqrt(torch.mean(x * x, dim=-1, keepdim=True) + 1e-6)
    x = x / nf
    pad = (c.hop - L % c.hop) % c.hop
    if pad:
        x = torch.cat((x, x[..., :pad]), dim=-1)                                       # wrap-around pad (:566-568)
    packed = stft_packed(spec, x)                                                      # (B, 2F, T)
    T = packed.shape[-1]
    P = folded if folded is not None else fold(sd, c, T)
    cin = packed.reshape(B, 2, c.n_bins, T)
    power = (cin * cin).sum(dim=1)
    mag = torch.pow(power, 0.15)
    comp = cin * torch.pow(power.clamp_min(torch.finfo(torch.float32).tiny), 0.15 - 0.5).unsqueeze(1)
    x = torch.cat((mag.unsqueeze(1), comp), dim=1).transpose(-1, -2)                   # (B, 3, T, F)

    x = F.conv2d(x, P["enc.c1_w"], P["enc.c1_b"])
    x = F.prelu(F.instance_norm(x, None, None, P["enc.n1_w"], P["enc.n1_b"], True, 0.1, 1e-5), P["enc.p1"])
    x = _dense_block(x, P, "enc.dd", c)
    x = F.conv2d(x, P["enc.c2_w"], P["enc.c2_b"], stride=(1, 2), padding=(0, 1))
    x = F.prelu(F.instance_norm(x, None, None, P["enc.n2_w"], P["enc.n2_b"], True, 0.1, 1e-5), P["enc.p2"])
    if dbg is not None:
        dbg["feat"], dbg["enc"] = torch.cat((mag.unsqueeze(1), comp), dim=1), x

    Fq = c.n_freqs
    for i in range(c.layers):
        o = f"B{i}"
        # intra path: sequences over frequency, one per (window, frame)
        t = F.conv2d(_norm_ch(x), P[f"{o}.intra.fconv_w"], P[f"{o}.intra.fconv_b"], groups=c.emb)     # (B, emb*ks, T, F-ks+1)
        t = t.permute(0, 2, 3, 1).reshape(B * T, Fq - c.emb_ks + 1, c.path_in)
        t = _path(P, f"{o}.intra", c, t, Fq, T, B, dbg)
        t = t.reshape(B, T, Fq, c.emb).permute(0, 3, 1, 2)
        t = _se(P, f"{o}.intra", t) + x
        # inter path: sequences over time, one per (window, sub-band)
        inp = t
        s = _norm_ch(inp).permute(0, 3, 1, 2).reshape(B * Fq, c.emb, T)
        s = F.conv1d(s, P[f"{o}.inter.unfold_w"], P[f"{o}.inter.unfold_b"], stride=c.emb_hs, groups=c.emb).transpose(1, 2)
        s = _path(P, f"{o}.inter", c, s, T, Fq, B, dbg)
        s = s.reshape(B, Fq, T, c.emb).permute(0, 3, 1, 2)                              # (B, C, F, T): SE runs in this layout
        inter = _se(P, f"{o}.inter", s).transpose(-1, -2) + inp
        x = _triple_attention(P, f"{o}.att", c, inter) + inter
        if dbg is not None:
            dbg[f"{o}.intra"], dbg[f"{o}.inter"], dbg[f"{o}.x"] = t, inter, x

    def upsample(y, pre):
        y = F.conv2d(y, P[f"{pre}.sp_w"], P[f"{pre}.sp_b"], padding=(0, 1))
        return y.reshape(B, c.sp_r, c.emb, T, Fq).permute(0, 2, 3, 4, 1).reshape(B, c.emb, T, Fq * c.sp_r)

    xm = upsample(_dense_block(x, P, "md.dd", c), "md")
    xm = F.conv2d(xm, P["md.c1_w"], P["md.c1_b"])
    xm = F.prelu(F.instance_norm(xm, None, None, P["md.nw"], P["md.nb"], True, 0.1, 1e-5), P["md.pa"])
    xm = F.conv2d(xm, P["md.fin_w"], P["md.fin_b"]).permute(0, 3, 2, 1).squeeze(-1)                  # (B, 201, T)
    mask = F.prelu(xm, P["md.pout"])
    xc = upsample(_dense_block(x, P, "cd.dd", c), "cd")
    xc = F.prelu(F.instance_norm(xc, None, None, P["cd.nw"], P["cd.nb"], True, 0.1, 1e-5), P["cd.pa"])
    cplx = F.conv2d(xc, P["cd.c_w"], P["cd.c_b"]).transpose(-1, -2)                                  # (B, 2, 201, T)

    fin = mask.unsqueeze(1) * comp + cplx
    fin = fin * torch.pow((fin * fin).sum(dim=1), float(0.5 / 0.3) - 0.5).unsqueeze(1)               # undo the 0.3 compression
    if dbg is not None:
        dbg["mask"], dbg["complex"], dbg["spec_out"] = mask, cplx, fin
    y = istft_packed(spec, fin.reshape(B, 2 * c.n_bins, T))[..., :L] * nf
    if out_rate != 16000:
        y = F.interpolate(y, size=out_len(L_in, out_rate, in_rate), mode="linear", align_corners=False)
    if "int" in out_dtype.lower():
        return y.clamp(min=-32768.0, max=32767.0).to(torch.int16)
    y = y * INV_INT16
    return y if "32" in out_dtype else y.to(torch.float16)


def mfgan_forward_batch(sd, audio, c: GanConfig = GanConfig(), in_dtype="F32", out_dtype="F32", chunk: int = 4,
                        in_rate: int = 16000, out_rate: int = 16000):
    outs = [mfgan_forward(sd, audio[s:s + chunk], c, in_dtype, out_dtype, in_rate=in_rate, out_rate=out_rate)
            for s in range(0, audio.shape[0], chunk)]
    return torch.cat(outs, dim=0)

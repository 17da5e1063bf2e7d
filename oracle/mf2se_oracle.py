"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference MossFormer2-SE-48K path.

Restates `MOSSFORMER_SE.__init__` (weight folds) and `MOSSFORMER_SE.forward`
(reference `MossFormer2_SE_48K/Export_MossFormer_SE.py:74-507`) as plain functions over a
flat `state_dict`.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline
legs may import it.

The reference wrapper reads its parameters from the un-vendored `clearvoice` package
(`clearvoice.models.mossformer2_se`, no pinned version).  Only attribute
*shapes* are needed: `skeleton()` below builds a parameter holder with exactly the attribute
paths the wrapper dereferences, and its `state_dict()` keys are this build's checkpoint
naming.  Two sub-modules are *called* by the wrapper rather than restated
(`mossformer.norm`, `mossformer.conv1d_encoder`, :348-349); they are taken here as
GroupNorm(1, 180, eps=1e-8) and a bias-free 1x1 Conv1d(180, 512), the speechbrain-style
`select_norm('ln')` / encoder the wrapper's own `F.group_norm(... intra_norm ...)` (:477)
implies.  For those two ops the parity is therefore self-referential.

Pinned (tests/test_oracle_pinning.py): against the reference wrapper itself, executed from
/root/reference around the skeleton on identical seeded weights (container only), and
against the committed fixtures tests/golden/mf2se_*.npz generated from that execution.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from stft_oracle import SPECS, forward_basis, istft_packed


@dataclass(frozen=True)
class Mf2Config:
    layers: int = 24
    dim: int = 512
    vu: int = 1024            # FLASH hidden (v and u each)
    qk: int = 128
    group: int = 256
    dw_kernel: int = 17
    fsmn_inner: int = 256
    lorder: int = 20
    rot_freqs: int = 16       # rotary over the first 2*16 channels
    n_mels: int = 60
    out_bins: int = 961
    num_spks: int = 2
    sample_rate: int = 48000
    nfft: int = 1920
    hop: int = 384

    @property
    def feat_in(self) -> int:
        return 3 * self.n_mels

    def n_frames(self, length: int) -> int:
        return (length - self.nfft) // self.hop + 1


INV_INT16 = float(1.0 / 32768.0)
LOG_INT16_POWER = float(2.0 * np.log(32768.0))


# ----------------------------------------------------------------------------- parameter holder
class _ScaleNorm(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.scale = dim ** -0.5
        self.eps = 1e-5
        self.g = nn.Parameter(torch.ones(1))


class _DwConv(nn.Module):
    """ConvModule: .sequential[1].conv is the depthwise Conv1d the wrapper reads."""

    def __init__(self, ch, k):
        super().__init__()
        holder = nn.Module()
        holder.conv = nn.Conv1d(ch, ch, k, padding=(k - 1) // 2, groups=ch, bias=False)
        self.sequential = nn.ModuleList([nn.Identity(), holder])


class _FFConvM(nn.Module):
    def __init__(self, din, dout, k, norm):
        super().__init__()
        self.mdl = nn.ModuleList([norm, nn.Linear(din, dout), nn.SiLU(), _DwConv(dout, k)])


class _OffsetScale(nn.Module):
    def __init__(self, dim, heads=4):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(heads, dim))
        self.beta = nn.Parameter(torch.zeros(heads, dim))


class _Rotary(nn.Module):
    def __init__(self, n):
        super().__init__()
        # lucidrains rotary 'lang' frequencies for dim = 2n
        self.freqs = nn.Parameter(1.0 / (10000 ** (torch.arange(0, 2 * n, 2).float() / (2 * n))), requires_grad=False)


class _Flash(nn.Module):
    def __init__(self, c: Mf2Config):
        super().__init__()
        self.group_size = c.group
        self.to_hidden = _FFConvM(c.dim, 2 * c.vu, c.dw_kernel, _ScaleNorm(c.dim))
        self.to_qk = _FFConvM(c.dim, c.qk, c.dw_kernel, _ScaleNorm(c.dim))
        self.to_out = _FFConvM(c.vu, c.dim, c.dw_kernel, _ScaleNorm(c.vu))
        self.qk_offset_scale = _OffsetScale(c.qk)
        self.rotary_pos_emb = _Rotary(c.rot_freqs)


class _UniDeepFsmn(nn.Module):
    def __init__(self, c: Mf2Config):
        super().__init__()
        d = c.fsmn_inner
        self.output_dim = d
        self.lorder = c.lorder
        self.linear = nn.Linear(d, d)
        self.project = nn.Linear(d, d, bias=False)
        self.conv1 = nn.Conv2d(d, d, [2 * c.lorder - 1, 1], [1, 1], groups=d, bias=False)


class _GatedFsmn(nn.Module):
    def __init__(self, c: Mf2Config):
        super().__init__()
        d = c.fsmn_inner
        self.to_u = _FFConvM(d, d, c.dw_kernel, nn.LayerNorm(d))
        self.to_v = _FFConvM(d, d, c.dw_kernel, nn.LayerNorm(d))
        self.fsmn = _UniDeepFsmn(c)


class _FsmnBlock(nn.Module):
    def __init__(self, c: Mf2Config):
        super().__init__()
        d = c.fsmn_inner
        self.conv1 = nn.Sequential(nn.Conv1d(c.dim, d, 1), nn.PReLU())
        self.norm1 = nn.LayerNorm(d)
        self.gated_fsmn = _GatedFsmn(c)
        self.norm2 = nn.LayerNorm(d)
        self.conv2 = nn.Conv1d(d, c.dim, 1)


class _PosEnc(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.scale = nn.Parameter(torch.ones(1))
        self.register_buffer("inv_freq", 1.0 / (10000 ** (torch.arange(0, dim, 2).float() / dim)))


def skeleton(c: Mf2Config = Mf2Config()) -> nn.Module:
    """Parameter holder with the attribute paths `MOSSFORMER_SE` dereferences (:76-283, :348-349,
    :358-362, :482-486).  Torch default initialisation."""
    m = nn.Module()
    m.pos_enc = _PosEnc(c.dim)
    m.norm = nn.GroupNorm(1, c.feat_in, eps=1e-8)
    m.conv1d_encoder = nn.Conv1d(c.feat_in, c.dim, 1, bias=False)
    core = nn.Module()
    core.layers = nn.ModuleList([_Flash(c) for _ in range(c.layers)])
    core.fsmn = nn.ModuleList([_FsmnBlock(c) for _ in range(c.layers)])
    intra = nn.Module()
    intra.mossformerM = core
    intra.norm = nn.LayerNorm(c.dim)
    m.mdl = nn.Module()
    m.mdl.intra_mdl = intra
    m.mdl.intra_norm = nn.GroupNorm(1, c.dim, eps=1e-8)
    m.prelu = nn.PReLU()
    m.conv1d_out = nn.Conv1d(c.dim, c.dim * c.num_spks, 1)
    m.output = nn.Sequential(nn.Conv1d(c.dim, c.dim, 1), nn.Tanh())
    m.output_gate = nn.Sequential(nn.Conv1d(c.dim, c.dim, 1), nn.Sigmoid())
    m.conv1_decoder = nn.Conv1d(c.dim, c.out_bins, 1, bias=False)
    return m


def random_state_dict(c: Mf2Config = Mf2Config(), seed: int = 0) -> dict[str, torch.Tensor]:
    """Seeded weights: default inits, then every gain / bias / offset-scale is perturbed so that
    each fold in `fold()` is exercised with non-trivial values."""
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    sd = {k: v.clone().float() for k, v in skeleton(c).state_dict().items()}
    for k, v in sd.items():
        if k.endswith("inv_freq") or k.endswith("freqs"):
            continue
        r = torch.randn(v.shape, generator=g)
        if k.endswith(".g") or "qk_offset_scale.gamma" in k or k.endswith("pos_enc.scale"):
            sd[k] = 1.0 + 0.25 * r
        elif "qk_offset_scale.beta" in k:
            sd[k] = 0.1 * r
        elif ("norm" in k or ".mdl.0." in k) and k.endswith("weight") and v.ndim == 1:
            sd[k] = 1.0 + 0.2 * r
        elif k.endswith("bias"):
            sd[k] = v + 0.05 * r
        elif k.startswith("prelu") or ".conv1.1." in k:
            sd[k] = 0.25 + 0.05 * r
    return sd


# ----------------------------------------------------------------------------- frontend tables
def kaldi_frontend_basis(win_len: int = 1920, preemph: float = 0.97) -> torch.Tensor:
    """(2*(P/2+1), win_len) fp32: DC removal, pre-emphasis (first sample replicated), symmetric
    Hamming window and the P-point real DFT (P = next power of two) as one matrix (:228-252)."""
    padded = 1 << (win_len - 1).bit_length()
    nb = padded // 2 + 1
    win = torch.hamming_window(win_len, periodic=False, alpha=0.54, beta=0.46, dtype=torch.float64)
    ang = (2.0 * torch.pi / padded) * torch.arange(nb, dtype=torch.float64)[:, None] * torch.arange(win_len, dtype=torch.float64)[None, :]
    re = torch.cos(ang) * win
    im = -torch.sin(ang) * win
    pre = torch.eye(win_len, dtype=torch.float64)
    pre[0, 0] -= preemph
    idx = torch.arange(1, win_len)
    pre[idx, idx - 1] -= preemph
    dc = torch.eye(win_len, dtype=torch.float64) - 1.0 / win_len
    lin = pre @ dc
    return torch.cat([re @ lin, im @ lin], dim=0).float()


def kaldi_mel_matrix(n_mels: int = 60, padded: int = 2048, fs: float = 48000.0, f_lo: float = 20.0) -> torch.Tensor:
    """(n_mels, padded/2+1) triangular Kaldi mel filters, last column zero (:254-275)."""
    n_bins = padded // 2
    width = fs / padded

    def mel(f):
        return 1127.0 * float(np.log(1.0 + f / 700.0))

    m_lo, m_hi = mel(f_lo), mel(0.5 * fs)
    step = (m_hi - m_lo) / (n_mels + 1)
    k = torch.arange(n_mels, dtype=torch.float64)[:, None]
    left, mid, right = m_lo + k * step, m_lo + (k + 1.0) * step, m_lo + (k + 2.0) * step
    m = (1127.0 * torch.log(1.0 + (width * torch.arange(n_bins, dtype=torch.float64)) / 700.0))[None, :]
    tri = torch.clamp(torch.minimum((m - left) / (mid - left), (right - m) / (right - mid)), min=0.0)
    return F.pad(tri, (0, 1)).float()


# ----------------------------------------------------------------------------- weight folds
def fold(sd: dict, c: Mf2Config, n_frames: int) -> dict[str, torch.Tensor]:
    """Raw state_dict -> the fused tensors the forward uses (restates :96-226).  Linear weights
    stay (N, K) row-major."""
    P: dict[str, torch.Tensor] = {}
    spec = SPECS["mossformer2_se_48k"]
    P["frontend"] = torch.cat([kaldi_frontend_basis(c.nfft), forward_basis(spec)], dim=0).contiguous()
    P["mel_banks"] = kaldi_mel_matrix(c.n_mels, 1 << (c.nfft - 1).bit_length(), float(c.sample_rate))
    P["norm.w"], P["norm.b"] = sd["norm.weight"].float(), sd["norm.bias"].float()
    P["enc.w"] = sd["conv1d_encoder.weight"][:, :, 0].float()
    t = torch.arange(n_frames, dtype=torch.float32)
    sinu = t[:, None] * sd["pos_enc.inv_freq"].float()
    emb = torch.cat((sinu.sin(), sinu.cos()), dim=-1) * sd["pos_enc.scale"].float()
    P["emb_pos"] = emb.half().float().contiguous()                       # (T, dim), fp16 storage round trip (:117)
    fr = sd["mdl.intra_mdl.mossformerM.layers.0.rotary_pos_emb.freqs"]
    ang = torch.arange(n_frames, dtype=fr.dtype)[:, None] * fr
    ang = torch.stack((ang, ang), dim=-1).flatten(-2)
    P["rot_cos"] = ang.cos().half().float().contiguous()                 # (T, 2*rot_freqs)
    P["rot_sin"] = ang.sin().half().float().contiguous()

    sn_in = 1.0 / (c.dim ** -0.5)                                        # ScaleNorm 1/scale folded into the weights
    sn_out = 1.0 / (c.vu ** -0.5)
    for i in range(c.layers):
        f = f"mdl.intra_mdl.mossformerM.layers.{i}"
        wh = sd[f"{f}.to_hidden.mdl.1.weight"].double() * sd[f"{f}.to_hidden.mdl.0.g"].double() * sn_in
        wq = sd[f"{f}.to_qk.mdl.1.weight"].double() * sd[f"{f}.to_qk.mdl.0.g"].double() * sn_in
        P[f"L{i}.in_w"] = torch.cat((wh, wq), 0).float().contiguous()
        P[f"L{i}.in_b"] = torch.cat((sd[f"{f}.to_hidden.mdl.1.bias"], sd[f"{f}.to_qk.mdl.1.bias"]), 0).float()
        P[f"L{i}.in_c"] = torch.cat((sd[f"{f}.to_hidden.mdl.3.sequential.1.conv.weight"],
                                     sd[f"{f}.to_qk.mdl.3.sequential.1.conv.weight"]), 0)[:, 0, :].float().contiguous()
        P[f"L{i}.out_w"] = (sd[f"{f}.to_out.mdl.1.weight"].double() * sd[f"{f}.to_out.mdl.0.g"].double() * sn_out).float()
        P[f"L{i}.out_b"] = sd[f"{f}.to_out.mdl.1.bias"].float()
        P[f"L{i}.out_c"] = sd[f"{f}.to_out.mdl.3.sequential.1.conv.weight"][:, 0, :].float().contiguous()
        head_scale = torch.ones(4, 1, dtype=torch.float64)
        head_scale[0, 0] = 1.0 / c.group                                 # quadratic query: 1/group_size
        head_scale[3, 0] = 1.0 / float(n_frames)                         # linear key: 1/n (static export)
        P[f"L{i}.qk_gamma"] = (sd[f"{f}.qk_offset_scale.gamma"].double() * head_scale).float()
        P[f"L{i}.qk_beta"] = (sd[f"{f}.qk_offset_scale.beta"].double() * head_scale).float()

        b = f"mdl.intra_mdl.mossformerM.fsmn.{i}"
        P[f"L{i}.c1_w"] = sd[f"{b}.conv1.0.weight"][:, :, 0].float()
        P[f"L{i}.c1_b"] = sd[f"{b}.conv1.0.bias"].float()
        P[f"L{i}.c1_a"] = sd[f"{b}.conv1.1.weight"].float()
        P[f"L{i}.n1_w"], P[f"L{i}.n1_b"] = sd[f"{b}.norm1.weight"].float(), sd[f"{b}.norm1.bias"].float()
        ws, bs, cs = [], [], []
        for br in ("to_u", "to_v"):
            lw, lb = sd[f"{b}.gated_fsmn.{br}.mdl.0.weight"].double(), sd[f"{b}.gated_fsmn.{br}.mdl.0.bias"].double()
            w = sd[f"{b}.gated_fsmn.{br}.mdl.1.weight"].double()
            ws.append(w * lw[None, :])
            bs.append(w @ lb + sd[f"{b}.gated_fsmn.{br}.mdl.1.bias"].double())
            cs.append(sd[f"{b}.gated_fsmn.{br}.mdl.3.sequential.1.conv.weight"][:, 0, :])
        P[f"L{i}.uv_w"] = torch.cat(ws, 0).float().contiguous()
        P[f"L{i}.uv_b"] = torch.cat(bs, 0).float()
        P[f"L{i}.uv_c"] = torch.cat(cs, 0).float().contiguous()
        P[f"L{i}.ul_w"] = sd[f"{b}.gated_fsmn.fsmn.linear.weight"].float()
        P[f"L{i}.ul_b"] = sd[f"{b}.gated_fsmn.fsmn.linear.bias"].float()
        P[f"L{i}.up_w"] = sd[f"{b}.gated_fsmn.fsmn.project.weight"].float()
        P[f"L{i}.mem_c"] = sd[f"{b}.gated_fsmn.fsmn.conv1.weight"][:, 0, :, 0].float().contiguous()
        P[f"L{i}.n2_w"], P[f"L{i}.n2_b"] = sd[f"{b}.norm2.weight"].float(), sd[f"{b}.norm2.bias"].float()
        P[f"L{i}.c2_w"] = sd[f"{b}.conv2.weight"][:, :, 0].float()
        P[f"L{i}.c2_b"] = sd[f"{b}.conv2.bias"].float()

    P["mm_norm.w"], P["mm_norm.b"] = sd["mdl.intra_mdl.norm.weight"].float(), sd["mdl.intra_mdl.norm.bias"].float()
    P["intra_norm.w"], P["intra_norm.b"] = sd["mdl.intra_norm.weight"].float(), sd["mdl.intra_norm.bias"].float()
    P["prelu_a"] = sd["prelu.weight"].float()
    spk_w = sd["conv1d_out.weight"][:c.dim, :, 0].double()               # speaker 0 only (:209-224)
    spk_b = sd["conv1d_out.bias"][:c.dim].double()
    gw = torch.cat((sd["output.0.weight"], sd["output_gate.0.weight"]), 0)[:, :, 0].double()
    gb = torch.cat((sd["output.0.bias"], sd["output_gate.0.bias"]), 0).double()
    P["gate_w"] = (gw @ spk_w).float().contiguous()
    P["gate_b"] = (gw @ spk_b + gb).float()
    P["dec_w"] = sd["conv1_decoder.weight"][:, :, 0].float()
    return P


# ----------------------------------------------------------------------------- forward
def _dwconv_res(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """x (B, T, C) + depthwise 'same' conv over T with taps w (C, k)."""
    k = w.shape[-1]
    return x + F.conv1d(x.transpose(1, 2), w.unsqueeze(1), None, padding=(k - 1) // 2, groups=w.shape[0]).transpose(1, 2)


def _deltas(x: torch.Tensor) -> torch.Tensor:
    """5-tap regression deltas along T with replicated edges (:304-310); x (B, C, T)."""
    n = 2
    taps = torch.arange(-n, n + 1, dtype=torch.float32).view(1, 1, -1) * float(1.0 / (n * (n + 1) * (2 * n + 1) / 3))
    b, ch, t = x.shape
    y = F.conv1d(F.pad(x.reshape(-1, 1, t), (n, n), mode="replicate"), taps)
    return y.reshape(b, ch, t)


def features(P: dict, c: Mf2Config, x: torch.Tensor):
    """x (B,1,L) normalised PCM -> (log-mel + deltas (B,180,T), STFT rows (B,1922,T))."""
    fr = F.conv1d(x, P["frontend"].unsqueeze(1), stride=c.hop)
    nk = P["frontend"].shape[0] - 2 * c.out_bins
    kal, st = fr[:, :nk], fr[:, nk:]
    sq = kal * kal
    power = sq[:, :nk // 2] + sq[:, nk // 2:]
    floor = float(torch.finfo(torch.float32).eps * INV_INT16 * INV_INT16)
    mel = torch.matmul(P["mel_banks"].unsqueeze(0), power).clamp(min=floor).log() + LOG_INT16_POWER
    d1 = _deltas(mel)
    d2 = _deltas(d1)
    return torch.cat([mel, d1, d2], dim=1), st


def flash_layer(P: dict, c: Mf2Config, i: int, h: torch.Tensor, dbg=None) -> torch.Tensor:
    """FLASH_ShareA_FFConvM with fused to_hidden||to_qk (:391-438); h (B, T, dim), T <= group."""
    B, T, D = h.shape
    half = D // 2
    shifted = torch.cat((torch.zeros(B, 1, half), h[:, :-1, :half]), dim=1)       # token shift of the first half
    xs = torch.cat((shifted, h[:, :, half:]), dim=-1)
    eps_in = 1e-5 / (c.dim ** -0.5)
    xs = xs / (torch.norm(xs, dim=-1, keepdim=True) + eps_in)
    proj = _dwconv_res(F.silu(F.linear(xs, P[f"L{i}.in_w"], P[f"L{i}.in_b"])), P[f"L{i}.in_c"])
    v, u, qk = torch.split(proj, [c.vu, c.vu, c.qk], dim=-1)
    vu = proj[..., :2 * c.vu]
    heads = qk.unsqueeze(-2) * P[f"L{i}.qk_gamma"] + P[f"L{i}.qk_beta"]            # (B,T,4,qk)
    r = 2 * c.rot_freqs
    mid = heads[..., :r]
    rot = torch.stack((-mid[..., 1::2], mid[..., 0::2]), dim=-1).flatten(-2)
    cos, sin = P["rot_cos"][None, :T, None, :], P["rot_sin"][None, :T, None, :]
    heads = torch.cat((mid * cos + rot * sin, heads[..., r:]), dim=-1)
    quad_q, lin_q, quad_k, lin_k = heads.unbind(dim=2)
    # one group (T <= group): zero-padded keys contribute nothing, padded queries are discarded
    attn = F.relu(torch.matmul(quad_q, quad_k.transpose(1, 2)))
    quad = torch.matmul(attn * attn, vu)
    kv = torch.matmul(lin_k.transpose(1, 2), vu)                                   # 1/n folded into lin_k
    att = quad + torch.matmul(lin_q, kv)
    att_v, att_u = torch.split(att, [c.vu, c.vu], dim=-1)
    gated = (att_u * v) * torch.sigmoid(att_v * u)
    eps_out = 1e-5 / (c.vu ** -0.5)
    y = gated / (torch.norm(gated, dim=-1, keepdim=True) + eps_out)
    y = _dwconv_res(F.silu(F.linear(y, P[f"L{i}.out_w"], P[f"L{i}.out_b"])), P[f"L{i}.out_c"])
    if dbg is not None:
        dbg[f"L{i}.proj"] = proj
        dbg[f"L{i}.att"] = att
        dbg[f"L{i}.gated"] = gated
    return h + y


def fsmn_layer(P: dict, c: Mf2Config, i: int, h: torch.Tensor, dbg=None) -> torch.Tensor:
    """Gated_FSMN_Block with fused to_u||to_v (:440-469); h (B, T, dim)."""
    d = c.fsmn_inner
    c1 = F.prelu(F.linear(h, P[f"L{i}.c1_w"], P[f"L{i}.c1_b"]), P[f"L{i}.c1_a"])
    g_in = F.layer_norm(c1, (d,), P[f"L{i}.n1_w"], P[f"L{i}.n1_b"], 1e-5)
    xn = F.layer_norm(g_in, (d,), None, None, 1e-5)
    uv = _dwconv_res(F.silu(F.linear(xn, P[f"L{i}.uv_w"], P[f"L{i}.uv_b"])), P[f"L{i}.uv_c"])
    xu, xv = torch.split(uv, [d, d], dim=-1)
    f1 = F.relu(F.linear(xu, P[f"L{i}.ul_w"], P[f"L{i}.ul_b"]))
    xp = F.linear(f1, P[f"L{i}.up_w"])
    mem = P[f"L{i}.mem_c"]                                                          # (d, 2*lorder-1), zero pad lorder-1 both sides
    conv = F.conv1d(xp.transpose(1, 2), mem.unsqueeze(1), None, padding=c.lorder - 1, groups=d).transpose(1, 2)
    xu = xu + xp + conv
    y = F.layer_norm(xv * xu + g_in, (d,), P[f"L{i}.n2_w"], P[f"L{i}.n2_b"], 1e-5)
    if dbg is not None:
        dbg[f"L{i}.uv"] = uv
        dbg[f"L{i}.y"] = y
    return F.linear(y, P[f"L{i}.c2_w"], P[f"L{i}.c2_b"]) + h


def mf2se_core(P: dict, c: Mf2Config, feats: torch.Tensor, dbg=None) -> torch.Tensor:
    """(B,180,T) features -> (B,T,out_bins) non-negative mask (:348-486)."""
    z = F.group_norm(feats, 1, P["norm.w"], P["norm.b"], 1e-8)
    z = (F.conv1d(z, P["enc.w"].unsqueeze(-1)).transpose(1, 2) + P["emb_pos"][None, :feats.shape[-1]]).contiguous()
    h = z
    if dbg is not None:
        dbg["z"] = z
    for i in range(c.layers):
        h = flash_layer(P, c, i, h, dbg)
        if dbg is not None:
            dbg[f"L{i}.flash"] = h
        h = fsmn_layer(P, c, i, h, dbg)
        if dbg is not None:
            dbg[f"L{i}.h"] = h
    h = F.layer_norm(h, (c.dim,), P["mm_norm.w"], P["mm_norm.b"], 1e-5)
    h = F.group_norm(h.transpose(1, 2), 1, P["intra_norm.w"], P["intra_norm.b"], 1e-8).transpose(1, 2) + z
    h = F.prelu(h, P["prelu_a"])
    gate = F.linear(h, P["gate_w"], P["gate_b"])
    t = torch.tanh(gate[..., :c.dim]) * torch.sigmoid(gate[..., c.dim:])
    if dbg is not None:
        dbg["tail"] = t
    return F.relu(F.linear(t, P["dec_w"]))


def model_len(length: int, in_rate: int, c: Mf2Config = Mf2Config()) -> int:
    """MODEL_AUDIO_LENGTH (:48): the window length at the 48 kHz model rate."""
    return int(round(length * c.sample_rate / in_rate))


def mf2se_forward(sd: dict, audio: torch.Tensor, c: Mf2Config = Mf2Config(), in_dtype: str = "F32",
                  out_dtype: str = "F32", dbg=None, folded: dict | None = None, in_rate: int | None = None,
                  out_rate: int | None = None) -> torch.Tensor:
    """audio (B,1,L) in `in_dtype` -> (B,1,L_out) in `out_dtype`; every window independent.  in_rate / out_rate != 48 kHz:
    linear resampling to MODEL_AUDIO_LENGTH behind the PCM scale (:318-325) and to OUTPUT_AUDIO_LENGTH behind the
    ISTFT (:491-498)."""
    B, _, L_in = audio.shape
    in_rate, out_rate = in_rate or c.sample_rate, out_rate or c.sample_rate
    x = audio.float()
    if "int" in in_dtype.lower():
        x = x * INV_INT16
    if in_rate != c.sample_rate:
        x = F.interpolate(x, size=model_len(L_in, in_rate, c), mode="linear", align_corners=False)
    L = x.shape[-1]
    T = c.n_frames(L)
    if T > c.group:
        raise ValueError("restatement covers one FLASH group (frames <= group_size)")
    P = folded if folded is not None else fold(sd, c, T)
    feats, st = features(P, c, x)
    mask = mf2se_core(P, c, feats, dbg)                                             # (B,T,bins)
    masked = (st.reshape(B, 2, c.out_bins, T) * mask.transpose(1, 2).unsqueeze(1)).reshape(B, 2 * c.out_bins, T)
    if dbg is not None:
        dbg["feats"], dbg["stft"], dbg["mask"] = feats, st, mask
    y = istft_packed(SPECS["mossformer2_se_48k"], masked)
    if out_rate != c.sample_rate:
        y = F.interpolate(y, size=int(round(L_in * out_rate / in_rate)), mode="linear", align_corners=False)
    if "int" in out_dtype.lower():
        y = y.clamp(min=-1.0, max=32767.0 / 32768.0) * 32768.0                     # int32 staging cast (:499-504)
        return y.to(torch.int32).clamp(min=-32768, max=32767).to(torch.int16)
    return y if "32" in out_dtype else y.to(torch.float16)


def mf2se_forward_batch(sd, audio, c: Mf2Config = Mf2Config(), in_dtype="F32", out_dtype="F32", chunk: int = 8):
    """Same as `mf2se_forward`, folds computed once and the batch walked in slices (CPU baseline)."""
    T = c.n_frames(audio.shape[-1])
    P = fold(sd, c, T)
    outs = [mf2se_forward(sd, audio[s:s + chunk], c, in_dtype, out_dtype, folded=P) for s in range(0, audio.shape[0], chunk)]
    return torch.cat(outs, dim=0)


def flops_per_window(c: Mf2Config, length: int) -> float:
    """Dense-contraction FLOPs (2*MAC) of one window: linear layers, attention, DFT front/back."""
    T = c.n_frames(length)
    lin = c.dim * (2 * c.vu + c.qk) + c.vu * c.dim + c.dim * c.fsmn_inner * 2 + 2 * c.fsmn_inner * c.fsmn_inner + c.fsmn_inner * c.dim
    att = T * c.qk + T * 2 * c.vu + 2 * c.qk * 2 * c.vu
    tail = c.dim * 2 * c.dim + c.dim * c.out_bins + c.feat_in * c.dim
    dft = (2 * (1 << (c.nfft - 1).bit_length()) // 2 + 2 + 2 * c.out_bins) * c.nfft + 2 * c.out_bins * c.nfft
    return 2.0 * T * (c.layers * (lin + att) + tail + dft)

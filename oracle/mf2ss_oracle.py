"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference MossFormer2-SS-16K path
(two-speaker separation; BASELINE.json configs[4], SURVEY.md 8 rows a2/a7/a10/a12 for C5b).

Restates `MOSSFORMER_SS.__init__` (weight folds) and `norm_audio` / `_run_mdl` / `forward`
(reference `MossFormer2_SS_16K/Export_MossFormer2_SS_16K.py:84-662`) as plain functions over a
flat `state_dict`.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs
may import it.

The reference wrapper reads its parameters from the un-vendored `clearvoice` package
(`clearvoice.models.mossformer2_ss.mossformer2`, no pinned version) and then drops
the module (:394-396): every executed op is a leaf torch op on packed buffers.  `skeleton()`
builds a parameter holder with exactly the attribute paths the constructor dereferences
(:102-392); its `state_dict()` keys are this build's checkpoint naming.  Layer geometry follows
the shapes in the wrapper's own comments (:166-177, :275-276) and the upstream DilatedDenseNet
(depthwise (2*lorder-1, 1) Conv2d with `in*(j+1) -> in` channels, `groups = in`, dilation 2**j,
InstanceNorm2d(affine), PReLU(in)); the wrapper validates that geometry itself (:284-291).

Pinned (tests/test_oracle_pinning.py): against the reference wrapper executed from
/root/reference around the skeleton on identical seeded weights (container only), and against
the committed fixtures tests/golden/mf2ss_*.npz generated from that execution.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F

from mf2se_oracle import _Flash, _FFConvM, _PosEnc


@dataclass(frozen=True)
class SsConfig:
    layers: int = 24
    dim: int = 512
    vu: int = 1024
    qk: int = 128
    group: int = 256
    dw_kernel: int = 17
    fsmn_inner: int = 256
    lorder: int = 20
    mem_depth: int = 2
    rot_freqs: int = 16
    num_spks: int = 2
    enc_kernel: int = 16
    enc_stride: int = 8
    sample_rate: int = 16000

    def n_frames(self, length: int) -> int:
        return (length - self.enc_kernel) // self.enc_stride + 1

    def out_len(self, length: int) -> int:
        return (self.n_frames(length) - 1) * self.enc_stride + self.enc_kernel


INV_INT16 = float(1.0 / 32768.0)
NORM_FACTOR = float(10.0 ** (-25.0 / 20.0))


# ----------------------------------------------------------------------------- parameter holder
class _DilatedDense(nn.Module):
    def __init__(self, c: SsConfig):
        super().__init__()
        d, k = c.fsmn_inner, 2 * c.lorder - 1
        for j in range(c.mem_depth):
            setattr(self, f"conv{j + 1}", nn.Conv2d(d * (j + 1), d, (k, 1), dilation=(2 ** j, 1), groups=d, bias=False))
            setattr(self, f"norm{j + 1}", nn.InstanceNorm2d(d, affine=True))
            setattr(self, f"prelu{j + 1}", nn.PReLU(d))


class _UniDeepFsmnDilated(nn.Module):
    def __init__(self, c: SsConfig):
        super().__init__()
        d = c.fsmn_inner
        self.depth = c.mem_depth
        self.lorder = c.lorder
        self.linear = nn.Linear(d, d)
        self.project = nn.Linear(d, d, bias=False)
        self.conv = _DilatedDense(c)


class _GatedFsmnDilated(nn.Module):
    def __init__(self, c: SsConfig):
        super().__init__()
        d = c.fsmn_inner
        self.to_u = _FFConvM(d, d, c.dw_kernel, nn.LayerNorm(d))
        self.to_v = _FFConvM(d, d, c.dw_kernel, nn.LayerNorm(d))
        self.fsmn = _UniDeepFsmnDilated(c)


class _FsmnBlockDilated(nn.Module):
    def __init__(self, c: SsConfig):
        super().__init__()
        d = c.fsmn_inner
        self.conv1 = nn.Sequential(nn.Conv1d(c.dim, d, 1), nn.PReLU())
        self.norm1 = nn.LayerNorm(d)
        self.gated_fsmn = _GatedFsmnDilated(c)
        self.norm2 = nn.LayerNorm(d)
        self.conv2 = nn.Conv1d(d, c.dim, 1)


def skeleton(c: SsConfig = SsConfig()) -> nn.Module:
    """Parameter holder with the attribute paths `MOSSFORMER_SS.__init__` dereferences (:90-392)."""
    m = nn.Module()
    m.num_spks = c.num_spks
    m.enc = nn.Module()
    m.enc.conv1d = nn.Conv1d(1, c.dim, c.enc_kernel, stride=c.enc_stride, bias=False)
    m.dec = nn.ConvTranspose1d(c.dim, 1, c.enc_kernel, stride=c.enc_stride, bias=False)
    k = nn.Module()
    k.norm = nn.GroupNorm(1, c.dim, eps=1e-8)
    k.conv1d_encoder = nn.Conv1d(c.dim, c.dim, 1, bias=False)
    k.pos_enc = _PosEnc(c.dim)
    core = nn.Module()
    core.layers = nn.ModuleList([_Flash(c) for _ in range(c.layers)])
    core.fsmn = nn.ModuleList([_FsmnBlockDilated(c) for _ in range(c.layers)])
    intra = nn.Module()
    intra.mossformerM = core
    intra.norm = nn.LayerNorm(c.dim)
    k.mdl = nn.Module()
    k.mdl.intra_mdl = intra
    k.mdl.intra_norm = nn.GroupNorm(1, c.dim, eps=1e-8)
    k.prelu = nn.PReLU()
    k.conv1d_out = nn.Conv1d(c.dim, c.dim * c.num_spks, 1)
    k.output = nn.Sequential(nn.Conv1d(c.dim, c.dim, 1), nn.Tanh())
    k.output_gate = nn.Sequential(nn.Conv1d(c.dim, c.dim, 1), nn.Sigmoid())
    k.conv1_decoder = nn.Conv1d(c.dim, c.dim, 1, bias=False)
    m.mask_net = k
    return m


def random_state_dict(c: SsConfig = SsConfig(), seed: int = 0) -> dict[str, torch.Tensor]:
    """Seeded weights: default inits with every gain / bias / slope perturbed so each fold is exercised."""
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
        elif ".fsmn.conv.prelu" in k:
            sd[k] = 0.25 + 0.05 * r
        elif ("norm" in k or ".mdl.0." in k) and k.endswith("weight") and v.ndim == 1:
            sd[k] = 1.0 + 0.2 * r
        elif k.endswith("bias"):
            sd[k] = v + 0.05 * r
        elif k.startswith("mask_net.prelu") or ".conv1.1." in k:
            sd[k] = 0.25 + 0.05 * r
    return sd


# ----------------------------------------------------------------------------- weight folds
def fold(sd: dict, c: SsConfig, n_frames: int) -> dict[str, torch.Tensor]:
    """Raw state_dict -> the fused tensors the forward uses (restates :119-392)."""
    P: dict[str, torch.Tensor] = {}
    mn = "mask_net"
    P["enc_w"] = sd["enc.conv1d.weight"][:, 0, :].float().contiguous()                 # (dim, 16)
    P["enc_b"] = sd["enc.conv1d.bias"].float() if "enc.conv1d.bias" in sd else torch.zeros(c.dim)
    P["dec_w"] = sd["dec.weight"][:, 0, :].float().contiguous()                        # (dim, 16)
    P["dec_b"] = sd["dec.bias"].float() if "dec.bias" in sd else torch.zeros(1)
    fw = sd[f"{mn}.conv1d_encoder.weight"].double()
    P["front_w"] = (fw * sd[f"{mn}.norm.weight"].double().reshape(1, -1, 1)).float()[:, :, 0].contiguous()   # :222-224
    fb = fw.squeeze(-1) @ sd[f"{mn}.norm.bias"].double()
    if f"{mn}.conv1d_encoder.bias" in sd:
        fb = fb + sd[f"{mn}.conv1d_encoder.bias"].double()
    P["front_b"] = fb.float()
    t = torch.arange(n_frames, dtype=torch.float32)
    sinu = t.unsqueeze(-1) * sd[f"{mn}.pos_enc.inv_freq"].float()
    P["emb_pos"] = (torch.cat((sinu.sin(), sinu.cos()), dim=-1) * sd[f"{mn}.pos_enc.scale"].float()).contiguous()   # fp32 (:156-162)
    core = f"{mn}.mdl.intra_mdl.mossformerM"
    fr = sd[f"{core}.layers.0.rotary_pos_emb.freqs"]
    ang = torch.arange(n_frames, dtype=fr.dtype).unsqueeze(-1) * fr
    ang = torch.stack((ang, ang), dim=-1).flatten(-2)
    P["rot_cos"] = ang.cos().contiguous()                                              # fp32 (:198-206)
    P["rot_sin"] = ang.sin().contiguous()

    sn_in = float(1.0 / (c.dim ** -0.5))
    sn_out = float(1.0 / (c.vu ** -0.5))
    for i in range(c.layers):
        f = f"{core}.layers.{i}"
        wh = sd[f"{f}.to_hidden.mdl.1.weight"].double() * sd[f"{f}.to_hidden.mdl.0.g"].double() * sn_in
        wq = sd[f"{f}.to_qk.mdl.1.weight"].double() * sd[f"{f}.to_qk.mdl.0.g"].double() * sn_in
        P[f"L{i}.in_w"] = torch.cat((wh, wq), 0).float().contiguous()
        P[f"L{i}.in_b"] = torch.cat((sd[f"{f}.to_hidden.mdl.1.bias"], sd[f"{f}.to_qk.mdl.1.bias"]), 0).float()
        P[f"L{i}.in_c"] = torch.cat((sd[f"{f}.to_hidden.mdl.3.sequential.1.conv.weight"],
                                     sd[f"{f}.to_qk.mdl.3.sequential.1.conv.weight"]), 0)[:, 0, :].float().contiguous()
        P[f"L{i}.out_w"] = (sd[f"{f}.to_out.mdl.1.weight"].double() * sd[f"{f}.to_out.mdl.0.g"].double() * sn_out).float()
        P[f"L{i}.out_b"] = sd[f"{f}.to_out.mdl.1.bias"].float()
        P[f"L{i}.out_c"] = sd[f"{f}.to_out.mdl.3.sequential.1.conv.weight"][:, 0, :].float().contiguous()
        hs = torch.ones(4, 1, dtype=torch.float64)
        hs[0, 0] = float(1.0 / c.group)
        hs[3, 0] = float(1.0 / n_frames)
        P[f"L{i}.qk_gamma"] = (sd[f"{f}.qk_offset_scale.gamma"].double() * hs).float()
        P[f"L{i}.qk_beta"] = (sd[f"{f}.qk_offset_scale.beta"].double() * hs).float()

        b = f"{core}.fsmn.{i}"
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
        fs = f"{b}.gated_fsmn.fsmn"
        P[f"L{i}.ul_w"] = sd[f"{fs}.linear.weight"].float()
        P[f"L{i}.ul_b"] = sd[f"{fs}.linear.bias"].float()
        P[f"L{i}.up_w"] = sd[f"{fs}.project.weight"].float()
        for j in range(c.mem_depth):
            P[f"L{i}.mem{j}_w"] = sd[f"{fs}.conv.conv{j + 1}.weight"][:, :, :, 0].float().contiguous()   # (d, j+1, 39)
            P[f"L{i}.mem{j}_nw"] = sd[f"{fs}.conv.norm{j + 1}.weight"].float()
            P[f"L{i}.mem{j}_nb"] = sd[f"{fs}.conv.norm{j + 1}.bias"].float()
            P[f"L{i}.mem{j}_a"] = sd[f"{fs}.conv.prelu{j + 1}.weight"].float()
        P[f"L{i}.n2_w"], P[f"L{i}.n2_b"] = sd[f"{b}.norm2.weight"].float(), sd[f"{b}.norm2.bias"].float()
        P[f"L{i}.c2_w"] = sd[f"{b}.conv2.weight"][:, :, 0].float()
        P[f"L{i}.c2_b"] = sd[f"{b}.conv2.bias"].float()

    P["mm_norm.w"], P["mm_norm.b"] = sd[f"{mn}.mdl.intra_mdl.norm.weight"].float(), sd[f"{mn}.mdl.intra_mdl.norm.bias"].float()
    P["intra_norm.w"], P["intra_norm.b"] = sd[f"{mn}.mdl.intra_norm.weight"].float(), sd[f"{mn}.mdl.intra_norm.bias"].float()
    P["prelu_a"] = sd[f"{mn}.prelu.weight"].float()
    gw = torch.cat((sd[f"{mn}.output.0.weight"], sd[f"{mn}.output_gate.0.weight"]), 0)[:, :, 0].double()
    gb = torch.cat((sd[f"{mn}.output.0.bias"], sd[f"{mn}.output_gate.0.bias"]), 0).double()
    tw, tb = [], []
    for s in range(c.num_spks):                                                        # :381-389
        sw = sd[f"{mn}.conv1d_out.weight"][s * c.dim:(s + 1) * c.dim, :, 0].double()
        sb = sd[f"{mn}.conv1d_out.bias"][s * c.dim:(s + 1) * c.dim].double()
        tw.append((gw @ sw).float())
        tb.append((gw @ sb + gb).float())
    P["gate_w"] = torch.cat(tw, 0).contiguous()                                        # (spks*2*dim, dim)
    P["gate_b"] = torch.cat(tb, 0).contiguous()
    P["mask_w"] = sd[f"{mn}.conv1_decoder.weight"][:, :, 0].float()
    return P


# ----------------------------------------------------------------------------- forward
def norm_audio(x: torch.Tensor, eps: float = 1e-6):
    """Two-stage per-window RMS normalisation (:403-423).  x (B,1,L) int16-scale -> (x_norm, rms_in (B,1,1))."""
    x = x * INV_INT16
    p = x * x
    avg = p.mean(dim=(1, 2), keepdim=True)
    rms = torch.sqrt(avg)
    scalar = NORM_FACTOR / (rms + eps)
    m = (p > avg).to(p.dtype)
    high = torch.sqrt((p * m).sum(dim=(1, 2), keepdim=True) / m.sum(dim=(1, 2), keepdim=True).clamp(min=1.0))
    scalarx = NORM_FACTOR / (high * scalar + eps)
    x = (x * scalar) * scalarx
    gp = scalar * scalarx
    undo = 1.0 / (gp + eps)
    return x, rms * gp * undo * 32767.0


def _dwconv_res(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    k = w.shape[-1]
    return x + F.conv1d(x.transpose(1, 2), w.unsqueeze(1), None, padding=(k - 1) // 2, groups=w.shape[0]).transpose(1, 2)


def flash_layer(P: dict, c: SsConfig, i: int, h: torch.Tensor, dbg=None) -> torch.Tensor:
    """FLASH_ShareA_FFConvM (:461-514): grouped quadratic + global linear attention; h (B, n, dim)."""
    B, n, D = h.shape
    half = D // 2
    shifted = torch.cat((torch.zeros(B, 1, half), h[:, :-1, :half]), dim=1)
    xs = torch.cat((shifted, h[:, :, half:]), dim=-1)
    eps_in = float(1e-5 / (c.dim ** -0.5))
    xs = xs / torch.clamp(torch.norm(xs, dim=-1, keepdim=True), min=eps_in)        # clamp, not add (:467)
    proj = _dwconv_res(F.silu(F.linear(xs, P[f"L{i}.in_w"], P[f"L{i}.in_b"])), P[f"L{i}.in_c"])
    vu, qk = proj[..., :2 * c.vu], proj[..., 2 * c.vu:]
    v, u = vu[..., :c.vu], vu[..., c.vu:]
    heads = qk.unsqueeze(-2) * P[f"L{i}.qk_gamma"] + P[f"L{i}.qk_beta"]
    r = 2 * c.rot_freqs
    mid = heads[..., :r]
    rot = torch.stack((-mid[..., 1::2], mid[..., 0::2]), dim=-1).flatten(-2)
    cos, sin = P["rot_cos"][None, :n, None, :], P["rot_sin"][None, :n, None, :]
    heads = torch.cat((mid * cos + rot * sin, heads[..., r:]), dim=-1)
    G = c.group
    pad = (G - n % G) % G
    ng = (n + pad) // G
    heads = F.pad(heads, (0, 0, 0, 0, 0, pad)).reshape(B, ng, G, 4, c.qk)
    vug = F.pad(vu, (0, 0, 0, pad)).reshape(B, ng, G, 2 * c.vu)
    quad_q, lin_q, quad_k, lin_k = heads.unbind(dim=3)
    attn = F.relu(torch.matmul(quad_q, quad_k.transpose(-1, -2)))
    quad = torch.matmul(attn * attn, vug)
    kv = torch.matmul(lin_k.reshape(B, ng * G, c.qk).transpose(1, 2), vug.reshape(B, ng * G, 2 * c.vu))   # 1/n folded into lin_k
    lin = torch.matmul(lin_q, kv.unsqueeze(1))
    att = (quad + lin).reshape(B, ng * G, 2 * c.vu)[:, :n]
    att_v, att_u = att[..., :c.vu], att[..., c.vu:]
    gated = (att_u * v) * torch.sigmoid(att_v * u)
    eps_out = float(1e-5 / (c.vu ** -0.5))
    y = gated / torch.clamp(torch.norm(gated, dim=-1, keepdim=True), min=eps_out)
    y = _dwconv_res(F.silu(F.linear(y, P[f"L{i}.out_w"], P[f"L{i}.out_b"])), P[f"L{i}.out_c"])
    if dbg is not None:
        dbg[f"L{i}.proj"], dbg[f"L{i}.att"], dbg[f"L{i}.gated"] = proj, att, gated
    return h + y


def fsmn_layer(P: dict, c: SsConfig, i: int, h: torch.Tensor, dbg=None) -> torch.Tensor:
    """Gated_FSMN_Block_Dilated (:516-550); h (B, n, dim)."""
    d = c.fsmn_inner
    c1 = F.prelu(F.linear(h, P[f"L{i}.c1_w"], P[f"L{i}.c1_b"]), P[f"L{i}.c1_a"])
    g_in = F.layer_norm(c1, (d,), P[f"L{i}.n1_w"], P[f"L{i}.n1_b"], 1e-5)
    xn = F.layer_norm(g_in, (d,), None, None, 1e-5)
    uv = _dwconv_res(F.silu(F.linear(xn, P[f"L{i}.uv_w"], P[f"L{i}.uv_b"])), P[f"L{i}.uv_c"])
    xu, xv = uv[..., :d], uv[..., d:]
    f1 = F.relu(F.linear(xu, P[f"L{i}.ul_w"], P[f"L{i}.ul_b"]))
    dense = F.linear(f1, P[f"L{i}.up_w"]).transpose(1, 2)                              # (B, d, n)
    mem = None
    for j in range(c.mem_depth):
        dil = 2 ** j
        padj = c.lorder + (dil - 1) * (c.lorder - 1) - 1
        mem = F.conv1d(dense, P[f"L{i}.mem{j}_w"], None, padding=padj, dilation=dil, groups=d)
        mem = F.instance_norm(mem, None, None, P[f"L{i}.mem{j}_nw"], P[f"L{i}.mem{j}_nb"], True, 0.1, 1e-5)
        mem = F.prelu(mem, P[f"L{i}.mem{j}_a"])
        if dbg is not None:
            dbg[f"L{i}.mem{j}"] = mem.transpose(1, 2)
        if j + 1 < c.mem_depth:
            dense = torch.cat((mem, dense), dim=1)
    xu = xu + mem.transpose(1, 2)
    y = F.layer_norm(xv * xu + g_in, (d,), P[f"L{i}.n2_w"], P[f"L{i}.n2_b"], 1e-5)
    if dbg is not None:
        dbg[f"L{i}.uv"], dbg[f"L{i}.y"] = uv, y
    return F.linear(y, P[f"L{i}.c2_w"], P[f"L{i}.c2_b"]) + h


def model_len(length: int, in_rate: int, c: SsConfig = SsConfig()) -> int:
    """MODEL_AUDIO_LENGTH (:36): the window length at the model rate."""
    return int(round(length * c.sample_rate / in_rate))


def mf2ss_forward(sd: dict, audio: torch.Tensor, c: SsConfig = SsConfig(), in_dtype: str = "F32", out_dtype: str = "F32",
                  dbg=None, folded: dict | None = None, in_rate: int | None = None, out_rate: int | None = None):
    """audio (B,1,L) int16-SCALE samples in `in_dtype` (:411 scales by 1/32768 whatever the dtype) ->
    tuple of `num_spks` tensors (B,1,L_out); every window independent.  in_rate / out_rate != 16 kHz: linear
    resampling to MODEL_AUDIO_LENGTH in front (:564-579) and to OUTPUT_AUDIO_LENGTH behind the gain (:633-648)."""
    B, _, L_in = audio.shape
    in_rate, out_rate = in_rate or c.sample_rate, out_rate or c.sample_rate
    x = audio.float()
    if in_rate != c.sample_rate:
        x = F.interpolate(x, size=model_len(L_in, in_rate, c), mode="linear", align_corners=False)
    L = x.shape[-1]
    n = c.n_frames(L)
    P = folded if folded is not None else fold(sd, c, n)
    x, rms_in = norm_audio(x)
    x_enc = F.relu(F.conv1d(x, P["enc_w"].unsqueeze(1), P["enc_b"], stride=c.enc_stride))          # (B, dim, n)
    z = F.group_norm(x_enc, 1, None, None, 1e-8)
    z = (F.conv1d(z, P["front_w"].unsqueeze(-1), P["front_b"]).transpose(1, 2) + P["emb_pos"][None, :n]).contiguous()
    h = z
    if dbg is not None:
        dbg["x"], dbg["x_enc"], dbg["z"] = x, x_enc.transpose(1, 2), z
    for i in range(c.layers):
        h = flash_layer(P, c, i, h, dbg)
        if dbg is not None:
            dbg[f"L{i}.flash"] = h
        h = fsmn_layer(P, c, i, h, dbg)
        if dbg is not None:
            dbg[f"L{i}.h"] = h
    h = F.layer_norm(h, (c.dim,), P["mm_norm.w"], P["mm_norm.b"], 1e-5)
    h = F.group_norm(h.transpose(1, 2), 1, P["intra_norm.w"], P["intra_norm.b"], 1e-8).transpose(1, 2) + z
    h = F.leaky_relu(h, negative_slope=float(P["prelu_a"]))
    gate = F.linear(h, P["gate_w"], P["gate_b"]).reshape(B, n, c.num_spks, 2 * c.dim)
    t = torch.tanh(gate[..., :c.dim]) * torch.sigmoid(gate[..., c.dim:])                             # (B, n, spk, dim)
    mask = F.relu(F.linear(t, P["mask_w"]))
    sep = (x_enc.transpose(1, 2).unsqueeze(2) * mask).permute(0, 2, 3, 1).reshape(B * c.num_spks, c.dim, n)
    wav = F.conv_transpose1d(sep, P["dec_w"].unsqueeze(1), P["dec_b"], stride=c.enc_stride).reshape(B, c.num_spks, -1)
    rms_out = torch.sqrt((wav * wav).mean(dim=2, keepdim=True))
    gain = torch.where(rms_out > 0.0, rms_in / rms_out, torch.zeros_like(rms_out))
    out = wav * gain
    if out_rate != c.sample_rate:
        size = int(round(L_in * out_rate / in_rate))                                                  # OUTPUT_AUDIO_LENGTH (:37)
        out = F.interpolate(out.reshape(1, B * c.num_spks, -1), size=size, mode="linear", align_corners=False).reshape(B, c.num_spks, -1)
    if dbg is not None:
        dbg["tail"], dbg["mask"], dbg["wav"] = t, mask, wav
    if "int" in out_dtype.lower():
        out = out.to(torch.int32).clamp(min=-32768, max=32767).to(torch.int16)
    else:
        out = out * INV_INT16
        if "16" in out_dtype:
            out = out.to(torch.float16)
    return tuple(out[:, s:s + 1].contiguous() for s in range(c.num_spks))


def mf2ss_forward_batch(sd, audio, c: SsConfig = SsConfig(), in_dtype="F32", out_dtype="F32", chunk: int = 2,
                        in_rate: int | None = None, out_rate: int | None = None):
    P = fold(sd, c, c.n_frames(model_len(audio.shape[-1], in_rate or c.sample_rate, c)))
    outs = [mf2ss_forward(sd, audio[s:s + chunk], c, in_dtype, out_dtype, folded=P, in_rate=in_rate, out_rate=out_rate)
            for s in range(0, audio.shape[0], chunk)]
    return tuple(torch.cat([o[s] for o in outs], dim=0) for s in range(c.num_spks))


def flops_per_window(c: SsConfig, length: int) -> float:
    """Dense-contraction FLOPs (2*MAC) of one window: linear layers + attention + encoder/decoder/tail."""
    n = c.n_frames(length)
    G = c.group
    ng = (n + G - 1) // G
    lin = c.dim * (2 * c.vu + c.qk) + c.vu * c.dim + c.dim * c.fsmn_inner * 2 + 2 * c.fsmn_inner * c.fsmn_inner + c.fsmn_inner * c.dim
    att_mac = ng * (G * G * c.qk + G * G * 2 * c.vu) + 2 * c.qk * (ng * G) * 2 * c.vu
    tail = c.dim * c.dim + c.dim * c.num_spks * 2 * c.dim + c.num_spks * c.dim * c.dim + (1 + c.num_spks) * c.dim * c.enc_kernel
    return 2.0 * (n * (c.layers * lin + tail) + c.layers * att_mac)

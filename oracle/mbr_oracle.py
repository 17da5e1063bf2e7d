"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference Mel-Band-Roformer (stereo) path.

Restates `MelBandRoformer.forward` / `_core` and the weight fusions of `__init__`
(reference `Mel_Band_Roformer/Stereo/Export_MelBandRoformer.py:262-680`) as plain functions
over a raw checkpoint-shaped `state_dict`.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s CPU-baseline legs may import it.

Pinned (tests/test_oracle_pinning.py): against the reference's own module executed from
/root/reference on identical seeded weights (container only) and against the committed
fixture tests/golden/mbr_*.npz generated from that execution (oracle/make_golden.py).
The reference has no golden vectors of its own for this path; the upstream
YAML is absent, so the hyper-parameters are the ones the in-file comments imply.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

from stft_oracle import SPECS, istft_packed, stft_packed


@dataclass(frozen=True)
class MbrConfig:
    dim: int = 384
    depth: int = 6
    heads: int = 8
    dim_head: int = 64
    num_bands: int = 60
    sample_rate: int = 44100
    nfft: int = 2048
    hop: int = 441
    stereo: bool = True
    mlp_expansion_factor: int = 4

    @property
    def channels(self) -> int:
        return 2 if self.stereo else 1

    @property
    def num_freqs(self) -> int:
        return self.nfft // 2 + 1

    @property
    def dim_inner(self) -> int:
        return self.heads * self.dim_head


# ----------------------------------------------------------------------------- mel band layout
def _hz_to_mel(f):
    """Slaney mel scale (Export_MelBandRoformer.py:68-88, htk=False branch)."""
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if f.ndim:
        hi = f >= min_log_hz
        mels[hi] = min_log_mel + np.log(f[hi] / min_log_hz) / logstep
    elif f >= min_log_hz:
        mels = min_log_mel + np.log(f / min_log_hz) / logstep
    return mels


def _mel_to_hz(m):
    """Export_MelBandRoformer.py:91-110."""
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if m.ndim:
        hi = m >= min_log_mel
        freqs[hi] = min_log_hz * np.exp(logstep * (m[hi] - min_log_mel))
    elif m >= min_log_mel:
        freqs = min_log_hz * np.exp(logstep * (m - min_log_mel))
    return freqs


def mel_filter_bank(sr, n_fft, n_mels):
    """`create_mel_filter_bank` (Export_MelBandRoformer.py:119-142), slaney norm, fp32."""
    fmax = float(sr) / 2
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    fftfreqs = np.linspace(0, fmax, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights.astype(np.float32, copy=False)


def band_layout(cfg: MbrConfig):
    """freq_indices (over the (freq,chan)-interleaved axis), dim_inputs, per-source averaging
    scale -- Export_MelBandRoformer.py:350-368, 472-473."""
    fb = torch.from_numpy(mel_filter_bank(cfg.sample_rate, cfg.nfft, cfg.num_bands))
    fb[0][0] = 1.0
    fb[-1, -1] = 1.0
    per_band = fb > 0
    idx = torch.arange(cfg.num_freqs).expand(cfg.num_bands, -1)[per_band]
    ch = cfg.channels
    if cfg.stereo:
        idx = (idx.unsqueeze(1).expand(-1, ch) * 2 + torch.arange(ch)).flatten()
    n_per_band = per_band.sum(dim=1)
    bands_per_freq = per_band.sum(dim=0)
    denom = 1.0 / bands_per_freq.repeat_interleave(ch).clamp(min=1e-8).double()
    dim_inputs = tuple(int(2 * f * ch) for f in n_per_band.tolist())
    denom_val = denom[idx.long()].repeat_interleave(2)         # per (source, re/im) value-branch scale
    return idx.to(torch.int64), dim_inputs, denom_val


# ----------------------------------------------------------------------------- weights
def state_dict_shapes(cfg: MbrConfig) -> dict:
    """Checkpoint key names / shapes of the holder module `m` (Export_MelBandRoformer.py:333-389)."""
    _, dim_inputs, _ = band_layout(cfg)
    d, di, h = cfg.dim, cfg.dim_inner, cfg.heads
    hid = d * cfg.mlp_expansion_factor
    s = {}
    for i in range(cfg.depth):
        for j in (0, 1):
            p = f"layers.{i}.{j}"
            s[f"{p}.layers.0.0.norm.gamma"] = (d,)
            s[f"{p}.layers.0.0.to_qkv.weight"] = (3 * di, d)
            s[f"{p}.layers.0.0.to_gates.weight"] = (h, d)
            s[f"{p}.layers.0.0.to_gates.bias"] = (h,)
            s[f"{p}.layers.0.0.to_out.0.weight"] = (d, di)
            s[f"{p}.layers.0.1.net.0.gamma"] = (d,)
            s[f"{p}.layers.0.1.net.1.weight"] = (4 * d, d)
            s[f"{p}.layers.0.1.net.1.bias"] = (4 * d,)
            s[f"{p}.layers.0.1.net.4.weight"] = (d, 4 * d)
            s[f"{p}.layers.0.1.net.4.bias"] = (d,)
            s[f"{p}.norm.gamma"] = (d,)
    for b, din in enumerate(dim_inputs):
        s[f"band_split.to_features.{b}.0.gamma"] = (din,)
        s[f"band_split.to_features.{b}.1.weight"] = (d, din)
        s[f"band_split.to_features.{b}.1.bias"] = (d,)
        q = f"mask_estimators.0.to_freqs.{b}.0"
        s[f"{q}.0.weight"] = (hid, d)
        s[f"{q}.0.bias"] = (hid,)
        s[f"{q}.2.weight"] = (hid, hid)
        s[f"{q}.2.bias"] = (hid,)
        s[f"{q}.4.weight"] = (2 * din, hid)
        s[f"{q}.4.bias"] = (2 * din,)
    return s


def random_state_dict(cfg: MbrConfig, seed: int = 0) -> dict:
    """Seeded synthetic weights (no checkpoint exists): nn.Linear-style uniform
    init, gammas around 1."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in state_dict_shapes(cfg).items():
        if name.endswith("gamma"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = shape[1] if len(shape) > 1 else shape[0]
            bound = 1.0 / np.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        sd[name] = t.float().contiguous()
    return sd


def fuse(sd: dict, cfg: MbrConfig) -> dict:
    """The weight fusions of `MelBandRoformer.__init__` (:455-531): RMSNorm gains folded into the
    consuming Linear (float64 product, stored fp32), attention scale folded into the Q rows,
    scatter-average denominator folded into the GLU value rows."""
    _, dim_inputs, denom_val = band_layout(cfg)
    d, di = cfg.dim, cfg.dim_inner
    out = {}
    for b, din in enumerate(dim_inputs):
        g = (din ** 0.5) * sd[f"band_split.to_features.{b}.0.gamma"].double()
        out[f"bs_w_{b}"] = (sd[f"band_split.to_features.{b}.1.weight"].double() * g.unsqueeze(0)).float().contiguous()
        out[f"bs_b_{b}"] = sd[f"band_split.to_features.{b}.1.bias"].float().contiguous()
    scale = cfg.dim_head ** -0.5
    for i in range(cfg.depth):
        for j, kind in ((0, "time"), (1, "freq")):
            p = f"layers.{i}.{j}"
            a = f"{p}.layers.0.0"
            f = f"{p}.layers.0.1"
            g_in = (d ** 0.5) * sd[f"{a}.norm.gamma"].double()
            wqkv = sd[f"{a}.to_qkv.weight"].double()
            wq, wk, wv = wqkv[:di], wqkv[di:2 * di], wqkv[2 * di:3 * di]
            wg = sd[f"{a}.to_gates.weight"].double()
            n = f"{kind}{i}"
            out[f"{n}_in_w"] = (torch.cat([wq * scale, wk, wv, wg], dim=0) * g_in.unsqueeze(0)).float().contiguous()
            out[f"{n}_in_b"] = torch.cat([torch.zeros(3 * di, dtype=torch.float64),
                                          sd[f"{a}.to_gates.bias"].double()]).float().contiguous()
            out[f"{n}_out_w"] = sd[f"{a}.to_out.0.weight"].float().contiguous()
            g_ff = (d ** 0.5) * sd[f"{f}.net.0.gamma"].double()
            out[f"{n}_ff1_w"] = (sd[f"{f}.net.1.weight"].double() * g_ff.unsqueeze(0)).float().contiguous()
            out[f"{n}_ff1_b"] = sd[f"{f}.net.1.bias"].float().contiguous()
            out[f"{n}_ff2_w"] = sd[f"{f}.net.4.weight"].float().contiguous()
            out[f"{n}_ff2_b"] = sd[f"{f}.net.4.bias"].float().contiguous()
            out[f"{n}_out_g"] = ((d ** 0.5) * sd[f"{p}.norm.gamma"].double()).float().contiguous()
    off = 0
    w1, b1, w2, b2 = [], [], [], []
    for b, din in enumerate(dim_inputs):
        q = f"mask_estimators.0.to_freqs.{b}.0"
        w1.append(sd[f"{q}.0.weight"]); b1.append(sd[f"{q}.0.bias"])
        w2.append(sd[f"{q}.2.weight"]); b2.append(sd[f"{q}.2.bias"])
        dv = denom_val[off:off + din]
        off += din
        w3 = sd[f"{q}.4.weight"].double().clone()
        b3 = sd[f"{q}.4.bias"].double().clone()
        w3[:din] *= dv.unsqueeze(1)
        b3[:din] *= dv
        out[f"me_w3_{b}"] = w3.float().contiguous()
        out[f"me_b3_{b}"] = b3.float().contiguous()
    out["me_w1t"] = torch.stack(w1, 0).transpose(1, 2).float().contiguous()
    out["me_b1"] = torch.stack(b1, 0).unsqueeze(1).float().contiguous()
    out["me_w2t"] = torch.stack(w2, 0).transpose(1, 2).float().contiguous()
    out["me_b2"] = torch.stack(b2, 0).unsqueeze(1).float().contiguous()
    return out


def rotary_tables(cfg: MbrConfig, n_frames: int):
    """Interleaved-pair rotary tables with the GPT-J sign folded into sin (:371-378, 438-452).
    The time tables go through an fp16 round trip, the freq tables stay fp32 (App. C.13)."""
    table_len = max(int(n_frames), int(cfg.num_bands))
    pos = torch.arange(table_len, dtype=torch.float32).unsqueeze(-1)
    inv_freq = 10000.0 ** -(torch.arange(0, cfg.dim_head, 2, dtype=torch.float32) / cfg.dim_head)
    rot = torch.repeat_interleave(pos * inv_freq, repeats=2, dim=-1)          # (len, dim_head)
    cos, sin = torch.cos(rot), torch.sin(rot)
    sign = torch.ones(cfg.dim_head)
    sign[0::2] = -1.0
    tcos = cos.half()[:n_frames].float()
    tsin = (sin.half().float() * sign)[:n_frames]
    fcos = cos[:cfg.num_bands]
    fsin = (sin[:cfg.num_bands] * sign)
    return tcos.contiguous(), tsin.contiguous(), fcos.contiguous(), fsin.contiguous()


# ----------------------------------------------------------------------------- forward
def _normalize(x):
    """:533-538."""
    n = torch.linalg.vector_norm(x, ord=2, dim=-1, keepdim=True)
    return x / torch.maximum(n, torch.tensor(1e-12))


def _rotate_half(x):
    """pair swap [1,0,3,2,...] (:449-452, 540-543)."""
    idx = torch.arange(x.shape[-1]).reshape(-1, 2).flip(1).flatten()
    return torch.index_select(x, -1, idx)


def _attention(cfg, x, p, rcos, rsin, b, n):
    """:545-563."""
    qkvg = F.linear(_normalize(x), p["in_w"], p["in_b"])
    di = cfg.dim_inner
    qkv_flat, gates = qkvg.split([3 * di, cfg.heads], dim=-1)
    qkv = qkv_flat.reshape(b, n, 3, cfg.heads, cfg.dim_head).permute(2, 0, 3, 1, 4)
    qk, v = qkv.split([2, 1], dim=0)
    qk = qk * rcos + _rotate_half(qk) * rsin
    q, k = qk.unbind(dim=0)
    v = v.squeeze(0)
    attn = torch.matmul(q, k.transpose(-1, -2)).softmax(dim=-1)
    out = torch.matmul(attn, v).transpose(1, 2)
    out = (out * gates.unsqueeze(-1).sigmoid()).reshape(b, n, di)
    return F.linear(out, p["out_w"])


def _transformer(cfg, x, p, rcos, rsin, b, n):
    """:568-571."""
    x = x + _attention(cfg, x, p, rcos, rsin, b, n)
    h = F.gelu(F.linear(_normalize(x), p["ff1_w"], p["ff1_b"]))
    x = x + F.linear(h, p["ff2_w"], p["ff2_b"])
    return _normalize(x) * p["out_g"]


def mbr_core(cfg: MbrConfig, fw: dict, stft_repr: torch.Tensor, dbg: dict | None = None):
    """`_core` (:585-627). stft_repr (B*chan, F, T, 2) -> masked (real, imag) (B*chan, F, T)."""
    idx, dim_inputs, _ = band_layout(cfg)
    ch, nf, nb, d = cfg.channels, cfg.num_freqs, cfg.num_bands, cfg.dim
    t = stft_repr.shape[-2]
    B = stft_repr.shape[0] // ch
    fc = nf * ch
    rep = stft_repr.reshape(B, ch, nf, t, 2).transpose(1, 2).reshape(B, fc, t, 2)
    x = torch.index_select(rep, 1, idx).transpose(1, 2).reshape(B, t, idx.numel() * 2)
    parts = x.split(dim_inputs, dim=-1)
    x = torch.stack([F.linear(_normalize(parts[i]), fw[f"bs_w_{i}"], fw[f"bs_b_{i}"]) for i in range(nb)], dim=0)
    if dbg is not None:
        dbg["band_split"] = x                                   # (nb, B, t, d)
    tcos, tsin, fcos, fsin = rotary_tables(cfg, t)
    for i in range(cfg.depth):
        pt = {k: fw[f"time{i}_{k}"] for k in ("in_w", "in_b", "out_w", "ff1_w", "ff1_b", "ff2_w", "ff2_b", "out_g")}
        pf = {k: fw[f"freq{i}_{k}"] for k in ("in_w", "in_b", "out_w", "ff1_w", "ff1_b", "ff2_w", "ff2_b", "out_g")}
        x = x.reshape(nb * B, t, d)
        x = _transformer(cfg, x, pt, tcos, tsin, nb * B, t)
        if dbg is not None:
            dbg[f"time{i}"] = x.reshape(nb, B, t, d)
        x = x.reshape(nb, B, t, d).permute(2, 1, 0, 3).reshape(t * B, nb, d)
        x = _transformer(cfg, x, pf, fcos, fsin, t * B, nb)
        x = x.reshape(t, B, nb, d).permute(2, 1, 0, 3)
        if dbg is not None:
            dbg[f"freq{i}"] = x                                  # (nb, B, t, d)
    xm = x.reshape(nb, B * t, d)
    h = torch.tanh(torch.baddbmm(fw["me_b1"], xm, fw["me_w1t"]))
    h = torch.tanh(torch.baddbmm(fw["me_b2"], h, fw["me_w2t"]))
    outs = [F.glu(F.linear(h[i], fw[f"me_w3_{i}"], fw[f"me_b3_{i}"]), dim=-1) for i in range(nb)]
    masks = torch.cat(outs, dim=-1).view(B, t, idx.numel(), 2).transpose(1, 2)
    if dbg is not None:
        dbg["masks"] = masks
    base = torch.zeros(B, fc, t, 2)
    avg = base.scatter_add_(1, idx.view(1, -1, 1, 1).expand(B, -1, t, 2), masks)
    re_in, im_in = rep.split(1, dim=-1)
    mr, mi = avg.split(1, dim=-1)
    out_r = re_in * mr - im_in * mi
    out_i = re_in * mi + im_in * mr
    real = out_r.reshape(B, nf, ch, t).permute(0, 2, 1, 3).reshape(B * ch, nf, t)
    imag = out_i.reshape(B, nf, ch, t).permute(0, 2, 1, 3).reshape(B * ch, nf, t)
    return real, imag


def mbr_forward(cfg: MbrConfig, fw: dict, audio: torch.Tensor, in_dtype="F32", out_dtype="F32",
                dbg: dict | None = None, in_rate: int = 44100, out_rate: int = 44100) -> torch.Tensor:
    """`MelBandRoformer.forward` for ONE window, no fold (:629-680).  audio (1, chan, W) -> (1, chan, W_out).
    in_rate / out_rate != 44.1 kHz: F.interpolate(scale_factor=...) on the raw input (:631-644); on the output before the
    x32767 PCM scale when down-sampling and after it when up-sampling (:662-675)."""
    spec = SPECS["mel_band_roformer"]
    x = audio.float()
    if in_rate != 44100:
        x = F.interpolate(x, scale_factor=float(44100 / in_rate), mode="linear", align_corners=False)
    x = x.squeeze(0).unsqueeze(1).contiguous()                                 # (chan,1,W)
    s = stft_packed(spec, x, input_scale=(1.0 / 32768.0) if "int" in in_dtype.lower() else 1.0)
    nf = cfg.num_freqs
    rep = torch.stack((s[:, :nf], s[:, nf:]), dim=-1)                          # (chan,F,T,2)
    if dbg is not None:
        dbg["stft"] = rep
    real, imag = mbr_core(cfg, fw, rep, dbg)
    y = istft_packed(spec, torch.cat((real, imag), dim=1)).transpose(0, 1).contiguous()
    scale = float(out_rate / 44100)
    if out_rate < 44100:
        y = F.interpolate(y, scale_factor=scale, mode="linear", align_corners=False)
    if "int" in out_dtype.lower():
        y = y * 32767.0
    if out_rate > 44100:
        y = F.interpolate(y, scale_factor=scale, mode="linear", align_corners=False)
    if "int" in out_dtype.lower():
        return y.clamp(min=-32768.0, max=32767.0).to(torch.int16)
    return y


def mbr_forward_batch(cfg, fw, audio, in_dtype="F32", out_dtype="F32", in_rate: int = 44100, out_rate: int = 44100):
    return torch.cat([mbr_forward(cfg, fw, audio[i:i + 1], in_dtype, out_dtype, in_rate=in_rate, out_rate=out_rate)
                      for i in range(audio.shape[0])], dim=0)

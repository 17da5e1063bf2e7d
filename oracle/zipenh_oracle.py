"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference ZipEnhancer path
(BASELINE.json configs[1]: ZipEnhancer 16 kHz, 64 x 1 s windows, fp32).

Restates the weight folds of `ZipEnhancer.__init__` and its `forward` with every block below it
(reference `ZipEnhancer/Export_ZipEnhancer.py:118-339` forward overrides, `:357-699` folds, `:701-927`
dense encoder / dual-path encoders / decoders / forward) as plain functions over a flat `state_dict`.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import it.  The CUDA path it
checks is csrc/zipenh_ops.cuh + csrc/zipenh.cu (model family `zipenhancer`); `dbg` collects the stage dumps
the host-harness test and the GPU stage test compare against.

Third-party arithmetic.  The reference wrapper drives the un-vendored `modelscope` package
(`modelscope.models.audio.ans.zipenhancer_layers.{scaling, zipformer, generator}`, no version pinned
anywhere in the reference).  Every forward with real arithmetic in it is OVERRIDDEN in the
reference file itself (BiasNorm, ActivationDropoutAndLinear, Zipformer2EncoderLayer, BypassModule,
SimpleDownsample / SimpleUpsample, RelPositionMultiheadAttentionWeights, SelfAttention, NonlinAttention,
ConvolutionModule, `:118-355`) or inlined into the wrapper (dense blocks, decoders, dual-path plumbing).
What the reference takes from modelscope unchanged is: constructors (i.e. layer SHAPES), `FeedforwardModule.forward`
(in_proj -> out_proj, the Balancer / Whiten diagnostics being identities in eval), `torch.nn` leaves, and the
`CompactRelPositionalEncoding.pe` table.  `skeleton()` builds a parameter holder with exactly the attribute
paths the wrapper dereferences; `ref_loader.load_zipenh` registers these skeleton classes under the modelscope
module names, lets the reference's own `apply_onnx_export_patches()` install ITS forwards on them and executes the
reference wrapper around the holder.  Hyper-parameters (`ZipConfig`) are the published
speech_zipenhancer_ans_multiloss_16k_base ones as far as the reference constrains them (4 heads and a fused
attention projection of 112 = 4 x (12 + 12 + 4) columns: `ZipEnhancer/Optimize_ONNX.py:70-71`; F' = 101 sub-bands
so that the x2 sub-pixel decoder + (1,2) conv return 201 bins; four encoders of one f-layer + one t-layer,
the middle two down-sampled, `:577-592, :863-867`) and otherwise the upstream defaults (64 channels,
feed-forward 256, depthwise kernel 15, value head 12, pos_dim 24): parity is SELF-REFERENTIAL in those free
dimensions and in the `pe` table construction (icefall's CompactRelPositionalEncoding restated in `compact_rel_pe`).

Pinned (tests/test_oracle_pinning.py): against the reference wrapper executed from /root/reference around the
skeleton on identical seeded weights (container only), and against the committed fixtures
tests/golden/zipenh_*.npz generated from that execution.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F

import ends_oracle


@dataclass(frozen=True)
class ZipConfig:
    channels: int = 64
    heads: int = 4
    query_head_dim: int = 12
    pos_head_dim: int = 4
    value_head_dim: int = 12
    pos_dim: int = 24
    ff_dim: int = 256
    conv_kernel: int = 15
    downsample: tuple = (1, 2, 2, 1)     # time and frequency factor of each of the four encoders
    dense_depth: int = 4
    up_factor: int = 2                   # sub-pixel upscale of the decoders
    nfft: int = 400
    hop: int = 100
    pe_max_len: int = 1000

    @property
    def n_bins(self) -> int:
        return self.nfft // 2 + 1

    @property
    def n_sub(self) -> int:             # sub-bands after the (1,3) stride-2 conv with pad 1
        return (self.n_bins + 2 - 3) // 2 + 1

    @property
    def nonlin_hidden(self) -> int:
        return 3 * self.channels // 4

    def n_frames(self, length: int) -> int:
        return length // self.hop + 1

    def ff_dims(self):
        return (self.ff_dim * 3 // 4, self.ff_dim, self.ff_dim * 5 // 4)


SWOOSH_L_OFFSET = 0.035
SWOOSH_R_OFFSET = 0.313261687


def compact_rel_pe(embed_dim: int, max_len: int) -> torch.Tensor:
    """icefall / modelscope `CompactRelPositionalEncoding.extend_pe` (length_factor 1): rows are the relative offsets
    -(max_len-1) .. max_len-1; cos / sin of atan(compressed offset) harmonics, last column 1 (bias)."""
    x = torch.arange(-(max_len - 1), max_len, dtype=torch.float32).unsqueeze(1)
    freqs = 1 + torch.arange(embed_dim // 2)
    clen = embed_dim ** 0.5
    xc = clen * x.sign() * ((x.abs() + clen).log() - math.log(clen))
    lscale = embed_dim / (2.0 * math.pi)
    xa = (xc / lscale).atan()
    pe = torch.zeros(x.shape[0], embed_dim)
    pe[:, 0::2] = (xa * freqs).cos()
    pe[:, 1::2] = (xa * freqs).sin()
    pe[:, -1] = 1.0
    return pe


# ----------------------------------------------------------------------------- parameter holder (shapes only)
class _Id(nn.Module):
    def forward(self, x):
        return x


class BiasNorm(nn.Module):
    def __init__(self, num_channels, channel_dim=-1):
        super().__init__()
        self.num_channels, self.channel_dim = num_channels, channel_dim
        self.log_scale = nn.Parameter(torch.tensor(1.0))
        self.bias = nn.Parameter(torch.zeros(num_channels))


class ActivationDropoutAndLinear(nn.Module):
    def __init__(self, cin, cout, activation):
        super().__init__()
        self.activation = activation
        lin = nn.Linear(cin, cout)
        self.weight, self.bias = lin.weight, lin.bias


class BypassModule(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.bypass_scale = nn.Parameter(torch.full((c,), 0.5))


class SimpleDownsample(nn.Module):
    def __init__(self, ds):
        super().__init__()
        self.downsample = ds
        self.bias = nn.Parameter(torch.zeros(ds))


class SimpleUpsample(nn.Module):
    def __init__(self, us):
        super().__init__()
        self.upsample = us


class CompactRelPositionalEncoding(nn.Module):
    def __init__(self, embed_dim, max_len):
        super().__init__()
        self.embed_dim = embed_dim
        self.pe = compact_rel_pe(embed_dim, max_len)


class RelPositionMultiheadAttentionWeights(nn.Module):
    def __init__(self, c: ZipConfig):
        super().__init__()
        self.num_heads, self.query_head_dim, self.pos_head_dim = c.heads, c.query_head_dim, c.pos_head_dim
        self.in_proj = nn.Linear(c.channels, (2 * c.query_head_dim + c.pos_head_dim) * c.heads)
        self.linear_pos = nn.Linear(c.pos_dim, c.heads * c.pos_head_dim, bias=False)


class SelfAttention(nn.Module):
    def __init__(self, c: ZipConfig):
        super().__init__()
        self.in_proj = nn.Linear(c.channels, c.heads * c.value_head_dim)
        self.out_proj = nn.Linear(c.heads * c.value_head_dim, c.channels)
        self.whiten = _Id()


class NonlinAttention(nn.Module):
    def __init__(self, c: ZipConfig):
        super().__init__()
        self.hidden_channels = c.nonlin_hidden
        self.in_proj = nn.Linear(c.channels, 3 * self.hidden_channels)
        self.balancer, self.whiten1, self.whiten2 = _Id(), _Id(), _Id()
        self.tanh = nn.Tanh()
        self.out_proj = nn.Linear(self.hidden_channels, c.channels)


class ConvolutionModule(nn.Module):
    def __init__(self, c: ZipConfig):
        super().__init__()
        self.in_proj = nn.Linear(c.channels, 2 * c.channels)
        self.balancer1, self.activation1, self.activation2, self.balancer2, self.whiten = _Id(), _Id(), _Id(), _Id(), _Id()
        self.sigmoid = nn.Sigmoid()
        self.depthwise_conv = nn.Conv1d(c.channels, c.channels, c.conv_kernel, groups=c.channels, padding=c.conv_kernel // 2)
        self.out_proj = ActivationDropoutAndLinear(c.channels, c.channels, "SwooshR")


class FeedforwardModule(nn.Module):
    def __init__(self, cin, hidden):
        super().__init__()
        self.in_proj = nn.Linear(cin, hidden)
        self.hidden_balancer, self.out_whiten = _Id(), _Id()
        self.out_proj = ActivationDropoutAndLinear(hidden, cin, "SwooshL")

    def forward(self, x):        # modelscope's own forward (not overridden by the reference): in_proj -> out_proj
        return self.out_whiten(self.out_proj(self.hidden_balancer(self.in_proj(x))))


class Zipformer2EncoderLayer(nn.Module):
    def __init__(self, c: ZipConfig):
        super().__init__()
        f1, f2, f3 = c.ff_dims()
        self.bypass, self.bypass_mid = BypassModule(c.channels), BypassModule(c.channels)
        self.self_attn_weights = RelPositionMultiheadAttentionWeights(c)
        self.self_attn1, self.self_attn2 = SelfAttention(c), SelfAttention(c)
        self.feed_forward1 = FeedforwardModule(c.channels, f1)
        self.feed_forward2 = FeedforwardModule(c.channels, f2)
        self.feed_forward3 = FeedforwardModule(c.channels, f3)
        self.nonlin_attention = NonlinAttention(c)
        self.conv_module1, self.conv_module2 = ConvolutionModule(c), ConvolutionModule(c)
        self.norm = BiasNorm(c.channels)


class DualPathZipformer2Encoder(nn.Module):
    def __init__(self, c: ZipConfig):
        super().__init__()
        self.encoder_pos = CompactRelPositionalEncoding(c.pos_dim, c.pe_max_len)
        self.f_layers = nn.ModuleList([Zipformer2EncoderLayer(c)])
        self.t_layers = nn.ModuleList([Zipformer2EncoderLayer(c)])
        self.bypass_layers = nn.ModuleList([BypassModule(c.channels), BypassModule(c.channels)])


class DualPathDownsampledZipformer2Encoder(nn.Module):
    def __init__(self, c: ZipConfig, ds: int):
        super().__init__()
        self.t_downsample_factor = self.f_downsample_factor = ds
        self.encoder = DualPathZipformer2Encoder(c)
        self.downsample_t, self.downsample_f = SimpleDownsample(ds), SimpleDownsample(ds)
        self.upsample_t, self.upsample_f = SimpleUpsample(ds), SimpleUpsample(ds)
        self.out_combiner = BypassModule(c.channels)


def _conv_norm_act(cin, cout, k, stride=(1, 1), padding=(0, 0), dilation=(1, 1), pad=None):
    mods = [] if pad is None else [nn.ConstantPad2d(pad, 0.0)]
    mods += [nn.Conv2d(cin, cout, k, stride, padding=padding, dilation=dilation), nn.InstanceNorm2d(cout, affine=True), nn.PReLU(cout)]
    return nn.Sequential(*mods)


class _DenseBlock(nn.Module):
    def __init__(self, c: ZipConfig):
        super().__init__()
        self.dense_block = nn.ModuleList(
            [_conv_norm_act(c.channels * (i + 1), c.channels, (2, 3), dilation=(2 ** i, 1), pad=(1, 1, 2 ** i, 0)) for i in range(c.dense_depth)])


class _SubPixel(nn.Module):
    def __init__(self, c: ZipConfig):
        super().__init__()
        self.upscale_width_factor = c.up_factor
        self.conv1 = nn.Conv2d(c.channels, c.channels * c.up_factor, (1, 3), padding=(0, 1))


class _DenseEncoder(nn.Module):
    def __init__(self, c: ZipConfig):
        super().__init__()
        self.dense_conv_1 = _conv_norm_act(2, c.channels, (1, 1))
        self.dense_block = _DenseBlock(c)
        self.dense_conv_2 = _conv_norm_act(c.channels, c.channels, (1, 3), (1, 2), padding=(0, 1))


class _MaskDecoder(nn.Module):
    def __init__(self, c: ZipConfig):
        super().__init__()
        self.dense_block = _DenseBlock(c)
        self.relu = nn.ReLU()
        self.mask_conv = nn.Sequential(_SubPixel(c), nn.InstanceNorm2d(c.channels, affine=True), nn.PReLU(c.channels),
                                       nn.Conv2d(c.channels, 1, (1, 2)))


class _PhaseDecoder(nn.Module):
    def __init__(self, c: ZipConfig):
        super().__init__()
        self.dense_block = _DenseBlock(c)
        self.phase_conv = nn.Sequential(_SubPixel(c), nn.InstanceNorm2d(c.channels, affine=True), nn.PReLU(c.channels))
        self.phase_conv_r = nn.Conv2d(c.channels, 1, (1, 2))
        self.phase_conv_i = nn.Conv2d(c.channels, 1, (1, 2))


class _TSConformer(nn.Module):
    def __init__(self, c: ZipConfig):
        super().__init__()
        self.encoders = nn.ModuleList(
            [DualPathZipformer2Encoder(c) if ds == 1 else DualPathDownsampledZipformer2Encoder(c, ds) for ds in c.downsample])


class _ZipEnhancerHolder(nn.Module):
    def __init__(self, c: ZipConfig):
        super().__init__()
        self.dense_encoder = _DenseEncoder(c)
        self.TSConformer = _TSConformer(c)
        self.mask_decoder = _MaskDecoder(c)
        self.phase_decoder = _PhaseDecoder(c)


def skeleton(cfg: ZipConfig = ZipConfig(), seed: int = 0) -> nn.Module:
    """Parameter holder with seeded, non-degenerate values in every parameter the folds touch."""
    torch.manual_seed(seed)
    m = _ZipEnhancerHolder(cfg).eval()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, BypassModule):
                mod.bypass_scale.copy_(0.2 + 0.7 * torch.rand(mod.bypass_scale.shape, generator=g))
            elif isinstance(mod, BiasNorm):
                mod.bias.copy_(0.1 * torch.randn(mod.bias.shape, generator=g))
                mod.log_scale.copy_(0.3 * torch.randn((), generator=g))
            elif isinstance(mod, SimpleDownsample):
                mod.bias.copy_(torch.randn(mod.bias.shape, generator=g))
            elif isinstance(mod, nn.InstanceNorm2d):
                mod.weight.copy_(1.0 + 0.2 * torch.randn(mod.weight.shape, generator=g))
                mod.bias.copy_(0.1 * torch.randn(mod.bias.shape, generator=g))
            elif isinstance(mod, nn.PReLU):
                mod.weight.copy_(0.25 + 0.1 * torch.randn(mod.weight.shape, generator=g))
        # a mask around |X|^0.3 of a unit-RMS window, so the synthetic model's output is at signal level (the parity
        # tolerance is absolute) instead of the near-silence a zero-mean random mask head produces
        m.mask_decoder.mask_conv[3].bias.fill_(1.6)
    return m


def random_state_dict(cfg: ZipConfig = ZipConfig(), seed: int = 0) -> dict:
    return {k: v.detach().clone() for k, v in skeleton(cfg, seed).state_dict().items()}


# ----------------------------------------------------------------------------- restated forward
def _inorm_prelu(sd, pre, x, ni, ai):
    x = F.instance_norm(x, weight=sd[f"{pre}.{ni}.weight"], bias=sd[f"{pre}.{ni}.bias"], eps=1e-5)
    return F.prelu(x, sd[f"{pre}.{ai}.weight"])


def _dense_block(sd, pre, x, depth):
    """DenseBlockV2, causal over frames (`_dense_block` :701-723): layer i sees [d_{i-1}, ..., d_0, x]; the frame taps are
    (t - 2^i, t), the sub-band taps (f-1, f, f+1) with zero padding."""
    skip = x
    for i in range(depth):
        d = 2 ** i
        p = f"{pre}.dense_block.{i}"
        y = F.conv2d(F.pad(skip, (1, 1, d, 0)), sd[f"{p}.1.weight"], sd[f"{p}.1.bias"], dilation=(d, 1))
        y = _inorm_prelu(sd, p, y, 2, 3)
        skip = torch.cat((y, skip), dim=1)
    return y


def _swoosh_linear(sd, pre, x, left: bool):
    """ActivationDropoutAndLinear (:131-140) with the activation's constant folded into the bias in double (:446-455)."""
    w, b = sd[f"{pre}.weight"], sd[f"{pre}.bias"]
    off = SWOOSH_L_OFFSET if left else SWOOSH_R_OFFSET
    fb = (b.double() - off * w.double().sum(dim=1)).to(b.dtype)
    x = F.softplus(x - (4.0 if left else 1.0)) - 0.08 * x
    return F.linear(x, w, fb)


def _ff(sd, pre, x):
    return _swoosh_linear(sd, f"{pre}.out_proj", F.linear(x, sd[f"{pre}.in_proj.weight"], sd[f"{pre}.in_proj.bias"]), True)


def attention_weights(sd, pre, x, cfg: ZipConfig, pe):
    """RelPositionMultiheadAttentionWeights (:232-296): softmax_j(q_i.k_j + p_i.R[j-i]) per head; R = linear_pos(pe rows of the
    offsets -(S-1) .. S-1).  The reference realises the j-i lookup with a pad / reshape skew; this is the direct gather."""
    n, s, _ = x.shape
    h, q, pd = cfg.heads, cfg.query_head_dim, cfg.pos_head_dim
    proj = F.linear(x, sd[f"{pre}.in_proj.weight"], sd[f"{pre}.in_proj.bias"])
    qq = proj[..., :h * q].reshape(n, s, h, q).permute(0, 2, 1, 3)
    kk = proj[..., h * q:2 * h * q].reshape(n, s, h, q).permute(0, 2, 3, 1)
    pp = proj[..., 2 * h * q:].reshape(n, s, h, pd).permute(0, 2, 1, 3)
    half = pe.shape[0] // 2
    rel = F.linear(pe[half - s + 1:half + s], sd[f"{pre}.linear_pos.weight"])          # (2S-1, H*pd)
    rel = rel.reshape(2 * s - 1, h, pd).permute(1, 2, 0)                                   # (H, pd, 2S-1)
    pos = torch.matmul(pp, rel.unsqueeze(0))                                               # (N, H, S, 2S-1)
    i = torch.arange(s)
    idx = (s - 1 - i).unsqueeze(1) + i.unsqueeze(0)                                        # [i, j] -> S-1-i+j
    pos = pos.gather(-1, idx.expand(n, h, s, s))
    return torch.softmax(torch.matmul(qq, kk) + pos, dim=-1)


def _self_attn(sd, pre, x, aw, cfg):
    n, s, _ = x.shape
    v = F.linear(x, sd[f"{pre}.in_proj.weight"], sd[f"{pre}.in_proj.bias"]).reshape(n, s, cfg.heads, cfg.value_head_dim)
    o = torch.matmul(aw, v.permute(0, 2, 1, 3)).permute(0, 2, 1, 3).reshape(n, s, -1)
    return F.linear(o, sd[f"{pre}.out_proj.weight"], sd[f"{pre}.out_proj.bias"])


def _nonlin_attn(sd, pre, x, aw0, cfg):
    hc = cfg.nonlin_hidden
    p = F.linear(x, sd[f"{pre}.in_proj.weight"], sd[f"{pre}.in_proj.bias"])
    s, m, y = p[..., :hc], p[..., hc:2 * hc], p[..., 2 * hc:]
    m = torch.matmul(aw0, m * torch.tanh(s)) * y
    return F.linear(m, sd[f"{pre}.out_proj.weight"], sd[f"{pre}.out_proj.bias"])


def _conv_module(sd, pre, x, cfg):
    c = cfg.channels
    p = F.linear(x, sd[f"{pre}.in_proj.weight"], sd[f"{pre}.in_proj.bias"])
    m = p[..., :c] * torch.sigmoid(p[..., c:])
    m = F.conv1d(m.transpose(1, 2), sd[f"{pre}.depthwise_conv.weight"], sd[f"{pre}.depthwise_conv.bias"], padding=cfg.conv_kernel // 2,
                 groups=c).transpose(1, 2)
    return _swoosh_linear(sd, f"{pre}.out_proj", m, False)


def zipformer_layer(sd, pre, x, outer_scale, cfg: ZipConfig, pe, dbg=None, tag=""):
    """Zipformer2EncoderLayer (:143-187) on (sequences, positions, channels), with the final BiasNorm, the layer bypass and the
    enclosing dual-path bypass in the reference's fused form (:659-676)."""
    x0 = x
    aw = attention_weights(sd, f"{pre}.self_attn_weights", x, cfg, pe)
    x = x + _ff(sd, f"{pre}.feed_forward1", x)
    if dbg is not None:
        dbg[f"{tag}.aw"] = aw
        dbg[f"{tag}.ff1"] = x
    x = x + _nonlin_attn(sd, f"{pre}.nonlin_attention", x, aw[:, 0], cfg)
    if dbg is not None:
        dbg[f"{tag}.nla"] = x
    x = x + _self_attn(sd, f"{pre}.self_attn1", x, aw, cfg)
    if dbg is not None:
        dbg[f"{tag}.sa1"] = x
    x = x + _conv_module(sd, f"{pre}.conv_module1", x, cfg)
    if dbg is not None:
        dbg[f"{tag}.cv1"] = x
    x = x + _ff(sd, f"{pre}.feed_forward2", x)
    x = x0 + (x - x0) * sd[f"{pre}.bypass_mid.bypass_scale"]
    if dbg is not None:
        dbg[f"{tag}.mid"] = x
    x = x + _self_attn(sd, f"{pre}.self_attn2", x, aw, cfg)
    x = x + _conv_module(sd, f"{pre}.conv_module2", x, cfg)
    x = x + _ff(sd, f"{pre}.feed_forward3", x)
    if dbg is not None:
        dbg[f"{tag}.ff3"] = x
    cs = sd[f"{pre}.bypass.bypass_scale"].double() * outer_scale.double()
    l2 = sd[f"{pre}.norm.log_scale"].double().exp() * math.sqrt(cfg.channels)
    nscale, rscale = (cs * l2).float(), (1.0 - cs).float()
    nrm = torch.linalg.vector_norm(x - sd[f"{pre}.norm.bias"], ord=2, dim=-1, keepdim=True)
    return (x / nrm) * nscale + x0 * rscale


def _dual_path(sd, pre, x, cfg, pe, dbg=None, tag=""):
    """x (B, T, F, C) channel-last: one layer over sub-bands (sequences = frames), one over frames (sequences = sub-bands)."""
    b, t, f, c = x.shape
    y = zipformer_layer(sd, f"{pre}.f_layers.0", x.reshape(b * t, f, c), sd[f"{pre}.bypass_layers.0.bypass_scale"], cfg, pe, dbg, tag + ".f")
    y = y.reshape(b, t, f, c).permute(0, 2, 1, 3).reshape(b * f, t, c)
    if dbg is not None:
        dbg[tag + ".f.out"] = y.reshape(b, f, t, c).permute(0, 2, 1, 3)
    y = zipformer_layer(sd, f"{pre}.t_layers.0", y, sd[f"{pre}.bypass_layers.1.bypass_scale"], cfg, pe, dbg, tag + ".t")
    return y.reshape(b, f, t, c).permute(0, 2, 1, 3)


def _downsample(x, dim, ds, bias):
    """SimpleDownsample (:194-220) along `dim`: repeat the last position up to a multiple of ds, softmax(bias)-weighted sum."""
    n = x.shape[dim]
    dn = (n + ds - 1) // ds
    if dn * ds > n:
        last = x.narrow(dim, n - 1, 1)
        x = torch.cat([x] + [last] * (dn * ds - n), dim=dim)
    w = bias.softmax(dim=0)
    shp = list(x.shape)
    shp[dim:dim + 1] = [dn, ds]
    x = x.reshape(shp)
    wshape = [1] * len(shp)
    wshape[dim + 1] = ds
    return (x * w.reshape(wshape)).sum(dim=dim + 1)


def _downsampled(sd, pre, x, ds, cfg, pe, dbg=None, tag=""):
    b, t, f, c = x.shape
    y = _downsample(x, 1, ds, sd[f"{pre}.downsample_t.bias"])
    y = _downsample(y, 2, ds, sd[f"{pre}.downsample_f.bias"])
    if dbg is not None:
        dbg[tag + ".down"] = y
    y = _dual_path(sd, f"{pre}.encoder", y, cfg, pe, dbg, tag)
    scale = sd[f"{pre}.out_combiner.bypass_scale"]
    y = y * scale
    y = y.repeat_interleave(ds, dim=2)[:, :, :f].repeat_interleave(ds, dim=1)[:, :t]
    return x * (1.0 - scale.double()).to(scale.dtype) + y


def _sub_pixel(sd, pre, x, r):
    """conv (1,3) to r*C channels, channel c*r+u -> sub-band f*r+u (`_decoder_upsample_pair` :754-768)."""
    y = F.conv2d(x, sd[f"{pre}.conv1.weight"], sd[f"{pre}.conv1.bias"], padding=(0, 1))
    b, cr, t, f = y.shape
    return y.reshape(b, cr // r, r, t, f).permute(0, 1, 3, 4, 2).reshape(b, cr // r, t, f * r)


def backbone(sd, feat, cfg: ZipConfig = ZipConfig(), dbg=None):
    """feat (B, 2, T, F) [compressed magnitude, phase] -> (mask-decoder output (B,1,T,F) before ReLU, phase (B,2,T,F))."""
    pe = compact_rel_pe(cfg.pos_dim, cfg.pe_max_len)
    de = "dense_encoder"
    x = F.conv2d(feat, sd[f"{de}.dense_conv_1.0.weight"], sd[f"{de}.dense_conv_1.0.bias"])
    x = _inorm_prelu(sd, f"{de}.dense_conv_1", x, 1, 2)
    if dbg is not None:
        dbg["enc0"] = x.permute(0, 2, 3, 1)
    x = _dense_block(sd, f"{de}.dense_block", x, cfg.dense_depth)
    if dbg is not None:
        dbg["enc_dense"] = x.permute(0, 2, 3, 1)
    x = F.conv2d(x, sd[f"{de}.dense_conv_2.0.weight"], sd[f"{de}.dense_conv_2.0.bias"], stride=(1, 2), padding=(0, 1))
    x = _inorm_prelu(sd, f"{de}.dense_conv_2", x, 1, 2)
    x = x.permute(0, 2, 3, 1)                                        # (B, T, F', C)
    if dbg is not None:
        dbg["enc"] = x
    for k, ds in enumerate(cfg.downsample):
        pre = f"TSConformer.encoders.{k}"
        x = _dual_path(sd, pre, x, cfg, pe, dbg, f"ts{k}") if ds == 1 else _downsampled(sd, pre, x, ds, cfg, pe, dbg, f"ts{k}")
        if dbg is not None:
            dbg[f"ts{k}"] = x
    x = x.permute(0, 3, 1, 2)                                        # (B, C, T, F')
    md, pd = "mask_decoder", "phase_decoder"
    m = _dense_block(sd, f"{md}.dense_block", x, cfg.dense_depth)
    m = _sub_pixel(sd, f"{md}.mask_conv.0", m, cfg.up_factor)
    m = _inorm_prelu(sd, f"{md}.mask_conv", m, 1, 2)
    if dbg is not None:
        dbg["mask_up"] = m.permute(0, 2, 3, 1)
    m = F.conv2d(m, sd[f"{md}.mask_conv.3.weight"], sd[f"{md}.mask_conv.3.bias"])
    p = _dense_block(sd, f"{pd}.dense_block", x, cfg.dense_depth)
    p = _sub_pixel(sd, f"{pd}.phase_conv.0", p, cfg.up_factor)
    p = _inorm_prelu(sd, f"{pd}.phase_conv", p, 1, 2)
    if dbg is not None:
        dbg["phase_up"] = p.permute(0, 2, 3, 1)
    ri = torch.cat((F.conv2d(p, sd[f"{pd}.phase_conv_r.weight"], sd[f"{pd}.phase_conv_r.bias"]),
                    F.conv2d(p, sd[f"{pd}.phase_conv_i.weight"], sd[f"{pd}.phase_conv_i.bias"])), dim=1)
    if dbg is not None:
        dbg["mx"], dbg["phase_ri"] = m, ri
    return m, ri


def zipenh_forward(sd, audio, cfg: ZipConfig = ZipConfig(), in_dtype: str = "F32", out_dtype: str = "F32", dbg=None, feat=None):
    """audio (B, 1, L) -> (B, 1, hop * (L // hop)); `ZipEnhancer.forward` (:818-927) without batch fold, at the model rate.

    `feat` (B, 2, T, F), when given, replaces the features computed here.  The phase feature atan2(im, re + 1e-5) (:844) sits on
    its branch cut wherever im is zero up to rounding and re < 0 -- structurally in the first frame and (for a hop-multiple
    length) the last one, whose reflect-padded, symmetrically windowed samples have an exactly real spectrum up to a sign: there
    the reference's own result is +pi or -pi by summation-order noise, and the backbone is not invariant to the flip.  Tests
    that compare another implementation end to end pass its features in, so both sides take the same +-pi decisions, and
    check separately that the features differ only by such flips."""
    feat0, nf = ends_oracle.zip_front(audio, in_dtype)
    feat = feat0 if feat is None else feat
    if dbg is not None:
        dbg["feat"], dbg["nf"] = feat, nf
    mx, ri = backbone(sd, feat, cfg, dbg)
    length = cfg.hop * (audio.shape[-1] // cfg.hop)
    return ends_oracle.zip_back(mx, ri, nf, length, out_dtype)


def zipenh_forward_batch(sd, audio, cfg: ZipConfig = ZipConfig(), in_dtype: str = "F32", out_dtype: str = "F32"):
    with torch.inference_mode():
        return torch.cat([zipenh_forward(sd, audio[i:i + 1], cfg, in_dtype, out_dtype) for i in range(audio.shape[0])], dim=0)

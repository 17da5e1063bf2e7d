"""-m gpu: the stand-alone conditioning / feature / recombine / output operators (ZipEnhancer,
MossFormerGAN-SE-16K, MossFormer2-SS-16K ends, linear resampler) through the C ABI vs the CPU
restatements in oracle/ends_oracle.py."""
import numpy as np
import pytest
import torch

import ends_oracle as eo
from make_golden import synth_audio

pytestmark = pytest.mark.gpu


def _audio(L, dt, batch=3, seed=21):
    x = synth_audio(L, seed, batch=batch)
    x[-1] = 0.0                                   # silent window: exercises every epsilon / guard
    return x if dt == "F32" else torch.round(x * 32767.0).to(torch.int16)


@pytest.mark.parametrize("L,kw", [(16000, dict(scale_factor=1 / 3)), (16000, dict(size=48000)), (8000, dict(scale_factor=2.0)),
                                  (44000, dict(size=16000)), (22500, dict(scale_factor=16000 / 22500)),
                                  (48000, dict(size=16000)), (16000, dict(size=16000))])
@pytest.mark.parametrize("dt", ["F32", "INT16", "F16"])
def test_resample_linear_matches_interpolate(L, kw, dt, libadn):
    from adn import ends

    x = synth_audio(L, 3, batch=2)
    x = {"F32": x, "INT16": torch.round(x * 32767).to(torch.int16), "F16": x.half()}[dt]
    ref = eo.resample_linear(x, **kw)
    got = ends.resample_linear(x.cuda(), **kw).cpu()
    assert got.shape == ref.shape
    tol = 2e-6 * float(ref.abs().max())           # one fp32 rounding of the two-term blend
    assert float((got - ref).abs().max()) <= tol
    assert float((eo.resample_linear_explicit(x, **kw) - ref).abs().max()) <= tol


@pytest.mark.parametrize("dt", ["F32", "INT16"])
def test_zipenhancer_ends(dt, libadn):
    from adn import ends

    L = 16000
    a = _audio(L, dt)
    e = ends.ZipEnds(L, dt, dt)
    feat, nf = e.analyse(a.cuda())
    rfeat, rnf = eo.zip_front(a, dt)
    assert feat.shape == rfeat.shape == (3, 2, 161, 201)
    assert torch.allclose(nf.cpu(), rnf, rtol=1e-6, atol=0)
    mag, rmag = feat[:, 0].cpu(), rfeat[:, 0]
    # p^0.15 is ill-conditioned at p -> 0 (d/dp = 0.15 p^-0.85): near-silent bins amplify the 1e-7 summation-order
    # differences of the DFT, so the tight bound applies to bins that carry energy
    loud = rmag > 1.0                                # power > 1 (of a maximum ~1e4)
    assert float((mag - rmag)[loud].abs().max()) <= 2e-5 * float(rmag.abs().max())
    assert float((mag - rmag).abs().max()) <= 1e-3
    # phase: compare as unit vectors where the bin carries energy (atan2 of a ~0 bin is noise on both sides)
    pha, rpha = feat[:, 1].cpu(), rfeat[:, 1]
    # (`loud`, not a fraction of max|X|^0.3: the compression maps a 1e-2 ratio to 2e-7 of the raw magnitude, where the
    # angle of the rounding noise is all that is left)
    d = torch.remainder(pha - rpha + torch.pi, 2 * torch.pi) - torch.pi
    assert float(d[loud].abs().max()) <= 2e-4
    g = torch.Generator().manual_seed(1)
    mx = torch.randn(3, 1, 161, 201, generator=g)
    ri = torch.randn(3, 2, 161, 201, generator=g)
    ri[0, :, :5] = 0.0                             # zero-phase guard (:885-888)
    y = e.synthesise(mx.cuda(), ri.cuda(), nf).cpu()
    yr = eo.zip_back(mx, ri, rnf, L, dt)
    assert y.shape == yr.shape == (3, 1, L) and y.dtype == yr.dtype
    if dt == "INT16":
        assert int((y.int() - yr.int()).abs().max()) <= 1
    else:
        assert float((y - yr).abs().max()) <= 1e-4 * max(1.0, float(yr.abs().max()))
    e.close()


@pytest.mark.parametrize("dt,L", [("F32", 16000), ("INT16", 15950)])
def test_mossformergan_ends(dt, L, libadn):
    from adn import ends

    a = _audio(L, dt)
    e = ends.GanEnds(L, dt, dt)
    feat, keep, nf = e.analyse(a.cuda())
    rfeat, rcc, rnf = eo.gan_front(a, dt)
    assert feat.shape == rfeat.shape and keep.shape == rcc.shape
    assert torch.allclose(nf.cpu(), rnf, rtol=1e-6, atol=0)
    loud = (rfeat[:, :1] > 1.0).expand_as(rfeat)     # see test_zipenhancer_ends: p^0.15, p^-0.35 near p = 0
    d = (feat.cpu() - rfeat).abs()
    assert float(d[loud].max()) <= 2e-5 * float(rfeat.abs().max()) and float(d.max()) <= 1e-3
    dk = (keep.cpu() - rcc).abs()
    assert float(dk[loud[:, 1:].transpose(-1, -2)].max()) <= 2e-5 * float(rcc.abs().max()) and float(dk.max()) <= 1e-3
    g = torch.Generator().manual_seed(2)
    T, F = e.frames, e.fbins
    mask = torch.rand(3, F, T, generator=g)
    cout = 0.1 * torch.randn(3, 2, F, T, generator=g)
    y = e.synthesise(mask.cuda(), cout.cuda(), keep, nf).cpu()
    yr = eo.gan_back(mask, cout, rcc, rnf, L, dt)
    assert y.shape == yr.shape == (3, 1, L) and y.dtype == yr.dtype
    if dt == "INT16":
        assert int((y.int() - yr.int()).abs().max()) <= 1
    else:
        assert float((y - yr).abs().max()) <= 1e-4 * max(1.0, float(yr.abs().max()))
    e.close()


@pytest.mark.parametrize("dt", ["F32", "INT16"])
def test_mossformer2_ss_ends(dt, libadn):
    from adn import ends

    L = 16000
    a = torch.round(synth_audio(L, 8, batch=3) * 32767.0)
    a[-1] = 0.0
    a = a.to(torch.int16) if dt == "INT16" else a
    e = ends.SsEnds(L, dt)
    x, rms_in = e.analyse(a.cuda())
    xr, rr = eo.ss_front(a)
    assert float((x.cpu() - xr).abs().max()) <= 2e-6 * max(1e-3, float(xr.abs().max()))
    assert torch.allclose(rms_in.cpu(), rr, rtol=2e-6, atol=1e-6)
    g = torch.Generator().manual_seed(4)
    wav = 0.05 * torch.randn(3, 2, L, generator=g)
    wav[1, 1] = 0.0                                # silent speaker: rms_out > 0 guard (:631)
    y = e.synthesise(wav.cuda(), rms_in).cpu()
    yr = eo.ss_back(wav, rr, dt)
    assert y.shape == yr.shape and y.dtype == yr.dtype
    if dt == "INT16":
        assert int((y.int() - yr.int()).abs().max()) <= 1 and not y[1, 1].any()
    else:
        assert float((y - yr).abs().max()) <= 1e-5 * max(1.0, float(yr.abs().max()))

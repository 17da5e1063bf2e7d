"""-m gpu: MossFormerGAN-SE-16K with in / out sample rates != 16 kHz inside the model (`F.interpolate(size=...)` behind the
int16 lift and behind the x norm_factor, MossFormerGAN_SE_16K/Export_MossFormer_SE.py:542-549, :884-891) vs the oracle, whose
resampling is pinned to the executed reference (tests/test_oracle_pinning.py).  Written after this round's GPU minutes were
spent: first GPU run at round end."""
import pytest
import torch

import mfgan_oracle as go

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("L,in_rate,out_rate,dt", [(1200, 8000, 48000, "F32"), (7200, 48000, 8000, "INT16"), (3300, 22500, 16000, "F32"),
                                                   (2400, 16000, 24000, "INT16")])
def test_resampled_io(L, in_rate, out_rate, dt, libadn):
    from adn import export, mfgan_params

    cfg = go.GanConfig(layers=2)
    sd = go.random_state_dict(cfg, 0)
    g = torch.Generator().manual_seed(4)
    x = (torch.rand(2, 1, L, generator=g) * 2 - 1) * 0.5
    xin = x if dt == "F32" else torch.round(x * 32767).to(torch.int16)
    with torch.inference_mode():
        y_ref = go.mfgan_forward_batch(sd, xin, cfg, dt, dt, chunk=1, in_rate=in_rate, out_rate=out_rate)
    m = export.mfgan_model(sd, mfgan_params.GanHyper(layers=cfg.layers), L, dt, dt, in_rate=in_rate, out_rate=out_rate)
    assert m.input.length == L and m.outputs[0].length == y_ref.shape[-1]
    y = m.run(xin.cuda()).cpu()
    assert y.shape == y_ref.shape and y.dtype == y_ref.dtype
    if dt == "INT16":
        d = (y.int() - y_ref.int()).abs()
        print(f"[{in_rate}->{out_rate}] int16 max LSB diff {int(d.max())}")
        assert int(d.max()) <= 1
    else:
        err = float((y - y_ref).abs().max())
        print(f"[{in_rate}->{out_rate}] max-abs err {err:.3e}")
        assert err <= 1e-4
    assert int(m.debug_read("launches")[0]) == m.launches_per_run(2)
    m.close()


def test_window_independence_at_the_bench_size(libadn):
    """BASELINE configs[4] shape: 6 blocks, 64 x 1 s windows in one run (four backbone passes of 16): every probed window equals
    its own single-window run bit for bit, an all-zero window stays finite, and the batch is finite."""
    from adn import export, mfgan_params

    cfg = go.GanConfig(layers=6)
    sd = go.random_state_dict(cfg, 0)
    L, B = 16000, 64
    g = torch.Generator().manual_seed(21)
    x = (torch.rand(B, 1, L, generator=g) * 2 - 1) * 0.4
    x[17] = 0.0
    m = export.mfgan_model(sd, mfgan_params.GanHyper(layers=6), L)
    xc = x.cuda()
    yb = m.run(xc).cpu()
    assert torch.isfinite(yb).all()
    for i in (0, 15, 16, 17, 47, 63):
        yi = m.run(xc[i:i + 1].contiguous()).cpu()
        assert torch.equal(yi[0], yb[i]), i
    m.close()

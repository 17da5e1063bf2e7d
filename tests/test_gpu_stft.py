"""-m gpu: CUDA STFT / ISTFT (through the C ABI) vs the CPU oracle and the fixtures generated
from the reference's own STFT_Process modules, for every in-scope geometry."""
import numpy as np
import pytest
import torch

import stft_oracle as so

pytestmark = pytest.mark.gpu

# fp32 dot products of length nfft in a different summation order than the reference's
# conv1d: tolerance = 2e-6 * nfft^0.5 * max|X| class.  Written out per case below.
TOL = {"gtcrn": 1e-4, "zipenhancer": 1e-4, "mossformergan_se_16k": 1e-4,
       "mossformer2_se_48k": 3e-4, "mel_band_roformer": 3e-4}


@pytest.mark.parametrize("name", list(so.SPECS))
def test_stft_istft_vs_golden_and_oracle(name, golden_dir, libadn):
    from adn import stft_tables
    from adn.stft_op import StftOp

    spec = so.SPECS[name]
    g = np.load(golden_dir / f"stft_{name}.npz")
    L = int(g["length"])
    op = StftOp(stft_tables.GEOMETRY[name], L)
    x = torch.from_numpy(g["x"]).cuda()
    s = op.forward(x).cpu()
    assert s.shape == g["spec"].shape                       # bit-exact frame indexing
    scale = float(np.abs(g["spec"]).max())
    err_g = float(np.abs(s.numpy() - g["spec"]).max())
    err_o = float((s - so.stft_packed(spec, torch.from_numpy(g["x"]))).abs().max())
    print(f"[stft {name}] max|X|={scale:.1f} err vs reference fixture {err_g:.3e}, vs oracle {err_o:.3e}")
    assert err_g <= TOL[name] * max(1.0, scale / 30.0)
    assert err_o <= TOL[name] * max(1.0, scale / 30.0)

    y = op.inverse(torch.from_numpy(g["spec_in"]).cuda()).cpu()
    assert y.shape == g["y"].shape
    err_y = float(np.abs(y.numpy() - g["y"]).max())
    ymax = float(np.abs(g["y"]).max())
    print(f"[istft {name}] max|y|={ymax:.2f} err vs reference fixture {err_y:.3e}")
    assert err_y <= 2e-5 * max(1.0, ymax)
    op.close()


@pytest.mark.parametrize("name", ["gtcrn", "zipenhancer"])
def test_round_trip_full_size(name, libadn):
    """Size-independent property at the BASELINE size: ISTFT(STFT(x)) == x on the kept range
    (the reference's own round-trip check, GTCRN/STFT_Process.py:580-600), batch 64."""
    from adn import stft_tables
    from adn.stft_op import StftOp

    geo = stft_tables.GEOMETRY[name]
    L = 16000
    op = StftOp(geo, L)
    gen = torch.Generator().manual_seed(1234)
    x = torch.randn(64, 1, L, generator=gen).cuda()
    y = op.inverse(op.forward(x))
    n = y.shape[-1]
    err = float((y - x[..., :n]).abs().max())
    print(f"[round trip {name}] B=64 max err {err:.3e}")
    assert err < 2e-4                                       # reference's own basis gives 1.6e-5..8.5e-5
    # linearity: STFT(a x1 + b x2) == a STFT(x1) + b STFT(x2)
    s1, s2 = op.forward(x[:32].contiguous()), op.forward(x[32:].contiguous())
    s12 = op.forward((0.5 * x[:32] - 2.0 * x[32:]).contiguous())
    assert float((s12 - (0.5 * s1 - 2.0 * s2)).abs().max()) < 1e-3
    op.close()

"""CPU: the host side of bench.py for every workload (no GPU work): descriptions, synthetic inputs, per-kernel algorithmic
work tables, model export (weight packing + metadata) -- so a typo in a workload is caught before the GPU box."""
import json

import numpy as np
import pytest

import bench


@pytest.mark.parametrize("name", sorted(bench.WORKLOADS))
def test_workload_host_side(name, tmp_path):
    from adn import modelfile

    wl = bench.WORKLOADS[name]()
    assert wl.name == name and " 4 " in wl.describe(4)
    assert wl.audio_seconds(4) == pytest.approx(4 * wl.chunk / wl.sr)
    x = wl.inputs(2, 1, 0)[0]
    assert tuple(x.shape) == (2, wl.channels, wl.chunk) and bool(np.isfinite(x.numpy()).all())
    work = wl.kernel_work()
    assert work and all(len(v) == 2 and v[0] >= 0 and v[1] >= 0 for v in work.values())
    assert isinstance(json.dumps(wl.describe(1)), str)
    if name in ("mf2se", "mf2ss", "mbr"):
        return                                          # 24-layer / depth-6 weight sets: packing is covered by test_host.py
    sd = wl.weights()
    path = tmp_path / f"{name}.adn"
    wl.export(sd, path)
    md, index, payload = modelfile.load(path)
    assert md["input_audio_length"] == str(wl.chunk) and md["input_audio_dtype"] == "F32"
    assert sum(e["count"] for e in index) <= payload.size


def test_cpu_baseline_leg_runs_for_the_new_families():
    for name in ("dfsmn",):
        wl = bench.WORKLOADS[name]()
        rate, dt = wl.cpu_rate(wl.weights(), 1, 2)
        assert rate > 0 and dt > 0

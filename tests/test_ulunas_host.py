"""CPU: the UL-UNAS launch sequence (csrc/ulunas_ops.cuh, the one libadn runs on the GPU) executed by a host loop
(tests/harness/ulunas_host.cpp) vs oracle/ulunas_oracle.py stage by stage (every encoder / dual-path / decoder block) and down to
the waveform (masked packed spectrum inverted with the oracle's ISTFT), on the raw state_dict of the reference fixtures."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import ulunas_oracle as uo
from stft_oracle import forward_basis, inverse_basis, pad_signal, window_sum

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "harness" / "ulunas_host.cpp"
HDRS = [ROOT / "audio-denoiser-onnx_b200" / "csrc" / n for n in ("ulunas_ops.cuh", "mfgan_gemm.cuh", "mfgan_ops.cuh")]
LIB = ROOT / "tests" / "_build" / "libulunas_host.so"
DUMP = C.CFUNCTYPE(None, C.c_char_p, C.POINTER(C.c_float), C.c_longlong)


@pytest.fixture(scope="module")
def host_lib():
    LIB.parent.mkdir(exist_ok=True)
    if not LIB.exists() or LIB.stat().st_mtime < max(p.stat().st_mtime for p in [SRC, *HDRS]):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-fopenmp", "-shared", "-I", str(HDRS[0].parent), str(SRC), "-o", str(LIB)],
                       check=True)
    lib = C.CDLL(str(LIB))
    lib.ulunas_host_forward.restype = C.c_int
    return lib


def run_host(lib, blob, spec: torch.Tensor):
    from adn import modelfile

    index, payload = modelfile.flatten(blob)
    n = len(index)
    names = (C.c_char_p * n)(*[e["name"].encode() for e in index])
    offs = (C.c_ulonglong * n)(*[e["offset"] for e in index])
    cnts = (C.c_ulonglong * n)(*[e["count"] for e in index])
    B, _, T = spec.shape
    s = np.ascontiguousarray(spec.numpy(), dtype=np.float32)
    out = np.zeros_like(s)
    dumps = {}

    def cb(name, ptr, count):
        dumps[name.decode()] = np.ctypeslib.as_array(ptr, shape=(count,)).copy()

    err = C.create_string_buffer(256)
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    rc = lib.ulunas_host_forward(names, offs, cnts, n, fp(payload), B, T, fp(s), fp(out), DUMP(cb), err, 256)
    assert rc > 0, err.value.decode()
    return out, dumps, rc


@pytest.mark.parametrize("L,B,dt", [(8192, 2, "F32"), (5000, 1, "INT16")])
def test_host_sequence_matches_oracle(L, B, dt, host_lib, golden_dir):
    from adn import ulunas_params as up

    g = np.load(golden_dir / "ulunas_f32_L16000.npz")
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd/")}
    gen = torch.Generator().manual_seed(5)
    x = (torch.rand(B, 1, L, generator=gen) * 2 - 1) * 0.4
    xin = x if dt == "F32" else torch.round(x * 32767).to(torch.int16)
    dbg = {}
    with torch.inference_mode():
        y_ref = uo.ulunas_forward(sd, xin, dt, dt, dbg=dbg)
        blob = up.pack(sd, L, dt, dt)
        spec = F.conv1d(pad_signal(uo.SPEC, xin.float()), torch.from_numpy(blob["stft.fwd"]).unsqueeze(1), stride=256)
    assert np.array_equal(blob["stft.fwd"], forward_basis(uo.SPEC, uo.INV_INT16 if dt == "INT16" else 1.0).numpy())
    T = spec.shape[-1]
    out, d, launches = run_host(host_lib, blob, spec)
    bad = []
    for k in sorted(dbg):                                   # the oracle dumps window 0 only
        ref = dbg[k]
        got = torch.from_numpy(d[k]).reshape((B,) + tuple(ref.shape[1:]))[:1]
        e, r = float((got - ref).abs().max()), float(ref.abs().max())
        print(f"{k:6s} {e:11.3e} {r:11.3e}")
        if not e <= 2e-5 * max(1.0, r):
            bad.append(k)
    with torch.inference_mode():
        inv = F.conv_transpose1d(torch.from_numpy(out), inverse_basis(uo.SPEC).unsqueeze(1), stride=256)
        y = inv[..., 256:inv.shape[-1] - 256] * torch.from_numpy(blob["stft.norm"])
    assert np.allclose(blob["stft.norm"], ((32767.0 if dt == "INT16" else 1.0) / window_sum(uo.SPEC, T)).numpy(), rtol=1e-6, atol=0)
    if dt == "INT16":
        yi = y.clamp(min=-32768.0, max=32767.0).to(torch.int16)
        assert int((yi.int() - y_ref.int()).abs().max()) <= 1
    else:
        e = float((y - y_ref).abs().max())
        print(f"wave   {e:11.3e} {float(y_ref.abs().max()):11.3e}")
        assert e <= 2e-6
    assert not bad, f"stages out of tolerance: {bad}"


def test_host_sequence_matches_reference_fixture(host_lib, golden_dir):
    """The same launch sequence vs the output of the EXECUTED reference (tests/golden/ulunas_f32_L16000.npz: 63 frames, one
    all-zero window) -- no oracle in between except its STFT / ISTFT tables, which are pinned bit-equal to the reference's."""
    from adn import ulunas_params as up

    g = np.load(golden_dir / "ulunas_f32_L16000.npz")
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd/")}
    x = torch.from_numpy(g["x"])
    blob = up.pack(sd, x.shape[-1], "F32", "F32")
    with torch.inference_mode():
        spec = F.conv1d(pad_signal(uo.SPEC, x), torch.from_numpy(blob["stft.fwd"]).unsqueeze(1), stride=256)
        out, _, _ = run_host(host_lib, blob, spec)
        inv = F.conv_transpose1d(torch.from_numpy(out), inverse_basis(uo.SPEC).unsqueeze(1), stride=256)
        y = inv[..., 256:inv.shape[-1] - 256] * torch.from_numpy(blob["stft.norm"])
    assert y.shape == tuple(g["y"].shape) and float((y - torch.from_numpy(g["y"])).abs().max()) <= 2e-6
    assert not y[2].any()                                   # the all-zero window stays exactly zero

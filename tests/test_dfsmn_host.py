"""CPU: the DFSMN launch sequence (csrc/dfsmn_ops.cuh, the one libadn runs on the GPU) executed by a host loop
(tests/harness/dfsmn_host.cpp) vs the CPU oracle, stage by stage and down to the waveform (the masked packed spectrum the
sequence hands to the ISTFT, inverted here with the oracle's ISTFT); plus adn/dfsmn_params.py vs the oracle's folds."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch

import dfsmn_oracle as do
from stft_oracle import istft_packed

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "harness" / "dfsmn_host.cpp"
HDRS = [ROOT / "audio-denoiser-onnx_b200" / "csrc" / n for n in ("dfsmn_ops.cuh", "mfgan_gemm.cuh", "mfgan_ops.cuh")]
LIB = ROOT / "tests" / "_build" / "libdfsmn_host.so"
DUMP = C.CFUNCTYPE(None, C.c_char_p, C.POINTER(C.c_float), C.c_longlong)


@pytest.fixture(scope="module")
def host_lib():
    LIB.parent.mkdir(exist_ok=True)
    if not LIB.exists() or LIB.stat().st_mtime < max(p.stat().st_mtime for p in [SRC, *HDRS]):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-fopenmp", "-shared", "-I", str(HDRS[0].parent), str(SRC), "-o", str(LIB)],
                       check=True)
    lib = C.CDLL(str(LIB))
    lib.dfsmn_host_forward.restype = C.c_int
    return lib


def run_host(lib, blob, h, audio: torch.Tensor):
    from adn import modelfile

    index, payload = modelfile.flatten(blob)
    n = len(index)
    names = (C.c_char_p * n)(*[e["name"].encode() for e in index])
    offs = (C.c_ulonglong * n)(*[e["offset"] for e in index])
    cnts = (C.c_ulonglong * n)(*[e["count"] for e in index])
    B, _, L = audio.shape
    T = h.n_frames(L)
    a = np.ascontiguousarray(audio.numpy())
    spec = np.zeros((B, 2 * 961, T), np.float32)
    dumps = {}

    def cb(name, ptr, count):
        dumps[name.decode()] = np.ctypeslib.as_array(ptr, shape=(count,)).copy()

    err = C.create_string_buffer(256)
    rc = lib.dfsmn_host_forward(names, offs, cnts, n, payload.ctypes.data_as(C.POINTER(C.c_float)), h.layers, h.lorder, B, L,
                                a.ctypes.data_as(C.c_void_p), int(a.dtype == np.int16), spec.ctypes.data_as(C.POINTER(C.c_float)),
                                DUMP(cb), err, 256)
    assert rc > 0, err.value.decode()
    return spec, dumps, rc


@pytest.mark.parametrize("L,B,dt", [(1920 + 960 * 6, 2, "F32"), (1920 + 960 * 11, 1, "INT16"), (1920, 1, "F32")])
def test_host_sequence_matches_oracle(L, B, dt, host_lib):
    from adn import dfsmn_params as dp

    cfg = do.DfsmnConfig(layers=3)
    sd = do.random_state_dict(cfg, 0)
    g = torch.Generator().manual_seed(7)
    x = (torch.rand(B, 1, L, generator=g) * 2 - 1) * 0.5
    xin = x if dt == "F32" else torch.round(x * 32767).to(torch.int16)
    dbg = {}
    with torch.inference_mode():
        y_ref = do.dfsmn_forward(sd, xin, cfg, dt, dt, dbg=dbg)
    h = dp.DfsmnHyper(layers=3)
    T = h.n_frames(L)
    spec, d, launches = run_host(host_lib, dp.pack(sd, h, L), h, xin)
    assert launches == 1 + 6 + 3 * h.layers                              # launches() of csrc/dfsmn.cu minus ISTFT / output

    def seq(name, ch):                                                     # (B*T, ch) dump -> (B, ch, T)
        return torch.from_numpy(d[name]).reshape(B, T, ch).transpose(1, 2)

    rows = [("feat", seq("feat", 120), dbg["feat"], 2e-5), ("lin1", seq("lin1", 256), dbg["lin1"], 2e-5)]
    rows += [(f"uf{i}", seq(f"uf{i}", 256), dbg[f"uf{i}"], 5e-5) for i in range(h.layers)]
    rows.append(("mask", seq("mask", 961), dbg["mask"], 2e-5))
    with torch.inference_mode():
        y = istft_packed(do.SYNTHESIS, torch.from_numpy(spec))
    if dt == "INT16":
        out = np.zeros(y.numel(), np.int16)
        yc = np.ascontiguousarray(y.numpy().reshape(-1))
        host_lib.dfsmn_host_out_i16(yc.ctypes.data_as(C.POINTER(C.c_float)), out.ctypes.data_as(C.POINTER(C.c_short)),
                                    C.c_longlong(out.size))
        assert int(np.abs(out.reshape(y_ref.shape).astype(np.int32) - y_ref.numpy().astype(np.int32)).max()) <= 1
    else:
        rows.append(("wave", y, y_ref, 2e-5))
    bad = []
    for n, got, ref, tol in rows:
        e, r = float((got - ref).abs().max()), float(ref.abs().max())
        print(f"{n:8s} {e:11.3e} {r:11.3e}")
        if not e <= tol * max(1.0, r):
            bad.append(n)
    assert not bad, f"stages out of tolerance: {bad}"


def test_pack_matches_oracle_fold():
    """Product-side folds and layouts (adn/dfsmn_params.py) vs the oracle's `fold` (bit-equal to the executed reference)."""
    from adn import dfsmn_params as dp

    cfg = do.DfsmnConfig(layers=2)
    sd = do.random_state_dict(cfg, 3)
    P = do.fold(sd, cfg)
    L = 1920 + 960 * 4
    h = dp.DfsmnHyper(layers=2)
    blob = dp.pack(sd, h, L)
    assert np.array_equal(blob["analysis_w"], P["analysis_w"].numpy())
    assert np.array_equal(blob["mel_t"], P["mel_banks"].numpy().T)
    for k in ("lin1", "lin2"):
        assert np.array_equal(blob[f"{k}_w"], P[f"{k}_w"].numpy().T) and np.array_equal(blob[f"{k}_b"], P[f"{k}_b"].numpy())
    for i in range(2):
        assert np.array_equal(blob[f"uf{i}.lin_w"], P[f"uf{i}.lin_w"].numpy().T)
        assert np.array_equal(blob[f"uf{i}.proj_w"], P[f"uf{i}.proj_w"].numpy().T)
        assert np.array_equal(blob[f"uf{i}.conv_w"], P[f"uf{i}.conv_w"].numpy().T)
    md = dp.metadata(h, L)
    assert md["model_family"] == "dfsmn" and md["max_signal_length"] == "5" and md["dfsmn_lorder"] == "20"
    with pytest.raises(ValueError):
        dp.pack(sd, h, 1920 + 100)


def test_host_sequence_matches_reference_fixture(host_lib, golden_dir):
    """The same launch sequence vs the output of the EXECUTED reference (tests/golden/dfsmn_f32_L9600_l3.npz, one all-zero window)."""
    from adn import dfsmn_params as dp

    g = np.load(golden_dir / "dfsmn_f32_L9600_l3.npz")
    cfg = do.DfsmnConfig(layers=int(g["layers"]))
    sd = do.random_state_dict(cfg, int(g["seed"]))
    h = dp.DfsmnHyper(layers=cfg.layers)
    x = torch.from_numpy(g["x"])
    spec, _, _ = run_host(host_lib, dp.pack(sd, h, x.shape[-1]), h, x)
    with torch.inference_mode():
        y = istft_packed(do.SYNTHESIS, torch.from_numpy(spec))
    assert y.shape == tuple(g["y"].shape) and float((y - torch.from_numpy(g["y"])).abs().max()) <= 1e-5
    assert not y[2].any()

"""CPU: the ZipEnhancer launch sequence (csrc/zipenh_ops.cuh, the one libadn runs on the GPU: functors + LinOps) executed by
a host loop (tests/harness/zipenh_host.cpp) vs the CPU oracle, stage by stage.  This checks the functors' index arithmetic,
the implicit-GEMM addressing of the dense-block / stride / sub-pixel convs, the weight layouts of adn/zipenh_params.py and
the buffer plumbing without a GPU; the -m gpu tests (test_gpu_zipenh.py) then check the same sequence as CUDA launches."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch

import zipenh_oracle as zo

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "harness" / "zipenh_host.cpp"
HDR = ROOT / "audio-denoiser-onnx_b200" / "csrc" / "zipenh_ops.cuh"
LIB = ROOT / "tests" / "_build" / "libzipenh_host.so"
DUMP = C.CFUNCTYPE(None, C.c_char_p, C.POINTER(C.c_float), C.c_longlong)


@pytest.fixture(scope="module")
def host_lib():
    LIB.parent.mkdir(exist_ok=True)
    if not LIB.exists() or LIB.stat().st_mtime < max(SRC.stat().st_mtime, HDR.stat().st_mtime):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-fopenmp", "-shared", "-I", str(HDR.parent), str(SRC), "-o", str(LIB)],
                       check=True)
    lib = C.CDLL(str(LIB))
    lib.zipenh_host_forward.restype = C.c_int
    return lib


def run_host(lib, blob: dict, ds, feat: torch.Tensor, T: int, want=None):
    from adn import modelfile

    index, payload = modelfile.flatten(blob)
    n = len(index)
    names = (C.c_char_p * n)(*[e["name"].encode() for e in index])
    offs = (C.c_ulonglong * n)(*[e["offset"] for e in index])
    cnts = (C.c_ulonglong * n)(*[e["count"] for e in index])
    B = feat.shape[0]
    f = np.ascontiguousarray(feat.numpy(), dtype=np.float32)
    mx = np.zeros((B, 1, T, 201), np.float32)
    ri = np.zeros((B, 2, T, 201), np.float32)
    dumps = {}

    def cb(name, ptr, count):
        k = name.decode()
        if want is None or k in want:
            dumps[k] = np.ctypeslib.as_array(ptr, shape=(count,)).copy()

    err = C.create_string_buffer(256)
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    dsa = (C.c_int * 4)(*ds)
    rc = lib.zipenh_host_forward(names, offs, cnts, n, fp(payload), dsa, B, T, fp(f), fp(mx), fp(ri), DUMP(cb), err, 256)
    assert rc > 0, err.value.decode()
    return mx, ri, dumps, rc


@pytest.mark.parametrize("L,B", [(1200, 2), (700, 1)])
def test_host_sequence_matches_oracle(L, B, host_lib):
    from adn import zipenh_params

    cfg = zo.ZipConfig()
    sd = zo.random_state_dict(cfg, 0)
    g = torch.Generator().manual_seed(7)
    x = (torch.rand(B, 1, L, generator=g) * 2 - 1) * 0.5
    dbg = {}
    with torch.inference_mode():
        zo.zipenh_forward(sd, x, cfg, dbg=dbg)
    h = zipenh_params.ZipHyper()
    T = h.n_frames(L)
    blob = zipenh_params.pack(sd, h, L)
    mx, ri, d, launches = run_host(host_lib, blob, cfg.downsample, dbg["feat"], T)
    rows = []

    def cmp(name, got, ref, tol=2e-5):
        ref = ref.detach().numpy().reshape(-1)
        got = np.asarray(got).reshape(-1)
        assert got.shape == ref.shape, (name, got.shape, ref.shape)
        err = float(np.abs(got - ref).max() / max(1.0, np.abs(ref).max()))
        rows.append((name, err))
        assert np.isfinite(got).all(), name
        assert err <= tol, (name, err, rows)

    def padded(a, Wp, lo, n):
        return a.reshape(B, T, Wp, 64)[:, :, lo:lo + n]

    cmp("enc0", padded(d["enc0"], 204, 1, 201), dbg["enc0"])
    pads = d["enc0"].reshape(B, T, 204, 64)
    assert not pads[:, :, 0].any() and not pads[:, :, 202:].any()
    cmp("enc.d3", padded(d["enc.d3"], 204, 1, 201), dbg["enc_dense"])
    cmp("enc", d["enc"], dbg["enc"])
    for k in ("f.aw", "f.ff1", "f.nla", "f.sa1", "f.cv1", "f.mid", "f.ff3", "f.out", "t.aw", "t.ff1", "t.nla", "t.cv1", "t.ff3"):
        ref = dbg[f"ts0.{k}"]
        got = d[f"ts0.{k}"]
        if k.startswith("t.") and not k.endswith("aw"):          # oracle: (B*F, T, C) sequence-major; here token order (B, T, F, C)
            ref = ref.reshape(B, 101, T, 64).permute(0, 2, 1, 3)
        cmp(f"ts0.{k}", got, ref)
    cmp("ts0", d["ts0"], dbg["ts0"])
    cmp("ts1.down", d["ts1.down"], dbg["ts1.down"])
    for k in range(1, 4):
        cmp(f"ts{k}", d[f"ts{k}"], dbg[f"ts{k}"], 5e-5)
    cmp("mask_up", d["mask_up"], dbg["mask_up"], 5e-5)
    cmp("phase_up", d["phase_up"], dbg["phase_up"], 5e-5)
    cmp("mx", mx, dbg["mx"], 5e-5)
    cmp("phase_ri", ri, dbg["phase_ri"], 5e-5)
    assert launches == 263

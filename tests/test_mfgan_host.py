"""CPU: the MossFormerGAN-SE-16K launch sequence (csrc/mfgan_ops.cuh, the one libadn runs on the GPU, every operator a
one-output-per-thread functor) executed by a host loop (tests/harness/mfgan_host.cpp) vs the CPU oracle, stage by stage.
This checks the functors' index arithmetic, the weight layouts of adn/mfgan_params.py and the buffer plumbing without a
GPU; the -m gpu tests (test_gpu_mfgan.py) then check the same sequence as CUDA launches."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch

import mfgan_oracle as go

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "harness" / "mfgan_host.cpp"
HDR = ROOT / "audio-denoiser-onnx_b200" / "csrc" / "mfgan_ops.cuh"
HDR2 = HDR.parent / "mfgan_gemm.cuh"
LIB = ROOT / "tests" / "_build" / "libmfgan_host.so"
DUMP = C.CFUNCTYPE(None, C.c_char_p, C.POINTER(C.c_float), C.c_longlong)


@pytest.fixture(scope="module")
def host_lib():
    LIB.parent.mkdir(exist_ok=True)
    if not LIB.exists() or LIB.stat().st_mtime < max(SRC.stat().st_mtime, HDR.stat().st_mtime, HDR2.stat().st_mtime):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-fopenmp", "-shared", "-I", str(HDR.parent), str(SRC), "-o", str(LIB)],
                       check=True)
    lib = C.CDLL(str(LIB))
    lib.mfgan_host_forward.restype = C.c_int
    return lib


def run_host(lib, blob: dict, layers: int, feat: torch.Tensor, T: int, use_gemm: bool = False):
    from adn import modelfile

    index, payload = modelfile.flatten(blob)
    n = len(index)
    names = (C.c_char_p * n)(*[e["name"].encode() for e in index])
    offs = (C.c_ulonglong * n)(*[e["offset"] for e in index])
    cnts = (C.c_ulonglong * n)(*[e["count"] for e in index])
    B = feat.shape[0]
    f = np.ascontiguousarray(feat.numpy(), dtype=np.float32)
    mask = np.zeros((B, 201, T), np.float32)
    cplx = np.zeros((B, 2, 201, T), np.float32)
    dumps = {}

    def cb(name, ptr, count):
        dumps[name.decode()] = np.ctypeslib.as_array(ptr, shape=(count,)).copy()

    err = C.create_string_buffer(256)
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    rc = lib.mfgan_host_forward(names, offs, cnts, n, fp(payload), layers, B, T, fp(f), fp(mask), fp(cplx), DUMP(cb), err, 256, int(use_gemm))
    assert rc > 0, err.value.decode()
    return mask, cplx, dumps, rc


@pytest.mark.parametrize("L,B,use_gemm", [(2400, 2, False), (1250, 1, False), (2400, 2, True), (850, 1, True)])
def test_host_sequence_matches_oracle(L, B, use_gemm, host_lib):
    """use_gemm: the contractions (Linear, the three attention branches, the convs) run as the GemmOps that
    csrc/mfgan_gemm.cuh `translate()` hands to the tiled CUDA GEMM, evaluated by plain loops."""
    from adn import mfgan_params

    cfg = go.GanConfig(layers=2)
    sd = go.random_state_dict(cfg, 0)
    g = torch.Generator().manual_seed(7)
    x = (torch.rand(B, 1, L, generator=g) * 2 - 1) * 0.5
    dbg = {}
    with torch.inference_mode():
        go.mfgan_forward(sd, x, cfg, dbg=dbg)
    h = mfgan_params.GanHyper(layers=2)
    T = h.n_frames(L)
    assert T == cfg.n_frames(L)
    blob = mfgan_params.pack(sd, h, L)
    feat = dbg["feat"].transpose(-1, -2).contiguous()                  # (B, 3, T, 201)
    mask, cplx, d, launches = run_host(host_lib, blob, 2, feat, T, use_gemm)
    Fq, E = 101, 64
    rows = []

    def cmp(name, got, ref, tol=2e-5):
        got = torch.from_numpy(np.asarray(got)).reshape(ref.shape)
        rows.append((name, float((got - ref).abs().max()), float(ref.abs().max()), tol))

    def px(a):                                                         # channel-last dump -> (B, C, T, F)
        return torch.from_numpy(a).reshape(B, T, Fq, E).permute(0, 3, 1, 2)

    cmp("enc", px(d["enc"]), dbg["enc"])
    for i in range(2):
        for p in ("intra", "inter"):
            for k in ("huv", "lin", "mf.huv", "mf.att"):
                cmp(f"B{i}.{p}.{k}", d[f"B{i}.{p}.{k}"], dbg[f"B{i}.{p}.{k}"], 5e-5)
            cmp(f"B{i}.{p}", px(d[f"B{i}.{p}"]), dbg[f"B{i}.{p}"], 5e-5)
        cmp(f"B{i}.x", px(d[f"B{i}.x"]), dbg[f"B{i}.x"], 5e-5)
    cmp("mask", mask, dbg["mask"], 5e-5)
    cmp("complex", cplx, dbg["complex"], 5e-5)
    print("\nstage                max|err|     max|ref|")
    for n, e, r, t in rows:
        print(f"{n:20s} {e:11.3e} {r:11.3e} {'OK' if e <= t * max(1.0, r) else 'FAIL'}")
    bad = [n for n, e, r, t in rows if not e <= t * max(1.0, r)]
    assert not bad, f"stages out of tolerance: {bad}"


def test_launch_count_matches_model_claim(host_lib):
    """`launches()` of csrc/mfgan.cu (the gpu_launches claim of bench.py) == what the sequence really launches."""
    from adn import mfgan_params

    cfg = go.GanConfig(layers=1)
    sd = go.random_state_dict(cfg, 1)
    h = mfgan_params.GanHyper(layers=1)
    L = 800
    T = h.n_frames(L)
    feat = torch.zeros(1, 3, T, 201)
    feat[:, 0] = 1.0
    _, _, _, launches = run_host(host_lib, mfgan_params.pack(sd, h, L), 1, feat, T)
    assert launches == 8 + 28 + 1 * (2 * 26 + 11) + 2 * (6 + 28)      # + 2 per path on the GPU: Att is three GEMMs

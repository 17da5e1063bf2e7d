"""ctypes binding of libadn.so (include/adn.h).  There is no CPU fallback: importing the
binding without the built library, or creating a model without an sm_100 GPU, raises."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

ADN_F32, ADN_I16, ADN_F16 = 0, 1, 2
DTYPE_NAMES = {ADN_F32: "F32", ADN_I16: "INT16", ADN_F16: "F16"}
NP_DTYPES = {ADN_F32: np.float32, ADN_I16: np.int16, ADN_F16: np.float16}
ORT_TYPES = {ADN_F32: "tensor(float)", ADN_I16: "tensor(int16)", ADN_F16: "tensor(float16)"}

LIB_PATH = Path(__file__).resolve().parent.parent / "libadn.so"

EXPORTED = [
    "adn_version", "adn_create", "adn_destroy", "adn_io_info", "adn_run", "adn_run_host",
    "adn_workspace_bytes", "adn_launches_per_run", "adn_debug_read", "adn_debug_stop_after", "adn_set_profiling",
    "adn_last_kernel_times", "adn_last_error", "adn_stft_create", "adn_stft_forward",
    "adn_stft_inverse", "adn_stft_destroy", "adn_resample_linear", "adn_rms_normalize", "adn_two_stage_rms",
    "adn_spec_features", "adn_spec_recombine", "adn_condition_output",
]


class TensorEntry(C.Structure):
    _fields_ = [("name", C.c_char_p), ("offset", C.c_uint64), ("count", C.c_uint64)]


class Desc(C.Structure):
    _fields_ = [("keys", C.POINTER(C.c_char_p)), ("values", C.POINTER(C.c_char_p)), ("n_kv", C.c_int32),
                ("tensors", C.POINTER(TensorEntry)), ("n_tensors", C.c_int32)]


class TensorInfo(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("dtype", C.c_int32), ("channels", C.c_int32), ("length", C.c_int32)]


class StftGeom(C.Structure):
    _fields_ = [("nfft", C.c_int32), ("hop", C.c_int32), ("center", C.c_int32), ("pad_reflect", C.c_int32),
                ("norm_multiply", C.c_int32)]


_lib = None


def lib():
    """Loads libadn.so (built by adn.build / __graft_entry__.build()); fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("ADN_LIB", LIB_PATH))
    if not path.exists():
        raise RuntimeError(
            f"libadn.so not found at {path}: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "The CUDA extension is mandatory; there is no CPU path.")
    L = C.CDLL(str(path))
    vp, i32, sz = C.c_void_p, C.c_int32, C.c_size_t
    L.adn_version.restype = C.c_char_p
    L.adn_last_error.restype = C.c_char_p
    L.adn_last_error.argtypes = [vp]
    L.adn_create.argtypes = [C.POINTER(vp), C.POINTER(Desc), vp, sz, C.c_int]
    L.adn_create.restype = i32
    L.adn_destroy.argtypes = [vp]
    L.adn_destroy.restype = None
    L.adn_io_info.argtypes = [vp, C.POINTER(TensorInfo), C.POINTER(TensorInfo), C.POINTER(i32)]
    L.adn_io_info.restype = i32
    L.adn_run.argtypes = [vp, vp, C.POINTER(vp), i32, vp]
    L.adn_run.restype = i32
    L.adn_run_host.argtypes = [vp, vp, C.POINTER(vp), i32]
    L.adn_run_host.restype = i32
    L.adn_workspace_bytes.argtypes = [vp, i32]
    L.adn_workspace_bytes.restype = sz
    L.adn_launches_per_run.argtypes = [vp, i32]
    L.adn_launches_per_run.restype = i32
    L.adn_debug_read.argtypes = [vp, C.c_char_p, vp, sz, C.POINTER(sz)]
    L.adn_debug_read.restype = i32
    L.adn_debug_stop_after.argtypes = [vp, i32]
    L.adn_debug_stop_after.restype = i32
    L.adn_set_profiling.argtypes = [vp, i32]
    L.adn_set_profiling.restype = i32
    L.adn_last_kernel_times.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_float), i32, C.POINTER(i32)]
    L.adn_last_kernel_times.restype = i32
    L.adn_stft_create.argtypes = [C.POINTER(vp), C.POINTER(StftGeom), vp, vp, vp, i32, C.c_int]
    L.adn_stft_create.restype = i32
    L.adn_stft_forward.argtypes = [vp, vp, vp, i32, i32, vp]
    L.adn_stft_forward.restype = i32
    L.adn_stft_inverse.argtypes = [vp, vp, vp, i32, i32, vp]
    L.adn_stft_inverse.restype = i32
    L.adn_stft_destroy.argtypes = [vp]
    L.adn_stft_destroy.restype = None
    L.adn_resample_linear.argtypes = [vp, i32, vp, i32, i32, i32, C.c_double, vp]
    L.adn_rms_normalize.argtypes = [vp, i32, C.c_float, C.c_float, vp, vp, i32, i32, i32, vp]
    L.adn_two_stage_rms.argtypes = [vp, i32, C.c_float, C.c_float, vp, vp, i32, i32, vp]
    L.adn_spec_features.argtypes = [i32, vp, vp, vp, i32, i32, i32, vp]
    L.adn_spec_recombine.argtypes = [i32, vp, vp, vp, vp, i32, i32, i32, vp]
    L.adn_condition_output.argtypes = [i32, vp, i32, vp, i32, vp, i32, i32, i32, vp]
    for fn in (L.adn_resample_linear, L.adn_rms_normalize, L.adn_two_stage_rms, L.adn_spec_features,
               L.adn_spec_recombine, L.adn_condition_output):
        fn.restype = i32
    _lib = L
    return L


class AdnError(RuntimeError):
    pass


def check(status: int, handle=None, what: str = "libadn"):
    if status != 0:
        msg = lib().adn_last_error(handle)
        raise AdnError(f"{what} failed (status {status}): {msg.decode() if msg else ''}")


def make_desc(metadata: dict[str, str], index: list[dict]):
    """Builds the adn_desc; returns (desc, keepalive) -- keepalive must outlive the call."""
    keys = [str(k).encode() for k in metadata]
    vals = [str(v).encode() for v in metadata.values()]
    karr = (C.c_char_p * len(keys))(*keys)
    varr = (C.c_char_p * len(vals))(*vals)
    names = [t["name"].encode() for t in index]
    tarr = (TensorEntry * len(index))()
    for i, t in enumerate(index):
        tarr[i].name = names[i]
        tarr[i].offset = int(t["offset"])
        tarr[i].count = int(t["count"])
    d = Desc(C.cast(karr, C.POINTER(C.c_char_p)), C.cast(varr, C.POINTER(C.c_char_p)), len(keys),
             C.cast(tarr, C.POINTER(TensorEntry)), len(index))
    return d, (keys, vals, karr, varr, names, tarr)

"""`Model`: thin Python owner of an `adn_model*` handle.

PyTorch tensors are used only as device-memory containers (pointer + stream); all
compute happens inside libadn.so."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, modelfile


class IOInfo:
    def __init__(self, info: _lib.TensorInfo):
        self.name = info.name.decode()
        self.dtype = int(info.dtype)
        self.channels = int(info.channels)
        self.length = int(info.length)
        self.np_dtype = _lib.NP_DTYPES[self.dtype]

    def __repr__(self):
        return f"IOInfo({self.name!r}, {_lib.DTYPE_NAMES[self.dtype]}, (1,{self.channels},{self.length}))"


class Model:
    def __init__(self, metadata: dict[str, str], index: list[dict], payload: np.ndarray, device_id: int = 0):
        self._h = C.c_void_p()
        self.metadata = dict(metadata)
        self.device_id = device_id
        L = _lib.lib()
        payload = np.ascontiguousarray(payload, dtype=np.float32)
        desc, keep = _lib.make_desc(self.metadata, index)
        st = L.adn_create(C.byref(self._h), C.byref(desc), payload.ctypes.data_as(C.c_void_p), payload.size,
                          device_id)
        del keep
        _lib.check(st, None, "adn_create")
        tin, tout, n = _lib.TensorInfo(), (_lib.TensorInfo * 4)(), C.c_int32(0)
        _lib.check(L.adn_io_info(self._h, C.byref(tin), tout, C.byref(n)), self._h, "adn_io_info")
        self.input = IOInfo(tin)
        self.outputs = [IOInfo(tout[i]) for i in range(n.value)]

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_file(cls, path, device_id: int = 0) -> "Model":
        md, index, payload = modelfile.load(path)
        return cls(md, index, payload, device_id)

    @classmethod
    def from_tensors(cls, metadata: dict[str, str], tensors: dict[str, np.ndarray], device_id: int = 0) -> "Model":
        index, payload = modelfile.flatten(tensors)
        return cls(metadata, index, payload, device_id)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            _lib.lib().adn_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ running
    def run(self, audio, out=None, stream=None):
        """audio: CUDA torch tensor (B, C, L) of the model's input dtype.  Returns the CUDA
        output tensor (B, C, L_out).  Asynchronous on the current torch stream."""
        import torch

        if not (audio.is_cuda and audio.is_contiguous()):
            raise ValueError("adn.Model.run needs a contiguous CUDA tensor")
        B = audio.shape[0]
        tdt = {np.float32: torch.float32, np.int16: torch.int16, np.float16: torch.float16}
        if audio.dtype != tdt[self.input.np_dtype] or tuple(audio.shape[1:]) != (self.input.channels, self.input.length) or B < 1:
            raise ValueError(f"adn.Model.run: input {tuple(audio.shape)} {audio.dtype} does not match {self.input}")
        multi = len(self.outputs) > 1            # MossFormer2-SS: one tensor per speaker, returned as a tuple
        if out is None:
            out = tuple(torch.empty((B, o.channels, o.length), dtype=tdt[o.np_dtype], device=audio.device)
                        for o in self.outputs)
        elif not multi and not isinstance(out, (tuple, list)):
            out = (out,)
        if len(out) != len(self.outputs):
            raise ValueError(f"adn.Model.run: {len(out)} output buffers for {len(self.outputs)} model outputs")
        for t, o in zip(out, self.outputs):      # the library writes B * C * L elements through these pointers
            if not (t.is_cuda and t.device == audio.device and t.is_contiguous() and t.dtype == tdt[o.np_dtype]
                    and tuple(t.shape) == (B, o.channels, o.length)):
                raise ValueError(f"adn.Model.run: output buffer {tuple(t.shape)} {t.dtype} on {t.device} does not match "
                                 f"({B}, {o.channels}, {o.length}) {o} on {audio.device}")
        st = torch.cuda.current_stream(audio.device).cuda_stream if stream is None else stream
        outs = (C.c_void_p * len(out))(*[t.data_ptr() for t in out])
        _lib.check(_lib.lib().adn_run(self._h, C.c_void_p(audio.data_ptr()), outs, B, C.c_void_p(st)),
                   self._h, "adn_run")
        return tuple(out) if multi else out[0]

    def run_host(self, audio: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """audio: host array (B, C, L); synchronous (H2D + kernels + D2H inside the call)."""
        a = np.ascontiguousarray(audio, dtype=self.input.np_dtype)
        B = a.shape[0]
        multi = len(self.outputs) > 1
        if a.ndim != 3 or tuple(a.shape[1:]) != (self.input.channels, self.input.length) or B < 1:
            raise ValueError(f"adn.Model.run_host: input {a.shape} does not match {self.input}")
        if out is None:
            out = tuple(np.empty((B, o.channels, o.length), dtype=o.np_dtype) for o in self.outputs)
        elif not multi and not isinstance(out, (tuple, list)):
            out = (out,)
        if len(out) != len(self.outputs):
            raise ValueError(f"adn.Model.run_host: {len(out)} output buffers for {len(self.outputs)} model outputs")
        for t, o in zip(out, self.outputs):
            if not (isinstance(t, np.ndarray) and t.flags.c_contiguous and t.flags.writeable and t.dtype == o.np_dtype
                    and t.shape == (B, o.channels, o.length)):
                raise ValueError(f"adn.Model.run_host: output buffer does not match ({B}, {o.channels}, {o.length}) {o}")
        outs = (C.c_void_p * len(out))(*[t.ctypes.data for t in out])
        _lib.check(_lib.lib().adn_run_host(self._h, C.c_void_p(a.ctypes.data), outs, B), self._h, "adn_run_host")
        return tuple(out) if multi else out[0]

    def run_host_ptr(self, in_ptr: int, out_ptr, batch: int):
        """out_ptr: one host address, or a sequence with one address per model output."""
        ptrs = list(out_ptr) if isinstance(out_ptr, (list, tuple)) else [out_ptr]
        outs = (C.c_void_p * len(ptrs))(*ptrs)
        _lib.check(_lib.lib().adn_run_host(self._h, C.c_void_p(in_ptr), outs, batch), self._h, "adn_run_host")

    # ------------------------------------------------------------------ diagnostics
    def workspace_bytes(self, batch: int) -> int:
        return int(_lib.lib().adn_workspace_bytes(self._h, batch))

    def launches_per_run(self, batch: int) -> int:
        return int(_lib.lib().adn_launches_per_run(self._h, batch))

    def debug_read(self, name: str) -> np.ndarray:
        n = C.c_size_t(0)
        L = _lib.lib()
        _lib.check(L.adn_debug_read(self._h, name.encode(), None, 0, C.byref(n)), self._h, "adn_debug_read")
        buf = np.empty(n.value, np.float32)
        _lib.check(L.adn_debug_read(self._h, name.encode(), buf.ctypes.data_as(C.c_void_p), n.value, C.byref(n)),
                   self._h, "adn_debug_read")
        return buf

    def debug_stop_after(self, n_launches: int):
        _lib.check(_lib.lib().adn_debug_stop_after(self._h, int(n_launches)), self._h, "adn_debug_stop_after")

    def set_profiling(self, on: bool):
        _lib.check(_lib.lib().adn_set_profiling(self._h, 1 if on else 0), self._h, "adn_set_profiling")

    def kernel_times(self) -> list[tuple[str, float]]:
        cap = 4096
        names = (C.c_char_p * cap)()
        ms = (C.c_float * cap)()
        n = C.c_int32(0)
        _lib.check(_lib.lib().adn_last_kernel_times(self._h, names, ms, cap, C.byref(n)), self._h,
                   "adn_last_kernel_times")
        return [(names[i].decode(), float(ms[i])) for i in range(n.value)]

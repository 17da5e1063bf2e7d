"""Stand-alone STFT / ISTFT operators with the reference `STFT_Process` call surface
(`_stft_B_packed_forward` / `_istft_B_packed_forward`, GTCRN/STFT_Process.py:303-336)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, stft_tables


class StftOp:
    def __init__(self, geometry: stft_tables.StftGeometry, length: int, device_id: int = 0):
        self.g = geometry
        self.length = length
        self.n_frames = geometry.n_frames(length)
        self.out_length = geometry.out_length(self.n_frames)
        fwd = np.ascontiguousarray(stft_tables.forward_basis(geometry).numpy())
        inv = np.ascontiguousarray(stft_tables.inverse_basis(geometry).numpy())
        nrm = np.ascontiguousarray(stft_tables.norm_table(geometry, self.n_frames).numpy())
        assert nrm.size == self.out_length
        geom = _lib.StftGeom(geometry.nfft, geometry.hop, int(geometry.center), int(geometry.pad_mode == "reflect"),
                             int(geometry.norm == "multiply"))
        self._h = C.c_void_p()
        L = _lib.lib()
        _lib.check(L.adn_stft_create(C.byref(self._h), C.byref(geom), fwd.ctypes.data_as(C.c_void_p),
                                     inv.ctypes.data_as(C.c_void_p), nrm.ctypes.data_as(C.c_void_p),
                                     self.n_frames, device_id), None, "adn_stft_create")

    def forward(self, x):
        """x: CUDA fp32 (B,1,L) -> (B,2F,T)."""
        import torch

        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.shape[-1] == self.length
        B = x.shape[0]
        out = torch.empty((B, 2 * self.g.fbins, self.n_frames), dtype=torch.float32, device=x.device)
        st = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(_lib.lib().adn_stft_forward(self._h, C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), B,
                                               self.length, C.c_void_p(st)), None, "adn_stft_forward")
        return out

    def inverse(self, spec):
        """spec: CUDA fp32 (B,2F,T) -> (B,1,L_out)."""
        import torch

        assert spec.is_cuda and spec.dtype == torch.float32 and spec.is_contiguous()
        B = spec.shape[0]
        out = torch.empty((B, 1, self.out_length), dtype=torch.float32, device=spec.device)
        st = torch.cuda.current_stream(spec.device).cuda_stream
        _lib.check(_lib.lib().adn_stft_inverse(self._h, C.c_void_p(spec.data_ptr()), C.c_void_p(out.data_ptr()), B,
                                               spec.shape[-1], C.c_void_p(st)), None, "adn_stft_inverse")
        return out

    def close(self):
        if self._h and self._h.value:
            _lib.lib().adn_stft_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

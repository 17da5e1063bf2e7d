"""onnxruntime-shaped facade over libadn: the slice of the ORT Python API that the
reference's `Inference_*_ONNX.py` scripts touch (reference
`GTCRN/Inference_GTCRN_ONNX.py:142-177, 193-214, 237-267, 306-317`).

A maintainer switches the reference's inference script to the B200 path with

    import adn.ort_shim as onnxruntime

and points `onnx_model_A` at a `.adn` file.  Session/run options and
provider tables are accepted and ignored; I/O goes through caller-owned `OrtValue`
buffers exactly as with ORT's io_binding.

Extension over ORT: the leading dimension of a bound OrtValue may be a batch B of
independent chunks (the reference graphs are fixed at 1); the whole
batch runs in one `adn_run_host` call.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

from . import _lib
from .model import Model


# ----------------------------------------------------------------------------- option objects
class ExecutionMode:
    ORT_SEQUENTIAL = 0
    ORT_PARALLEL = 1


class GraphOptimizationLevel:
    ORT_DISABLE_ALL = 0
    ORT_ENABLE_BASIC = 1
    ORT_ENABLE_EXTENDED = 2
    ORT_ENABLE_ALL = 99


class SessionOptions:
    def __init__(self):
        self.log_severity_level = 4
        self.log_verbosity_level = 4
        self.inter_op_num_threads = 0
        self.intra_op_num_threads = 0
        self.execution_mode = ExecutionMode.ORT_SEQUENTIAL
        self.graph_optimization_level = GraphOptimizationLevel.ORT_ENABLE_ALL
        self._entries: dict[str, str] = {}

    def add_session_config_entry(self, key: str, value: str):
        self._entries[str(key)] = str(value)


class RunOptions:
    def __init__(self):
        self.log_severity_level = 4
        self.log_verbosity_level = 4
        self._entries: dict[str, str] = {}

    def add_run_config_entry(self, key: str, value: str):
        self._entries[str(key)] = str(value)


class _OrtDevice:
    """Stand-in for onnxruntime.capi._pybind_state.OrtDevice (Inference_GTCRN_ONNX.py:68,92,105,112)."""

    def __init__(self, device_type=0, memory=0, device_id=0):
        self.device_type, self.memory, self.device_id = device_type, memory, device_id

    @staticmethod
    def cpu():
        return 0

    @staticmethod
    def cuda():
        return 1

    @staticmethod
    def dml():
        return 2

    @staticmethod
    def default_memory():
        return 0


class _Capi:
    class _pybind_state:
        OrtDevice = _OrtDevice


capi = _Capi()


# ----------------------------------------------------------------------------- values
class OrtValue:
    """Caller-owned buffer.  `device_type` 'cpu' keeps a numpy array; 'cuda' keeps a torch
    CUDA tensor (zero-copy binding)."""

    def __init__(self, array, device_type="cpu", device_id=0):
        self._device_type = device_type
        self._device_id = device_id
        if device_type == "cuda":
            import torch

            self._t = torch.as_tensor(np.ascontiguousarray(array)).to(f"cuda:{device_id}")
            self._a = None
        else:
            self._a = np.ascontiguousarray(array)
            self._t = None

    @staticmethod
    def ortvalue_from_numpy(array, device_type="cpu", device_id=0):
        return OrtValue(array, device_type, device_id)

    def update_inplace(self, array):
        array = np.ascontiguousarray(array)
        if self._t is not None:
            import torch

            self._t.copy_(torch.as_tensor(array).reshape(self._t.shape))
        else:
            if array.shape != self._a.shape or array.dtype != self._a.dtype:
                raise ValueError(f"update_inplace: shape/dtype mismatch {array.shape}/{array.dtype} vs "
                                 f"{self._a.shape}/{self._a.dtype}")
            np.copyto(self._a, array)

    def numpy(self):
        if self._t is not None:
            return self._t.cpu().numpy()
        return self._a

    def shape(self):
        return list(self._t.shape) if self._t is not None else list(self._a.shape)

    def device_name(self):
        return self._device_type


class _NodeArg:
    def __init__(self, name, type_, shape):
        self.name, self.type, self.shape = name, type_, shape


class _ModelMeta:
    def __init__(self, md):
        self.custom_metadata_map = dict(md)


class IOBinding:
    def __init__(self, session):
        self._session = session
        self.inputs: dict[str, OrtValue] = {}
        self.outputs: dict[str, OrtValue] = {}

    def bind_ortvalue_input(self, name, value: OrtValue):
        self._session._check_name(name, True)
        self.inputs[name] = value

    def bind_ortvalue_output(self, name, value: OrtValue):
        self._session._check_name(name, False)
        self.outputs[name] = value

    def clear_binding_inputs(self):
        self.inputs.clear()

    def clear_binding_outputs(self):
        self.outputs.clear()


# ----------------------------------------------------------------------------- session
class InferenceSession:
    """`onnxruntime.InferenceSession(path, sess_options=, providers=, provider_options=,
    disabled_optimizers=)` (Inference_GTCRN_ONNX.py:213-214,230-237).

    `path` is a `.adn` model file (any file name: the container is recognised by its magic).  A path ending in
    `_Metadata.<ext>` is the metadata sidecar `adn.modelfile.save` writes beside every model (the reference opens it only to
    read `custom_metadata_map`, audio_onnx_metadata.py:290-303); if it is absent the main model's header is read instead.
    No device work is done for a sidecar session."""

    def __init__(self, path, sess_options=None, providers=None, provider_options=None,
                 disabled_optimizers=None, device_id: int | None = None, **_ignored):
        p = Path(str(path))
        self._metadata_only = False
        if p.stem.endswith("_Metadata"):
            self._metadata_only = True
            if not (p.exists() and p.stat().st_size > 8):      # no sidecar written: read the main file's header instead
                p = p.with_name(p.stem[: -len("_Metadata")] + p.suffix)
        if p.suffix == ".onnx" and not p.exists():
            p = p.with_suffix(".adn")
        if not p.exists():
            raise FileNotFoundError(f"model file not found: {p}")
        self._path = p
        if device_id is None:
            device_id = 0
            for po in provider_options or []:
                if isinstance(po, dict) and "device_id" in po:
                    device_id = int(po["device_id"])
        self._providers = ["AdnB200ExecutionProvider"]
        if self._metadata_only:
            from . import modelfile

            md, _, _ = modelfile.load(p)
            self._model = None
            self._md = md
            self._inputs_meta, self._outputs_meta = [], []
            return
        self._model = Model.from_file(p, device_id)
        self._md = self._model.metadata
        i = self._model.input
        self._inputs_meta = [_NodeArg(i.name, _lib.ORT_TYPES[i.dtype], [1, i.channels, i.length])]
        self._outputs_meta = [_NodeArg(o.name, _lib.ORT_TYPES[o.dtype], [1, o.channels, o.length])
                              for o in self._model.outputs]

    # -- introspection used by the reference scripts
    def get_inputs(self):
        return list(self._inputs_meta)

    def get_outputs(self):
        return list(self._outputs_meta)

    def get_providers(self):
        return list(self._providers)

    def get_modelmeta(self):
        return _ModelMeta(self._md)

    def io_binding(self):
        return IOBinding(self)

    def _check_name(self, name, is_input):
        metas = self._inputs_meta if is_input else self._outputs_meta
        if name not in [m.name for m in metas]:
            raise ValueError(f"unknown {'input' if is_input else 'output'} name {name!r}; "
                             f"model has {[m.name for m in metas]}")

    # -- execution
    def run_with_iobinding(self, binding: IOBinding, run_options=None):
        if self._model is None:
            raise RuntimeError("metadata-only session cannot run")
        m = self._model
        vin = binding.inputs.get(m.input.name)
        vouts = [binding.outputs.get(o.name) for o in m.outputs]     # MossFormer2-SS binds separated_0 and separated_1
        if vin is None or any(v is None for v in vouts):              # (MossFormer2_SS_16K/Inference_MossFormer_SS_ONNX.py:312-317)
            raise RuntimeError("run_with_iobinding: input and every output must be bound")
        if vin._t is not None:            # device-resident binding
            if any(v._t is None for v in vouts):
                raise RuntimeError("input is bound on cuda but output is on cpu")
            import torch

            m.run(vin._t, out=vouts[0]._t if len(vouts) == 1 else tuple(v._t for v in vouts))
            torch.cuda.current_stream(vin._t.device).synchronize()   # ORT runs synchronously (:146)
            return
        a = vin._a
        exp_in = (m.input.channels, m.input.length)
        if a.dtype != m.input.np_dtype or tuple(a.shape[-2:]) != exp_in:
            raise ValueError(f"input buffer {a.shape}/{a.dtype} does not match model input "
                             f"(B,{exp_in[0]},{exp_in[1]})/{np.dtype(m.input.np_dtype)}")
        batch = int(np.prod(a.shape[:-2])) if a.ndim > 2 else 1
        for meta, v in zip(m.outputs, vouts):
            o, exp_out = v._a, (meta.channels, meta.length)
            if o.dtype != meta.np_dtype or tuple(o.shape[-2:]) != exp_out or o.size != batch * exp_out[0] * exp_out[1]:
                raise ValueError(f"output buffer {o.shape}/{o.dtype} does not match model output "
                                 f"({batch},{exp_out[0]},{exp_out[1]})/{np.dtype(meta.np_dtype)}")
        m.run_host_ptr(a.ctypes.data, [v._a.ctypes.data for v in vouts], batch)

    def run(self, output_names, input_feed: dict, run_options=None):
        m = self._model
        a = np.ascontiguousarray(input_feed[m.input.name])
        r = m.run_host(a)
        return list(r) if isinstance(r, tuple) else [r]


def get_available_providers():
    return ["AdnB200ExecutionProvider"]


def get_device():
    return "GPU"

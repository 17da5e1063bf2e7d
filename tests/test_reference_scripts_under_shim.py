"""CPU, build container only: the reference's UNMODIFIED inference scripts executed with `onnxruntime := adn.ort_shim`.

`GTCRN/Inference_GTCRN_ONNX.py`, `ZipEnhancer/Inference_ZipEnhancer_ONNX.py`,
`MossFormer2_SS_16K/Inference_MossFormer_SS_ONNX.py` and `H-GTCRN/Inference_H_GTCRN_ONNX.py` are run as `__main__` from /root/reference, byte for byte, against a model
file written by `adn.export` into a temporary directory (passed as argv[1], the scripts' own override).  Everything the scripts
touch on the ORT side -- SessionOptions / RunOptions attributes, OrtDevice, provider tables, the metadata sidecar session,
`get_inputs()` / `_inputs_meta`, `OrtValue.ortvalue_from_numpy / update_inplace / numpy`, `io_binding`, `run_with_iobinding` --
must exist in the shim with the meaning the scripts rely on.  There is no GPU here, so the one thing replaced is the device
behind the shim: `adn.ort_shim.Model` is swapped for a stand-in with the same interface that evaluates the CPU oracle (test
infrastructure).  pydub / soundfile / onnx (absent in this image) are stubs: decoded audio in, written audio captured.
The result must equal the scripts' window loop transcribed around the same oracle."""
import ctypes
import runpy
import sys
import types
from pathlib import Path

import numpy as np
import pytest
import torch

import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="/root/reference not present")
REF = Path("/root/reference")


class _OracleModel:
    """adn.model.Model's interface (input / outputs / metadata / run_host_ptr) over a per-window CPU function."""
    registry = {}

    def __init__(self, path):
        from adn import _lib, modelfile
        from adn.model import IOInfo

        self.metadata, _, _ = modelfile.load(path)
        self.fn, in_name, out_names, out_len, *rest = self.registry[Path(path).name]
        self.in_channels = rest[0] if rest else 1
        code = {"F32": _lib.ADN_F32, "INT16": _lib.ADN_I16, "F16": _lib.ADN_F16}

        def info(name, dtype, length, channels=1):
            t = _lib.TensorInfo()
            t.name, t.dtype, t.channels, t.length = name.encode(), code[dtype], channels, length
            return IOInfo(t)

        md = self.metadata
        self.input = info(in_name, md["input_audio_dtype"], int(md["input_audio_length"]), self.in_channels)
        self.outputs = [info(n, md["output_audio_dtype"], out_len) for n in out_names]
        self.calls = 0

    @classmethod
    def from_file(cls, path, device_id=0):
        return cls(path)

    def run_host_ptr(self, in_ptr, out_ptrs, batch):
        self.calls += 1
        n = batch * self.in_channels * self.input.length
        ct = {np.int16: ctypes.c_int16, np.float32: ctypes.c_float}[self.input.np_dtype]
        x = np.ctypeslib.as_array((ct * n).from_address(in_ptr)).reshape(batch, self.in_channels, -1)
        ys = self.fn(torch.from_numpy(x.copy()))
        ys = ys if isinstance(ys, tuple) else (ys,)
        for ptr, y, o in zip(out_ptrs, ys, self.outputs):
            co = {np.int16: ctypes.c_int16, np.float32: ctypes.c_float}[o.np_dtype]
            dst = np.ctypeslib.as_array((co * (batch * o.length)).from_address(ptr))
            np.copyto(dst, y.numpy().reshape(-1))

    def close(self):
        pass


def _run_script(script: Path, model_dir: Path, audio: np.ndarray, monkeypatch):
    import adn.ort_shim as shim

    written = {}
    sf = types.ModuleType("soundfile")
    sf.write = lambda path, data, sr, subtype=None: written.__setitem__(Path(path).name, (np.array(data), sr, subtype))

    class _Seg:
        def set_channels(self, n):
            return self

        def set_frame_rate(self, sr):
            return self

        def get_array_of_samples(self):
            return audio

    pydub = types.ModuleType("pydub")
    pydub.AudioSegment = types.SimpleNamespace(from_file=lambda path, *a, **k: _Seg())
    capi = types.ModuleType("onnxruntime.capi")
    capi._pybind_state = shim.capi._pybind_state
    for name, mod in (("onnxruntime", shim), ("onnxruntime.capi", capi), ("soundfile", sf), ("pydub", pydub), ("onnx", types.ModuleType("onnx"))):
        monkeypatch.setitem(sys.modules, name, mod)
    for name in ("audio_onnx_metadata", "Example_Audio"):
        monkeypatch.delitem(sys.modules, name, raising=False)
    monkeypatch.setattr(shim, "Model", _OracleModel)
    monkeypatch.setattr(sys, "argv", [str(script), str(model_dir)])
    monkeypatch.chdir(model_dir)
    ns = runpy.run_path(str(script), run_name="__main__")
    return written, ns


def test_gtcrn_script(tmp_path, monkeypatch):
    import gtcrn_oracle as go
    from adn import export

    sd = go.random_state_dict(0)
    W = 4096
    export.export_gtcrn(sd, tmp_path / "GTCRN.onnx", W, "INT16", "INT16")          # the name the script looks for (argv[1] / GTCRN.onnx)
    assert (tmp_path / "GTCRN_Metadata.onnx").exists()
    fn = lambda x: go.gtcrn_forward_batch(sd, x, "INT16", "INT16")
    out_len = int(fn(torch.zeros(1, 1, W, dtype=torch.int16)).shape[-1])
    _OracleModel.registry["GTCRN.onnx"] = (fn, "noisy_audio", ["denoised_audio"], out_len)
    rng = np.random.default_rng(0)
    audio = rng.integers(-9000, 9000, size=2 * W + 777, dtype=np.int16)
    written, ns = _run_script(REF / "GTCRN" / "Inference_GTCRN_ONNX.py", tmp_path, audio, monkeypatch)
    y, sr, subtype = written["denoised.wav"]
    assert sr == 16000 and subtype == "PCM_16" and y.dtype == np.int16
    # the script's loop (:287-333) transcribed: stride = output length (in != out, equal rates), zero tail, trim to the input length
    n = len(audio)
    stride = out_len
    num = int(np.ceil((n - W) / stride)) + 1
    a = np.concatenate([audio, np.zeros((num - 1) * stride + W - n, np.int16)])
    want = np.concatenate([fn(torch.from_numpy(a[k * stride:k * stride + W].reshape(1, 1, -1))).numpy().reshape(-1) for k in range(num)])[:n]
    assert np.array_equal(y, want)
    assert ns["ort_session_A"].get_providers() == ["AdnB200ExecutionProvider"]


def test_zipenhancer_script(tmp_path, monkeypatch):
    import zipenh_oracle as zo
    from adn import export

    cfg = zo.ZipConfig()
    sd = zo.random_state_dict(cfg, 0)
    W = 1600
    export.export_zipenh(sd, tmp_path / "ZipEnhancer.onnx", None, W, "INT16", "INT16")
    fn = lambda x: zo.zipenh_forward_batch(sd, x, cfg, "INT16", "INT16")
    _OracleModel.registry["ZipEnhancer.onnx"] = (fn, "noisy_audio", ["denoised_audio"], W)
    rng = np.random.default_rng(1)
    audio = rng.integers(-9000, 9000, size=2 * W + 300, dtype=np.int16)
    written, _ = _run_script(REF / "ZipEnhancer" / "Inference_ZipEnhancer_ONNX.py", tmp_path, audio, monkeypatch)
    (y, sr, subtype), = written.values()
    n = len(audio)
    num = int(np.ceil((n - W) / W)) + 1
    a = np.concatenate([audio, np.zeros(num * W - n, np.int16)])
    want = np.concatenate([fn(torch.from_numpy(a[k * W:(k + 1) * W].reshape(1, 1, -1))).numpy().reshape(-1) for k in range(num)])[:n]
    assert sr == 16000 and y.dtype == np.int16 and np.array_equal(y, want)


def test_mossformer2_ss_script(tmp_path, monkeypatch):
    import mf2ss_oracle as so
    from adn import export, mf2ss_params

    cfg = so.SsConfig(layers=1)
    sd = so.random_state_dict(cfg, 0)
    W = 2408
    md = export.export_mf2ss(sd, tmp_path / "MossFormer2_SS_16K.onnx", mf2ss_params.SsHyper(layers=1), W, "INT16", "INT16")
    pad_head = int(md["pad_head"])
    fn = lambda x: so.mf2ss_forward_batch(sd, x, cfg, "INT16", "INT16")
    _OracleModel.registry["MossFormer2_SS_16K.onnx"] = (fn, "mix_audio", ["separated_0", "separated_1"], W)
    rng = np.random.default_rng(2)
    audio = rng.integers(-9000, 9000, size=W + 500, dtype=np.int16)
    np.random.seed(11)                                    # the un-folded script pads the tail with RMS-matched gaussian noise (:293-296)
    written, _ = _run_script(REF / "MossFormer2_SS_16K" / "Inference_MossFormer_SS_ONNX.py", tmp_path, audio, monkeypatch)
    assert len(written) == 2
    a = np.concatenate([np.zeros(pad_head, np.int16), audio])
    n = len(a)
    num = int(np.ceil((n - W) / W)) + 1 if n > W else 1
    pad = num * W - n
    np.random.seed(11)
    tail = a[-pad:].astype(np.float32)
    noise = (np.sqrt(np.mean(tail * tail, dtype=np.float32), dtype=np.float32) * np.random.normal(loc=0.0, scale=1.0, size=(1, 1, pad))).astype(np.int16)
    a = np.concatenate([a, noise.reshape(-1)])
    outs = [fn(torch.from_numpy(a[k * W:(k + 1) * W].reshape(1, 1, -1))) for k in range(num)]
    for s, (name, (y, sr, subtype)) in enumerate(sorted(written.items())):
        want = np.concatenate([o[s].numpy().reshape(-1) for o in outs])[pad_head:n]
        assert sr == 16000 and np.array_equal(y, want), name


def test_h_gtcrn_script(tmp_path, monkeypatch):
    """Two microphones in, one channel out: the script sizes its buffers from the session's channel counts (:306-309), pads the
    tail by reflection (`pad_audio_tail_with_context`, :138-153) and concatenates the windows (:369-371)."""
    import hgtcrn_oracle as ho
    from adn import chunker, export

    sd = ho.random_state_dict(0)
    W = 256 * 24
    export.export_hgtcrn(sd, tmp_path / "H_GTCRN.onnx", W, "INT16", "INT16")
    assert (tmp_path / "H_GTCRN_Metadata.onnx").exists()
    fn = lambda x: ho.hgtcrn_forward_batch(sd, x, "INT16", "INT16")
    _OracleModel.registry["H_GTCRN.onnx"] = (fn, "noisy_audio", ["denoised_audio"], W, 2)
    rng = np.random.default_rng(1)
    n = 2 * W + 1234
    src = rng.integers(-6000, 6000, size=n).astype(np.int16)
    stereo = np.stack((src, (0.7 * np.roll(src, 4)).astype(np.int16) + rng.integers(-2000, 2000, size=n).astype(np.int16)), axis=0)   # (2, n)
    written, ns = _run_script(REF / "H-GTCRN" / "Inference_H_GTCRN_ONNX.py", tmp_path, stereo.T.reshape(-1).copy(), monkeypatch)   # interleaved, as pydub gives it
    y, sr, subtype = written["denoised.wav"]
    assert sr == 16000 and subtype == "PCM_16" and y.dtype == np.int16 and y.shape == (n,)
    wins, stride = chunker.split(stereo, W, W, tail="reflect")
    assert stride == W and wins.shape == (3, 2, W)
    want = fn(torch.from_numpy(wins)).numpy().reshape(-1)[:n]
    assert np.array_equal(y, want)

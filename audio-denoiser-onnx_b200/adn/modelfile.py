"""`.adn` model file: the build's stand-in for `<Model>.onnx` + `<Model>_Metadata.onnx`.

The reference ships weights as ONNX initializers and its run-time constants as string
`metadata_props` (audio_onnx_metadata.py:83-112).  libadn needs neither a graph nor
protobuf, so the model file is a flat container:

    b"ADN1" | u32 header_len | JSON header | fp32 payload

header = {"metadata": {key: str}, "tensors": [{"name", "offset", "count", "shape"}]},
offsets/counts in floats.  The metadata keys are the reference's own
(REQUIRED_AUDIO_METADATA_KEYS, audio_onnx_metadata.py:8-26).
"""
from __future__ import annotations

import json
import struct
from pathlib import Path

import numpy as np

MAGIC = b"ADN1"


def metadata_path_for_model(path) -> Path:
    """`<Model>_Metadata.onnx` beside the model file, whatever the model's own suffix: the sidecar the reference's
    `load_runtime_metadata` opens and requires to exist (audio_onnx_metadata.py:37-39, :290-297).  The container is
    recognised by its magic, not by its name."""
    p = Path(path)
    return p.with_name(f"{p.stem}_Metadata.onnx")


def save(path, metadata: dict[str, str], tensors: dict[str, np.ndarray], sidecar: bool = True) -> None:
    """Writes the model file and, beside it, the metadata-only sidecar (same container, no tensors) -- the counterpart of the
    reference's `stamp_export_metadata` (Export_GTCRN.py:780-790)."""
    if sidecar and tensors:
        save(metadata_path_for_model(path), metadata, {}, sidecar=False)
    index, chunks, off = [], [], 0
    for name, arr in tensors.items():
        a = np.ascontiguousarray(arr, dtype=np.float32)
        pad = (-off) % 4                       # keep every tensor 16-byte aligned
        if pad:
            chunks.append(np.zeros(pad, np.float32))
            off += pad
        index.append({"name": name, "offset": off, "count": int(a.size), "shape": list(a.shape)})
        chunks.append(a.reshape(-1))
        off += a.size
    header = json.dumps({"metadata": {str(k): str(v) for k, v in metadata.items()}, "tensors": index}).encode()
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<I", len(header)))
        f.write(header)
        f.write(np.concatenate(chunks).tobytes() if chunks else b"")


def load(path):
    """Returns (metadata dict, tensor index list, flat fp32 payload)."""
    raw = Path(path).read_bytes()
    if raw[:4] != MAGIC:
        raise ValueError(f"{path}: not an ADN1 model file")
    (hlen,) = struct.unpack("<I", raw[4:8])
    header = json.loads(raw[8:8 + hlen].decode())
    if hlen > len(raw) - 8:
        raise ValueError(f"{path}: header length {hlen} exceeds the file")
    if (len(raw) - 8 - hlen) % 4:
        raise ValueError(f"{path}: payload is not a whole number of fp32 values")
    payload = np.frombuffer(raw, dtype=np.float32, offset=8 + hlen).copy()
    _validate_index(path, header.get("tensors"), payload.size)
    return header["metadata"], header["tensors"], payload


def _validate_index(path, index, nfloats: int):
    """Every tensor record must lie inside the payload (the C ABI re-checks offset / count without 64-bit wrap-around,
    api.cu adn_create; a malformed file should fail here with the tensor's name, not there with an index)."""
    if not isinstance(index, list):
        raise ValueError(f"{path}: tensor index missing")
    for t in index:
        name = t.get("name") if isinstance(t, dict) else None
        try:
            off, cnt = int(t["offset"]), int(t["count"])
            shape = [int(d) for d in t["shape"]]
        except (KeyError, TypeError, ValueError):
            raise ValueError(f"{path}: malformed tensor record {name!r}") from None
        if off < 0 or cnt < 0 or off > nfloats or cnt > nfloats - off or any(d < 0 for d in shape) or int(np.prod(shape, dtype=np.int64)) != cnt:
            raise ValueError(f"{path}: tensor {name!r} (offset {off}, count {cnt}, shape {shape}) does not fit the payload of {nfloats} values")


def flatten(tensors: dict[str, np.ndarray]):
    """In-memory equivalent of save()+load(): (index, payload)."""
    index, chunks, off = [], [], 0
    for name, arr in tensors.items():
        a = np.ascontiguousarray(arr, dtype=np.float32)
        pad = (-off) % 4
        if pad:
            chunks.append(np.zeros(pad, np.float32))
            off += pad
        index.append({"name": name, "offset": off, "count": int(a.size), "shape": list(a.shape)})
        chunks.append(a.reshape(-1))
        off += a.size
    return index, np.concatenate(chunks)

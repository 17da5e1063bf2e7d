"""Model export: reference `state_dict` -> `.adn` model file (the build's counterpart of the
`Export_*.py` main blocks, e.g. GTCRN/Export_GTCRN.py:705-792, minus ONNX)."""
from __future__ import annotations

from . import gtcrn_params, modelfile


def export_gtcrn(state_dict: dict, path, input_audio_length: int = 16000, in_dtype: str = "INT16",
                 out_dtype: str = "INT16") -> dict[str, str]:
    """Writes `path` (.adn) for one static chunk length; returns the metadata stamped."""
    md = gtcrn_params.metadata(input_audio_length, in_dtype, out_dtype)
    tensors = gtcrn_params.pack(state_dict, input_audio_length)
    modelfile.save(path, md, tensors)
    return md


def gtcrn_model(state_dict: dict, input_audio_length: int = 16000, in_dtype: str = "F32",
                out_dtype: str = "F32", device_id: int = 0):
    """In-memory shortcut: build a `Model` without touching the file system."""
    from .model import Model

    md = gtcrn_params.metadata(input_audio_length, in_dtype, out_dtype)
    return Model.from_tensors(md, gtcrn_params.pack(state_dict, input_audio_length), device_id)

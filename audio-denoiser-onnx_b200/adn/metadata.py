"""Run-time configuration from the model's string metadata.

The drop-in contract with the reference (`audio_onnx_metadata.py:8-26, 247-303, 354-386`) is a SCHEMA, and it is kept as one here:
the required-key list, each run-time constant's metadata key, type, required flag and default, and the two error messages a caller
of the reference can observe.  Everything else is this module's own: one generic typed lookup (`MetadataReader.get`) driven by the
table `RUNTIME_SCHEMA`; the `required_*` / `optional_*` method names of the reference's reader are generated from it for callers
that were written against them.
"""
from __future__ import annotations

REQUIRED_AUDIO_METADATA_KEYS = (
    "audio_metadata_version", "producer", "model_name", "task", "model_family", "dynamic_axes", "opset",
    "input_audio_dtype", "output_audio_dtype", "in_sample_rate", "out_sample_rate", "model_sample_rate",
    "input_audio_length", "input_to_output_scale", "max_dynamic_audio_seconds", "normalize_audio_default",
    "normalize_target_rms",
)

_TRUE, _FALSE = frozenset({"1", "true", "yes", "on"}), frozenset({"0", "false", "no", "off"})


def _to_bool(text, key):
    word = str(text).strip().lower()
    if word in _TRUE or word in _FALSE:
        return word in _TRUE
    raise ValueError(f"Metadata key {key} must be a boolean encoded as 1/0, got {text!r}.")


_CASTS = {"int": lambda t, k: int(t), "float": lambda t, k: float(t), "bool": _to_bool, "string": lambda t, k: t}


class MetadataReader:
    """Typed view of a `custom_metadata_map`.  An absent key and an empty string are the same thing."""

    def __init__(self, metadata):
        self.metadata = dict(metadata or {})

    def get(self, key, kind="string", required=False, default=None):
        text = self.metadata.get(key)
        if text is None or text == "":
            if required:
                raise KeyError(f"Required metadata key {key} is missing. "
                               "Re-export with the matching Export_*.py and rerun Optimize_ONNX.py.")
            return default
        return _CASTS[kind](text, key)

    def string(self, key, default=None, required=False):
        return self.get(key, "string", required, default)


def _install_typed_getters():
    for kind in ("int", "float", "bool"):
        setattr(MetadataReader, f"required_{kind}", lambda self, key, _k=kind: self.get(key, _k, required=True))
        setattr(MetadataReader, f"optional_{kind}", lambda self, key, default=None, _k=kind: self.get(key, _k, False, default))


_install_typed_getters()


def load_runtime_metadata(session, required_keys=REQUIRED_AUDIO_METADATA_KEYS) -> MetadataReader:
    reader = MetadataReader(session.get_modelmeta().custom_metadata_map or {})
    for key in required_keys:
        reader.get(key, required=True)
    return reader


# constant -> (metadata key, type, required, default).  A callable default is evaluated on the constants resolved so far
# (the table is ordered), which is how the reference derives three of its fall-backs from the sample rates.
def _fold_input_default(c):
    w = c["FOLD_WINDOW_LENGTH"]
    return max(1, int(round(w * c["IN_SAMPLE_RATE"] / c["MODEL_SAMPLE_RATE"]))) if w else 0


RUNTIME_SCHEMA = (
    ("IN_SAMPLE_RATE", "in_sample_rate", "int", True, None),
    ("OUT_SAMPLE_RATE", "out_sample_rate", "int", True, None),
    ("MODEL_SAMPLE_RATE", "model_sample_rate", "int", True, None),
    ("INPUT_TO_OUTPUT_SCALE", "input_to_output_scale", "float", True, None),
    ("BATCH_WINDOW_SECONDS", "batch_window_seconds", "float", False, 0.0),
    ("HOP_LENGTH", "hop_length", "int", False, 0),
    ("FOLD_WINDOW_LENGTH", "fold_window_length", "int", False, 0),
    ("FOLD_INPUT_LENGTH", "fold_input_length", "int", False, _fold_input_default),
    ("BATCH_FOLD_INFERENCE", "batch_fold_inference_default", "bool", False, False),
    ("MAX_DYNAMIC_AUDIO_SECONDS", "max_dynamic_audio_seconds", "int", True, None),
    ("NORMALIZE_AUDIO", "normalize_audio_default", "bool", True, None),
    ("NORMALIZE_TARGET_RMS", "normalize_target_rms", "float", True, None),
    ("INPUT_CHANNELS", "input_channels", "int", False, 1),
    ("OUTPUT_CHANNELS", "output_channels", "int", False, 1),
    ("N_CHANNELS", "input_channels", "int", False, 1),
    ("NUM_AUDIO_INPUTS", "num_audio_inputs", "int", False, 1),
    ("PAD_HEAD", "pad_head", "int", False, 0),
    ("ENC_STRIDE", "enc_stride", "int", False, 0),
    ("OUTPUT_SOURCES", "output_sources", "int", False, 1),
    ("ORIGINAL_SAMPLE_RATE", "original_sample_rate", "int", False, lambda c: c["IN_SAMPLE_RATE"]),
    ("SUPER_SAMPLE_RATE", "super_sample_rate", "int", False, lambda c: c["OUT_SAMPLE_RATE"]),
    ("SCALE_FACTOR", "scale_factor", "float", False, lambda c: float(c["OUT_SAMPLE_RATE"] / c["IN_SAMPLE_RATE"])),
)


def runtime_config_from_metadata(reader: MetadataReader) -> dict:
    cfg: dict = {}
    for constant, key, kind, required, default in RUNTIME_SCHEMA:
        value = reader.get(key, kind, required)
        if value is None:
            value = default(cfg) if callable(default) else default
        cfg[constant] = value
    return cfg

"""Run-time configuration from the model's string metadata -- mirrors the reference's
`MetadataReader` / `runtime_config_from_metadata` (audio_onnx_metadata.py:247-303,354-386):
same keys, defaults, error text for a missing required key, and the same 23 constants."""
from __future__ import annotations

REQUIRED_AUDIO_METADATA_KEYS = (
    "audio_metadata_version", "producer", "model_name", "task", "model_family", "dynamic_axes", "opset",
    "input_audio_dtype", "output_audio_dtype", "in_sample_rate", "out_sample_rate", "model_sample_rate",
    "input_audio_length", "input_to_output_scale", "max_dynamic_audio_seconds", "normalize_audio_default",
    "normalize_target_rms",
)


def _missing(key):
    return (f"Required metadata key {key} is missing. "
            "Re-export with the matching Export_*.py and rerun Optimize_ONNX.py.")


def _parse_bool(value, key):
    v = str(value).strip().lower()
    if v in {"1", "true", "yes", "on"}:
        return True
    if v in {"0", "false", "no", "off"}:
        return False
    raise ValueError(f"Metadata key {key} must be a boolean encoded as 1/0, got {value!r}.")


class MetadataReader:
    def __init__(self, metadata):
        self.metadata = dict(metadata or {})

    def string(self, key, default=None, required=False):
        value = self.metadata.get(key)
        if value is None or value == "":
            if required:
                raise KeyError(_missing(key))
            return default
        return value

    def required_int(self, key):
        return int(self.string(key, required=True))

    def optional_int(self, key, default=None):
        v = self.string(key)
        return default if v is None else int(v)

    def required_float(self, key):
        return float(self.string(key, required=True))

    def optional_float(self, key, default=None):
        v = self.string(key)
        return default if v is None else float(v)

    def required_bool(self, key):
        return _parse_bool(self.string(key, required=True), key)

    def optional_bool(self, key, default=None):
        v = self.string(key)
        return default if v is None else _parse_bool(v, key)


def load_runtime_metadata(session, required_keys=REQUIRED_AUDIO_METADATA_KEYS) -> MetadataReader:
    reader = MetadataReader(session.get_modelmeta().custom_metadata_map or {})
    for key in required_keys:
        reader.string(key, required=True)
    return reader


def runtime_config_from_metadata(reader: MetadataReader) -> dict:
    in_sr = reader.required_int("in_sample_rate")
    out_sr = reader.required_int("out_sample_rate")
    model_sr = reader.required_int("model_sample_rate")
    fold_w = reader.optional_int("fold_window_length", 0)
    return {
        "IN_SAMPLE_RATE": in_sr,
        "OUT_SAMPLE_RATE": out_sr,
        "MODEL_SAMPLE_RATE": model_sr,
        "INPUT_TO_OUTPUT_SCALE": reader.required_float("input_to_output_scale"),
        "BATCH_WINDOW_SECONDS": reader.optional_float("batch_window_seconds", 0.0),
        "HOP_LENGTH": reader.optional_int("hop_length", 0),
        "FOLD_WINDOW_LENGTH": fold_w,
        "FOLD_INPUT_LENGTH": reader.optional_int(
            "fold_input_length", max(1, int(round(fold_w * in_sr / model_sr))) if fold_w else 0),
        "BATCH_FOLD_INFERENCE": reader.optional_bool("batch_fold_inference_default", False),
        "MAX_DYNAMIC_AUDIO_SECONDS": reader.required_int("max_dynamic_audio_seconds"),
        "NORMALIZE_AUDIO": reader.required_bool("normalize_audio_default"),
        "NORMALIZE_TARGET_RMS": reader.required_float("normalize_target_rms"),
        "INPUT_CHANNELS": reader.optional_int("input_channels", 1),
        "OUTPUT_CHANNELS": reader.optional_int("output_channels", 1),
        "N_CHANNELS": reader.optional_int("input_channels", 1),
        "NUM_AUDIO_INPUTS": reader.optional_int("num_audio_inputs", 1),
        "PAD_HEAD": reader.optional_int("pad_head", 0),
        "ENC_STRIDE": reader.optional_int("enc_stride", 0),
        "OUTPUT_SOURCES": reader.optional_int("output_sources", 1),
        "ORIGINAL_SAMPLE_RATE": reader.optional_int("original_sample_rate", in_sr),
        "SUPER_SAMPLE_RATE": reader.optional_int("super_sample_rate", out_sr),
        "SCALE_FACTOR": reader.optional_float("scale_factor", float(out_sr / in_sr)),
    }

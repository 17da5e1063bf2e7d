"""Model export: reference `state_dict` -> `.adn` model file (the build's counterpart of the
`Export_*.py` main blocks, e.g. GTCRN/Export_GTCRN.py:705-792, minus ONNX)."""
from __future__ import annotations

from . import gtcrn_params, modelfile


def export_gtcrn(state_dict: dict, path, input_audio_length: int = 16000, in_dtype: str = "INT16",
                 out_dtype: str = "INT16", in_rate: int = 16000, out_rate: int = 16000) -> dict[str, str]:
    """Writes `path` (.adn) for one static chunk length; returns the metadata stamped."""
    md = gtcrn_params.metadata(input_audio_length, in_dtype, out_dtype, in_rate, out_rate)
    tensors = gtcrn_params.pack(state_dict, input_audio_length, in_rate)
    modelfile.save(path, md, tensors)
    return md


def gtcrn_model(state_dict: dict, input_audio_length: int = 16000, in_dtype: str = "F32",
                out_dtype: str = "F32", device_id: int = 0, in_rate: int = 16000, out_rate: int = 16000):
    """In-memory shortcut: build a `Model` without touching the file system."""
    from .model import Model

    md = gtcrn_params.metadata(input_audio_length, in_dtype, out_dtype, in_rate, out_rate)
    return Model.from_tensors(md, gtcrn_params.pack(state_dict, input_audio_length, in_rate), device_id)


def export_mbr(state_dict: dict, path, hyper=None, input_audio_length: int = 66150, in_dtype: str = "INT16",
               out_dtype: str = "INT16", in_rate: int = 44100, out_rate: int = 44100) -> dict[str, str]:
    """Mel-Band-Roformer (stereo) `.adn` for one static window length (counterpart of
    Mel_Band_Roformer/Stereo/Export_MelBandRoformer.py:690-735)."""
    from . import mbr_params

    hyper = hyper or mbr_params.MbrHyper()
    md = mbr_params.metadata(hyper, input_audio_length, in_dtype, out_dtype, in_rate, out_rate)
    modelfile.save(path, md, mbr_params.pack(state_dict, hyper, input_audio_length, in_rate))
    return md


def mbr_model(state_dict: dict, hyper=None, input_audio_length: int = 66150, in_dtype: str = "F32",
              out_dtype: str = "F32", device_id: int = 0, in_rate: int = 44100, out_rate: int = 44100):
    from . import mbr_params
    from .model import Model

    hyper = hyper or mbr_params.MbrHyper()
    md = mbr_params.metadata(hyper, input_audio_length, in_dtype, out_dtype, in_rate, out_rate)
    return Model.from_tensors(md, mbr_params.pack(state_dict, hyper, input_audio_length, in_rate), device_id)


def export_mf2se(state_dict: dict, path, hyper=None, input_audio_length: int = 48000, in_dtype: str = "INT16",
                 out_dtype: str = "INT16", matmul_dtype: str = "F32", in_rate: int | None = None,
                 out_rate: int | None = None) -> dict[str, str]:
    """MossFormer2-SE-48K `.adn` for one static window length (counterpart of
    MossFormer2_SE_48K/Export_MossFormer_SE.py:519-563).  `state_dict` keys: see adn/mf2se_params.py."""
    from . import mf2se_params

    hyper = hyper or mf2se_params.Mf2Hyper()
    md = mf2se_params.metadata(hyper, input_audio_length, in_dtype, out_dtype, matmul_dtype, in_rate, out_rate)
    modelfile.save(path, md, mf2se_params.pack(state_dict, hyper, input_audio_length, in_rate))
    return md


def mf2se_model(state_dict: dict, hyper=None, input_audio_length: int = 48000, in_dtype: str = "F32",
                out_dtype: str = "F32", device_id: int = 0, matmul_dtype: str = "F32", in_rate: int | None = None,
                out_rate: int | None = None):
    from . import mf2se_params
    from .model import Model

    hyper = hyper or mf2se_params.Mf2Hyper()
    md = mf2se_params.metadata(hyper, input_audio_length, in_dtype, out_dtype, matmul_dtype, in_rate, out_rate)
    return Model.from_tensors(md, mf2se_params.pack(state_dict, hyper, input_audio_length, in_rate), device_id)


def export_mf2ss(state_dict: dict, path, hyper=None, input_audio_length: int = 16000, in_dtype: str = "INT16",
                 out_dtype: str = "INT16", in_rate: int | None = None, out_rate: int | None = None) -> dict[str, str]:
    """MossFormer2-SS-16K `.adn` for one static window length (counterpart of
    MossFormer2_SS_16K/Export_MossFormer2_SS_16K.py:672-709).  `state_dict` keys: see adn/mf2ss_params.py."""
    from . import mf2ss_params

    hyper = hyper or mf2ss_params.SsHyper()
    md = mf2ss_params.metadata(hyper, input_audio_length, in_dtype, out_dtype, in_rate, out_rate)
    modelfile.save(path, md, mf2ss_params.pack(state_dict, hyper, input_audio_length, in_rate))
    return md


def mf2ss_model(state_dict: dict, hyper=None, input_audio_length: int = 16000, in_dtype: str = "F32",
                out_dtype: str = "F32", device_id: int = 0, in_rate: int | None = None, out_rate: int | None = None):
    from . import mf2ss_params
    from .model import Model

    hyper = hyper or mf2ss_params.SsHyper()
    md = mf2ss_params.metadata(hyper, input_audio_length, in_dtype, out_dtype, in_rate, out_rate)
    return Model.from_tensors(md, mf2ss_params.pack(state_dict, hyper, input_audio_length, in_rate), device_id)


def export_mfgan(state_dict: dict, path, hyper=None, input_audio_length: int = 16000, in_dtype: str = "INT16",
                 out_dtype: str = "INT16", in_rate: int | None = None, out_rate: int | None = None) -> dict[str, str]:
    """MossFormerGAN-SE-16K `.adn` for one static window length (counterpart of
    MossFormerGAN_SE_16K/Export_MossFormer_SE.py:900-950).  `state_dict` keys: see adn/mfgan_params.py."""
    from . import mfgan_params

    hyper = hyper or mfgan_params.GanHyper()
    md = mfgan_params.metadata(hyper, input_audio_length, in_dtype, out_dtype, in_rate, out_rate)
    modelfile.save(path, md, mfgan_params.pack(state_dict, hyper, input_audio_length, in_rate))
    return md


def mfgan_model(state_dict: dict, hyper=None, input_audio_length: int = 16000, in_dtype: str = "F32",
                out_dtype: str = "F32", device_id: int = 0, in_rate: int | None = None, out_rate: int | None = None):
    from . import mfgan_params
    from .model import Model

    hyper = hyper or mfgan_params.GanHyper()
    md = mfgan_params.metadata(hyper, input_audio_length, in_dtype, out_dtype, in_rate, out_rate)
    return Model.from_tensors(md, mfgan_params.pack(state_dict, hyper, input_audio_length, in_rate), device_id)


def export_zipenh(state_dict: dict, path, hyper=None, input_audio_length: int = 16000, in_dtype: str = "INT16",
                  out_dtype: str = "INT16") -> dict[str, str]:
    """ZipEnhancer `.adn` for one static window length at 16 kHz (counterpart of ZipEnhancer/Export_ZipEnhancer.py:938-1001).
    `state_dict` keys: see adn/zipenh_params.py."""
    from . import zipenh_params

    hyper = hyper or zipenh_params.ZipHyper()
    md = zipenh_params.metadata(hyper, input_audio_length, in_dtype, out_dtype)
    modelfile.save(path, md, zipenh_params.pack(state_dict, hyper, input_audio_length))
    return md


def zipenh_model(state_dict: dict, hyper=None, input_audio_length: int = 16000, in_dtype: str = "F32", out_dtype: str = "F32",
                 device_id: int = 0):
    from . import zipenh_params
    from .model import Model

    hyper = hyper or zipenh_params.ZipHyper()
    md = zipenh_params.metadata(hyper, input_audio_length, in_dtype, out_dtype)
    return Model.from_tensors(md, zipenh_params.pack(state_dict, hyper, input_audio_length), device_id)


def export_dfsmn(state_dict: dict, path, hyper=None, input_audio_length: int = 96000, in_dtype: str = "INT16",
                 out_dtype: str = "INT16") -> dict[str, str]:
    """DFSMN (48 kHz) `.adn` for one static window length (counterpart of DFSMN/Export_DFSMN.py:252-324).
    `state_dict` keys: see adn/dfsmn_params.py."""
    from . import dfsmn_params

    hyper = hyper or dfsmn_params.DfsmnHyper()
    md = dfsmn_params.metadata(hyper, input_audio_length, in_dtype, out_dtype)
    modelfile.save(path, md, dfsmn_params.pack(state_dict, hyper, input_audio_length))
    return md


def dfsmn_model(state_dict: dict, hyper=None, input_audio_length: int = 96000, in_dtype: str = "F32",
                out_dtype: str = "F32", device_id: int = 0):
    from . import dfsmn_params
    from .model import Model

    hyper = hyper or dfsmn_params.DfsmnHyper()
    md = dfsmn_params.metadata(hyper, input_audio_length, in_dtype, out_dtype)
    return Model.from_tensors(md, dfsmn_params.pack(state_dict, hyper, input_audio_length), device_id)


def export_ulunas(state_dict: dict, path, input_audio_length: int = 32000, in_dtype: str = "INT16", out_dtype: str = "INT16") -> dict[str, str]:
    """UL-UNAS `.adn` for one static window length (counterpart of UL-UNAS/Export_UL_UNAS.py:918-1012).  `state_dict`: the raw
    `ULUNAS().state_dict()` (the checkpoint after the reference's own `convert_state_dict`), see adn/ulunas_params.py."""
    from . import ulunas_params

    md = ulunas_params.metadata(input_audio_length, in_dtype, out_dtype)
    modelfile.save(path, md, ulunas_params.pack(state_dict, input_audio_length, in_dtype, out_dtype))
    return md


def ulunas_model(state_dict: dict, input_audio_length: int = 32000, in_dtype: str = "F32", out_dtype: str = "F32", device_id: int = 0):
    from . import ulunas_params
    from .model import Model

    md = ulunas_params.metadata(input_audio_length, in_dtype, out_dtype)
    return Model.from_tensors(md, ulunas_params.pack(state_dict, input_audio_length, in_dtype, out_dtype), device_id)


def export_hgtcrn(state_dict: dict, path, input_audio_length: int = 16128, in_dtype: str = "INT16", out_dtype: str = "INT16") -> dict[str, str]:
    """H-GTCRN (two microphones, WPE + AuxIVA front end) `.adn` for one static window of k * 256 samples at 16 kHz (counterpart of
    H-GTCRN/Export_H_GTCRN.py:1075-1186).  `state_dict`: the raw `GTCRN_IVA` state_dict (reference key names)."""
    from . import hgtcrn_params

    md = hgtcrn_params.metadata(input_audio_length, in_dtype, out_dtype)
    modelfile.save(path, md, hgtcrn_params.pack(state_dict, input_audio_length))
    return md


def hgtcrn_model(state_dict: dict, input_audio_length: int = 16128, in_dtype: str = "F32", out_dtype: str = "F32", device_id: int = 0):
    from . import hgtcrn_params
    from .model import Model

    md = hgtcrn_params.metadata(input_audio_length, in_dtype, out_dtype)
    return Model.from_tensors(md, hgtcrn_params.pack(state_dict, input_audio_length), device_id)

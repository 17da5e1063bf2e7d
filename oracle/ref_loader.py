"""TEST INFRASTRUCTURE ONLY -- loader that executes the *reference's own* PyTorch
definitions from /root/reference (read-only, present in the build container only).

It is used by `oracle/make_golden.py` to generate the committed fixtures under
`tests/golden/` and by the `-m "not gpu"` pinning tests (skipped when
/root/reference is absent, e.g. on the GPU box).  Nothing in the product path
(`audio-denoiser-onnx_b200/`), `bench.py` or `smoke()` imports this file.

Recipe: the reference's `Export_*.py` files run their ONNX
export at import time and import `onnx`/`onnxruntime`/`onnxslim`, none of which is
installed here.  We therefore (1) register stub modules for those names, (2) read the
source, cut it before `def _run_inference_demo` (i.e. keep only the class
definitions), (3) text-patch the module-level constants, (4) `exec` it.
No reference source is copied into this repository.
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

REFERENCE_ROOT = Path("/root/reference")


def reference_available() -> bool:
    return (REFERENCE_ROOT / "GTCRN" / "Export_GTCRN.py").exists()


def _install_stubs() -> None:
    for name in ("onnxruntime", "onnx", "onnxslim", "ml_collections"):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            if name == "onnxslim":
                mod.slim = lambda *a, **k: None
            if name == "ml_collections":
                mod.ConfigDict = dict
            sys.modules[name] = mod
    if "Rewrite_ONNX_GRU_Zero_State" not in sys.modules:
        mod = types.ModuleType("Rewrite_ONNX_GRU_Zero_State")
        mod.rewrite = lambda *a, **k: None
        sys.modules["Rewrite_ONNX_GRU_Zero_State"] = mod


def load_stft_module(model_dir: str):
    """Import `<model_dir>/STFT_Process.py` from the reference as a fresh module.

    Each model folder carries its own variant, so the module is
    loaded under a unique name."""
    import importlib.util

    _install_stubs()
    path = REFERENCE_ROOT / model_dir / "STFT_Process.py"
    name = "ref_stft_" + model_dir.replace("/", "_").replace("-", "_")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_export_namespace(model_dir: str, script: str, patches: dict[str, str]) -> dict:
    """Exec the class-definition part of an Export script with patched constants.

    `patches` maps an exact source line prefix (e.g. ``"INPUT_AUDIO_LENGTH   = 32000"``)
    to its replacement text."""
    _install_stubs()
    mdir = REFERENCE_ROOT / model_dir
    for p in (str(REFERENCE_ROOT), str(mdir)):          # the model folder must come FIRST: every folder has its own STFT_Process.py
        while p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    # the per-model STFT_Process must win over a previously imported sibling variant
    sys.modules.pop("STFT_Process", None)
    src = (mdir / script).read_text()
    src = src.split("\ndef _run_inference_demo")[0]
    for old, new in patches.items():
        if old not in src:
            raise RuntimeError(f"patch anchor not found in {script}: {old!r}")
        src = src.replace(old, new, 1)
    ns: dict = {"__file__": str(mdir / script),
                "__name__": "ref_export_" + model_dir.replace("/", "_").replace("-", "_")}
    exec(compile(src, str(mdir / script), "exec"), ns)
    sys.modules.pop("STFT_Process", None)
    return ns


def load_gtcrn(input_audio_length: int = 16000, io_dtype: str = "F32", in_rate: int = 16000, out_rate: int = 16000,
               model_rate_frames: bool = False):
    """Build the reference GTCRN_CUSTOM wrapper (random init, eval) for one chunk length.

    Returns (namespace, build) where build(state_dict|None) -> wrapper module.
    Static buffers are baked per chunk length (Export_GTCRN.py:44-46,234-239).

    model_rate_frames: the static export sizes its frame count from the INPUT-rate length (:45), so as shipped the
    forward's reshape (:665-667) fails for IN_SAMPLE_RATE != 16 kHz.  With this flag that ONE constant is patched to the
    frame count of the resampled (model-rate) signal, floor(L * 16000 / in_rate) // HOP + 1 -- every line of arithmetic
    that then runs (resample / PCM scale / centring order, STFT, network, ISTFT) is the reference's own."""
    import torch

    patches = {
        "INPUT_AUDIO_LENGTH   = 32000": f"INPUT_AUDIO_LENGTH   = {int(input_audio_length)}",
        "IN_AUDIO_DTYPE       = 'INT16'": f"IN_AUDIO_DTYPE       = '{io_dtype}'",
        "OUT_AUDIO_DTYPE      = 'INT16'": f"OUT_AUDIO_DTYPE      = '{io_dtype}'",
        "IN_SAMPLE_RATE       = 16000": f"IN_SAMPLE_RATE       = {int(in_rate)}",
        "OUT_SAMPLE_RATE      = 16000": f"OUT_SAMPLE_RATE      = {int(out_rate)}",
    }
    if model_rate_frames:
        import math
        model_len = int(math.floor(float(input_audio_length) * (1.0 / (in_rate / 16000.0))))   # F.interpolate(scale_factor=...)
        patches["STATIC_SIGNAL_LENGTH = None if DYNAMIC_AXES else ((FOLD_WINDOW_LENGTH if USE_BATCH_FOLD else EXPORT_AUDIO_LENGTH) // HOP_LENGTH + 1)"] = \
            f"STATIC_SIGNAL_LENGTH = {model_len // 256 + 1}"
    ns = load_export_namespace("GTCRN", "Export_GTCRN.py", patches)

    def build(state_dict=None):
        with torch.inference_mode():
            STFT_Process = ns["STFT_Process"]
            stft = STFT_Process(
                model_type="stft_B", n_fft=ns["NFFT"], hop_len=ns["HOP_LENGTH"],
                win_length=ns["WINDOW_LENGTH"], max_frames=0, window_type=ns["WINDOW_TYPE"],
                center_pad=True, pad_mode=ns["PAD_MODE"], input_scale=1.0,
            ).eval()
            istft = STFT_Process(
                model_type="istft_B", n_fft=ns["NFFT"], hop_len=ns["HOP_LENGTH"],
                win_length=ns["WINDOW_LENGTH"], max_frames=ns["MAX_SIGNAL_LENGTH"],
                window_type=ns["WINDOW_TYPE"], center_pad=True, pad_mode=ns["PAD_MODE"],
                output_scale=1.0, static_norm=True,
            ).eval()
            g = ns["GTCRN"]().eval()
            if state_dict is not None:
                missing, unexpected = g.load_state_dict(state_dict, strict=False)
                assert not unexpected, unexpected
            g.prepare_for_export_()
            wrapper = ns["GTCRN_CUSTOM"](
                g.float(), stft, istft, ns["IN_SAMPLE_RATE"], ns["OUT_SAMPLE_RATE"],
                False, ns["FOLD_WINDOW_LENGTH"],
            ).eval()
        return wrapper

    return ns, build


def load_mbr(input_audio_length: int, io_dtype: str = "F32", out_rate: int = 44100, in_rate: int = 44100):
    """Reference Mel-Band-Roformer (stereo) wrapper for one un-folded window.

    Returns (namespace, build) with build(state_dict, **model_kwargs) -> module.  The
    checkpoint load inside `MelBandRoformer.__init__` (Export_MelBandRoformer.py:391-394) is
    redirected to the given state_dict by patching `torch.load` for the duration of the call."""
    import torch

    patches = {
        "INPUT_AUDIO_LENGTH  = 88200": f"INPUT_AUDIO_LENGTH  = {int(input_audio_length)}",
        "USE_BATCH_FOLD       = True": "USE_BATCH_FOLD       = False",
        "IN_AUDIO_DTYPE        = 'INT16'": f"IN_AUDIO_DTYPE        = '{io_dtype}'",
        "OUT_AUDIO_DTYPE       = 'INT16'": f"OUT_AUDIO_DTYPE       = '{io_dtype}'",
        "OUT_SAMPLE_RATE   = 44100": f"OUT_SAMPLE_RATE   = {int(out_rate)}",
    }
    if in_rate != 44100:
        # As shipped MAX_SIGNAL_LENGTH is sized from the INPUT-rate length (:50), so the static frame count does not match the
        # resampled signal; that ONE constant is patched to the model-rate frame count, floor(L * 44100 / in_rate) // HOP + 1,
        # and the reference's own forward (:629-680) is executed.
        import math
        model_len = int(math.floor(float(input_audio_length) * float(44100 / in_rate)))
        patches["IN_SAMPLE_RATE    = 44100"] = f"IN_SAMPLE_RATE    = {int(in_rate)}"
        patches["MAX_SIGNAL_LENGTH    = 2048 if DYNAMIC_AXES else (((FOLD_WINDOW_LENGTH if USE_BATCH_FOLD else INPUT_AUDIO_LENGTH) // HOP_LENGTH) + 1)"] = \
            f"MAX_SIGNAL_LENGTH    = {model_len // 441 + 1}"
    ns = load_export_namespace("Mel_Band_Roformer/Stereo", "Export_MelBandRoformer.py", patches)

    def build(state_dict, **model_kwargs):
        real_load = torch.load
        torch.load = lambda *a, **k: state_dict
        try:
            with torch.inference_mode():
                S = ns["STFT_Process"]
                stft = S(model_type="stft_B", n_fft=ns["NFFT"], hop_len=ns["HOP_LENGTH"], win_length=ns["WINDOW_LENGTH"],
                         max_frames=0, window_type=ns["WINDOW_TYPE"], center_pad=True, pad_mode="reflect").eval()
                istft = S(model_type="istft_B", n_fft=ns["NFFT"], hop_len=ns["HOP_LENGTH"],
                          win_length=ns["WINDOW_LENGTH"], max_frames=ns["MAX_SIGNAL_LENGTH"],
                          window_type=ns["WINDOW_TYPE"], center_pad=True, pad_mode="reflect", static_frames=True).eval()
                m = ns["MelBandRoformer"](stft, istft, ns["MAX_SIGNAL_LENGTH"], False, ns["FOLD_WINDOW_LENGTH"],
                                          ns["EXPORT_AUDIO_LENGTH"], **model_kwargs).eval()
        finally:
            torch.load = real_load
        return m

    return ns, build


def load_mf2se(input_audio_length: int, io_dtype: str = "F32", in_rate: int = 48000, out_rate: int = 48000):
    """Reference MossFormer2-SE-48K wrapper (`MOSSFORMER_SE`) for one un-folded window.

    The wrapper's forward is made of leaf ops; only its constructor reads the absent
    `clearvoice` model.  Returns (namespace, build) with
    build(holder) -> wrapper, where `holder` is `mf2se_oracle.skeleton()` carrying the weights."""
    import torch

    for name in ("clearvoice", "clearvoice.models", "clearvoice.models.mossformer2_se",
                 "clearvoice.models.mossformer2_se.mossformer2_se_wrapper"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["clearvoice.models.mossformer2_se.mossformer2_se_wrapper"].MossFormer2_SE_48K = object

    ns = load_export_namespace(
        "MossFormer2_SE_48K",
        "Export_MossFormer_SE.py",
        {
            "INPUT_AUDIO_LENGTH = 96000": f"INPUT_AUDIO_LENGTH = {int(input_audio_length)}",
            "IN_AUDIO_DTYPE     = 'INT16'": f"IN_AUDIO_DTYPE     = '{io_dtype}'",
            "OUT_AUDIO_DTYPE    = 'INT16'": f"OUT_AUDIO_DTYPE    = '{io_dtype}'",
            "IN_SAMPLE_RATE     = 48000": f"IN_SAMPLE_RATE     = {int(in_rate)}",
            "OUT_SAMPLE_RATE    = 48000": f"OUT_SAMPLE_RATE    = {int(out_rate)}",
        },
    )

    def build(holder):
        with torch.inference_mode():
            S = ns["STFT_Process"]
            stft = S(model_type="stft_B", n_fft=ns["NFFT"], hop_len=ns["HOP_LENGTH"], win_length=ns["WINDOW_LENGTH"],
                     max_frames=0, window_type=ns["WINDOW_TYPE"], center_pad=False, pad_mode="constant").eval()
            istft = S(model_type="istft_B", n_fft=ns["NFFT"], hop_len=ns["HOP_LENGTH"], win_length=ns["WINDOW_LENGTH"],
                      max_frames=ns["MAX_SIGNAL_LENGTH"], window_type=ns["WINDOW_TYPE"], center_pad=False,
                      pad_mode="constant", static_frames=True).eval()
            outer = types.SimpleNamespace(mossformer=holder.eval().float())
            return ns["MOSSFORMER_SE"](outer, stft, istft, ns["NFFT"], ns["N_MELS"], ns["IN_SAMPLE_RATE"],
                                       ns["OUT_SAMPLE_RATE"], ns["MAX_SIGNAL_LENGTH"], False, ns["FOLD_WINDOW_LENGTH"]).eval()

    return ns, build


def load_mf2ss(input_audio_length: int, io_dtype: str = "F32", in_rate: int = 16000, out_rate: int = 16000):
    """Reference MossFormer2-SS-16K wrapper (`MOSSFORMER_SS`) for one un-folded window.

    The wrapper's forward is made of leaf ops on packed buffers; only its constructor reads the
    absent `clearvoice` model.  Returns (namespace, build) with
    build(holder) -> wrapper, where `holder` is `mf2ss_oracle.skeleton()` carrying the weights."""
    import torch

    for name in ("clearvoice", "clearvoice.models", "clearvoice.models.mossformer2_ss",
                 "clearvoice.models.mossformer2_ss.mossformer2"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["clearvoice.models.mossformer2_ss.mossformer2"].MossFormer2_SS_16K = object

    ns = load_export_namespace(
        "MossFormer2_SS_16K",
        "Export_MossFormer2_SS_16K.py",
        {
            "INPUT_AUDIO_LENGTH      = 32000": f"INPUT_AUDIO_LENGTH      = {int(input_audio_length)}",
            "IN_AUDIO_DTYPE          = 'INT16'": f"IN_AUDIO_DTYPE          = '{io_dtype}'",
            "OUT_AUDIO_DTYPE         = 'INT16'": f"OUT_AUDIO_DTYPE         = '{io_dtype}'",
            "USE_BATCH_FOLD          = True": "USE_BATCH_FOLD          = False",
            "IN_SAMPLE_RATE          = 16000": f"IN_SAMPLE_RATE          = {int(in_rate)}",
            "OUT_SAMPLE_RATE         = 16000": f"OUT_SAMPLE_RATE         = {int(out_rate)}",
        },
    )

    def build(holder):
        with torch.inference_mode():
            return ns["MOSSFORMER_SS"](holder.eval().float(), ns["INPUT_AUDIO_LENGTH"], ns["IN_SAMPLE_RATE"],
                                       ns["OUT_SAMPLE_RATE"], False, ns["FOLD_WINDOW_LENGTH"]).eval()

    return ns, build


def load_mfgan(input_audio_length: int, io_dtype: str = "F32", in_rate: int = 16000, out_rate: int = 16000):
    """Reference MossFormerGAN-SE-16K wrapper (`MOSSFORMER_SE` of MossFormerGAN_SE_16K/Export_MossFormer_SE.py)
    for one un-folded window.  The wrapper's forward is made of leaf ops; its constructor and forward read
    parameters off the absent `clearvoice` generator.  Returns (namespace, build) with
    build(holder) -> wrapper, where `holder` is `mfgan_oracle.skeleton()` carrying the weights."""
    import torch

    for name in ("clearvoice", "clearvoice.models", "clearvoice.models.mossformer_gan_se",
                 "clearvoice.models.mossformer_gan_se.generator"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["clearvoice.models.mossformer_gan_se.generator"].MossFormerGAN_SE_16K = object
    if "Rewrite_ONNX_Asymmetric_Padding" not in sys.modules:
        mod = types.ModuleType("Rewrite_ONNX_Asymmetric_Padding")
        mod.rewrite_asymmetric_causal_convs = lambda *a, **k: None
        sys.modules["Rewrite_ONNX_Asymmetric_Padding"] = mod

    ns = load_export_namespace(
        "MossFormerGAN_SE_16K",
        "Export_MossFormer_SE.py",
        {
            "INPUT_AUDIO_LENGTH   = 32000": f"INPUT_AUDIO_LENGTH   = {int(input_audio_length)}",
            "IN_AUDIO_DTYPE       = 'INT16'": f"IN_AUDIO_DTYPE       = '{io_dtype}'",
            "OUT_AUDIO_DTYPE      = 'INT16'": f"OUT_AUDIO_DTYPE      = '{io_dtype}'",
            "USE_BATCH_FOLD       = True": "USE_BATCH_FOLD       = False",
            "IN_SAMPLE_RATE       = 16000": f"IN_SAMPLE_RATE       = {int(in_rate)}",
            "OUT_SAMPLE_RATE      = 16000": f"OUT_SAMPLE_RATE      = {int(out_rate)}",
        },
    )

    def build(holder):
        with torch.inference_mode():
            S = ns["STFT_Process"]
            stft = S(model_type="stft_C", n_fft=ns["NFFT"], hop_len=ns["HOP_LENGTH"], win_length=ns["WINDOW_LENGTH"],
                     max_frames=0, window_type=ns["WINDOW_TYPE"], center_pad=True, pad_mode="reflect").eval()
            istft = S(model_type="istft_C", n_fft=ns["NFFT"], hop_len=ns["HOP_LENGTH"], win_length=ns["WINDOW_LENGTH"],
                      max_frames=ns["MAX_SIGNAL_LENGTH"], window_type=ns["WINDOW_TYPE"], center_pad=True, pad_mode="reflect",
                      precompute_window_sum=True).eval()
            return ns["MOSSFORMER_SE"](holder.eval().float(), stft, istft, ns["IN_SAMPLE_RATE"], ns["OUT_SAMPLE_RATE"],
                                       False, ns["FOLD_WINDOW_LENGTH"]).eval()

    return ns, build


def load_dfsmn(input_audio_length: int, io_dtype: str = "F32"):
    """Reference DFSMN wrapper (`DFSMN` of DFSMN/Export_DFSMN.py) for one un-folded window at 48 kHz.  The wrapper's
    forward is made of leaf ops; its constructor reads the weights off the absent `modelscope` pipeline model.
    Returns (namespace, build) with build(holder) -> wrapper, where `holder` is `dfsmn_oracle.skeleton()`."""
    import torch

    for name in ("modelscope", "modelscope.pipelines", "modelscope.utils", "modelscope.utils.constant"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["modelscope.pipelines"].pipeline = lambda *a, **k: None
    sys.modules["modelscope.utils.constant"].Tasks = types.SimpleNamespace(acoustic_noise_suppression="ans")
    if "Rewrite_ONNX_Causal_Padding" not in sys.modules:
        mod = types.ModuleType("Rewrite_ONNX_Causal_Padding")
        mod.rewrite_causal_fsmn_padding = lambda *a, **k: None
        sys.modules["Rewrite_ONNX_Causal_Padding"] = mod

    ns = load_export_namespace(
        "DFSMN",
        "Export_DFSMN.py",
        {
            "INPUT_AUDIO_LENGTH    = 96000": f"INPUT_AUDIO_LENGTH    = {int(input_audio_length)}",
            "IN_AUDIO_DTYPE        = 'INT16'": f"IN_AUDIO_DTYPE        = '{io_dtype}'",
            "OUT_AUDIO_DTYPE       = 'INT16'": f"OUT_AUDIO_DTYPE       = '{io_dtype}'",
        },
    )

    def build(holder):
        with torch.inference_mode():
            S = ns["STFT_Process"]
            stft = S(model_type="stft_B", n_fft=ns["NFFT_STFT"], win_length=ns["WINDOW_LENGTH"], hop_len=ns["HOP_LENGTH"],
                     max_frames=0, window_type=ns["WINDOW_TYPE"], center_pad=False, pad_mode="constant").eval()
            istft = S(model_type="istft_B", n_fft=ns["NFFT_STFT"], win_length=ns["WINDOW_LENGTH"], hop_len=ns["HOP_LENGTH"],
                      max_frames=ns["MAX_SIGNAL_LENGTH"], window_type=ns["ISTFT_WINDOW_TYPE"], center_pad=False,
                      pad_mode="constant", static_norm=True).eval()
            return ns["DFSMN"](holder.eval().float(), stft, istft, ns["NFFT_STFT"], ns["N_MELS"], ns["IN_SAMPLE_RATE"],
                               ns["OUT_SAMPLE_RATE"], use_batch_fold=False, fold_window=ns["FOLD_WINDOW_LENGTH"], static_batch=1).eval()

    return ns, build


def load_ulunas(input_audio_length: int = 16000, io_dtype: str = "F32"):
    """Reference UL-UNAS wrapper (`ULUNAS_CUSTOM` of UL-UNAS/Export_UL_UNAS.py) for one un-folded window at 16 kHz.
    The model definition is self-contained in the script (like GTCRN's).  Returns (namespace, build) with
    build(state_dict | None, seed) -> (wrapper, raw_state_dict): `ULUNAS()` with seeded default init and randomised BatchNorm
    statistics (so the BN fold is exercised) unless a raw state_dict is given, then `prepare_for_export_` and the wrapper exactly
    as the script's main does (:938-975)."""
    import torch

    ns = load_export_namespace(
        "UL-UNAS",
        "Export_UL_UNAS.py",
        {
            "INPUT_AUDIO_LENGTH    = 32000": f"INPUT_AUDIO_LENGTH    = {int(input_audio_length)}",
            "IN_AUDIO_DTYPE        = 'INT16'": f"IN_AUDIO_DTYPE        = '{io_dtype}'",
            "OUT_AUDIO_DTYPE       = 'INT16'": f"OUT_AUDIO_DTYPE       = '{io_dtype}'",
        },
    )

    def build(state_dict=None, seed: int = 0):
        is_int = "int" in io_dtype.lower()
        with torch.inference_mode():
            S = ns["STFT_Process"]
            stft = S(model_type="stft_B", n_fft=ns["NFFT"], hop_len=ns["HOP_LENGTH"], win_length=ns["WINDOW_LENGTH"], max_frames=0,
                     window_type=ns["WINDOW_TYPE"], center_pad=True, pad_mode=ns["STFT_PAD_MODE"],
                     input_scale=ns["INV_INT16"] if is_int else 1.0).eval()
            istft = S(model_type="istft_B", n_fft=ns["NFFT"], hop_len=ns["HOP_LENGTH"], win_length=ns["WINDOW_LENGTH"],
                      max_frames=ns["MAX_SIGNAL_LENGTH"], window_type=ns["WINDOW_TYPE"], center_pad=True,
                      pad_mode=ns["STFT_PAD_MODE"], output_scale=32767.0 if is_int else 1.0, static_norm=True).eval()
            torch.manual_seed(seed)
            net = ns["ULUNAS"]().eval()
            if state_dict is None:
                g = torch.Generator().manual_seed(seed + 1)
                for m in net.modules():
                    if isinstance(m, torch.nn.BatchNorm2d):
                        m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=g))
                        m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
                        m.weight.copy_(1.0 + 0.2 * torch.randn(m.weight.shape, generator=g))
                        m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
            else:
                missing, unexpected = net.load_state_dict(state_dict, strict=False)
                assert not unexpected, unexpected
            raw = {k: v.clone() for k, v in net.state_dict().items()}
            net.prepare_for_export_()
            wrapper = ns["ULUNAS_CUSTOM"](net.float(), stft, istft, ns["IN_SAMPLE_RATE"], ns["OUT_SAMPLE_RATE"],
                                          remove_dc_offset=ns["REMOVE_DC_OFFSET"], use_batch_fold=False,
                                          fold_window=ns["FOLD_WINDOW_LENGTH"], input_scale_folded=is_int,
                                          output_scale_folded=is_int).eval()
        return wrapper, raw

    return ns, build


def load_hgtcrn(input_audio_length: int = 16128, io_dtype: str = "F32"):
    """Reference H-GTCRN wrapper (`H_GTCRN_CUSTOM` of H-GTCRN/Export_H_GTCRN.py: 2-channel STFT -> WPE -> AuxIVA -> GTCRN_IVA ->
    ISTFT) for one un-folded stereo window at 16 kHz; the model definition is self-contained in the script.  Returns
    (namespace, build) with build(state_dict | None, seed) -> (wrapper, raw_state_dict), constructed as the script's main does
    (:1075-1140) around `GTCRN_IVA()` with seeded default init and randomised BatchNorm statistics."""
    import torch

    ns = load_export_namespace(
        "H-GTCRN",
        "Export_H_GTCRN.py",
        {
            "INPUT_AUDIO_LENGTH   = 32000": f"INPUT_AUDIO_LENGTH   = {int(input_audio_length)}",
            "IN_AUDIO_DTYPE       = 'INT16'": f"IN_AUDIO_DTYPE       = '{io_dtype}'",
            "OUT_AUDIO_DTYPE      = 'INT16'": f"OUT_AUDIO_DTYPE      = '{io_dtype}'",
        },
    )

    def build(state_dict=None, seed: int = 0):
        with torch.inference_mode():
            S = ns["STFT_Process"]
            T = ns["MAX_SIGNAL_LENGTH"]
            stft = S(model_type="stft_B", n_fft=ns["NFFT"], hop_len=ns["HOP_LENGTH"], win_length=ns["WINDOW_LENGTH"], max_frames=0,
                     window_type=ns["WINDOW_TYPE"], center_pad=True, pad_mode=ns["PAD_MODE"], input_scale=1.0).eval()
            istft = S(model_type="istft_B", n_fft=ns["NFFT"], hop_len=ns["HOP_LENGTH"], win_length=ns["WINDOW_LENGTH"], max_frames=T,
                      window_type=ns["WINDOW_TYPE"], center_pad=True, pad_mode=ns["PAD_MODE"], output_scale=1.0, static_cola=True).eval()
            wpe = ns["OnnxFriendlyWPE"](n_channels=2, rt60=ns["WPE_RT60"], hop_length=ns["HOP_LENGTH"], delay=ns["WPE_DELAY"],
                                        sample_rate=16000, num_iter=ns["WPE_ITER"], ns_iter=ns["CG_SOLVE_ITER"], n_freq_bins=257,
                                        max_frames=T, batch_size=1, dynamic_frames=False).eval()
            iva = ns["OnnxFriendlyAuxIVA"](n_iter=ns["IVA_ITER"], n_channels=2, batch_size=1, n_frames=T).eval()
            torch.manual_seed(seed)
            net = ns["GTCRN_IVA"](batch_size=1, n_frames=T).eval()
            if state_dict is None:
                g = torch.Generator().manual_seed(seed + 1)
                for m in net.modules():
                    if isinstance(m, torch.nn.BatchNorm2d):
                        m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=g))
                        m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
                        m.weight.copy_(1.0 + 0.2 * torch.randn(m.weight.shape, generator=g))
                        m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
            else:
                missing, unexpected = net.load_state_dict(state_dict, strict=False)
                assert not unexpected, unexpected
            raw = {k: v.clone() for k, v in net.state_dict().items()}
            net.fuse_bn_()
            wrapper = ns["H_GTCRN_CUSTOM"](net, stft, istft, wpe, iva, n_fft=512, in_sample_rate=16000, out_sample_rate=16000,
                                           use_batch_fold=False, fold_window=ns["FOLD_WINDOW_LENGTH"],
                                           model_audio_length=ns["MODEL_AUDIO_LENGTH"], n_frames=T, frontend_batch=1,
                                           fold_input_pcm_scale=False, fold_output_pcm_scale=False).eval()
        return wrapper, raw

    return ns, build


def load_zipenh(input_audio_length: int = 16000, io_dtype: str = "F32"):
    """Reference ZipEnhancer wrapper (`ZipEnhancer` of ZipEnhancer/Export_ZipEnhancer.py) for one un-folded window at 16 kHz.
    The wrapper drives the un-vendored `modelscope` Zipformer2 modules; every forward with arithmetic in it is overridden in the
    reference file itself (`apply_onnx_export_patches`, :341-355).  The skeleton classes of `zipenh_oracle` (shapes only) are
    registered under the modelscope module names, the reference installs ITS forwards on them, and the reference wrapper runs
    around the holder.  Returns (namespace, build) with build(holder) -> wrapper."""
    import torch

    import zipenh_oracle as zo

    base = "modelscope.models.audio.ans.zipenhancer_layers"
    for name in ("modelscope", "modelscope.models", "modelscope.models.base", "modelscope.models.audio", "modelscope.models.audio.ans",
                 base, base + ".scaling", base + ".zipformer"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["modelscope.models.base"].Model = object
    scaling, zipformer = sys.modules[base + ".scaling"], sys.modules[base + ".zipformer"]
    sys.modules[base].scaling, sys.modules[base].zipformer = scaling, zipformer
    for cls in ("BiasNorm", "ActivationDropoutAndLinear"):
        setattr(scaling, cls, getattr(zo, cls))
    for cls in ("Zipformer2EncoderLayer", "BypassModule", "SimpleDownsample", "SimpleUpsample", "RelPositionMultiheadAttentionWeights",
                "SelfAttention", "NonlinAttention", "ConvolutionModule", "CompactRelPositionalEncoding"):
        setattr(zipformer, cls, getattr(zo, cls))
    if "Rewrite_ONNX_Asymmetric_Padding" not in sys.modules:
        mod = types.ModuleType("Rewrite_ONNX_Asymmetric_Padding")
        mod.rewrite_asymmetric_causal_convs = lambda *a, **k: None
        sys.modules["Rewrite_ONNX_Asymmetric_Padding"] = mod

    ns = load_export_namespace(
        "ZipEnhancer",
        "Export_ZipEnhancer.py",
        {
            "INPUT_AUDIO_LENGTH = 32000": f"INPUT_AUDIO_LENGTH = {int(input_audio_length)}",
            "IN_AUDIO_DTYPE  = 'INT16'": f"IN_AUDIO_DTYPE  = '{io_dtype}'",
            "OUT_AUDIO_DTYPE = 'INT16'": f"OUT_AUDIO_DTYPE = '{io_dtype}'",
            "USE_BATCH_FOLD        = True": "USE_BATCH_FOLD        = False",
        },
    )
    ns["apply_onnx_export_patches"]()

    def build(holder):
        with torch.inference_mode():
            S = ns["STFT_Process"]
            stft = S(model_type="stft_B", n_fft=ns["NFFT"], hop_len=ns["HOP_LENGTH"], win_length=ns["WINDOW_LENGTH"], max_frames=0,
                     window_type=ns["WINDOW_TYPE"], center_pad=True, pad_mode="reflect").eval()
            istft = S(model_type="istft_B", n_fft=ns["NFFT"], hop_len=ns["HOP_LENGTH"], win_length=ns["WINDOW_LENGTH"],
                      max_frames=ns["MAX_SIGNAL_LENGTH"], window_type=ns["WINDOW_TYPE"], center_pad=True, pad_mode="reflect",
                      static_norm=ns["STATIC_SHAPE"]).eval()
            return ns["ZipEnhancer"](holder.eval().float(), stft, istft, ns["IN_SAMPLE_RATE"], ns["OUT_SAMPLE_RATE"],
                                     use_batch_fold=False, fold_window=ns["FOLD_WINDOW_LENGTH"],
                                     use_rectangular_istft=ns["USE_RECTANGULAR_ISTFT"]).eval()

    return ns, build

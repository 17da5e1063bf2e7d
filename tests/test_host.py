"""CPU-side tests: C-ABI library loads and exports every symbol include/adn.h declares (no
compute calls without a GPU), host logic (model file, metadata, chunk planner, weight
packing) and the loud failure when no GPU is usable."""
import re
from pathlib import Path

import numpy as np
import pytest
import torch

import gtcrn_oracle as go

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(libadn):
    header = (ROOT / "include" / "adn.h").read_text()
    declared = sorted(set(re.findall(r"\b(adn_[a-z_0-9]+)\s*\(", header)))
    from adn import _lib

    assert sorted(_lib.EXPORTED) == declared, (sorted(_lib.EXPORTED), declared)
    for name in declared:
        assert hasattr(libadn, name), f"libadn.so does not export {name}"
    assert b"sm_100a" in libadn.adn_version()


def test_library_is_sm100a_only(libadn):
    import subprocess

    from adn import build

    out = subprocess.run(["cuobjdump", "--list-elf", str(build.LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly(libadn):
    from adn import _lib, export

    with pytest.raises(_lib.AdnError, match="no CPU fallback"):
        export.gtcrn_model(go.random_state_dict(0), 16000)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from adn import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setenv("ADN_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU path"):
        _lib.lib()


def test_modelfile_round_trip(tmp_path):
    from adn import export, gtcrn_params, modelfile

    sd = go.random_state_dict(0)
    md = export.export_gtcrn(sd, tmp_path / "m.adn", 16000, "INT16", "F32")
    md2, index, payload = modelfile.load(tmp_path / "m.adn")
    assert md2 == md and md2["model_family"] == "gtcrn" and md2["output_audio_length"] == "15872"
    blob = gtcrn_params.pack(sd, 16000)
    assert [t["name"] for t in index] == list(blob)
    for t in index:
        assert t["offset"] % 4 == 0
        a = payload[t["offset"]: t["offset"] + t["count"]]
        assert np.array_equal(a, blob[t["name"]].reshape(-1))
    with pytest.raises(ValueError):
        (tmp_path / "bad.adn").write_bytes(b"ONNX....")
        modelfile.load(tmp_path / "bad.adn")


def test_modelfile_rejects_an_index_that_leaves_the_payload(tmp_path):
    """A tensor record whose offset / count / shape does not fit the payload fails at load with the tensor's name (ADVICE r1:
    the C ABI dereferences blob + offset; its own bound check is the second line of defence)."""
    import json
    import struct

    from adn import modelfile

    good = {"a": np.arange(6, dtype=np.float32).reshape(2, 3), "b": np.ones(5, np.float32)}
    modelfile.save(tmp_path / "ok.adn", {"model_family": "gtcrn"}, good)
    raw = (tmp_path / "ok.adn").read_bytes()
    (hlen,) = struct.unpack("<I", raw[4:8])
    header = json.loads(raw[8:8 + hlen].decode())
    md, index, payload = modelfile.load(tmp_path / "ok.adn")
    assert [t["name"] for t in index] == ["a", "b"] and payload.size >= 11

    def rewrite(mutate):
        h = json.loads(json.dumps(header))
        mutate(h["tensors"])
        hb = json.dumps(h).encode()
        (tmp_path / "bad.adn").write_bytes(raw[:4] + struct.pack("<I", len(hb)) + hb + raw[8 + hlen:])
        return tmp_path / "bad.adn"

    for mutate in (lambda t: t[1].__setitem__("offset", 2**63),            # far outside (the uint64 wrap-around case)
                   lambda t: t[1].__setitem__("count", payload.size + 1),
                   lambda t: t[0].__setitem__("shape", [2, 4]),              # shape disagrees with count
                   lambda t: t[0].__setitem__("offset", -4),
                   lambda t: t[1].pop("count")):
        with pytest.raises(ValueError, match="tensor"):
            modelfile.load(rewrite(mutate))
    with pytest.raises(ValueError):
        (tmp_path / "trunc.adn").write_bytes(raw[:-2])                        # payload not a whole number of floats
        modelfile.load(tmp_path / "trunc.adn")


def test_weight_packing_matches_oracle_fold():
    """BN fold + deconv->conv rewrite: the packed GTConv block reproduces the oracle's
    (reference-order) block on random input, computed with plain torch ops."""
    import torch.nn.functional as F

    from adn import gtcrn_params

    sd = go.random_state_dict(4)
    blob = gtcrn_params.pack(sd, 16000)
    for name, prefix, dil, deconv in (("enc_gt.1", "encoder.en_convs.3", 2, False),
                                      ("dec_gt.0", "decoder.de_convs.0", 5, True)):
        p = torch.from_numpy(blob[name])
        w1, b1 = p[:384].reshape(16, 24), p[384:400]
        wd, bd = p[400:544].reshape(16, 3, 3), p[544:560]
        w2, b2 = p[560:688].reshape(16, 8).T.contiguous(), p[688:696]      # stored [c][o]
        a1, ad = p[696], p[697]
        g = torch.Generator().manual_seed(0)
        x = torch.randn(1, 16, 20, 33, generator=g)
        dbg = {}
        go._gtconv(sd, prefix, x, dil, deconv, dbg)
        ref = dbg[f"{prefix}.h1"]
        s = go._sfe(x[:, :8])
        h = F.prelu(F.conv2d(s, w1.reshape(16, 24, 1, 1), b1), a1.reshape(1))
        h = F.pad(h, [0, 0, 2 * dil, 0])
        h = F.prelu(F.conv2d(h, wd.reshape(16, 1, 3, 3), bd, padding=(0, 1), dilation=(dil, 1), groups=16),
                    ad.reshape(1))
        h = F.conv2d(h, w2.reshape(8, 16, 1, 1), b2)
        assert (h - ref).abs().max() < 1e-5, name
    # ERB nonzero ranges cover every nonzero of the dense matrices
    bm = blob["erb.bm"]
    for j in range(64):
        nz = np.nonzero(bm[:, j])[0]
        assert nz.min() >= blob["erb.bm_lo"][j] and nz.max() < blob["erb.bm_hi"][j]
    bs = blob["erb.bs"]
    for i in range(192):
        nz = np.nonzero(bs[:, i])[0]
        assert nz.min() >= blob["erb.bs_lo"][i] and nz.max() < blob["erb.bs_hi"][i]


def test_metadata_reader_matches_reference_contract():
    from adn import gtcrn_params, metadata

    md = gtcrn_params.metadata(16000)
    r = metadata.MetadataReader(md)
    for k in metadata.REQUIRED_AUDIO_METADATA_KEYS:
        assert r.string(k, required=True)
    cfg = metadata.runtime_config_from_metadata(r)
    assert len(cfg) == 22      # the reference's 22 typed constants (audio_onnx_metadata.py:359-385)
    assert cfg["FOLD_WINDOW_LENGTH"] == 24064 and cfg["BATCH_FOLD_INFERENCE"] is False
    assert cfg["PAD_HEAD"] == 0 and cfg["OUTPUT_SOURCES"] == 1 and cfg["SCALE_FACTOR"] == 1.0
    bad = dict(md)
    del bad["in_sample_rate"]
    with pytest.raises(KeyError, match="Required metadata key in_sample_rate is missing"):
        metadata.runtime_config_from_metadata(metadata.MetadataReader(bad))
    bad = dict(md, normalize_audio_default="maybe")
    with pytest.raises(ValueError, match="must be a boolean"):
        metadata.runtime_config_from_metadata(metadata.MetadataReader(bad))


@pytest.mark.parametrize("n,in_len,out_len,exp", [
    (52800, 16000, 15872, (15872, 4, 63616)),      # GTCRN: stride = out length (:289-290)
    (16000, 16000, 15872, (16000, 1, 16000)),
    (100, 16000, 15872, (16000, 1, 16000)),        # short input zero-padded (:296-298)
    (40000, 16000, 16000, (16000, 3, 48000)),      # length-preserving model
])
def test_chunk_planner(n, in_len, out_len, exp):
    from adn import chunker

    assert chunker.plan_windows(n, in_len, out_len) == exp
    a = np.arange(n, dtype=np.int32)
    w, stride = chunker.split(a, in_len, out_len)
    assert w.shape == (exp[1], 1, in_len) and stride == exp[0]
    assert w[0, 0, 0] == 0 and (n < in_len or w[-1, 0, 0] == (exp[1] - 1) * stride)


def test_ort_shim_surface_without_gpu(tmp_path):
    """Metadata sidecar sessions need no device (audio_onnx_metadata.py:290-303)."""
    import adn.ort_shim as onnxruntime
    from adn import export, metadata

    export.export_gtcrn(go.random_state_dict(0), tmp_path / "GTCRN.adn", 16000)
    s = onnxruntime.InferenceSession(str(tmp_path / "GTCRN_Metadata.onnx"))
    r = metadata.load_runtime_metadata(s)
    assert r.required_int("input_audio_length") == 16000
    with pytest.raises(FileNotFoundError):
        onnxruntime.InferenceSession(str(tmp_path / "missing.adn"))
    opts = onnxruntime.SessionOptions()
    opts.add_session_config_entry("session.set_denormal_as_zero", "1")
    ro = onnxruntime.RunOptions()
    ro.add_run_config_entry("disable_synchronize_execution_providers", "0")
    assert onnxruntime.capi._pybind_state.OrtDevice.cpu() == 0


def test_fold_split_is_reference_fold():
    """Host-side fold == the graph's `audio.reshape(C, n, W).transpose(0, 1)` (Export_MelBandRoformer.py:648)
    plus the zero tail pad of the folded loop (Inference_MelBandRoformer_ONNX.py:298-300)."""
    from adn import chunker

    W, C, n = 441 * 4, 2, 3
    a = np.arange(C * (n * W - 100), dtype=np.float32).reshape(C, -1)
    w, stride = chunker.split(a, W, W)
    assert w.shape == (n, C, W) and stride == W
    padded = np.concatenate((a, np.zeros((C, 100), np.float32)), axis=-1)
    ref = torch.from_numpy(padded).reshape(C, n, W).transpose(0, 1).numpy()
    assert np.array_equal(w, ref)
    # mono -> both channels, extra channels dropped (:273-287)
    assert np.array_equal(chunker.match_channels(np.arange(5), 2), np.stack([np.arange(5)] * 2))
    assert chunker.match_channels(np.zeros((3, 7)), 2).shape == (2, 7)
    # un-folded script: RMS-matched gaussian tail (:301-303)
    x = np.random.default_rng(0).normal(size=(2, 1000)).astype(np.float32)
    p = chunker.tail_pad(x, 400, "noise", np.random.default_rng(1))
    assert p.shape == (2, 1400) and np.array_equal(p[:, :1000], x)
    assert abs(float(p[:, 1000:].std()) - float(np.sqrt(np.mean(x[:, -400:] ** 2)))) < 0.15


def test_mbr_packing_and_band_layout_match_oracle():
    """Product host code (adn/mbr_params.py) vs the oracle's restatement of the reference fusions."""
    import mbr_oracle as mo
    from adn import mbr_params as mp

    cfg, h = mo.MbrConfig(depth=1), mp.MbrHyper(depth=1)
    i1, w1, d1 = mo.band_layout(cfg)
    i2, w2, d2 = mp.band_layout(h)
    assert torch.equal(i1, i2) and w1 == w2 and torch.equal(d1, d2)
    assert len(w1) == 60 and sum(w1) == 7916 and i1.numel() == 3958
    sd = mo.random_state_dict(cfg, 0)
    fw, blob = mo.fuse(sd, cfg), mp.pack(sd, h, 4410)
    for b in (0, 17, 59):
        assert np.array_equal(blob[f"bs_w.{b}"], fw[f"bs_w_{b}"].numpy())
        assert np.array_equal(blob[f"me_w3.{b}"], fw[f"me_w3_{b}"].numpy())
    assert np.array_equal(blob["tf.0.in_w"], fw["time0_in_w"].numpy())
    assert np.array_equal(blob["tf.1.ff2_w"], fw["freq0_ff2_w"].numpy())
    assert np.array_equal(blob["me_w2"], fw["me_w2t"].transpose(1, 2).numpy())
    tc, ts, fc, fs = mo.rotary_tables(cfg, 11)
    assert np.array_equal(blob["rope.tcos"], tc.numpy()) and np.array_equal(blob["rope.fsin"], fs.numpy())
    md = mp.metadata(h, 4410)
    assert md["model_family"] == "mel_band_roformer" and md["input_channels"] == "2"


def test_mf2se_packing_matches_oracle_fold():
    """adn.mf2se_params.pack (product) == mf2se_oracle.fold (restated reference __init__), bit for bit;
    depthwise taps are stored tap-major in the blob."""
    import mf2se_oracle as mo
    from adn import mf2se_params as mp

    cfg = mo.Mf2Config(layers=2)
    sd = mo.random_state_dict(cfg, 1)
    L = 1920 + 384 * 30
    P = mo.fold(sd, cfg, cfg.n_frames(L))
    blob = mp.pack(sd, mp.Mf2Hyper(layers=2), L)
    for k, v in P.items():
        ref = v.numpy()
        if k.split(".")[-1] in ("in_c", "out_c", "uv_c", "mem_c"):
            ref = ref.T
        assert np.array_equal(blob[k].reshape(ref.shape), ref), k
    md = mp.metadata(mp.Mf2Hyper(layers=2), L)
    assert md["model_family"] == "mossformer2_se" and md["center_pad"] == "0" and md["max_signal_length"] == "31"
    with pytest.raises(ValueError):
        mp.pack(sd, mp.Mf2Hyper(layers=2), 48001)            # not nfft + k*hop
    with pytest.raises(ValueError):
        mp.pack(sd, mp.Mf2Hyper(layers=2), 1920 + 384 * 256)  # more than one FLASH group


def test_mf2ss_packing_matches_oracle_fold():
    """adn.mf2ss_params.pack (product) == mf2ss_oracle.fold (restated reference __init__), bit for bit; depthwise
    taps are stored tap-major, the dilated memory kernels (channel, slot, tap) as (slot, tap, channel)."""
    import mf2ss_oracle as so
    from adn import mf2ss_params as sp

    cfg = so.SsConfig(layers=2)
    sd = so.random_state_dict(cfg, 1)
    L = 16 + 8 * 299
    P = so.fold(sd, cfg, cfg.n_frames(L))
    blob = sp.pack(sd, sp.SsHyper(layers=2), L)
    seen = set()
    for k, v in P.items():
        ref, name = v.numpy(), k
        leaf = k.split(".")[-1]
        if leaf in ("in_c", "out_c", "uv_c"):
            ref = ref.T
        elif leaf in ("mem0_w", "mem1_w"):
            ref, name = ref.transpose(1, 2, 0), k[:-1] + "c"
        assert np.array_equal(blob[name].reshape(ref.shape), ref), k
        seen.add(name)
    assert seen == set(blob)
    h = sp.SsHyper(layers=2)
    md = sp.metadata(h, L)
    assert md["model_family"] == "mossformer2_ss" and md["output_sources"] == "2" and md["enc_stride"] == "8"
    assert "nfft" not in md and md["pad_head"] == "8000" and md["output_audio_length"] == str(L)
    assert h.n_frames(16000) == 1999 and h.out_len(16000) == 16000 and h.out_len(16005) == 16000
    with pytest.raises(ValueError):
        sp.pack(sd, h, 8)


def test_mfgan_pack_matches_oracle_fold():
    """adn/mfgan_params.py (product-side folds + kernel layouts) vs the oracle's `fold` (pinned bit-equal to the executed
    reference wrapper's fused buffers): every folded tensor is found in the blob, transposed as csrc/mfgan_ops.cuh reads it."""
    import mfgan_oracle as go
    from adn import mfgan_params as gp

    cfg = go.GanConfig(layers=2)
    sd = go.random_state_dict(cfg, 5)
    L = 2450
    h = gp.GanHyper(layers=2)
    T = h.n_frames(L)
    assert T == cfg.n_frames(L) == 26 and h.padded(L) == 2500
    P = go.fold(sd, cfg, T)
    blob = gp.pack(sd, h, L)
    lin_t = {"fl_w", "fp_w", "uv_w", "rl_w", "rp_w", "in_w", "out_w", "p_w"}
    taps_t = {"fm_w", "uv_c", "rm_w", "in_c", "out_c"}
    conv = {"conv_w", "c2_w", "sp_w", "c1_w", "c_w"}
    seen = set()
    for k, v in P.items():
        leaf, ref, name = k.split(".")[-1], v.numpy(), k
        if leaf == "cross_scale":
            continue                                   # recomputed in csrc/mfgan_ops.cuh bind() as float(Q / BT)
        if leaf in ("fconv_w", "unfold_w"):
            ref, name = ref.reshape(128, 2), k.rsplit(".", 1)[0] + ".gw"
        elif leaf in ("fconv_b", "unfold_b"):
            name = k.rsplit(".", 1)[0] + ".gb"
        elif k.startswith("enc.c1_w"):
            ref = ref[:, :, 0, 0]
        elif leaf in conv:
            ref = ref.transpose(2, 3, 1, 0)
        elif leaf == "w" and ".att." in k:
            ref = ref.T
        elif leaf in lin_t or leaf in taps_t:
            ref = ref.T
        elif leaf == "lin_w":
            ref = ref.transpose(2, 0, 1)
        elif leaf in ("q_g", "k_g", "v_g", "q_b", "k_b", "v_b"):
            continue                                   # checked below as one (112, n_freqs) table
        elif leaf == "p_a":
            ref = np.broadcast_to(ref, (64,))
        elif leaf == "fin_w":
            ref = ref.reshape(-1)
        assert np.array_equal(blob[name].reshape(ref.shape), ref), k
        seen.add(name)
    for i in range(2):
        g = np.concatenate([P[f"B{i}.att.{t}_g"].numpy().reshape(-1, 101) for t in "qkv"], 0)
        b = np.concatenate([P[f"B{i}.att.{t}_b"].numpy().reshape(-1, 101) for t in "qkv"], 0)
        assert np.array_equal(blob[f"B{i}.att.g"], g) and np.array_equal(blob[f"B{i}.att.beta"], b)
        seen |= {f"B{i}.att.g", f"B{i}.att.beta"}
    assert set(blob) - seen == {"stft.fwd", "stft.inv", "stft.norm"}
    md = gp.metadata(h, L)
    assert md["model_family"] == "mossformergan_se" and md["gan_layers"] == "2" and md["max_signal_length"] == "26"
    with pytest.raises(ValueError):
        gp.pack(sd, h, 399)


def test_wavio_roundtrip_and_reference_examples(tmp_path):
    """stdlib-wave PCM16 reader / writer: round trip, stereo layout, mono fold."""
    from adn import wavio

    rng = np.random.default_rng(0)
    st = rng.integers(-30000, 30000, size=(2, 4411), dtype=np.int16)
    wavio.write_wav(tmp_path / "s.wav", st, 44100)
    back, sr = wavio.read_wav(tmp_path / "s.wav")
    assert sr == 44100 and back.dtype == np.int16 and np.array_equal(back, st)
    mono = wavio.to_mono(st)
    assert mono.shape == (4411,) and np.array_equal(mono, (st.astype(np.int32).sum(0) // 2).astype(np.int16))
    wavio.write_wav(tmp_path / "m.wav", mono, 16000)
    back, sr = wavio.read_wav(tmp_path / "m.wav")
    assert back.shape == (1, 4411) and sr == 16000
    with pytest.raises(ValueError):
        wavio.write_wav(tmp_path / "f.wav", mono.astype(np.float32), 16000)


class _FakeSession:
    """Stand-in with the slice of the ORT session surface `chunker` touches: a per-window function in place of the GPU."""

    class _Arg:
        def __init__(self, name, shape):
            self.name, self.type, self.shape = name, "tensor(int16)", shape

    class _Meta:
        def __init__(self, md):
            self.custom_metadata_map = md

    class _Binding:
        def __init__(self):
            self.inputs, self.outputs = {}, {}

        def bind_ortvalue_input(self, n, v):
            self.inputs[n] = v

        def bind_ortvalue_output(self, n, v):
            self.outputs[n] = v

    def __init__(self, in_len, out_len, fns, md):
        self._in = [self._Arg("mix_audio", [1, 1, in_len])]
        self._out = [self._Arg(f"separated_{i}", [1, 1, out_len]) for i in range(len(fns))]
        self._fns, self._md, self.calls = fns, md, []

    def get_inputs(self):
        return self._in

    def get_outputs(self):
        return self._out

    def get_modelmeta(self):
        return self._Meta(self._md)

    def io_binding(self):
        return self._Binding()

    def run_with_iobinding(self, b, run_options=None):
        x = b.inputs["mix_audio"].numpy()
        self.calls.append(x.shape[0])
        for arg, fn in zip(self._out, self._fns):
            np.copyto(b.outputs[arg.name].numpy(), fn(x)[..., :arg.shape[-1]])


@pytest.mark.parametrize("n,win,out_len", [(10000, 4808, 4808), (3000, 4808, 4808), (9616, 4808, 4808), (7000, 2411, 2408)])
def test_separate_is_the_reference_ss_loop(n, win, out_len):
    """`chunker.separate` == the run section of Inference_MossFormer_SS_ONNX.py:269-340 transcribed as a batch-1 loop:
    PAD_HEAD zeros in front, stride = window, zero tail (fold mode), every output concatenated and cut to
    [pad_head : pad_head + len(audio)] -- but issued as ONE batched run."""
    from adn import chunker

    pad_head = 800
    rng = np.random.default_rng(n)
    audio = rng.integers(-20000, 20000, size=n, dtype=np.int16)
    fns = [lambda x: (x // 2).astype(np.int16), lambda x: (-(x // 3) + np.arange(x.shape[-1], dtype=np.int16) % 7).astype(np.int16)]
    sess = _FakeSession(win, out_len, fns, {"pad_head": str(pad_head)})
    got = chunker.separate(sess, audio)
    assert len(sess.calls) == 1                                  # one batched run, not a loop
    # the reference's loop, one window at a time
    a = np.concatenate([np.zeros(pad_head, np.int16), audio])
    audio_len = len(a)
    if audio_len > win:
        num = int(np.ceil((audio_len - win) / win)) + 1
        total = (num - 1) * win + win
    else:
        num, total = 1, win
    a = np.concatenate([a, np.zeros(total - audio_len, np.int16)])
    assert sess.calls[0] == num
    for k, fn in enumerate(fns):
        saved, s, e = [], 0, win
        while e <= total:
            saved.append(fn(a[s:e].reshape(1, 1, -1))[..., :out_len])
            s += win
            e = s + win
        want = np.concatenate(saved, axis=-1).reshape(-1)[pad_head:audio_len]
        assert got[k].dtype == np.int16 and np.array_equal(got[k], want)


@pytest.mark.parametrize("in_sr,out_sr,win,out_len", [(16000, 48000, 1600, 4800), (48000, 16000, 4800, 1600), (16000, 16000, 1600, 1472)])
def test_denoise_and_separate_with_resampled_io(in_sr, out_sr, win, out_len):
    """A model whose output window differs from its input window: with different I/O sample rates the stride stays the input
    window and the result is cut to int(n * out / in) samples (Inference_GTCRN_ONNX.py:288-290, :303, :332); with equal rates
    the windows overlap by in - out samples.  `separate` scales PAD_HEAD and the length by input_to_output_scale
    (Inference_MossFormer_SS_ONNX.py:308-309, :339-340).  Compared with the reference loop transcribed window by window."""
    from adn import chunker

    n = 5 * win + 123
    rng = np.random.default_rng(5)
    audio = rng.integers(-20000, 20000, size=n, dtype=np.int16)

    def fn(x):                                            # nearest-neighbour "resampler" standing in for the model
        idx = np.minimum((np.arange(out_len) * win) // out_len, win - 1)
        return x[..., idx]

    md = {"in_sample_rate": str(in_sr), "out_sample_rate": str(out_sr), "input_to_output_scale": str(float(out_sr / in_sr)), "pad_head": "400"}
    sess = _FakeSession(win, out_len, [fn], md)
    sess._in[0].name = "mix_audio"
    got = chunker.denoise(sess, audio)
    stride = out_len if (win != out_len and in_sr == out_sr) else win
    num = int(np.ceil((n - win) / stride)) + 1
    total = (num - 1) * stride + win
    a = np.concatenate([audio, np.zeros(total - n, np.int16)])
    saved, s = [], 0
    while s + win <= total:
        saved.append(fn(a[s:s + win].reshape(1, 1, -1)))
        s += stride
    want = np.concatenate(saved, axis=-1).reshape(-1)[:int(n * out_sr / in_sr)]
    assert np.array_equal(got, want)
    if in_sr != out_sr:
        sess = _FakeSession(win, out_len, [fn, fn], md)
        sep = chunker.separate(sess, audio)
        scale = out_sr / in_sr
        a = np.concatenate([np.zeros(400, np.int16), audio])
        m = len(a)
        numw = int(np.ceil((m - win) / win)) + 1
        a = np.concatenate([a, np.zeros((numw - 1) * win + win - m, np.int16)])
        full = np.concatenate([fn(a[k * win:(k + 1) * win].reshape(1, 1, -1)) for k in range(numw)], axis=-1).reshape(-1)
        want = full[int(round(400 * scale)):int(round(m * scale))]
        assert np.array_equal(sep[0], want) and np.array_equal(sep[1], want)


def test_hgtcrn_packing_matches_oracle_fold():
    """adn/hgtcrn_params.py: the ConvBlock-wrapper key names map onto GTCRN's packer, en_convs.0 has 18 input channels, the
    decoder's GTConv blocks stay plain convolutions (no tap flip), and the folds equal the oracle's."""
    import hgtcrn_oracle as ho
    from adn import hgtcrn_params as hp

    sd = ho.random_state_dict(3)
    blob = hp.pack(sd, 256 * 20)
    w0, b0 = ho._fold(sd, "encoder.en_convs.0")
    assert np.array_equal(blob["enc_front_h"][:1440], w0[:, :, 0, :].permute(2, 1, 0).reshape(-1).numpy())
    assert np.array_equal(blob["enc_front_h"][1440:1456], b0.numpy())
    wd, bd = ho._fold(sd, "decoder.de_convs.0.depth_conv")
    assert np.array_equal(blob["dec_gt.0"][400:544], wd[:, 0].reshape(-1).numpy())          # un-flipped taps
    w2, b2 = ho._fold(sd, "decoder.de_convs.0.point_conv2")
    assert np.array_equal(blob["dec_gt.0"][560:688], w2[:, :, 0, 0].T.reshape(-1).numpy())   # stored [c][o]
    assert blob["istft.norm"].shape == (256 * 20,) and blob["stft.fwd"].shape == (514, 512)
    md = hp.metadata(256 * 20, "INT16", "F32")
    assert md["model_family"] == "h_gtcrn" and md["input_channels"] == "2" and md["output_channels"] == "1"
    assert md["max_signal_length"] == "21" and md["feature_kind"] == "stft_wpe_auxiva" and md["cg_solve_iter"] == "6"
    with pytest.raises(ValueError):
        hp.metadata(1000)


@pytest.mark.parametrize("n,pad", [(0, 5), (1, 4), (7, 5), (40, 13)])
def test_reflect_tail_is_the_reference_context_pad(n, pad):
    """chunker.tail_pad(mode='reflect') == `pad_audio_tail_with_context` of the un-folded H-GTCRN script
    (H-GTCRN/Inference_H_GTCRN_ONNX.py:138-153): mirror about the last sample, one sample repeats, nothing -> zeros."""
    from adn import chunker

    a = (np.arange(2 * n, dtype=np.int16).reshape(2, n) * 3 - 7)
    got = chunker.tail_pad(a, pad, "reflect")
    assert got.shape == (2, n + pad) and got.dtype == a.dtype and np.array_equal(got[:, :n], a)
    if n == 0:
        assert not got.any()
    elif n == 1:
        assert np.array_equal(got[:, n:], np.repeat(a[:, -1:], pad, axis=-1))
    else:
        for k in range(pad):
            assert np.array_equal(got[:, n + k], a[:, n - 2 - k])

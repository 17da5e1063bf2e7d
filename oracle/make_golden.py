"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by EXECUTING THE REFERENCE
(/root/reference, build container only) on seeded inputs / weights.

    python oracle/make_golden.py

The fixtures travel to the GPU box; /root/reference does not.  Every fixture stores the
inputs, the reference outputs and the seeds, so the `-m gpu` tests can compare the CUDA
path with the reference itself, not only with the restated oracle.
"""
from __future__ import annotations

import hashlib
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))

import gtcrn_oracle as go           # noqa: E402
import ref_loader                   # noqa: E402
from stft_oracle import SPECS       # noqa: E402

GOLDEN = HERE.parent / "tests" / "golden"

# (oracle spec name, reference folder, stft model_type, istft model_type, istft static kwarg, L)
STFT_CASES = [
    ("gtcrn", "GTCRN", "stft_B", "istft_B", {"static_norm": True}, 4096),
    ("zipenhancer", "ZipEnhancer", "stft_B", "istft_B", {"static_norm": True}, 3200),
    ("mossformer2_se_48k", "MossFormer2_SE_48K", "stft_B", "istft_B", {"static_frames": True}, 9600),
    ("mel_band_roformer", "Mel_Band_Roformer/Stereo", "stft_B", "istft_B", {"static_frames": True}, 8820),
    ("mossformergan_se_16k", "MossFormerGAN_SE_16K", "stft_C", "istft_C", {"precompute_window_sum": True}, 3200),
]


def synth_audio(n: int, seed: int = 1234, batch: int = 1) -> torch.Tensor:
    """Band-limited speech-like noise + sinusoids, peak 0.5 FS."""
    g = torch.Generator().manual_seed(seed)
    x = 0.1 * torch.randn(batch, 1, n, generator=g)
    k = torch.hann_window(33, periodic=False)
    k = (k / k.sum()).reshape(1, 1, -1)
    x = torch.nn.functional.conv1d(x, k, padding=16)
    t = torch.arange(n, dtype=torch.float32) / 16000.0
    for i, f0 in enumerate((220.0, 587.0, 1310.0)):
        amp = 0.3 / (i + 1)
        ph = torch.rand(batch, 1, 1, generator=g) * 6.2831853
        x = x + amp * torch.sin(2 * torch.pi * f0 * t.reshape(1, 1, -1) + ph)
    x = x / x.abs().amax(dim=-1, keepdim=True) * 0.5
    return x.float().contiguous()


def sd_digest(sd: dict) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().numpy().tobytes())
    return h.hexdigest()


def ref_stft_pair(spec_name: str, folder: str, st_type: str, ist_type: str, static_kw: dict, length: int):
    spec = SPECS[spec_name]
    mod = ref_loader.load_stft_module(folder)
    wt = {"hamming_sym": "hamming"}.get(spec.window_type, spec.window_type)
    t = spec.n_frames(length)
    stft = mod.STFT_Process(model_type=st_type, n_fft=spec.nfft, hop_len=spec.hop, win_length=spec.win_length,
                            max_frames=0, window_type=wt, center_pad=spec.center, pad_mode=spec.pad_mode).eval()
    istft = mod.STFT_Process(model_type=ist_type, n_fft=spec.nfft, hop_len=spec.hop, win_length=spec.win_length,
                             max_frames=t, window_type=wt, center_pad=spec.center, pad_mode=spec.pad_mode,
                             **static_kw).eval()
    return stft, istft


def ref_stft_forward(stft, x):
    out = stft(x)
    if isinstance(out, (tuple, list)):
        out = torch.cat(out, dim=1)
    return out


def ref_istft_forward(istft, s, fbins):
    if istft.model_type == "istft_C":
        return istft(s)
    return istft(s[:, :fbins], s[:, fbins:])


def main():
    assert ref_loader.reference_available(), "/root/reference is required to generate fixtures"
    GOLDEN.mkdir(parents=True, exist_ok=True)
    torch.manual_seed(1234)
    with torch.inference_mode():
        # ---- STFT / ISTFT, all in-scope variants
        for name, folder, st, ist, kw, L in STFT_CASES:
            spec = SPECS[name]
            stft, istft = ref_stft_pair(name, folder, st, ist, kw, L)
            g = torch.Generator().manual_seed(1234)
            x = torch.rand(2, 1, L, generator=g) * 2 - 1          # uniform[-1,1] like SURVEY App. B
            s = ref_stft_forward(stft, x)
            # a spectrum that is NOT a valid STFT exercises the ISTFT alone
            s2 = s * (0.5 + torch.rand(s.shape, generator=g))
            y = ref_istft_forward(istft, s2, spec.fbins)
            np.savez_compressed(GOLDEN / f"stft_{name}.npz", x=x.numpy(), spec=s.numpy(), spec_in=s2.numpy(),
                                y=y.numpy(), length=L)
            print(f"stft_{name}: x{tuple(x.shape)} -> spec{tuple(s.shape)} -> y{tuple(y.shape)}")

        # ---- GTCRN end to end, F32 I/O and INT16 I/O (SURVEY.md 8a a1-a5,a10-a12)
        sd = go.random_state_dict(0)
        digest = sd_digest(sd)
        x = synth_audio(16000, 1234, batch=3)
        x[2] = 0.0                                               # all-zero chunk edge case
        for dt in ("F32", "INT16"):
            _, build = ref_loader.load_gtcrn(16000, dt)
            w = build(sd)
            xin = x if dt == "F32" else torch.round(x * 32767.0).to(torch.int16)
            y = torch.cat([w(xin[i:i + 1]) for i in range(xin.shape[0])], dim=0)
            np.savez_compressed(GOLDEN / f"gtcrn_{dt.lower()}_L16000.npz", x=xin.numpy(), y=y.numpy(),
                                seed=0, sd_digest=digest)
            print(f"gtcrn {dt}: {tuple(xin.shape)} -> {tuple(y.shape)}")
        # second chunk length (different T, exercises partial tiles): 2 s like the reference default
        _, build = ref_loader.load_gtcrn(8000, "F32")
        w = build(sd)
        x8 = synth_audio(8000, 99, batch=2)
        y8 = torch.cat([w(x8[i:i + 1]) for i in range(2)], dim=0)
        np.savez_compressed(GOLDEN / "gtcrn_f32_L8000.npz", x=x8.numpy(), y=y8.numpy(), seed=0, sd_digest=digest)
        print(f"gtcrn F32 L8000: {tuple(x8.shape)} -> {tuple(y8.shape)}")


def mbr_kwargs(cfg):
    return dict(dim=cfg.dim, depth=cfg.depth, stereo=True, num_stems=1, time_transformer_depth=1,
                freq_transformer_depth=1, num_bands=cfg.num_bands, dim_head=cfg.dim_head, heads=cfg.heads,
                mask_estimator_depth=2, stft_n_fft=cfg.nfft, stft_hop_length=cfg.hop, stft_win_length=cfg.nfft,
                sample_rate=cfg.sample_rate)


def main_mbr():
    """Mel-Band-Roformer (stereo) fixtures: the reference's own module on seeded weights, one
    un-folded window per run (depth 2 keeps the fixture generation and the CPU oracle quick; the
    mask estimator -- 92 % of the parameters -- is depth independent)."""
    import mbr_oracle as mo

    assert ref_loader.reference_available()
    cfg = mo.MbrConfig(depth=2)
    sd = mo.random_state_dict(cfg, 0)
    with torch.inference_mode():
        for L, dt in ((4410, "F32"), (13230, "INT16")):
            _, build = ref_loader.load_mbr(L, dt)
            m = build(sd, **mbr_kwargs(cfg))
            g = torch.Generator().manual_seed(1234)
            x = (torch.rand(2, 2, L, generator=g) * 2 - 1) * 0.5
            x[1, :, L // 2:] = 0.0
            xin = x if dt == "F32" else torch.round(x * 32767.0).to(torch.int16)
            y = torch.cat([m(xin[i:i + 1]) for i in range(2)], dim=0)
            np.savez_compressed(GOLDEN / f"mbr_{dt.lower()}_L{L}_d2.npz", x=xin.numpy(), y=y.numpy(), seed=0, depth=2)
            print(f"mbr {dt} L{L}: {tuple(xin.shape)} -> {tuple(y.shape)} max|y| {y.abs().max().item():.4f}")


def main_mf2se():
    """MossFormer2-SE-48K fixtures: the reference wrapper executed around `mf2se_oracle.skeleton()`
    on seeded weights, 2 FLASH+FSMN layers (the layer count is the only reduced hyper-parameter),
    windows of 31 and 26 frames, one all-zero window."""
    import mf2se_oracle as mo

    assert ref_loader.reference_available()
    cfg = mo.Mf2Config(layers=2)
    sd = mo.random_state_dict(cfg, 0)
    hold = mo.skeleton(cfg)
    hold.load_state_dict(sd)
    with torch.inference_mode():
        for L, dt in ((13440, "F32"), (11520, "INT16")):
            _, build = ref_loader.load_mf2se(L, dt)
            w = build(hold)
            x = synth_audio(L, 4321, batch=3)
            x[2] = 0.0
            xin = x if dt == "F32" else torch.round(x * 32767.0).to(torch.int16)
            y = torch.cat([w(xin[i:i + 1].clone()) for i in range(3)], dim=0)
            np.savez_compressed(GOLDEN / f"mf2se_{dt.lower()}_L{L}_l2.npz", x=xin.numpy(), y=y.numpy(), seed=0, layers=2)
            print(f"mf2se {dt} L{L}: {tuple(xin.shape)} -> {tuple(y.shape)} max|y| {y.float().abs().max().item():.4f}")


def main_fulldepth():
    """FULL-DEPTH reference-executed fixtures (the other fixtures reduce the layer count): MossFormer2-SE-48K with 24 layers
    (BASELINE configs[2]), MossFormer2-SS-16K with 24 layers, Mel-Band-Roformer at depth 6 -- the depths bench.py runs.  One or
    two short windows each: the files carry the input, the reference's output and the seed of the weights only."""
    import mbr_oracle as bo
    import mf2se_oracle as mo
    import mf2ss_oracle as so

    assert ref_loader.reference_available()
    with torch.inference_mode():
        cfg = mo.Mf2Config(layers=24)
        hold = mo.skeleton(cfg)
        hold.load_state_dict(mo.random_state_dict(cfg, 0))
        L = 13440
        _, build = ref_loader.load_mf2se(L, "F32")
        w = build(hold)
        x = synth_audio(L, 4321, batch=2)
        y = torch.cat([w(x[i:i + 1].clone()) for i in range(2)], dim=0)
        np.savez_compressed(GOLDEN / f"mf2se_f32_L{L}_l24.npz", x=x.numpy(), y=y.numpy(), seed=0, layers=24)
        print(f"mf2se full depth: {tuple(x.shape)} -> {tuple(y.shape)} max|y| {y.abs().max().item():.4f}")

        cfg = so.SsConfig(layers=24)
        hold = so.skeleton(cfg)
        hold.load_state_dict(so.random_state_dict(cfg, 0))
        L = 4808
        _, build = ref_loader.load_mf2ss(L, "F32")
        w = build(hold)
        x = synth_audio(L, 4321, batch=2) * 32767.0
        ys = [torch.cat([w(x[i:i + 1].clone())[s] for i in range(2)], dim=0) for s in range(2)]
        np.savez_compressed(GOLDEN / f"mf2ss_f32_L{L}_l24.npz", x=x.numpy(), y0=ys[0].numpy(), y1=ys[1].numpy(), seed=0, layers=24)
        print(f"mf2ss full depth: {tuple(x.shape)} -> 2 x {tuple(ys[0].shape)} max|y| {ys[0].abs().max().item():.4f}")

        cfg = bo.MbrConfig(depth=6)
        sd = bo.random_state_dict(cfg, 0)
        L = 4410
        _, build = ref_loader.load_mbr(L, "F32")
        m = build(sd, **mbr_kwargs(cfg))
        g = torch.Generator().manual_seed(1234)
        x = (torch.rand(1, 2, L, generator=g) * 2 - 1) * 0.5
        y = m(x.clone())
        np.savez_compressed(GOLDEN / f"mbr_f32_L{L}_d6.npz", x=x.numpy(), y=y.numpy(), seed=0, depth=6)
        print(f"mbr full depth: {tuple(x.shape)} -> {tuple(y.shape)} max|y| {y.abs().max().item():.4f}")


def main_mf2ss():
    """MossFormer2-SS-16K fixtures: the reference wrapper (`MOSSFORMER_SS`) executed around
    `mf2ss_oracle.skeleton()` on seeded weights, 2 FLASH + dilated-FSMN layers (the layer count is the
    only reduced hyper-parameter); 600 frames = 3 FLASH groups with 168 padded keys, and 300 frames;
    int16-scale samples; one all-zero window (the `rms_out > 0` guard, :631)."""
    import mf2ss_oracle as so

    assert ref_loader.reference_available()
    cfg = so.SsConfig(layers=2)
    sd = so.random_state_dict(cfg, 0)
    hold = so.skeleton(cfg)
    hold.load_state_dict(sd)
    with torch.inference_mode():
        for L, dt in ((4808, "F32"), (2408, "INT16")):
            _, build = ref_loader.load_mf2ss(L, dt)
            w = build(hold)
            x = synth_audio(L, 4321, batch=3) * 32767.0
            x[2] = 0.0
            xin = x if dt == "F32" else torch.round(x).to(torch.int16)
            ys = [torch.cat([w(xin[i:i + 1].clone())[s] for i in range(3)], dim=0) for s in range(2)]
            np.savez_compressed(GOLDEN / f"mf2ss_{dt.lower()}_L{L}_l2.npz", x=xin.numpy(), y0=ys[0].numpy(), y1=ys[1].numpy(),
                                seed=0, layers=2)
            print(f"mf2ss {dt} L{L}: {tuple(xin.shape)} -> 2 x {tuple(ys[0].shape)} max|y| {ys[0].float().abs().max().item():.4f}")


def main_mfgan():
    """MossFormerGAN-SE-16K fixtures: the reference wrapper (`MOSSFORMER_SE` of MossFormerGAN_SE_16K) executed around
    `mfgan_oracle.skeleton()` on seeded weights, 2 SyncANet blocks (the block count is the only reduced
    hyper-parameter); 3150 samples (wrap-around pad to 3200, 33 frames) in F32 and 2400 samples (25 frames) in INT16;
    one all-zero window."""
    import mfgan_oracle as go

    assert ref_loader.reference_available()
    cfg = go.GanConfig(layers=2)
    sd = go.random_state_dict(cfg, 0)
    hold = go.skeleton(cfg)
    hold.load_state_dict(sd)
    with torch.inference_mode():
        for L, dt in ((3150, "F32"), (2400, "INT16")):
            _, build = ref_loader.load_mfgan(L, dt)
            w = build(hold)
            x = synth_audio(L, 4321, batch=3)
            x[2] = 0.0
            xin = x if dt == "F32" else torch.round(x * 32767.0).to(torch.int16)
            y = torch.cat([w(xin[i:i + 1].clone()) for i in range(3)], dim=0)
            np.savez_compressed(GOLDEN / f"mfgan_{dt.lower()}_L{L}_l2.npz", x=xin.numpy(), y=y.numpy(), seed=0, layers=2)
            print(f"mfgan {dt} L{L}: {tuple(xin.shape)} -> {tuple(y.shape)} max|y| {y.float().abs().max().item():.4f}")


def main_zipenh():
    """ZipEnhancer fixtures (BASELINE configs[1]): the reference wrapper (`ZipEnhancer` of
    ZipEnhancer/Export_ZipEnhancer.py, with the reference's own forward overrides installed on the skeleton classes) executed
    around `zipenh_oracle.skeleton()` on seeded weights -- the full model (four dual-path encoders), 3200 samples (33 frames) in
    F32 and 2400 samples (25 frames) in INT16; one all-zero window.  The weights are regenerated from the seed.  The inputs carry a
    broadband noise floor (noisy speech does): in a bin without energy the phase feature atan2(im, re + 1e-5) (:844) is the angle of
    rounding noise and the reference's own output is not reproducible from one BLAS to the next."""
    import zipenh_oracle as zo

    assert ref_loader.reference_available()
    cfg = zo.ZipConfig()
    with torch.inference_mode():
        for L, dt in ((3200, "F32"), (2400, "INT16")):
            _, build = ref_loader.load_zipenh(L, dt)
            w = build(zo.skeleton(cfg, 0))
            x = synth_audio(L, 1357, batch=4)
            x = (x + 0.03 * torch.randn(x.shape, generator=torch.Generator().manual_seed(97))).clamp(-1.0, 1.0)
            x[2] = 0.0
            xin = x if dt == "F32" else torch.round(x * 32767.0).to(torch.int16)
            y = torch.cat([w(xin[i:i + 1].clone()) for i in range(4)], dim=0)
            # the phase feature the reference computed (:839-844, same modules, same machine): its +-pi branch-cut decisions in
            # the edge frames are rounding noise, so an implementation can only be compared where it took the same ones
            phas = []
            for i in range(4):                                                     # one window per call, as the forward saw it
                a = xin[i:i + 1].float() * (1.0 if dt == "INT16" else 32768.0)
                a = a / torch.sqrt(torch.mean(a * a, dim=-1, keepdim=True) + 1e-6)
                re, im = w.stft_model(a)
                phas.append(torch.atan2(im, re + 1e-5).transpose(1, 2))
            pha = torch.cat(phas, dim=0).contiguous()                              # (B, T, F)
            np.savez_compressed(GOLDEN / f"zipenh_{dt.lower()}_L{L}.npz", x=xin.numpy(), y=y.numpy(), pha=pha.numpy(), seed=0)
            print(f"zipenh {dt} L{L}: {tuple(xin.shape)} -> {tuple(y.shape)} max|y| {y.float().abs().max().item():.4f}")


def main_dfsmn():
    """DFSMN (48 kHz) fixtures: the reference wrapper (`DFSMN` of DFSMN/Export_DFSMN.py) executed around
    `dfsmn_oracle.skeleton()` on seeded weights, 3 FSMN layers (the depth is the only reduced hyper-parameter);
    1920 + 960 * 8 samples (9 frames) in F32 and 1920 + 960 * 5 (6 frames) in INT16; one all-zero window."""
    import dfsmn_oracle as do

    assert ref_loader.reference_available()
    cfg = do.DfsmnConfig(layers=3)
    sd = do.random_state_dict(cfg, 0)
    hold = do.skeleton(cfg)
    hold.load_state_dict(sd)
    with torch.inference_mode():
        for L, dt in ((1920 + 960 * 8, "F32"), (1920 + 960 * 5, "INT16")):
            _, build = ref_loader.load_dfsmn(L, dt)
            w = build(hold)
            x = synth_audio(L, 2468, batch=3)
            x[2] = 0.0
            xin = x if dt == "F32" else torch.round(x * 32767.0).to(torch.int16)
            y = torch.cat([w(xin[i:i + 1].clone()) for i in range(3)], dim=0)
            np.savez_compressed(GOLDEN / f"dfsmn_{dt.lower()}_L{L}_l3.npz", x=xin.numpy(), y=y.numpy(), seed=0, layers=3)
            print(f"dfsmn {dt} L{L}: {tuple(xin.shape)} -> {tuple(y.shape)} max|y| {y.float().abs().max().item():.4f}")


def main_ulunas():
    """UL-UNAS fixtures: the reference `ULUNAS_CUSTOM` executed on
    seeded default-init weights with randomised BatchNorm statistics.  The RAW (pre-fold) state_dict travels in the fixture
    (`sd/<key>`), because the model class itself cannot run on the GPU box; 16000 samples -> 15872 (63 frames), F32 and INT16,
    one all-zero window."""
    assert ref_loader.reference_available()
    for dt in ("F32", "INT16"):
        _, build = ref_loader.load_ulunas(16000, dt)
        w, raw = build(None, 0)
        x = synth_audio(16000, 1357, batch=3)
        x[2] = 0.0
        xin = x if dt == "F32" else torch.round(x * 32767.0).to(torch.int16)
        with torch.inference_mode():
            y = torch.cat([w(xin[i:i + 1].clone()) for i in range(3)], dim=0)
        np.savez_compressed(GOLDEN / f"ulunas_{dt.lower()}_L16000.npz", x=xin.numpy(), y=y.numpy(), seed=0,
                            **{f"sd/{k}": v.numpy() for k, v in raw.items()})
        print(f"ulunas {dt}: {tuple(xin.shape)} -> {tuple(y.shape)} max|y| {y.float().abs().max().item():.4f}")


def main_hgtcrn():
    """H-GTCRN fixtures: the reference `H_GTCRN_CUSTOM` (2-channel
    STFT -> WPE -> AuxIVA -> GTCRN_IVA -> ISTFT) executed on seeded default-init weights with randomised BatchNorm statistics;
    the RAW state_dict of `GTCRN_IVA` travels in the fixture (`sd/<key>`).  Stereo windows of 16128 samples (64 frames), F32 and
    INT16; correlated channels (a delayed, attenuated copy plus independent noise), as a two-microphone pickup would give.
    The WPE and AuxIVA stage outputs of the executed reference (forward hooks; (2, 2, 257, 64) real / imaginary planes) travel
    too: the six conjugate-gradient steps of the WPE solve amplify one-ulp differences by up to 1e5 in single bins, so the
    oracle's later stages are pinned ON the reference's WPE output (tests/test_oracle_pinning.py)."""
    assert ref_loader.reference_available()
    L = 16128
    for dt in ("F32", "INT16"):
        _, build = ref_loader.load_hgtcrn(L, dt)
        w, raw = build(None, 0)
        a = synth_audio(L, 97, batch=2)[:, 0]
        n = synth_audio(L, 98, batch=2)[:, 0]
        x = torch.stack((a, 0.7 * torch.roll(a, 3, dims=-1) + 0.3 * n), dim=1)           # (2 windows, 2 channels, L)
        xin = x if dt == "F32" else torch.round(x * 32767.0).to(torch.int16)
        cap = {"wpe": [], "iva": []}
        hooks = [w.wpe.register_forward_hook(lambda m, i, o: cap["wpe"].append(torch.stack(o, dim=0)[:, 0].clone())),
                 w.iva.register_forward_hook(lambda m, i, o: cap["iva"].append(torch.stack(o, dim=0)[:, 0].clone()))]
        with torch.inference_mode():
            y = torch.cat([w(xin[i:i + 1].clone()) for i in range(2)], dim=0)
        for h in hooks:
            h.remove()
        stage = {k: torch.stack(v, dim=0).numpy() for k, v in cap.items()}             # (window, re|im, mic, F, T)
        np.savez_compressed(GOLDEN / f"hgtcrn_{dt.lower()}_L{L}.npz", x=xin.numpy(), y=y.numpy(), seed=0,
                            ref_wpe=stage["wpe"], ref_iva=stage["iva"],
                            **{f"sd/{k}": v.numpy() for k, v in raw.items()})
        print(f"hgtcrn {dt}: {tuple(xin.shape)} -> {tuple(y.shape)} max|y| {y.float().abs().max().item():.4f}")


if __name__ == "__main__":
    if "--hgtcrn" in sys.argv:
        main_hgtcrn()
    elif "--ulunas" in sys.argv:
        main_ulunas()
    elif "--dfsmn" in sys.argv:
        main_dfsmn()
    elif "--zipenh" in sys.argv:
        main_zipenh()
    elif "--mfgan" in sys.argv:
        main_mfgan()
    elif "--mf2ss" in sys.argv:
        main_mf2ss()
    elif "--mbr" in sys.argv:
        main_mbr()
    elif "--mf2se" in sys.argv:
        main_mf2se()
    elif "--fulldepth" in sys.argv:
        main_fulldepth()
    else:
        main()

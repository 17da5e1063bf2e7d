#!/usr/bin/env python
"""Benchmark of the north-star path: noisy chunk in -> clean chunk out (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl adn|reference]

A "step" is one pass of the hot path over one batch of B synthetic 1 s chunks per GPU.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field's definition.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for _p in (ROOT / "audio-denoiser-onnx_b200", ROOT / "oracle"):
    if str(_p) not in sys.path:
        sys.path.insert(0, str(_p))

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "audio_seconds_per_second"
UNIT = "audio-s/s"
CHUNK = 16000          # 1 s @ 16 kHz
SR = 16000
T_FRAMES = 63
L_OUT = 15872


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


# Algorithmic work per chunk of each kernel (DESIGN.md "Kernels"): (bytes moved to/from HBM
# if every operand is touched exactly once, fp32 flops).  T=63 frames, F=257.
def kernel_work():
    T = T_FRAMES
    f16 = 16 * 33 * 4 * T           # one (T,16,33) fp32 activation
    spec = 514 * 4 * T
    return {
        "prep": (CHUNK * 4 + 16512 * 4, 2 * CHUNK),
        "stft_gemm": (16512 * 4 + spec, 2 * 514 * 512 * T),
        "stft_gemm_tc": (2 * 16512 * 4 + spec, 2 * 514 * 512 * T),
        "istft_gemm_tc": (2 * spec + L_OUT * 4, 2 * 514 * 512 * T),
        "enc_front": (spec + 16 * 65 * 4 * T + f16, 2 * T * (65 * 720 + 33 * 16 * 40 + 3 * 2 * 192)),
        "gt_main": (f16 // 2 + f16 // 2 + 8 * 4 * T, 2 * T * 33 * (384 + 144 + 128)),
        "tra_gru": (2 * 8 * 4 * T, 2 * T * (3 * 16 * 24 + 128)),
        "tra_apply": (f16 // 2 + f16 // 2 + f16, T * 528),
        "dp_intra": (2 * f16 + 3 * f16, 2 * T * (33 * 16 * 3 * 12 + 2 * 33 * 256 + 33 * 48 * 8) + 8 * T * 528),
        "dp_inter": (3 * f16 + f16, 2 * T * 33 * 16 * 3 * 8),
        "ln_res": (4 * f16, 2 * T * 33 * 256 + 8 * T * 528),
        "dec_tail": (f16 + 16 * 65 * 4 * T + 2 * spec, 2 * T * (65 * 16 * 20 + 129 * 2 * 40 + 2 * 2 * 192 + 4 * 257)),
        "istft_gemm": (spec + L_OUT * 4, 2 * 514 * 512 * T),
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": sorted(reasons),
                "samples": len(self.rows)}


def make_inputs(batch: int, n_sets: int, seed: int = 1234):
    """Synthetic speech-like chunks.  n_sets distinct batches so consecutive
    steps never re-read the same input lines from L2."""
    from make_golden import synth_audio

    base = synth_audio(CHUNK, seed, min(batch, 64))
    sets = []
    for s in range(n_sets):
        g = torch.Generator().manual_seed(seed + 1 + s)
        idx = torch.randint(0, base.shape[0], (batch,), generator=g)
        gain = 0.25 + 0.75 * torch.rand(batch, 1, 1, generator=g)
        x = base[idx] * gain
        x = torch.roll(x, shifts=int(torch.randint(0, CHUNK, (1,), generator=g)), dims=-1)
        sets.append(x.contiguous())
    return sets


def cpu_oracle_rate(sd, n_chunks: int, threads: int):
    """Reference-arm / cpu_baseline: the oracle port of the reference's PyTorch graph, one
    (1,1,L) call per chunk like process_segment (Inference_GTCRN_ONNX.py:314-317)."""
    import gtcrn_oracle as go
    from make_golden import synth_audio

    torch.set_num_threads(threads)
    x = synth_audio(CHUNK, 4321, 4)
    with torch.inference_mode():
        for i in range(3):
            go.gtcrn_forward(sd, x[i % 4:i % 4 + 1])
        t0 = time.perf_counter()
        for i in range(n_chunks):
            go.gtcrn_forward(sd, x[i % 4:i % 4 + 1])
        dt = time.perf_counter() - t0
    return n_chunks * (CHUNK / SR) / dt, dt


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path on the host cores.
    onnxruntime / onnx are not installable here and /root/reference does not travel, so this
    arm times the oracle port (PyTorch eager, the modules the ONNX graph is traced from)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import gtcrn_oracle as go

    sd = go.random_state_dict(0)
    threads = os.cpu_count() or 1
    per_step = args.ref_chunks
    torch.set_num_threads(threads)
    for _ in range(max(args.warmup, 3)):
        cpu_oracle_rate(sd, 2, threads)
    t0 = time.perf_counter()
    total = 0
    for _ in range(args.steps):
        _, _ = cpu_oracle_rate(sd, per_step, threads)
        total += per_step
    dt = time.perf_counter() - t0
    val = total * (CHUNK / SR) / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"GTCRN 16 kHz, {args.batch} x 1 s chunks per GPU, F32 I/O (CPU arm: bounded sample of "
                               f"{per_step} chunks per step, chunk-at-a-time)", "model": "gtcrn"},
        "rtf": 1.0 / val,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{per_step} chunks/step x {args.steps} steps, oracle/gtcrn_oracle.py (PyTorch eager "
                                   f"restatement of Export_GTCRN.py; ORT itself is not installable offline)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=512, help="chunks per GPU per step")
    ap.add_argument("--impl", default="adn", choices=["adn", "reference"])
    ap.add_argument("--ref-chunks", type=int, default=24, help="CPU chunks per step for --impl reference")
    ap.add_argument("--cpu-baseline-chunks", type=int, default=600)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference(args)
        return

    import gtcrn_oracle as go   # seeded synthetic weights only (no compute from oracle/ in the timed path)
    from adn import _lib, build, export
    import adn.ort_shim as onnxruntime

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libadn has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    build.build()

    B = args.batch
    sd = go.random_state_dict(0)
    model = export.gtcrn_model(sd, CHUNK, "F32", "F32", device_id=local_rank)
    n_sets = 8                                           # 8 x 32 MiB inputs > 126 MB L2
    host_sets = make_inputs(B, n_sets, seed=1234 + rank)
    dev_sets = [x.to(dev) for x in host_sets]
    out = torch.empty((B, 1, L_OUT), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- device-resident throughput ("value")
    for i in range(args.warmup):
        model.run(dev_sets[i % n_sets], out=out)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clk = ClockSampler(local_rank)
    clk.__enter__()              # sampled across all GPU loops of this run (timed + per-kernel + e2e)
    e0.record(stream)
    for i in range(args.steps):
        model.run(dev_sets[i % n_sets], out=out)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    audio_s = world * B * args.steps * (CHUNK / SR)
    value = audio_s / (ms * 1e-3)

    # ---------------- per-kernel device times (CUDA events on the launching stream)
    model.set_profiling(True)
    acc: dict[str, list[float]] = {}
    for i in range(args.steps):
        model.run(dev_sets[i % n_sets], out=out)
        torch.cuda.synchronize(dev)
        for name, t_ms in model.kernel_times():
            acc.setdefault(name, []).append(t_ms)
    model.set_profiling(False)
    per_kernel = {k: (sum(v) / args.steps, len(v) // args.steps) for k, v in acc.items()}   # (ms per step, launches)
    step_ms_prof = sum(v[0] for v in per_kernel.values())
    pk = peaks()
    work = kernel_work()
    top = max(per_kernel, key=lambda k: per_kernel[k][0])
    top_ms, top_launches = per_kernel[top]
    launch_ms = top_ms / top_launches
    wb, wf = work[top]
    t_hbm = wb * B / (pk["hbm_gbs"] * 1e9)
    t_tc = wf * B / (pk["bf16_tflops_sustained"] * 1e12)
    if t_hbm >= t_tc:
        roof = {"bound": "hbm", "achieved": wb * B / (launch_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s"}
    else:
        roof = {"bound": "tensor", "achieved": wf * B / (launch_ms * 1e-3) / 1e12,
                "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof.update({"traffic": None, "kernel": top, "kernel_ms_per_launch": launch_ms,
                 "kernel_share_of_step": top_ms / step_ms_prof, "peak_source": pk["src"],
                 "algorithmic_bytes_per_launch": wb * B, "algorithmic_flops_per_launch": wf * B})

    # ---------------- end to end through the reference-facing API, host buffers
    # (OrtValue over pinned host memory -> run_with_iobinding -> adn_run_host: H2D, kernels, D2H)
    tmpdir = Path(os.environ.get("TMPDIR", "/tmp")) / f"adn_bench_{os.getpid()}"
    tmpdir.mkdir(parents=True, exist_ok=True)
    mpath = tmpdir / "GTCRN.adn"
    export.export_gtcrn(sd, mpath, CHUNK, "F32", "F32")
    sess = onnxruntime.InferenceSession(str(mpath), providers=["CPUExecutionProvider"], device_id=local_rank)
    pin_in = [x.pin_memory() for x in host_sets]
    pin_out = torch.empty((B, 1, L_OUT), dtype=torch.float32).pin_memory()
    vout = onnxruntime.OrtValue.ortvalue_from_numpy(pin_out.numpy())
    vout._a = pin_out.numpy()                                   # keep the pinned storage (no copy)
    vins = []
    for p in pin_in:
        v = onnxruntime.OrtValue.ortvalue_from_numpy(p.numpy())
        v._a = p.numpy()
        vins.append(v)
    bind = sess.io_binding()
    bind.bind_ortvalue_output("denoised_audio", vout)
    checksum = 0.0
    for i in range(args.warmup):
        bind.bind_ortvalue_input("noisy_audio", vins[i % n_sets])
        sess.run_with_iobinding(bind)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        bind.bind_ortvalue_input("noisy_audio", vins[i % n_sets])
        sess.run_with_iobinding(bind)                           # synchronous, result is in host memory
        checksum += float(pin_out[0, 0, 0])
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    if dist is not None:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = audio_s / (e2e_ms * 1e-3)
    clk.__exit__(None, None, None)

    # ---------------- CPU baseline on this box's host cores (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate, dt = cpu_oracle_rate(sd, args.cpu_baseline_chunks, threads)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{args.cpu_baseline_chunks} of the {B} chunks of one step, chunk-at-a-time, {dt:.1f} s; "
                         "oracle/gtcrn_oracle.py (PyTorch-eager restatement of the graph ORT would run)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"GTCRN 16 kHz, {B} x 1 s chunks per GPU per step, F32 in / F32 out",
                       "model": "gtcrn", "batch_per_gpu": B, "chunk_samples": CHUNK,
                       "l2_policy": f"{n_sets} distinct input batches rotated ({n_sets * B * CHUNK * 4 / 2**20:.0f} MiB) "
                                    f"+ {model.workspace_bytes(B) / 2**20:.0f} MiB workspace streamed per step, both > 126 MB L2",
                       "parallelism": f"batch-shard x{world}, weights replicated, no data-path collective"},
            "rtf": 1.0 / value,
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * CHUNK * 4,
                    "d2h_bytes_per_step": B * L_OUT * 4, "ms_per_step": e2e_ms / args.steps,
                    "api": "adn.ort_shim.InferenceSession.run_with_iobinding -> adn_run_host (pinned host buffers)"},
            "gpu_launches": model.launches_per_run(B) * args.steps,
            "roofline": roof,
            "cpu_baseline": cpu,
            "kernels_ms_per_step": {k: round(v[0], 4) for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1][0])},
            "lib": _lib.lib().adn_version().decode(),
        }
        print(json.dumps(line))
    model.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

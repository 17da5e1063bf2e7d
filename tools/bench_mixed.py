"""BASELINE.json configs[4]: MossFormerGAN-SE-16K + MossFormer2-SS-16K mixed stream, 1 s chunks at 16 kHz.

A stream of `--chunks` requests (alternating enhancement / separation) is routed by `adn.dist.run_mixed_stream`: grouped
per model, each model's batch sharded over all ranks (weights of both models on every GPU), results back in request
order.  One process per GPU (torchrun) or a single process (world size 1).  Device time (CUDA events, max over ranks)
for the whole stream -> audio-seconds per second; rank 0 prints one JSON line.

    python tools/bench_mixed.py [--chunks 512] [--steps 3] [--warmup 1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_mixed.py
"""
import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "audio-denoiser-onnx_b200"), str(ROOT / "oracle")]

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=512)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--ss-batch", type=int, default=64, help="MossFormer2-SS windows per launch (bounds its workspace)")
    args = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    import mf2ss_oracle as so           # seeded weights only
    import mfgan_oracle as go
    from adn import export, mf2ss_params, mfgan_params
    from adn.dist import run_mixed_stream

    L = 16000
    gan = export.mfgan_model(go.random_state_dict(go.GanConfig(layers=6), 0), mfgan_params.GanHyper(layers=6), L, "F32", "F32", device_id=local)
    ss = export.mf2ss_model(so.random_state_dict(so.SsConfig(layers=24), 0), mf2ss_params.SsHyper(layers=24), L, "F32", "F32", device_id=local)

    def run_ss(x):                       # bounded sub-batches: the SS workspace is ~0.5 GB per window
        outs = [ss.run(x[i:i + args.ss_batch].contiguous()) for i in range(0, x.shape[0], args.ss_batch)]
        return tuple(torch.cat([o[k] for o in outs], 0) for k in range(2))

    run_fns = {"enhance": gan.run, "separate": run_ss}
    specs = {"enhance": ((1, L), torch.float32), "separate": ((1, L), torch.float32)}
    requests = None
    if rank == 0:
        g = torch.Generator().manual_seed(1234)
        t = torch.arange(L, dtype=torch.float32) / 16000
        requests = []
        for i in range(args.chunks):
            x = 0.2 * torch.randn(1, L, generator=g) + 0.3 * torch.sin(2 * torch.pi * (200.0 + i) * t)
            x = x / x.abs().amax() * 0.5
            tag = "enhance" if i % 2 == 0 else "separate"
            requests.append((tag, (x if tag == "enhance" else x * 32767.0).to(dev)))

    def step():
        return run_mixed_stream(run_fns, requests, specs, dev)

    for _ in range(max(args.warmup, 1)):
        res = step()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = step()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        assert len(res) == args.chunks and res[0].shape == (1, L) and isinstance(res[1], tuple) and res[1][0].shape == (1, L)
        ok = all(torch.isfinite(r if not isinstance(r, tuple) else r[0]).all().item() for r in res[:8])
        audio_s = args.chunks * args.steps * 1.0
        val = audio_s / (float(ms) * 1e-3)
        print(json.dumps({"metric": "audio-seconds per second, mixed MossFormerGAN-SE-16K + MossFormer2-SS-16K stream",
                          "value": val, "unit": "audio-s/s", "rtf": 1.0 / val, "n_gpus": world, "steps": args.steps,
                          "ms_per_step": float(ms) / args.steps, "chunks_per_step": args.chunks, "finite": ok,
                          "config": {"workload": f"{args.chunks} x 1 s @16 kHz requests, alternating enhancement / separation, "
                                                 f"inputs resident on rank 0's GPU, scatter -> run -> gather per model",
                                     "parallelism": f"each model's batch sharded x{world}, both models' weights on every GPU"}}))
    gan.close()
    ss.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

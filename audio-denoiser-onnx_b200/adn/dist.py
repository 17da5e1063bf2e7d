"""Multi-GPU batch sharding: chunks are independent, so the batch is split into
contiguous blocks, one per rank (one process per GPU, weights replicated), with exactly one
exchange step each side -- scatter the inputs, gather the outputs -- over
`torch.distributed` (NCCL over NVLink on the GPU box, gloo in the CPU tests).  There is no
collective inside the model."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block partition; the first n % world ranks get one extra chunk."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def scatter_batch(audio: torch.Tensor | None, n: int, shape_tail: tuple, dtype, device, src: int = 0, group=None):
    """Rank `src` holds audio (n, *shape_tail); every rank receives its block."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(n, world, rank)
    cap = shard_bounds(n, world, 0)[1]                       # largest block
    recv = torch.zeros((max(cap, 1), *shape_tail), dtype=dtype, device=device)
    if rank == src:
        parts = []
        for r in range(world):
            a, b = shard_bounds(n, world, r)
            p = torch.zeros((max(cap, 1), *shape_tail), dtype=dtype, device=device)
            if b > a:
                p[: b - a] = audio[a:b].to(device)
            parts.append(p)
        dist.scatter(recv, parts, src=src, group=group)
    else:
        dist.scatter(recv, None, src=src, group=group)
    return recv[: hi - lo]


def gather_batch(out_local: torch.Tensor, n: int, dst: int = 0, group=None):
    """Inverse of scatter_batch: rank `dst` gets the (n, ...) concatenation, others None."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    cap = max(shard_bounds(n, world, 0)[1], 1)
    pad = torch.zeros((cap, *out_local.shape[1:]), dtype=out_local.dtype, device=out_local.device)
    pad[: out_local.shape[0]] = out_local
    if rank == dst:
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.gather(pad, bufs, dst=dst, group=group)
        parts = []
        for r in range(world):
            a, b = shard_bounds(n, world, r)
            parts.append(bufs[r][: b - a])
        return torch.cat(parts, dim=0)
    dist.gather(pad, None, dst=dst, group=group)
    return None


def run_sharded(run_fn, audio: torch.Tensor | None, n: int, shape_tail: tuple, dtype, device, group=None):
    """scatter -> run_fn(local block) -> gather.  `run_fn` maps (b, C, L) -> (b, C, L_out) on
    `device` (Model.run on the GPU box).  Ranks whose block is empty skip the run (B=1 => one
    GPU active)."""
    local = scatter_batch(audio, n, shape_tail, dtype, device, group=group)
    if local.shape[0] > 0:
        out = run_fn(local.contiguous())
    else:
        out = run_fn(torch.zeros((1, *shape_tail), dtype=dtype, device=device))
        out = tuple(o[:0] for o in out) if isinstance(out, (tuple, list)) else out[:0]
    if isinstance(out, (tuple, list)):                    # several outputs (MossFormer2-SS: one per speaker)
        parts = [gather_batch(o, n, group=group) for o in out]
        return None if parts[0] is None else tuple(parts)
    return gather_batch(out, n, group=group)


def run_mixed_stream(run_fns: dict, requests: list | None, specs: dict, device, group=None):
    """Mixed-model stream (BASELINE.json configs[4]: MossFormerGAN-SE + MossFormer2-SS chunks interleaved in one
    stream).  Every rank holds every model's weights (they all fit on one GPU), so routing is: group the
    requests by model tag, shard EACH model's batch over all ranks (scatter -> run -> gather, no other collective),
    and put the results back in request order.

    run_fns:  tag -> callable (b, C, L) -> tensor or tuple of tensors (Model.run on the GPU box)
    requests: on rank 0 a list of (tag, tensor (C, L)); None elsewhere
    specs:    tag -> ((C, L), dtype) of that model's input, known on every rank
    Returns on rank 0 a list, one entry per request (tensor (C', L') or tuple of them); None elsewhere."""
    rank = dist.get_rank(group)
    tags = sorted(run_fns)
    counts = torch.zeros(len(tags), dtype=torch.int64, device=device)      # (NCCL moves device tensors only)
    if rank == 0:
        for t, _ in requests:
            counts[tags.index(t)] += 1
    dist.broadcast(counts, src=0, group=group)
    results = [None] * (len(requests) if rank == 0 else 0)
    for ti, tag in enumerate(tags):
        n = int(counts[ti])
        if n == 0:
            continue
        shape_tail, dtype = specs[tag]
        batch, where = None, []
        if rank == 0:
            where = [i for i, (t, _) in enumerate(requests) if t == tag]
            batch = torch.stack([requests[i][1] for i in where], dim=0)
        out = run_sharded(run_fns[tag], batch, n, tuple(shape_tail), dtype, device, group=group)
        if rank == 0:
            for j, i in enumerate(where):
                results[i] = tuple(o[j] for o in out) if isinstance(out, tuple) else out[j]
    return results if rank == 0 else None

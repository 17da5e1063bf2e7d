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


def _exchange(ops):
    """One grouped launch of point-to-point transfers (ncclGroupStart / ncclSend / ncclRecv / ncclGroupEnd on the GPU box)."""
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def scatter_batch(audio: torch.Tensor | None, n: int, shape_tail: tuple, dtype, device, src: int = 0, group=None):
    """Rank `src` holds audio (n, *shape_tail) on `device`; every rank receives exactly its block: grouped sends of the
    contiguous row ranges, no padding and no staging copies (rank `src` keeps a view of its own rows)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(n, world, rank)
    if rank == src:
        audio = audio.to(device).contiguous()
        ops = []
        for r in range(world):
            a, b = shard_bounds(n, world, r)
            if r != src and b > a:
                ops.append(dist.P2POp(dist.isend, audio[a:b], r, group))
        _exchange(ops)
        return audio[lo:hi]
    recv = torch.empty((hi - lo, *shape_tail), dtype=dtype, device=device)
    if hi > lo:
        _exchange([dist.P2POp(dist.irecv, recv, src, group)])
    return recv


def gather_batch(out_local: torch.Tensor, n: int, dst: int = 0, group=None, out: torch.Tensor | None = None):
    """Inverse of scatter_batch: rank `dst` gets the (n, ...) concatenation (written straight into `out` when given), others
    None.  Every rank's block lands in its row range of the result: grouped receives, no padding, no concatenation copy."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(n, world, rank)
    if rank == dst:
        if out is None:
            out = torch.empty((n, *out_local.shape[1:]), dtype=out_local.dtype, device=out_local.device)
        out[lo:hi].copy_(out_local)
        ops = []
        for r in range(world):
            a, b = shard_bounds(n, world, r)
            if r != dst and b > a:
                ops.append(dist.P2POp(dist.irecv, out[a:b], r, group))
        _exchange(ops)
        return out
    if hi > lo:
        _exchange([dist.P2POp(dist.isend, out_local.contiguous(), dst, group)])
    return None


def run_sharded(run_fn, audio: torch.Tensor | None, n: int, shape_tail: tuple, dtype, device, group=None, marks=None):
    """scatter -> run_fn(local block) -> gather.  `run_fn` maps (b, C, L) -> (b, C, L_out) on
    `device` (Model.run on the GPU box).  Ranks whose block is empty skip the run (B=1 => one
    GPU active).  `marks`, when given, is called with "scattered" and "ran" between the three phases (bench.py records
    CUDA events there to report the exchange's share of the step)."""
    local = scatter_batch(audio, n, shape_tail, dtype, device, group=group)
    if marks:
        marks("scattered")
    if local.shape[0] > 0:
        out = run_fn(local.contiguous())
    else:
        out = run_fn(torch.zeros((1, *shape_tail), dtype=dtype, device=device))
        out = tuple(o[:0] for o in out) if isinstance(out, (tuple, list)) else out[:0]
    if marks:
        marks("ran")
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

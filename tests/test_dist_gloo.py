"""N>1 path on CPU: world_size-2 gloo run of the batch scatter / run / gather plumbing
(adn/dist.py) with a stand-in per-chunk function (the CUDA model cannot run here)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_model(x):
    # independent per chunk, length-changing like GTCRN (16000 -> 15872 becomes L -> L-8)
    return (x[..., :-8] * 2.0 + x.mean(dim=-1, keepdim=True)).contiguous()


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "audio-denoiser-onnx_b200"))
    from adn import dist as adist

    g = torch.Generator().manual_seed(0)
    full = torch.randn(n, 1, 64, generator=g)
    out = adist.run_sharded(_fake_model, full if rank == 0 else None, n, (1, 64), torch.float32, torch.device("cpu"))
    if rank == 0:
        q.put(out.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1, 2, 5, 8])
def test_sharded_run_matches_single_process(n):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(0)
    full = torch.randn(n, 1, 64, generator=g)
    assert got.shape == (n, 1, 56)
    assert torch.equal(torch.from_numpy(got), _fake_model(full))


def test_shard_bounds_cover_everything():
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "audio-denoiser-onnx_b200"))
    from adn.dist import shard_bounds

    for n in (0, 1, 7, 64, 512, 513):
        for world in (1, 2, 4, 8):
            cuts = [shard_bounds(n, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def _fake_two_output_model(x):
    # two outputs like MossFormer2-SS (one waveform per speaker), independent per chunk
    return (x * 0.5).contiguous(), (x.flip(-1) - x.amax(dim=-1, keepdim=True)).contiguous()


def _mixed_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "audio-denoiser-onnx_b200"))
    from adn import dist as adist

    fns = {"enhance": _fake_model, "separate": _fake_two_output_model}
    specs = {"enhance": ((1, 64), torch.float32), "separate": ((1, 48), torch.float32)}
    reqs = None
    if rank == 0:
        g = torch.Generator().manual_seed(3)
        order = ["separate", "enhance", "enhance", "separate", "separate", "enhance", "separate"]
        reqs = [(t, torch.randn(*specs[t][0], generator=g)) for t in order]
    out = adist.run_mixed_stream(fns, reqs, specs, torch.device("cpu"))
    if rank == 0:
        q.put([(t, x.numpy(), [o.numpy() for o in (r if isinstance(r, tuple) else (r,))]) for (t, x), r in zip(reqs, out)])
    dist.barrier()
    dist.destroy_process_group()


def test_mixed_stream_routing_two_models():
    """BASELINE configs[4]-style mixed stream on world_size 2: requests of two model families interleaved, each family's
    batch sharded over both ranks, results returned in request order (tuple outputs included)."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_mixed_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(got) == 7
    for tag, x, outs in got:
        xt = torch.from_numpy(x).unsqueeze(0)
        ref = _fake_model(xt) if tag == "enhance" else _fake_two_output_model(xt)
        ref = ref if isinstance(ref, tuple) else (ref,)
        assert len(outs) == len(ref)
        for o, r in zip(outs, ref):
            assert torch.equal(torch.from_numpy(o), r[0])

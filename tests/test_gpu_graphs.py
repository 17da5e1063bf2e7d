"""-m gpu: adn_run replays a run as one CUDA graph on a non-default stream (the batch-1, launch-bound regime).  The replayed
run must equal the kernel-by-kernel run bit for bit, survive a workspace re-allocation (stale graphs are dropped), and never
engage on the legacy default stream."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _graph_launches(m):
    return int(m.debug_read("graph_launches")[0])


def _check(m, make_input, batches=(1, 3, 1)):
    side = torch.cuda.Stream()
    for B in batches:
        x = make_input(B).cuda()
        ref = m.run(x)                                    # legacy default stream: always eager
        ref = tuple(r.clone() for r in ref) if isinstance(ref, tuple) else ref.clone()
        torch.cuda.synchronize()
        before = _graph_launches(m)
        outs = []
        with torch.cuda.stream(side):
            out = None
            for _ in range(4):                            # eager, capture + launch, replay, replay
                out = m.run(x, out=out)
                side.synchronize()
                outs.append(tuple(o.clone() for o in out) if isinstance(out, tuple) else out.clone())
        for o in outs:
            for a, b in zip(o if isinstance(o, tuple) else (o,), ref if isinstance(ref, tuple) else (ref,)):
                assert torch.equal(a, b)
        assert _graph_launches(m) - before >= 2, "the run was not replayed as a graph"
    return True


def test_gtcrn_graph_replay(libadn):
    import gtcrn_oracle as go
    from adn import export
    from make_golden import synth_audio

    m = export.gtcrn_model(go.random_state_dict(0), 16000, "F32", "F32")
    assert _graph_launches(m) == 0
    m.run(synth_audio(16000, 1, 2).cuda())
    torch.cuda.synchronize()
    assert _graph_launches(m) == 0                        # default stream: no graphs
    _check(m, lambda B: synth_audio(16000, 5 + B, B))
    m.close()


def test_zipenh_graph_replay(libadn):
    import zipenh_oracle as zo
    from adn import export

    m = export.zipenh_model(zo.random_state_dict(zo.ZipConfig(), 0), None, 1600, "F32", "F32")
    g = torch.Generator().manual_seed(1)
    _check(m, lambda B: (torch.rand(B, 1, 1600, generator=g) * 2 - 1) * 0.5)
    m.close()


def test_mf2se_graph_survives_replanning(libadn):
    """MossFormer2-SE re-plans its workspace on every batch change: graphs captured before hold stale addresses and must go."""
    import mf2se_oracle as so
    from adn import export, mf2se_params

    cfg = so.Mf2Config(layers=2)
    m = export.mf2se_model(so.random_state_dict(cfg, 0), mf2se_params.Mf2Hyper(layers=2), 11520, "F32", "F32")
    g = torch.Generator().manual_seed(2)
    _check(m, lambda B: (torch.rand(B, 1, 11520, generator=g) * 2 - 1) * 0.3, batches=(2, 1, 2, 5, 2))
    m.close()


def test_mf2ss_two_outputs_graph(libadn):
    import mf2ss_oracle as so
    from adn import export, mf2ss_params

    cfg = so.SsConfig(layers=2)
    m = export.mf2ss_model(so.random_state_dict(cfg, 0), mf2ss_params.SsHyper(layers=2), 2408, "F32", "F32")
    g = torch.Generator().manual_seed(3)
    _check(m, lambda B: (torch.rand(B, 1, 2408, generator=g) * 2 - 1) * 8000.0, batches=(1, 2))
    m.close()

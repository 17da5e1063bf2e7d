"""Experiment: one model over B windows on one stream vs two model instances over B/2 windows each on two
streams (tensor-bound GEMMs of one half can overlap the memory-bound kernels of the other)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / "audio-denoiser-onnx_b200", ROOT / "oracle"):
    sys.path.insert(0, str(p))
import torch

which, B, reps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
if which == "ss":
    import mf2ss_oracle as so
    from adn import export, mf2ss_params
    cfg = so.SsConfig(); sd = so.random_state_dict(cfg, 0); L = 16000
    make = lambda: export.mf2ss_model(sd, mf2ss_params.SsHyper(), L)
    x = ((torch.rand(B, 1, L) - 0.5) * 20000.0).cuda()
else:
    import mf2se_oracle as mo
    from adn import export, mf2se_params
    cfg = mo.Mf2Config(); sd = mo.random_state_dict(cfg, 0); L = 48000
    make = lambda: export.mf2se_model(sd, mf2se_params.Mf2Hyper(), L)
    x = (torch.rand(B, 1, L) - 0.5).cuda()

def timeit(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

m = make()
t1 = timeit(lambda: m.run(x))
m.close()
ma, mb = make(), make()
xa, xb = x[: B // 2].contiguous(), x[B // 2:].contiguous()
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
def two():
    cur = torch.cuda.current_stream()
    sa.wait_stream(cur); sb.wait_stream(cur)
    ma.run(xa, stream=sa.cuda_stream)
    mb.run(xb, stream=sb.cuda_stream)
    cur.wait_stream(sa); cur.wait_stream(sb)
t2 = timeit(two)
print(f"{which} B={B}: one stream {t1:.2f} ms, two streams x B/2 {t2:.2f} ms  ({t1 / t2:.2f}x)")

"""Per-kernel CUDA-event times of MossFormer2-SS (layers, batch, reps from argv): quick A/B tool."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / "audio-denoiser-onnx_b200", ROOT / "oracle"):
    sys.path.insert(0, str(p))
import torch

import mf2ss_oracle as so
from adn import export, mf2ss_params

layers, B, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
cfg = so.SsConfig(layers=layers)
sd = so.random_state_dict(cfg, 0)
m = export.mf2ss_model(sd, mf2ss_params.SsHyper(layers=layers), 16000)
xs = [((torch.rand(B, 1, 16000) - 0.5) * 20000.0).cuda() for _ in range(2)]
for i in range(2):
    m.run(xs[i % 2])
torch.cuda.synchronize()
m.set_profiling(True)
acc = {}
for i in range(reps):
    m.run(xs[i % 2])
    torch.cuda.synchronize()
    for name, ms in m.kernel_times():
        acc.setdefault(name, []).append(ms)
tot = 0.0
for k, v in sorted(acc.items(), key=lambda kv: -sum(kv[1])):
    per_launch = sum(v) / len(v)
    tot += sum(v) / reps
    print(f"{k:16s} {per_launch * 1e3:9.1f} us/launch  x{len(v) // reps:3d}  {sum(v) / reps:8.3f} ms/run")
print(f"total {tot:.3f} ms/run")

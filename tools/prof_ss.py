"""ncu target: a few MossFormer2-SS windows through the C ABI (layer count / batch / reps from argv)."""
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
L = int(sys.argv[4]) if len(sys.argv) > 4 else 16000
m = export.mf2ss_model(sd, mf2ss_params.SsHyper(layers=layers), L)
x = ((torch.rand(B, 1, L) - 0.5) * 20000.0).cuda()
for _ in range(reps):
    y = m.run(x)
torch.cuda.synchronize()
print(float(y[0].abs().max()))

"""ncu target: a few MossFormer2-SE windows through the C ABI (layer count / batch from argv)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / "audio-denoiser-onnx_b200", ROOT / "oracle"):
    sys.path.insert(0, str(p))
import torch

import mf2se_oracle as mo
from adn import export, mf2se_params

layers, B, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
cfg = mo.Mf2Config(layers=layers)
sd = mo.random_state_dict(cfg, 0)
mm = "BF16" if len(sys.argv) > 4 and sys.argv[4] == "bf16" else "F32"
m = export.mf2se_model(sd, mf2se_params.Mf2Hyper(layers=layers), 48000, matmul_dtype=mm)
x = (torch.rand(B, 1, 48000) - 0.5).cuda()
for _ in range(reps):
    y = m.run(x)
torch.cuda.synchronize()
print(float(y.abs().max()))

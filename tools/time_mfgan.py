"""Times the MossFormerGAN-SE-16K path on one GPU and prints the per-operator breakdown (CUDA events per launch).
usage: python tools/time_mfgan.py [batch] [layers]"""
import json
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "audio-denoiser-onnx_b200"), str(ROOT / "oracle")]
import torch

import mfgan_oracle as go            # seeded weights only
from adn import export, mfgan_params

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 6
cfg = go.GanConfig(layers=layers)
sd = go.random_state_dict(cfg, 0)
L = 16000
m = export.mfgan_model(sd, mfgan_params.GanHyper(layers=layers), L)
x = ((torch.rand(B, 1, L) * 2 - 1) * 0.3).cuda()
for _ in range(2):
    m.run(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    m.run(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
m.set_profiling(True)
m.run(x)
torch.cuda.synchronize()
agg, cnt = defaultdict(float), defaultdict(int)
for n, t in m.kernel_times():
    agg[n] += t
    cnt[n] += 1
tot = sum(agg.values())
print(json.dumps({"model": "mossformergan_se_16k", "batch": B, "layers": layers, "ms_per_run": ms, "audio_s_per_s": B / (ms / 1e3),
                  "rtf": (ms / 1e3) / B, "launches": m.launches_per_run(B), "workspace_MB": m.workspace_bytes(B) / 2 ** 20}))
for n in sorted(agg, key=agg.get, reverse=True):
    print(f"{n:22s} {cnt[n]:5d} launches {agg[n]:10.3f} ms {100 * agg[n] / tot:6.2f} %")

"""Runs one model a few times on cuda:0 (for ncu: a short command with a known launch sequence).
    python tools/run_once.py --model zipenh --batch 64 --runs 2"""
import argparse
import importlib.util
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "audio-denoiser-onnx_b200"), str(ROOT / "oracle")]
import torch

spec = importlib.util.spec_from_file_location("bench", ROOT / "bench.py")
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="zipenh")
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--runs", type=int, default=2)
a = ap.parse_args()
wl = bench.WORKLOADS[a.model]()
B = a.batch or wl.default_batch
m = wl.build(wl.weights(), 0)
x = wl.inputs(B, 1, 1234)[0].cuda()
for _ in range(a.runs):
    y = m.run(x)
torch.cuda.synchronize()
print(a.model, B, "launches per run", m.launches_per_run(B))
m.close()

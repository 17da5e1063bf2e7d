#!/bin/sh
# Round-2 multi-GPU measurements on one 8 x B200 box (run under `gpurun --gpus 8`): the real data path
# (adn.dist.run_sharded: NCCL grouped send/recv scatter -> run -> gather) on the clock.
#   strong scaling: ZipEnhancer 64 and 512 windows in total (BASELINE configs[1] shape), MossFormer2-SE 256 in total (configs[2])
#   configs[3]: Mel-Band-Roformer 128 x 8 s stereo segments folded into 768 windows, 8 GPUs
#   configs[4]: MossFormerGAN + MossFormer2-SS mixed stream, 512 x 1 s requests, 8 GPUs
# One JSON line per run -> gpurun_out/r2_multigpu.jsonl
OUT=gpurun_out/r2_multigpu.jsonl
: > $OUT
run() {  # nproc, then bench.py args
  n=$1; shift
  port=$((29600 + $(wc -l < $OUT)))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port "$@" 2>> gpurun_out/r2_multigpu.err | grep '^{' >> $OUT
}
for n in 1 2 4 8; do
  run $n bench.py --gpus $n --model zipenh --scaling strong --batch 64 --steps 5 --warmup 3
done
for n in 1 8; do
  run $n bench.py --gpus $n --model zipenh --scaling strong --batch 512 --steps 3 --warmup 3
done
for n in 1 2 4 8; do
  run $n bench.py --gpus $n --model mf2se --scaling strong --batch 256 --steps 10 --warmup 3
done
run 8 bench.py --gpus 8 --model mbr --scaling strong --segments 128x8s --steps 3 --warmup 3
run 8 tools/bench_mixed.py --chunks 512 --steps 2 --warmup 1
wc -l $OUT
python - <<'PY'
import json
for line in open("gpurun_out/r2_multigpu.jsonl"):
    d = json.loads(line)
    ex = d.get("exchange") or {}
    print(d.get("config", {}).get("model", "mixed"), d["n_gpus"], "GPUs", round(d["ms_per_step"], 2), "ms/step", round(d["value"], 1), d["unit"],
          "exchange share", round(ex.get("share_of_step", float("nan")), 4), "e2e", round((d.get("e2e") or {}).get("value", float("nan")), 1))
PY

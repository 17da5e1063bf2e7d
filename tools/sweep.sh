# RTF / audio-s/s per model at batch 1 / 64 / 512 on one B200 (BASELINE.json metric); one JSON line per run
mkdir -p gpurun_out
out=gpurun_out/sweep.jsonl
: > $out
run() { timeout 900 python bench.py --model $1 --batch $2 --steps $3 --warmup 3 --no-cpu-baseline 2>gpurun_out/sweep_err.log | tail -1 >> $out || echo "{\"model\": \"$1\", \"batch\": $2, \"failed\": true}" >> $out; }
run gtcrn 1 200; run gtcrn 64 100; run gtcrn 512 100
run mf2se 1 20; run mf2se 64 10; run mf2se 512 5
run mf2ss 1 10; run mf2ss 64 5; run mf2ss 512 3
run mbr 1 10; run mbr 64 5; run mbr 256 3
# families added late in round 1 (first-correct designs: small step counts; dfsmn / ulunas have not run on a GPU yet)
run mfgan 1 5; run mfgan 64 3; run mfgan 512 1
run dfsmn 1 20; run dfsmn 64 10; run dfsmn 512 5
run ulunas 1 5; run ulunas 64 3; run ulunas 512 1
python - <<PY
import json
for l in open("$out"):
    try: d = json.loads(l)
    except Exception: print("bad line", l[:200]); continue
    if d.get("failed"): print(d); continue
    print(f"{d['config']['model']:6s} B={d['config']['batch_per_gpu']:4d}  {d['ms_per_step']:10.3f} ms/step  {d['value']:12.1f} audio-s/s  RTF {d['rtf']:.3e}  e2e {d['e2e']['value']:12.1f}")
PY

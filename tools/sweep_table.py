"""Markdown table of a `bench.py --sweep` JSON line (profiles/r2*_batch_sweep.json) for DESIGN.md."""
import json
import sys

NAMES = {"gtcrn": "GTCRN 16 kHz (1 s chunks)", "zipenh": "ZipEnhancer 16 kHz (1 s windows)", "mf2se": "MossFormer2-SE-48K (1 s windows)",
         "mbr": "Mel-Band-Roformer stereo depth 6 (1.5 s windows)", "mfgan": "MossFormerGAN-SE-16K (1 s windows)",
         "mf2ss": "MossFormer2-SS-16K (1 s windows, 2 speakers)", "dfsmn": "DFSMN 48 kHz (1 s windows)", "ulunas": "UL-UNAS 16 kHz (1 s windows)"}
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("| model | batch | windows / launch | ms / step | audio-s / s | RTF | e2e audio-s / s |")
print("|---|---|---|---|---|---|---|")
for r in d["table"]:
    print(f"| {NAMES.get(r['model'], r['model'])} | {r['batch']} | {r['windows_per_launch']} | {r['ms_per_step']:.3f} | {r['audio_s_per_s']:,.0f} | "
          f"{r['rtf']:.2e} | {r['e2e_audio_s_per_s']:,.0f} |")

#!/bin/sh
# Per-kernel counts of the Blackwell tensor / TMA SASS mnemonics in the in-tree libadn.so (B200_PROFILING.md: UTCHMMA = tcgen05.mma,
# UTMALDG = TMA tensor load, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops).  Writes profiles/sass_summary.txt.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
LIB="$ROOT/audio-denoiser-onnx_b200/libadn.so"
OUT="$ROOT/profiles/sass_summary.txt"
{
  echo "# cuobjdump -sass $(basename "$LIB") ($(date -u +%Y-%m-%dT%H:%MZ)); build: nvcc -gencode arch=compute_100a,code=sm_100a"
  echo "# kernel | UTCHMMA | UTMALDG | LDTM | UTCBAR | SYNCS | HMMA(legacy mma.sync)"
  cuobjdump -sass "$LIB" | awk '
    /Function :/ { if (name != "") print name, mma, tma, ldtm, bar, syncs, hmma; name=$3; mma=tma=ldtm=bar=syncs=hmma=0 }
    /UTCHMMA|UTCQMMA|UTCIMMA/ { mma++ }
    /UTMALDG/ { tma++ }
    /LDTM/ { ldtm++ }
    /UTCBAR/ { bar++ }
    /SYNCS/ { syncs++ }
    / HMMA/ { hmma++ }
    END { if (name != "") print name, mma, tma, ldtm, bar, syncs, hmma }' | awk '$2 + $3 + $4 + $5 + $7 > 0' | while read n a b c d e f; do
      echo "$(echo "$n" | c++filt | cut -c1-110) | $a | $b | $c | $d | $e | $f"
    done
  echo "# totals over the library:"
  cuobjdump -sass "$LIB" | grep -o -E "UTCHMMA|UTMALDG\.[0-9]D|LDTM\.[x0-9a-zA-Z.]*|STTM\.[x0-9a-zA-Z.]*|UTCBAR|UTMASTG|R2UR\.BROADCAST|ELECT" | sort | uniq -c
} > "$OUT"
echo "wrote $OUT"

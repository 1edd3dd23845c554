#!/bin/bash
# GroupNorm chunk-geometry sweep (L2 budget x items per block) at the n_img of groups-per-call 1 and 5
cd "$(dirname "$0")/.." || exit 1
for n in ${NS:-16 80}; do
  for l2 in ${L2S:-32 48 64 96 128}; do
    for w in ${WAVES:-2 4}; do
      echo "== n=$n l2=$l2 waves=$w"
      CAP4D_GN_L2_MB=$l2 CAP4D_GN_WAVES=$w timeout 120 python scripts/dbg_norm.py $n | grep GN | awk '{printf "%s %s %s %s | ", $3, $4, $5, $6} END {print ""}'
    done
  done
done

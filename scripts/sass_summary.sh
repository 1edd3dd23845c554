#!/bin/bash
# SASS evidence of the Blackwell-native paths in the shipped library (B200_PROFILING.md, "What proves a Blackwell-native
# kernel"): per kernel, how many tcgen05 MMAs (UTC*MMA), TMA loads (UTMALDG), TMEM loads / stores (LDTM / STTM), tensor-core
# barriers (UTCBAR), MUFU.EX2, packed FFMA2 - and legacy HMMA (none expected).
#   scripts/sass_summary.sh > profiles/r02_sass_summary.txt
cd "$(dirname "$0")/.." || exit 1
LIB=cap4d_b200/libcap4d_b200.so
echo "# cuobjdump -sass $LIB ($(sha256sum $LIB | cut -c1-16)), source digest $(python -c 'from cap4d_b200 import build; print(build.kernel_digest()[:16])')"
printf "%-44s %8s %8s %6s %6s %7s %8s %7s %6s\n" kernel UTC*MMA UTMALDG LDTM STTM UTCBAR MUFU.EX2 FFMA2 HMMA
cuobjdump -sass "$LIB" 2>/dev/null | c++filt | awk '
  /Function :/ { if (name != "") emit(); name=$0; sub(/.*Function : /, "", name); mma=tma=ld=st=bar=ex2=f2=h=0; next }
  /UTC[A-Z]*MMA/ {mma++} /UTMALDG/ {tma++} /LDTM/ {ld++} /STTM/ {st++} /UTCBAR/ {bar++} /MUFU.EX2/ {ex2++} /FFMA2/ {f2++} / HMMA/ {h++}
  function emit() { if (mma+tma+ld+st+bar > 0) printf "%-44s %8d %8d %6d %6d %7d %8d %7d %6d\n", short(name), mma, tma, ld, st, bar, ex2, f2, h }
  function short(n,   s) { s=n; gsub(/cap4d::\(anonymous namespace\)::/, "", s); gsub(/\(.*$/, "", s); gsub(/^void /, "", s); gsub(/cap4d::/, "", s); return substr(s, 1, 44) }
  END { emit() }'

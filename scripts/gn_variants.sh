#!/bin/bash
# A/B GroupNorm builds: gpurun_variants/lib_gn_*.so are swapped in as the product library (norm parity tests, then the
# per-shape timings at 16 and 80 images).  Build them with scripts/build_variant.py gn_NAME -DCAP4D_GN_...=...
cd "$(dirname "$0")/.." || exit 1
cp cap4d_b200/libcap4d_b200.so /tmp/lib_orig.so
for v in gpurun_variants/lib_gn_*.so; do
  cp "$v" cap4d_b200/libcap4d_b200.so
  echo "== $v"
  [ -z "$SKIP_TESTS" ] && timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -p no:cacheprovider -k "norm" 2>&1 | tail -1
  for n in 16 80; do timeout 120 python scripts/dbg_norm.py $n | grep GN | awk -v n=$n '{printf "n=%s %s %s %s %s\n", n, $3, $4, $5, $6}'; done
done
cp /tmp/lib_orig.so cap4d_b200/libcap4d_b200.so

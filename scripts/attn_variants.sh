#!/bin/bash
# A/B attention-kernel builds: gpurun_variants/lib_*.so are swapped in as the product library
cd "$(dirname "$0")/.." || exit 1
cp cap4d_b200/libcap4d_b200.so /tmp/lib_orig.so
for v in gpurun_variants/lib_*.so; do
  case "$v" in *lib_gn_*) continue;; esac
  cp "$v" cap4d_b200/libcap4d_b200.so
  echo "== $v"
  [ -z "$SKIP_TESTS" ] && timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -p no:cacheprovider -k attention 2>&1 | tail -1
  for shape in "4096 320 16" "4096 320 80" "8192 640 2" "8192 640 10" "2048 1280 2" "2048 1280 10" "512 1280 10"; do timeout 100 python scripts/attn_trace.py $shape | tail -1; done
done
cp /tmp/lib_orig.so cap4d_b200/libcap4d_b200.so

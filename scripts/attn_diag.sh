#!/bin/bash
# A/B of attention builds (VARIANTS or gpurun_variants/lib_p_*.so): parity tests, clock64 trace + timing at the level-0/1 shapes.
cd "$(dirname "$0")/.." || exit 1
cp cap4d_b200/libcap4d_b200.so /tmp/lib_orig.so
for v in ${VARIANTS:-gpurun_variants/lib_p_*.so}; do
  cp "$v" cap4d_b200/libcap4d_b200.so
  echo "== $v"
  [ -z "$SKIP_TESTS" ] && timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -p no:cacheprovider -k attention 2>&1 | tail -1
  for shape in "4096 320 80" "8192 640 10" "2048 1280 10"; do timeout 100 python scripts/attn_trace.py $shape; done
done
cp /tmp/lib_orig.so cap4d_b200/libcap4d_b200.so

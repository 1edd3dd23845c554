#!/bin/bash
# A/B GEMM-kernel builds: gpurun_variants/lib_*.so are swapped in as the product library
cd "$(dirname "$0")/.." || exit 1
cp cap4d_b200/libcap4d_b200.so /tmp/lib_orig.so
for v in gpurun_variants/lib_*.so; do
  cp "$v" cap4d_b200/libcap4d_b200.so
  echo "== $v"
  [ -z "$SKIP_TESTS" ] && timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -p no:cacheprovider -k "gemm or geglu or conv" 2>&1 | tail -1
  for n in ${NIMGS:-16 80}; do N_IMG=$n timeout 300 python scripts/bench_shapes.py | awk -v n=$n '{print "n=" n " " $0}'; done
done
cp /tmp/lib_orig.so cap4d_b200/libcap4d_b200.so

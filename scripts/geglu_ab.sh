#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
cp cap4d_b200/libcap4d_b200.so /tmp/lib_orig.so
for v in gpurun_variants/lib_*.so; do
  cp "$v" cap4d_b200/libcap4d_b200.so
  echo "== $v"
  timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -p no:cacheprovider -k "geglu" 2>&1 | tail -1
  for shp in "65536 2560 320" "327680 2560 320" "81920 5120 640" "20480 10240 1280"; do timeout 100 python scripts/one_gemm.py $shp 2 0 10; done
done
cp /tmp/lib_orig.so cap4d_b200/libcap4d_b200.so

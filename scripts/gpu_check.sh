#!/bin/bash
# Run the GPU parity suite piecewise (one process per kernel family, so a CUDA fault in one
# family does not hide the others), logging under gpurun_out/.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh'
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
rc=0
run() {
  local name=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout 600 python -m pytest "$@" -q --tb=short -p no:cacheprovider -s > gpurun_out/$name.log 2>&1
  local r=$?
  tail -n 3 gpurun_out/$name.log | tee -a gpurun_out/summary.txt
  echo "exit $r" | tee -a gpurun_out/summary.txt
  [ $r -ne 0 ] && rc=1
}
: > gpurun_out/summary.txt
run gemm tests/test_gpu_kernels.py -m gpu -k "gemm"
run conv tests/test_gpu_kernels.py -m gpu -k "conv3x3"
run attn tests/test_gpu_kernels.py -m gpu -k "attention"
run norm tests/test_gpu_kernels.py -m gpu -k "groupnorm or layernorm or cfg_ddim"
run unet tests/test_gpu_unet.py -m gpu
grep -h -E "FAILED|ERROR|passed|failed|max-rel|PSNR" gpurun_out/*.log | head -80
exit $rc

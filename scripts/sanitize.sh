#!/bin/bash
# compute-sanitizer over the kernel parity tests at their small shapes: the only independent check
# of the hand-rolled mbarrier / TMEM / TMA protocols and of the GroupNorm image barrier.
#   gpurun --timeout 1500 -- scripts/sanitize.sh memcheck      (one tool per gpurun call: B200_PROFILING.md)
#   gpurun --timeout 1500 -- scripts/sanitize.sh racecheck
#   gpurun --timeout 1500 -- scripts/sanitize.sh synccheck
# Output: gpurun_out/r02_sanitizer_<tool>.log (copy the summary into profiles/).
cd "$(dirname "$0")/.." || exit 1
TOOL=${1:-memcheck}
mkdir -p gpurun_out
SMALL='(test_gemm_f32 and (128-64-64 or 256-128-128 or 100-96-192 or 32-64-128)) or test_gemm_geglu or test_gemm_operand_formats
 or (test_conv3x3_stride1 and (5-4-4-64 or 8-2-2-64 or 4-16-16-192)) or (test_conv3x3_stride2 and 8-4-4-64)
 or (test_upsample_conv3x3 and 8-2-2-64)
 or (test_attention and (1-128-64 or 2-256-128 or 3-64-64 or 2-200-64 or 8-16-256))
 or (test_groupnorm and (2-256-64 or 8-4-128 or 4-64-320)) or (test_layernorm and (7-640 or 1000-64)) or test_cfg_ddim'
SMALL=$(echo $SMALL | tr '\n' ' ')
# the same selection must pass without the tool first (a faulting program under a sanitizer can wedge the GPU)
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -p no:cacheprovider -k "$SMALL" > gpurun_out/r02_sanitizer_plain.log 2>&1 || { tail -5 gpurun_out/r02_sanitizer_plain.log; echo "plain run failed: not sanitizing"; exit 1; }
tail -1 gpurun_out/r02_sanitizer_plain.log
timeout 1300 compute-sanitizer --tool "$TOOL" --print-limit 50 --error-exitcode 3 \
  python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -p no:cacheprovider -k "$SMALL" > "gpurun_out/r02_sanitizer_$TOOL.log" 2>&1
echo "compute-sanitizer --tool $TOOL rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|hazard" "gpurun_out/r02_sanitizer_$TOOL.log" | head -30

#!/bin/bash
# `ncu --set full` captures of ONE launch each of the shipped conv3x3, GroupNorm and attention kernels at their
# level-0 production shapes (5 groups per call) -> gpurun_out/r02_full_{conv,gn,attn}.ncu-rep
# (read here with scripts/summarize_full.sh).  Every command runs plain first (B200_PROFILING.md).
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
cap() {  # name, kernel regex, command...
  local name=$1 rx=$2; shift 2
  "$@" > "gpurun_out/r02_full_${name}_plain.log" 2>&1 &&
  ncu --set full --clock-control none --import-source on -k "regex:$rx" -s 1 -c 1 -f -o "gpurun_out/r02_full_$name" "$@" > "gpurun_out/r02_full_${name}_ncu.log" 2>&1
  echo "$name rc=$? $(tail -n 1 gpurun_out/r02_full_${name}_plain.log)"
}
cap conv gemm_tc_kernel python scripts/one_conv.py 80 64 64 320 320 3
cap gn gn_fused_kernel python scripts/one_gn.py 80 4096 320 0 3
cap attn attn_tc_kernel python scripts/one_attn.py 4096 320 80 3
cap lin gemm_tc_kernel python scripts/one_gemm.py 327680 320 320 0 1 3
cap geglu gemm_tc_kernel python scripts/one_gemm.py 327680 2560 320 2 0 3
ls -la gpurun_out/r02_full_*.ncu-rep

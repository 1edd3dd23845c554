#!/bin/bash
# First GPU call of round 2: everything that was prepared blind at the end of round 1 gets its measurement.
#   scripts/prep_attn_variants.sh        (HERE, builds gpurun_variants/)
#   gpurun --timeout 1500 -- scripts/round2_first_call.sh
# Outputs: gpurun_out/r02_{l2_reuse.jsonl,attn_variants.log,gn_variants.log}
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "=== L2 producer->consumer reuse"; timeout 200 python scripts/l2_reuse.py > gpurun_out/r02_l2_reuse.jsonl 2> gpurun_out/r02_l2_reuse.err; cat gpurun_out/r02_l2_reuse.jsonl | cut -c1-400
echo "=== attention variants"; timeout 600 scripts/attn_variants.sh > gpurun_out/r02_attn_variants.log 2>&1; grep -E "^==|passed|failed|n_seq=80|n_seq=10|TFLOP" gpurun_out/r02_attn_variants.log
echo "=== GroupNorm variants"; timeout 400 scripts/gn_variants.sh > gpurun_out/r02_gn_variants.log 2>&1; grep -E "^==|passed|failed|hw=4096" gpurun_out/r02_gn_variants.log

#!/bin/bash
# ncu pass over bench.py ITSELF: DRAM bytes and duration of every kernel of the first two U-Net
# calls of the TIMED region (NVTX range "bench_timed") -> gpurun_out/r02_ncu_bench_traffic.csv;
# scripts/summarize_bench_traffic.py turns it into r02_traffic_bench.json (stamped with the digest of the kernel
# sources), which bench.py's roofline.traffic reads from profiles/.
#   gpurun --timeout 1200 -- scripts/ncu_bench_traffic.sh
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
ARGS="--steps 1 --warmup 3 --no-cuda-graph --no-cpu-baseline --no-gpu-eager-baseline ${BENCH_ARGS:-}"
COUNT=${COUNT:-540}
python bench.py $ARGS > gpurun_out/r02_ncu_bench_plain.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    --nvtx --nvtx-include "bench_timed/" -c "$COUNT" \
    --csv --log-file gpurun_out/r02_ncu_bench_traffic.csv python bench.py $ARGS > gpurun_out/r02_ncu_bench_run.log 2>&1
echo "ncu rc=$? count=$COUNT"
python scripts/summarize_bench_traffic.py gpurun_out/r02_ncu_bench_traffic.csv gpurun_out/r02_traffic_bench.json

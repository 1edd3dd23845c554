#!/bin/bash
# ncu pass over bench.py ITSELF: DRAM bytes and duration of every kernel of two U-Net calls of the
# timed region -> gpurun_out/r02_ncu_bench_traffic.csv; scripts/summarize_bench_traffic.py turns it into
# profiles/r02_traffic_bench.json (stamped with the digest of the kernel sources) which bench.py's roofline.traffic reads.
#   gpurun --timeout 1200 -- scripts/ncu_bench_traffic.sh
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
ARGS="--steps 1 --warmup 3 --no-cuda-graph --no-cpu-baseline --no-gpu-eager-baseline ${BENCH_ARGS:-}"
# launches before the timed region: 3 warm-up steps x calls per step x (kernels per call + gather + update)
SKIP=${SKIP:-$(python - <<'PY'
import json, subprocess, sys
out = subprocess.run([sys.executable, "scripts/count_launches.py"], capture_output=True, text=True).stdout.strip()
print(out or 19008)
PY
)}
COUNT=${COUNT:-540}
python bench.py $ARGS > gpurun_out/r02_ncu_bench_plain.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -s "$SKIP" -c "$COUNT" \
    --csv --log-file gpurun_out/r02_ncu_bench_traffic.csv python bench.py $ARGS > gpurun_out/r02_ncu_bench_run.log 2>&1
echo "ncu rc=$? skip=$SKIP count=$COUNT"
python scripts/summarize_bench_traffic.py gpurun_out/r02_ncu_bench_traffic.csv gpurun_out/r02_traffic_bench.json

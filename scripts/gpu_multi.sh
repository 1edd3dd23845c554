#!/bin/bash
# multi-GPU bench lines (N = $1): single_ref 840 views, multi_ref 840 and 80 views -> gpurun_out/r02_bench_${N}gpu_*.json
cd "$(dirname "$0")/.." || exit 1
N=${1:-2}
mkdir -p gpurun_out
run() {  # name, extra args
  python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus "$N" --steps "${STEPS:-3}" --warmup 3 $2 > "gpurun_out/r02_bench_${N}gpu_$1.json" 2> "gpurun_out/r02_bench_${N}gpu_$1.err"
  echo "$1 rc=$?"
  python - "gpurun_out/r02_bench_${N}gpu_$1.json" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print(" value", round(d["value"], 3), "e2e", round(d["e2e"]["value"], 3), "ms/step", round(d["ms_per_step"], 1), d["config"]["groups_per_call"],
          d["config"]["unet_calls_per_step_per_rank"], "graphs", d.get("cuda_graphs"), "checksum", d["e2e"]["checksum"])
except Exception as e:
    print(" no line:", e)
PY
}
for w in ${WORKLOADS:-single multi840 multi80}; do
  case $w in
    single) run single_ref "" ;;
    multi840) run multi_ref_840 "--workload multi_ref --n-gen 840" ;;
    multi80) run multi_ref_80 "--workload multi_ref --n-gen 80" ;;
  esac
done
for f in gpurun_out/r02_bench_${N}gpu_*.err; do tail -n 2 "$f"; done

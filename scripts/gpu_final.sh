#!/bin/bash
# Round-end GPU check: the driver's own commands (pytest -m gpu, smoke, bench) plus the conditioning-map bench and
# its ncu captures.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
set -o pipefail
echo "=== pytest"; timeout 600 python -m pytest tests/ -q -m gpu -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/final_pytest.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/final_smoke.log
echo "=== bench"; timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; cut -c1-300 gpurun_out/final_bench.json; tail -3 gpurun_out/final_bench.err
echo "=== cond bench"; timeout 100 python scripts/bench_cond.py > gpurun_out/final_cond_bench.json 2> gpurun_out/final_cond_bench.err; cut -c1-300 gpurun_out/final_cond_bench.json; tail -3 gpurun_out/final_cond_bench.err
echo "=== ncu cond launch list"
timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    -k regex:"cond_pos_enc|face_bins" --log-file gpurun_out/final_cond_launches.csv python scripts/bench_cond.py --iters 2 --no-cpu > gpurun_out/final_ncu_cond_list.log 2>&1
echo "ncu exit $?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"cond_pos_enc" -c 1 -o gpurun_out/r01e_cond_final \
    python scripts/bench_cond.py --iters 1 --no-cpu > gpurun_out/final_ncu_cond_full.log 2>&1
echo "ncu exit $?"

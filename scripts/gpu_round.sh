#!/bin/bash
# Full GPU round: the driver's own commands (pytest -m gpu, smoke), the per-class profile, an ncu
# launch list of one production forward, and a short bench run.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
set -o pipefail
echo "=== pytest" ; timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5
echo "=== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== profile" ; timeout 600 python scripts/profile_unet.py > gpurun_out/profile_unet.json 2> gpurun_out/profile_unet.err ; cat gpurun_out/profile_unet.json
echo "=== bench" ; timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err ; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "=== ncu launch list"
timeout 300 python scripts/profile_unet.py --ncu > gpurun_out/ncu_plain.log 2>&1 && \
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python scripts/profile_unet.py --ncu > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/launches.csv

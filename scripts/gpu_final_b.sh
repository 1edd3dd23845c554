#!/bin/bash
# Round-2b measurement call (after the row-quad attention and the coalescing residual epilogue): the driver's own
# commands (pytest -m gpu, smoke), bench (1 GPU), reference arm, ncu traffic pass over bench.py, attention
# micro-benchmark against SDPA, `ncu --set full` of one launch per kernel class.  Outputs under gpurun_out/r02b_*.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/r02b_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r02b_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02b_final_bench_1gpu.json 2> gpurun_out/r02b_final_bench_1gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r02b_final_bench_1gpu.json') if l.startswith('{')][-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'unet_ms', d['unet_step_ms'])
print({k: (round(v['ms_per_call'], 2), round(v.get('frac_of_bf16_peak', v.get('frac_of_hbm_peak', 0)), 3)) for k, v in d['kernels'].items()})
print('roofline', d['roofline']['frac'], d['roofline']['executed_frac'], d['roofline']['traffic'], d['roofline']['traffic_source'])
print('clocks', d['clocks'])
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/r02b_final_bench_ref.json 2> gpurun_out/r02b_final_bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/r02b_final_bench_ref.json
timeout 1200 scripts/ncu_bench_traffic.sh
timeout 300 python scripts/bench_attention.py > gpurun_out/r02b_attention_microbench.json 2> gpurun_out/r02b_attention_microbench.err; echo "attn micro rc=$?"
timeout 900 scripts/ncu_full.sh
timeout 600 python bench.py --workload multi_ref --n-gen 80 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/r02b_bench_1gpu_multi_ref_80.json 2> gpurun_out/r02b_bench_1gpu_multi_ref_80.err; echo "multi_ref rc=$?"; cut -c1-160 gpurun_out/r02b_bench_1gpu_multi_ref_80.json
CAP4D_GEMM_BIAS_SMEM=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/r02b_bench_bias_global.json 2>/dev/null; echo "bias-from-global A/B rc=$?"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/r02b_bench_bias_smem.json 2>/dev/null
python - <<'PY'
import json
for n in ("bias_global", "bias_smem"):
    d = json.loads([l for l in open(f"gpurun_out/r02b_bench_{n}.json") if l.startswith("{")][-1])
    print(n, round(d["value"], 4), round(d["unet_step_ms"], 2), d["clocks"]["sm_mhz"], {k: round(v["ms_per_call"], 2) for k, v in d["kernels"].items()})
PY

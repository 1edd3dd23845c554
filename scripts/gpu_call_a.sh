#!/bin/bash
# round 2, measurement call A: bench (1 GPU, graphs), reference arm smoke, ncu traffic pass over bench.py
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r02_bench_1gpu.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_1gpu.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'h2d/step',d['e2e']['h2d_bytes_per_step'],'unet_ms',d['unet_step_ms'])
print({k:(round(v['ms_per_call'],2), round(v.get('frac_of_bf16_peak',v.get('frac_of_hbm_peak',0)),3)) for k,v in d['kernels'].items()})
print('roofline',d['roofline']['frac'],d['roofline']['executed_frac'],d['roofline']['traffic_source'])
print('graphs',d['cuda_graphs'],'clocks',d['clocks'])
print('eager',json.dumps(d.get('gpu_eager_baseline')))
print('cpu',json.dumps(d.get('cpu_baseline')))
"
python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/r02_bench_ref.json
scripts/ncu_bench_traffic.sh

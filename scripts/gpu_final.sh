#!/bin/bash
# Round-end measurement call: smoke(), bench (1 GPU), reference arm, ncu traffic pass over bench.py, attention and VAE
# micro-benchmarks.  Outputs under gpurun_out/r02_final_*.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r02_final_smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_final_bench_1gpu.json 2> gpurun_out/r02_final_bench_1gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r02_final_bench_1gpu.json') if l.startswith('{')][-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'h2d/step', d['e2e']['h2d_bytes_per_step'], 'unet_ms', d['unet_step_ms'])
print({k: (round(v['ms_per_call'], 2), round(v.get('frac_of_bf16_peak', v.get('frac_of_hbm_peak', 0)), 3)) for k, v in d['kernels'].items()})
print('roofline', d['roofline']['frac'], d['roofline']['executed_frac'], d['roofline']['traffic'], d['roofline']['traffic_source'])
print('graphs', d['cuda_graphs'], 'clocks', d['clocks'])
print('eager', json.dumps(d.get('gpu_eager_baseline')))
print('cpu', json.dumps(d.get('cpu_baseline')))
PY
python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/r02_final_bench_ref.json 2> gpurun_out/r02_final_bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r02_final_bench_ref.json
scripts/ncu_bench_traffic.sh
python scripts/bench_attention.py > gpurun_out/r02_final_attention_microbench.json 2> gpurun_out/r02_final_attention_microbench.err; echo "attn micro rc=$?"
python scripts/bench_vae.py > gpurun_out/r02_final_vae_decode.json 2> gpurun_out/r02_final_vae_decode.err; echo "vae dec rc=$?"; tail -c 400 gpurun_out/r02_final_vae_decode.json
python scripts/bench_vae.py --encode 8 > gpurun_out/r02_final_vae_encode.json 2> gpurun_out/r02_final_vae_encode.err; echo "vae enc rc=$?"; tail -c 300 gpurun_out/r02_final_vae_encode.json

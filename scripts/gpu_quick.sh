#!/bin/bash
# kernel tests + per-class profile (quick iteration loop)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
[ -x scripts/microbench/pipes ] && [ -n "$RUN_MICRO" ] && scripts/microbench/pipes
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4
timeout 600 python scripts/profile_unet.py "$@" > gpurun_out/profile_unet.json 2> gpurun_out/profile_unet.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/profile_unet.json'))
print("forward ms", round(d["forward_ms_back_to_back"],3), "TF/s", round(d["achieved_tflops_back_to_back"],1), "launches", d["launches"])
for k,v in d["classes"].items():
    print(f"  {k:10s} {v['ms']:7.3f} ms", f"{v.get('tflops') or 0:7.1f} TF/s" if 'tflops' in v else "", f"{v.get('gbs') or 0:7.1f} GB/s" if 'gbs' in v else "")
PY
tail -3 gpurun_out/profile_unet.err

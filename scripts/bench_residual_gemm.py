#!/usr/bin/env python
"""Residual-epilogue GEMMs of the transformer blocks (to_out / proj_out: K = C, fp32 out; FF2: K = 4C, bf16 out) at the
production shapes of a 5-group call: time, algorithmic bytes per second and FLOP/s.
    python scripts/bench_residual_gemm.py [lib.so ...]      (no argument: the product library)
    FORCES="1,160;2,160" python scripts/bench_residual_gemm.py   (also: each tile configuration forced, CAP4D_GEMM_FORCE)"""
import json
import math
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = [  # (M, C): 10 x 8 views x 64^2, 32^2, 16^2 tokens; generated views only (10 x 7) at level 0 of the up path
    (327680, 320), (286720, 320), (81920, 640), (20480, 1280)]


def run_one():
    import torch
    from cap4d_b200 import ops
    dev = torch.device("cuda:0")
    rows = []
    for M, C in SHAPES:
        for K, mode, name, with_res in ((C, ops.OUT_F32, "to_out/proj_out", True), (4 * C, ops.OUT_BF16, "ff2", True),
                                        (C, ops.OUT_F32, "proj_in", False), (C, ops.OUT_BF16, "qkv", False),
                                        (C, ops.OUT_GEGLU, "ff1_geglu", False)):
            N = 3 * C if name == "qkv" else (8 * C if name == "ff1_geglu" else C)
            g = torch.Generator().manual_seed(1)
            a = torch.randn(M, K, device=dev).to(torch.bfloat16)
            w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(torch.bfloat16)
            bias = torch.randn(N, device=dev) if name != "qkv" else None
            res = torch.randn(M, N, device=dev) if with_res else None
            _, ms = ops.gemm(a, w, bias=bias, residual=res, out_mode=mode, time_iters=10)
            n_out = N // 2 if mode == ops.OUT_GEGLU else N
            nbytes = M * K * 2 + (M * N * 4 if with_res else 0) + M * n_out * (4 if mode == ops.OUT_F32 else 2)
            rows.append({"op": name, "M": M, "N": N, "K": K, "ms": round(ms, 4), "GBps": round(nbytes / ms / 1e6, 1),
                         "TFLOPs": round(2.0 * M * N * K / ms / 1e9, 1)})
    return rows


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        print(json.dumps(run_one()))
        sys.exit(0)
    libs = sys.argv[1:] or [None]
    forces = [None] + [f for f in os.environ.get("FORCES", "").split(";") if f]
    epis = [None] + [e for e in os.environ.get("EPIS", "").split(";") if e]   # CAP4D_GEMM_EPI values to force as well
    extra_env = [kv for kv in os.environ.get("ENVS", "").split(";") if kv]      # e.g. ENVS="CAP4D_GEMM_BIAS_SMEM=0"
    prod = os.path.join(ROOT, "cap4d_b200", "libcap4d_b200.so")
    keep = prod + ".keep"
    shutil.copy(prod, keep)
    try:
        for lib in libs:
            if lib:
                shutil.copy(lib, prod)
            for force in forces + [("epi", e) for e in epis[1:]] + [("env", kv) for kv in extra_env]:
                env = dict(os.environ)
                if isinstance(force, tuple) and force[0] == "env":
                    k_, v_ = force[1].split("=", 1)
                    env[k_] = v_
                    force = force[1]
                elif isinstance(force, tuple):
                    env["CAP4D_GEMM_EPI"] = force[1]
                    force = "CAP4D_GEMM_EPI=" + force[1]
                elif force:
                    env["CAP4D_GEMM_FORCE"] = force
                out = subprocess.run([sys.executable, __file__, "--child"], capture_output=True, text=True, env=env)
                print("==", lib or "product library", (force if "=" in force else "CAP4D_GEMM_FORCE=" + force) if force else "")
                if out.returncode != 0:
                    print(out.stderr[-2000:])
                    continue
                for r in json.loads(out.stdout.strip().splitlines()[-1]):
                    print(json.dumps(r))
    finally:
        shutil.move(keep, prod)

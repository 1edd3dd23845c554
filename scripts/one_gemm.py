#!/usr/bin/env python
"""One GEMM shape through the C ABI (for ncu captures):  python scripts/one_gemm.py M N K mode[0 f32|1 bf16|2 geglu] res[0|1] iters"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cap4d_b200 import ops  # noqa: E402

M, N, K, mode, res, iters = (int(v) for v in sys.argv[1:7])
dev = torch.device("cuda:0")
a = torch.randn(M, K, device=dev).to(torch.bfloat16)
w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
bias = torch.randn(N, device=dev)
r = torch.randn(M, N, device=dev) if res else None
_, ms = ops.gemm(a, w, bias=bias, residual=r, out_mode=mode, time_iters=iters)
print(f"M={M} N={N} K={K} mode={mode} res={res}: {ms * 1e3:.1f} us  {2.0 * M * N * K / ms / 1e9:.1f} TF/s  "
      f"force={os.environ.get('CAP4D_GEMM_FORCE', '-')}")

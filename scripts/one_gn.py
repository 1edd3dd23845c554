#!/usr/bin/env python
"""One GroupNorm(+SiLU) launch through the C ABI (for ncu captures):  python scripts/one_gn.py n_img hw C1 C2 iters"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cap4d_b200 import ops  # noqa: E402

n, hw, C1, C2, iters = (int(v) for v in sys.argv[1:6])
dev = torch.device("cuda:0")
x1 = torch.randn(n * hw, C1, device=dev)
x2 = torch.randn(n * hw, C2, device=dev) if C2 else None
g, b = torch.ones(C1 + C2, device=dev), torch.zeros(C1 + C2, device=dev)
r = ops.groupnorm(x1, x2, n, hw, g, b, 1e-5, True, want_raw=False, time_iters=iters)
nbytes = n * hw * (C1 + C2) * 6
print(f"groupnorm n={n} hw={hw} C={C1}+{C2}: {r[-1] * 1e3:.1f} us  {nbytes / r[-1] / 1e6:.0f} GB/s algorithmic")

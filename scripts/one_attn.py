#!/usr/bin/env python
"""One attention shape through the C ABI (for ncu captures):  python scripts/one_attn.py L C n_seq iters"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cap4d_b200 import ops  # noqa: E402

L, C, n_seq, iters = (int(v) for v in sys.argv[1:5])
dev = torch.device("cuda:0")
qkv = torch.randn(n_seq * L, 3 * C, device=dev).to(torch.bfloat16)
_, ms = ops.attention(qkv, C, L, time_iters=iters)
print(f"L={L} C={C} n_seq={n_seq}: {ms:.3f} ms {4.0 * n_seq * (C // 64) * L * L * 64 / ms / 1e9:.1f} TFLOP/s")

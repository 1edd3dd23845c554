#!/usr/bin/env python
"""Cycle timeline of one attention CTA (instrumented kernel build): where do the softmax
warpgroups and the MMA thread spend their time?   python scripts/attn_trace.py [L] [C] [n_seq]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cap4d_b200 import _lib, ops  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
C = int(sys.argv[2]) if len(sys.argv) > 2 else 320
n_seq = int(sys.argv[3]) if len(sys.argv) > 3 else 16
dev = torch.device("cuda:0")
qkv = torch.randn(n_seq * L, 3 * C, device=dev).to(torch.bfloat16)
out = torch.empty(n_seq * L, C, device=dev, dtype=torch.bfloat16)
trace = torch.zeros(3, 16, 8, device=dev, dtype=torch.int64)
lib = _lib.load()
for _ in range(2):
    _lib.check(lib.cap4d_b200_attention_trace(qkv.data_ptr(), out.data_ptr(), n_seq * L, C, L, 0.125, None,
                                              trace.data_ptr()), "trace")
torch.cuda.synchronize()
t = trace.cpu()
t0 = int(t[0, 0, 0])
names = ["start", "S ready", "S loaded", "max+pv", "exp done", "P stored", "arrived", "-"]
for wgi in range(2):
    print(f"softmax WG{wgi + 1}: per-tile stamps relative to the tile's start (cycles); first column = start since kernel t0")
    for j in range(1, 10):
        row = t[wgi, j]
        print(f"  j={j:2d} @{int(row[0]) - t0:7d} " + " ".join(f"{names[k]}:{int(row[k] - row[0]):5d}" for k in (1, 2, 3, 4, 5, 6)))
print("MMA thread of tile A: loop start, after QK(j+1) issue, after PV(j) issue (since t0)")
for j in range(1, 10):
    row = t[2, j]
    print(f"  j={j:2d} " + " ".join(f"{int(row[k]) - t0:7d}" for k in range(3)))
_, ms = ops.attention(qkv, C, L, time_iters=5)
flops = 4.0 * n_seq * (C // 64) * L * L * 64
print(f"L={L} C={C} n_seq={n_seq}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s")

#!/usr/bin/env python
"""Per-shape throughput of the tensor-core kernels at the production U-Net's shapes (B=2, V=8, 64x64)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cap4d_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
N_IMG = int(os.environ.get("N_IMG", "16"))


def conv(hw, cin, cout, count, iters=10):
    a = torch.randn(N_IMG, hw, hw, cin, device=dev).to(torch.bfloat16)
    w = (torch.randn(cout, 9 * cin, device=dev) / (3 * cin ** 0.5)).to(torch.bfloat16)
    bias = torch.randn(cout, device=dev)
    res = torch.randn(N_IMG * hw * hw, cout, device=dev)
    _, ms = ops.conv3x3(a, w, N_IMG, hw, hw, bias=bias, residual=res, time_iters=iters)
    fl = 2.0 * N_IMG * hw * hw * cout * 9 * cin
    print(f"conv  {hw:2d}x{hw:<2d} M={N_IMG*hw*hw:6d} N={cout:5d} K={9*cin:6d} x{count:2d}: {ms*1e3:8.1f} us {fl/ms/1e9:7.1f} TF/s  total {ms*count:6.3f} ms")
    return ms * count


def gemm(M, N, K, count, mode=0, res=False, iters=10):
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    r = torch.randn(M, N, device=dev) if res else None
    _, ms = ops.gemm(a, w, bias=bias, residual=r, out_mode=mode, time_iters=iters)
    fl = 2.0 * M * N * K
    tag = {0: "f32", 1: "bf16", 2: "geglu"}[mode]
    print(f"gemm {tag:5s} M={M:6d} N={N:5d} K={K:6d} x{count:2d}: {ms*1e3:8.1f} us {fl/ms/1e9:7.1f} TF/s  total {ms*count:6.3f} ms")
    return ms * count


tot = 0.0
print("== conv3x3 (ResBlocks / resample) ==")
for hw, cin, cout, cnt in [(64, 320, 320, 7), (64, 640, 320, 2), (64, 960, 320, 1), (64, 640, 640, 1),
                           (32, 640, 640, 6), (32, 320, 640, 1), (32, 960, 640, 1), (32, 1280, 640, 1), (32, 1920, 640, 1),
                           (32, 1280, 1280, 1),
                           (16, 1280, 1280, 7), (16, 2560, 1280, 2), (16, 640, 1280, 1), (16, 1920, 1280, 1),
                           (8, 1280, 1280, 11), (8, 2560, 1280, 3)]:
    tot += conv(hw, cin, cout, cnt)
print(f"conv total {tot:.3f} ms")
tot2 = 0.0
print("== linear (transformers) ==")
for hw, C, cnt in [(64, 320, 5), (32, 640, 5), (16, 1280, 5), (8, 1280, 1)]:
    M = N_IMG * hw * hw
    tot2 += gemm(M, C, C, cnt, 0)             # proj_in (fp32 out)
    tot2 += gemm(M, 3 * C, C, cnt, 1)         # qkv (bf16 out)
    tot2 += gemm(M, C, C, cnt, 0, res=True)   # to_out + residual
    tot2 += gemm(M, 8 * C, C, cnt, 2)         # ff1 geglu
    tot2 += gemm(M, C, 4 * C, cnt, 1, res=True)  # ff2 (+res, bf16 out)
    tot2 += gemm(M, C, C, cnt, 0, res=True)   # proj_out + residual
print(f"linear total {tot2:.3f} ms")

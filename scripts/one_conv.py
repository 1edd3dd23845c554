#!/usr/bin/env python
"""One 3x3 convolution through the C ABI (for ncu captures):  python scripts/one_conv.py n_img H W Cin Cout iters [f16]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cap4d_b200 import ops  # noqa: E402

n, H, W, Cin, Cout, iters = (int(v) for v in sys.argv[1:7])
dev = torch.device("cuda:0")
a = torch.randn(n, H, W, Cin, device=dev).to(torch.bfloat16)
w = (torch.randn(Cout, 9 * Cin, device=dev) / (9 * Cin) ** 0.5).to(torch.bfloat16)
bias = torch.randn(Cout, device=dev)
rowbias = torch.randn(n, Cout, device=dev)
_, ms = ops.conv3x3(a, w, n, H, W, bias=bias, rowbias=rowbias, time_iters=iters)
fl = 2.0 * n * H * W * Cout * 9 * Cin
print(f"conv3x3 n={n} {H}x{W} {Cin}->{Cout}: {ms * 1e3:.1f} us  {fl / ms / 1e9:.1f} TFLOP/s")

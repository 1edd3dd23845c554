"""GroupNorm / LayerNorm per-shape timings (the shapes the production U-Net launches)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cap4d_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 16
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
shapes = [(4096, 320, 0), (4096, 320, 320), (4096, 640, 320), (1024, 320, 0), (1024, 640, 0), (1024, 640, 640),
          (1024, 1280, 640), (1024, 640, 320), (256, 640, 0), (256, 1280, 0), (256, 1280, 1280), (256, 1280, 640),
          (64, 1280, 0), (64, 1280, 1280)]
for hw, C1, C2 in shapes:
    C = C1 + C2
    x1 = torch.randn(n_img * hw, C1, device=dev)
    x2 = torch.randn(n_img * hw, C2, device=dev) if C2 else None
    g = torch.ones(C, device=dev)
    b = torch.zeros(C, device=dev)
    r = ops.groupnorm(x1, x2, n_img, hw, g, b, 1e-5, True, want_raw=False, time_iters=iters)
    ms = r[-1]
    nbytes = n_img * hw * C * 6
    print(f"GN n={n_img} hw={hw} C={C1}+{C2}: {ms*1e3:7.1f} us  {nbytes/ms/1e6:7.1f} GB/s (algorithmic, {nbytes/1e6:.0f} MB)")
for M, C in [(4096 * n_img, 320), (1024 * n_img, 640), (256 * n_img, 1280)]:
    x = torch.randn(M, C, device=dev)
    g = torch.ones(C, device=dev)
    b = torch.zeros(C, device=dev)
    _, ms = ops.layernorm(x, g, b, 1e-5, time_iters=iters)
    print(f"LN M={M} C={C}: {ms*1e3:7.1f} us {M*C*6/ms/1e6:7.1f} GB/s")

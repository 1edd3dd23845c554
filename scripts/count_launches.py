#!/usr/bin/env python
"""Kernel launches bench.py makes before its timed region at the default workload (for ncu -s): warm-up steps x
U-Net calls per step x (launches per forward + gather + update).  Needs a GPU (asks the plan for its launch count)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cap4d_b200 import B200MMDMUnet  # noqa: E402
from cap4d_b200.config import MMDM_UNET_CONFIG  # noqa: E402

gpc, groups, warm = 5, 120, 3
dev = torch.device("cuda:0")
unet = B200MMDMUnet.random_init(MMDM_UNET_CONFIG, seed=0, device=dev)
B, V, H = 2 * gpc, 8, 64
x = torch.zeros(B, V, 4, H, H, device=dev)
ctrl = dict(z_input=torch.zeros(B, V, 4, H, H, device=dev), ref_mask=torch.zeros(B, V, 1, H, H, device=dev),
            pos_enc=torch.zeros(B, V, H, H, 50, device=dev))
ctrl["ref_mask"][:, :1] = 1.0
unet(x, timesteps=torch.zeros(B, V, dtype=torch.long, device=dev), control=ctrl, n_ref_views=1)
per_call = unet.num_launches() + 2
# random_init itself launches generator / pack kernels: those come before everything and are counted by ncu too, but
# bench.py makes exactly the same ones - count them by name instead of guessing: the caller adds nothing
print(warm * (groups // gpc) * per_call)

#!/usr/bin/env python
"""How much of the bf16 error budget is WEIGHT rounding?  The fp32 oracle is run on the GPU (TF32 off) with exact
weights, with weights rounded to bf16 and with weights rounded to fp16 (activations exact in all three): the
max-rel error of the noise prediction isolates the weights' share of the product path's total (tests/parity_budget.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mmdm_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
cfg = O.PRODUCTION_CONFIG
for seed, ts in ((1, 501), (0, 501)):
    sd = {k: v.to(dev) for k, v in O.init_state_dict(cfg, seed=seed).items()}
    x, t, ctrl = O.make_inputs(cfg, B=2, V=8, H=64, W=64, R=1, seed=100 * seed + ts, timestep=ts)
    x, t = x.to(dev), t.to(dev)
    ctrl = {k: v.to(dev) for k, v in ctrl.items()}
    is_mm = lambda k, v: v.dim() >= 2  # noqa: E731  (conv / linear weights; norms and biases stay fp32 in the product too)
    with torch.no_grad():
        ref = O.unet_forward(sd, cfg, x, t, ctrl)
        for name, dt in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
            sdr = {k: (v.to(dt).float() if is_mm(k, v) else v) for k, v in sd.items()}
            y = O.unet_forward(sdr, cfg, x, t, ctrl)
            print(f"seed {seed} t={ts}: weights rounded to {name}: eps max-rel {O.max_rel_err(y[:, 1:], ref[:, 1:]):.3e}")
    del sd
    torch.cuda.empty_cache()

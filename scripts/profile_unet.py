#!/usr/bin/env python
"""One production-shape U-Net forward (cap4d_mmdm_final, B=2 x V=8 x 64x64) for profiling.

    python scripts/profile_unet.py [--groups G] [--warm 3]                       # per-class CUDA-event timings
    ncu --profile-from-start off ... python scripts/profile_unet.py --ncu        # only the marked forward is profiled
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cap4d_b200 import B200MMDMUnet  # noqa: E402
from cap4d_b200.config import MMDM_UNET_CONFIG  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--groups", type=int, default=1)
    ap.add_argument("--warm", type=int, default=3)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--ncu", action="store_true")
    ap.add_argument("--hw", type=int, default=64)
    ap.add_argument("--views", type=int, default=8, help="views per group (the 3d attention spans views*h*w tokens)")
    ap.add_argument("--ref-views", type=int, default=1,
                    help="n_ref_views promise passed with every call, as the sampler does (0 = compute every view)")
    args = ap.parse_args()
    R = args.ref_views
    dev = torch.device("cuda:0")
    B, V, H = 2 * args.groups, args.views, args.hw
    unet = B200MMDMUnet.random_init(dict(MMDM_UNET_CONFIG, time_steps=V), seed=0, device=dev)
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(B, V, 4, H, H, generator=g, device=dev)
    ctrl = dict(z_input=torch.randn(B, V, 4, H, H, generator=g, device=dev),
                ref_mask=torch.zeros(B, V, 1, H, H, device=dev),
                pos_enc=torch.randn(B, V, H, H, 50, generator=g, device=dev))
    ctrl["ref_mask"][:, :1] = 1.0
    t = torch.full((B, V), 501, device=dev, dtype=torch.long)
    for _ in range(args.warm):
        unet(x, timesteps=t, control=ctrl, n_ref_views=R)
    torch.cuda.synchronize()
    if args.ncu:
        torch.cuda.profiler.start()
        unet(x, timesteps=t, control=ctrl, n_ref_views=R)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        unet(x, timesteps=t, control=ctrl, n_ref_views=R)
    e1.record()
    torch.cuda.synchronize()
    total_ms = e0.elapsed_time(e1) / args.iters
    acc = None
    for _ in range(args.iters):
        _, ms = unet.forward_timed(x, t, ctrl, n_ref_views=R)
        acc = ms if acc is None else {k: acc[k] + ms[k] for k in ms}
    ms = {k: v / args.iters for k, v in acc.items()}
    stats = unet.class_stats()
    out = {"shape": [B, V, H, H], "n_ref_views": R, "forward_ms_back_to_back": total_ms, "forward_ms_sum_of_launches": sum(ms.values()),
           "launches": unet.num_launches(), "classes": {}}
    flops_total = sum(s["flops"] for s in stats.values())
    out["tflops_total"] = flops_total / 1e12
    out["achieved_tflops_back_to_back"] = flops_total / (total_ms * 1e-3) / 1e12
    for c in ms:
        d = {"ms": ms[c], "launches": stats[c]["launches"]}
        if stats[c]["flops"] > 0:
            d["tflop"] = stats[c]["flops"] / 1e12
            d["tflops"] = stats[c]["flops"] / (ms[c] * 1e-3) / 1e12 if ms[c] > 0 else None
        if stats[c]["bytes"] > 0:
            d["gb"] = stats[c]["bytes"] / 1e9
            d["gbs"] = stats[c]["bytes"] / (ms[c] * 1e-3) / 1e9 if ms[c] > 0 else None
        out["classes"][c] = d
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""VAE decode throughput: production first_stage_config, 64x64 latents -> 512x512 images.
Prints one JSON line: images/s with latents resident in HBM, end to end from / to pinned host memory, the
executed TFLOP/s, and the reference's CPU path (oracle port) on the host cores.
    python scripts/bench_vae.py [--images 64] [--batch 8] [--no-cpu]"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cap4d_b200 import B200VAEDecoder  # noqa: E402
from cap4d_b200.vae import VAE_CONFIG  # noqa: E402


def decoder_flops(cfg, H, W):
    """2*MAC of Decoder.forward for one latent of H x W (un-folded convs, as the reference executes them)."""
    ch, mult, nrb = cfg["ch"], cfg["ch_mult"], cfg["num_res_blocks"]
    c = ch * mult[-1]
    hw = H * W
    fl = 2.0 * hw * c * 9 * cfg["z_channels"]
    fl += 2 * (2 * 2.0 * hw * c * 9 * c)                                  # mid block_1, block_2
    fl += 4 * 2.0 * hw * c * c + 2 * 2.0 * hw * hw * c                    # q, k, v, proj_out; scores and P.V
    cin = c
    for lvl in reversed(range(len(mult))):
        cout = ch * mult[lvl]
        for _ in range(nrb + 1):
            fl += 2.0 * hw * cout * 9 * cin + 2.0 * hw * cout * 9 * cout + (2.0 * hw * cout * cin if cin != cout else 0)
            cin = cout
        if lvl != 0:
            hw *= 4
            fl += 2.0 * hw * cin * 9 * cin
    return fl + 2.0 * hw * cfg["out_ch"] * 9 * cin


def bench_encode(args):
    """AutoencoderKL.encode (reference images -> latents, once per reference view): images/s on the device and the
    CPU path (oracle port) on one image."""
    dev = torch.device("cuda:0")
    vae = B200VAEDecoder.random_init(VAE_CONFIG, seed=0, device=dev, with_encoder=True)
    n = args.encode
    x = torch.tanh(torch.randn(n, 3, 512, 512, generator=torch.Generator().manual_seed(1))).to(dev)
    for _ in range(2):
        vae.encode_moments(x[:2], batch=2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    m = vae.encode_moments(x, batch=2)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    res = {"metric": "vae_encode_images_per_sec", "value": n / (ms * 1e-3), "unit": "images/s (512x512 -> 64x64 moments)",
           "images": n, "batch": 2, "ms_per_image": ms / n, "moments_absmax": float(m.abs().max()), "dtype": "bf16",
           "data": "synthetic"}
    if not args.no_cpu:
        from oracle import vae_oracle as VO  # cpu_baseline leg only

        sd = VO.init_vae_state_dict(VO.PRODUCTION_VAE, seed=0)
        sd.update(VO.init_vae_encoder_state_dict(VO.PRODUCTION_VAE, seed=0))
        t0 = time.perf_counter()
        VO.vae_encode_moments(sd, VO.PRODUCTION_VAE, x[:1].cpu())
        dt = time.perf_counter() - t0
        res["cpu_baseline"] = {"value": 1.0 / dt, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                               "sample": f"one 512x512 image through oracle/vae_oracle.py: {dt:.1f} s"}
    print(json.dumps(res))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=64)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--encode", type=int, default=0, help="also time AutoencoderKL.encode on this many 512x512 images")
    args = ap.parse_args()
    if args.encode:
        return bench_encode(args)
    dev = torch.device("cuda:0")
    vae = B200VAEDecoder.random_init(VAE_CONFIG, seed=0, device=dev)
    n, b = args.images, args.batch
    z_host = (torch.randn(n, 4, 64, 64, generator=torch.Generator().manual_seed(1)) * 0.8).pin_memory()
    z = z_host.to(dev)
    for _ in range(2):
        vae.decode_first_stage(z[:b], batch=b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y = vae.decode_first_stage(z, batch=b)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out_host = torch.empty((n, 3, 512, 512), dtype=torch.float32).pin_memory()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out_host.copy_(vae.decode_first_stage(z_host.to(dev, non_blocking=True), batch=b), non_blocking=True)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    fl = decoder_flops(VAE_CONFIG, 64, 64)
    line = {"metric": "vae_decoded_images_per_sec", "value": n / (ms * 1e-3), "unit": "images/s (512^2)", "n_gpus": 1,
            "images": n, "batch": b, "ms_per_image": ms / n, "launches_per_batch": vae.num_launches(),
            "tflop_per_image": fl / 1e12, "tflops": fl * n / (ms * 1e-3) / 1e12, "dtype": "bf16", "data": "synthetic",
            "e2e": {"value": n / e2e_s, "unit": "images/s", "h2d_bytes": z_host.numel() * 4, "d2h_bytes": out_host.numel() * 4},
            "checksum": float(y.abs().mean())}
    if not args.no_cpu:
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from oracle import vae_oracle as VO  # the reference's CPU path restated (bench baseline only)

        torch.set_num_threads(os.cpu_count() or 1)
        sd = VO.init_vae_state_dict(VO.PRODUCTION_VAE, seed=0)
        zc = z_host[:1].clone()
        VO.vae_decode(sd, VO.PRODUCTION_VAE, zc)
        t0 = time.perf_counter()
        VO.vae_decode(sd, VO.PRODUCTION_VAE, zc)
        t = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": 1.0 / t, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"one 64x64 latent decoded by the oracle on the host cores: {t:.2f} s"}
    print(json.dumps(line))


if __name__ == "__main__":
    main()

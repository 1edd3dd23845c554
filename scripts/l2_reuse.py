#!/usr/bin/env python
"""Does a consumer kernel find its producer's output in L2?  (decides the depth-first level-0 schedule)

For n_img in {2, 4, 8, 16, 80} images at the 64x64 level (C = 320, M = n_img * 4096 tokens) the pair the transformer runs
back to back - `to_out` (GEMM K = 320 + bias + fp32 residual -> fp32) followed by LayerNorm (fp32 -> bf16) - is timed
with CUDA events in three ways:

    cold   LayerNorm alone after a 512 MB buffer was written (L2 holds nothing of its input)
    warm   LayerNorm alone, repeated on the same input (upper bound of what L2 residency can give)
    chain  LayerNorm directly after the GEMM that produces its input (what a depth-first schedule would see)

and the GEMM is timed with its residual cold and right after a kernel that wrote it.  One JSON line per size.

    python scripts/l2_reuse.py [--iters 10] > gpurun_out/l2_reuse.jsonl
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cap4d_b200 import ops  # noqa: E402


def timed(fn, before=None, iters=10):
    """median CUDA-event time of fn() in ms; before() runs untimed ahead of every repetition"""
    ts = []
    for _ in range(iters):
        if before is not None:
            before()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--channels", type=int, default=320)
    ap.add_argument("--hw", type=int, default=4096)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    C, hw = args.channels, args.hw
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def flush():
        flush_buf.fill_(1)

    g = torch.Generator(device=dev).manual_seed(0)
    w = (torch.randn(C, C, generator=g, device=dev) * 0.05).to(torch.bfloat16)
    bias = torch.randn(C, generator=g, device=dev)
    gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    for n_img in (2, 4, 8, 16, 80):
        M = n_img * hw
        a = torch.randn(M, C, generator=g, device=dev).to(torch.bfloat16)
        res = torch.randn(M, C, generator=g, device=dev)
        x = ops.gemm(a, w, bias=bias, residual=res, out_mode=ops.OUT_F32)
        ln_bytes, gemm_bytes = M * C * 6, M * C * 10 + 2 * C * C
        for _ in range(3):
            ops.layernorm(x, gamma, beta)
        ln = lambda: ops.layernorm(x, gamma, beta)  # noqa: E731
        gemm = lambda: ops.gemm(a, w, bias=bias, residual=res, out_mode=ops.OUT_F32)  # noqa: E731
        t_cold = timed(ln, before=flush, iters=args.iters)
        t_warm = timed(ln, before=ln, iters=args.iters)
        # ops.gemm allocates a new output on every call: the chained LayerNorm must read that fresh buffer
        holder = {}

        def produce():
            flush()
            holder["x"] = ops.gemm(a, w, bias=bias, residual=res, out_mode=ops.OUT_F32)

        t_chain = timed(lambda: ops.layernorm(holder["x"], gamma, beta), before=produce, iters=args.iters)
        g_cold = timed(gemm, before=flush, iters=args.iters)
        g_after_writer = timed(gemm, before=lambda: (flush(), res.mul_(1.0)), iters=args.iters)
        print(json.dumps({
            "n_img": n_img, "M": M, "C": C, "tensor_MB_fp32": M * C * 4 / 1e6,
            "layernorm_us": {"cold": t_cold * 1e3, "warm": t_warm * 1e3, "after_producer": t_chain * 1e3},
            "layernorm_GBs": {"cold": ln_bytes / t_cold / 1e6, "warm": ln_bytes / t_warm / 1e6,
                              "after_producer": ln_bytes / t_chain / 1e6},
            "to_out_gemm_us": {"cold": g_cold * 1e3, "residual_just_written": g_after_writer * 1e3},
            "to_out_gemm_GBs": {"cold": gemm_bytes / g_cold / 1e6, "residual_just_written": gemm_bytes / g_after_writer / 1e6},
        }), flush=True)
        del a, res, x, holder


if __name__ == "__main__":
    main()

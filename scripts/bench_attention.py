#!/usr/bin/env python
"""Cross-view attention micro-benchmark: non-causal, no mask, head dim 64, scale 1/8, bf16
operands / fp32 accumulation.  Model-true shapes of the production U-Net (V = 8) plus a sweep over the views per
group at the 32x32 level ("3d" attention over V*h*w tokens).  Baselines on the same GPU: torch SDPA (library flash
attention, bf16) and the reference's legacy_attention arithmetic (cap4d/mmdm/net/attention.py:112-132: fp32
einsum / softmax / einsum with the score tensor materialised; skipped above 8192 tokens).
    python scripts/bench_attention.py > profiles/rNN_attention_microbench.json"""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cap4d_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def timed(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def case(name, n_seq, L, C):
    heads = C // 64
    qkv = torch.randn(n_seq * L, 3 * C, device=dev).to(torch.bfloat16)
    flops = 4.0 * n_seq * heads * L * L * 64
    out, ms = ops.attention(qkv, C, L, time_iters=5)
    q, k, v = (t.reshape(n_seq, L, heads, 64).permute(0, 2, 1, 3) for t in qkv.split(C, dim=1))
    ref = F.scaled_dot_product_attention(q, k, v, scale=0.125)
    err = float((out.reshape(n_seq, L, heads, 64).permute(0, 2, 1, 3).float() - ref.float()).abs().max())
    ms_sdpa = timed(lambda: F.scaled_dot_product_attention(q, k, v, scale=0.125))
    row = {"case": name, "seq_heads": n_seq * heads, "tokens": L, "ms": ms, "tflops": flops / ms / 1e9,
           "sdpa_ms": ms_sdpa, "sdpa_tflops": flops / ms_sdpa / 1e9, "speedup_vs_sdpa": ms_sdpa / ms,
           "max_abs_diff_vs_sdpa": err}
    if L <= 8192:
        qf, kf, vf = q.float(), k.float(), v.float()

        def legacy():  # attention.py:112-132
            s = torch.einsum("bhid,bhjd->bhij", qf, kf) * 0.125
            return torch.einsum("bhij,bhjd->bhid", s.softmax(dim=-1), vf)

        torch.backends.cuda.matmul.allow_tf32 = False
        ms_leg = timed(legacy, iters=2)
        row.update(legacy_fp32_ms=ms_leg, speedup_vs_legacy=ms_leg / ms)
    return row


rows = [case("L0 spatial (80 seq-heads x 4096)", 16, 4096, 320),
        case("L1 3d V=8 (20 x 8192)", 2, 8192, 640),
        case("L2 3d V=8 (40 x 2048)", 2, 2048, 1280),
        case("mid 3d V=8 (40 x 512)", 2, 512, 1280),
        case("L1 3d V=4 (20 x 4096)", 2, 4096, 640),
        case("L1 3d V=16 (20 x 16384)", 2, 16384, 640),
        case("literal BASELINE shape: 8 views x 64x64 tokens, 5 heads (10 x 32768)", 2, 32768, 320)]
print(json.dumps({"bench": "attention d=64 bf16, B200", "rows": rows}, indent=1))

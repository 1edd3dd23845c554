#!/usr/bin/env python
"""Mixed (tensor | HBM) roofline of the transformer linear layers of one U-Net call - analytic, runs without a GPU.

The bench reports the `linear` class against the TENSOR peak only (0.58).  The executor's data flow
(cap4d_b200/csrc/unet_exec.cu: plan_tf) keeps the residual stream in fp32, so the short-K GEMMs of the 64x64 level
move more bytes than their FLOPs can hide.  Per transformer with M tokens and C channels:

    GEMM       FLOPs        HBM bytes (A in | residual in | out)
    proj_in    2 M C C      2MC |  -  | 4MC (fp32 stream)
    qkv        6 M C C      2MC |  -  | 6MC (bf16)
    to_out     2 M C C      2MC | 4MC | 4MC
    ff1+GEGLU 16 M C C      2MC |  -  | 8MC (bf16, 4C wide)
    ff2        8 M C C      8MC | 4MC | 2MC (bf16: only proj_out reads it)
    proj_out   2 M C C      2MC | 4MC | 4MC
    + weights 12 C C x 2 B once per launch

The bound of a launch is max(FLOPs / tensor peak, bytes / HBM peak); the class bound is their sum (launches are
serialised on one stream).  Peaks: MEASURED_PEAKS.json (sustained bf16, copy bandwidth).  Measured class time: the
committed bench line.

    python scripts/linear_roofline.py [--groups 5] [--ref-views 1] > profiles/r01e_linear_roofline.json
"""
import argparse
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--groups", type=int, default=5, help="groups per U-Net call (bench default)")
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--ref-views", type=int, default=1, help="reference views dropped after the last 3d transformer")
    ap.add_argument("--bench", default=os.path.join(ROOT, "profiles", "r01e_bench_1gpu.json"))
    args = ap.parse_args()
    peaks = {"bf16_tflops_sustained": 1405.3, "hbm_gbs": 6447.2}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks.update(json.load(f))
    except OSError:
        pass
    tf_peak, hbm_peak = peaks["bf16_tflops_sustained"] * 1e12, peaks["hbm_gbs"] * 1e9
    n_img = 2 * args.groups * args.views
    n_gen = 2 * args.groups * (args.views - args.ref_views)
    # (level, C, tokens per image, transformers on all views, transformers on generated views only)
    # mmdm_unet.py:49-55 / openaimodel.py:544-774: 2 per level on the way down, 3 on the way up (levels 0-2), 1 in the
    # middle; the 64x64 up path runs after the reference views were dropped
    levels = [("L0 64x64", 320, 4096, 2, 3), ("L1 32x32", 640, 1024, 5, 0), ("L2 16x16", 1280, 256, 5, 0),
              ("mid 8x8", 1280, 64, 1, 0)]
    gemms = [  # name, FLOP coefficient of M*C*C, byte coefficient of M*C, weight matrices of C*C
        ("proj_in", 2, 6, 1), ("qkv", 6, 8, 3), ("to_out", 2, 10, 1), ("ff1_geglu", 16, 10, 8), ("ff2", 8, 14, 4),
        ("proj_out", 2, 10, 1)]
    rows, tot = [], {"flops": 0.0, "bytes": 0.0, "bound_s": 0.0, "tensor_s": 0.0, "hbm_s": 0.0}
    for name, C, hw, n_all, n_genonly in levels:
        for images, count in ((n_img, n_all), (n_gen, n_genonly)):
            if count == 0:
                continue
            M = images * hw
            for g, fc, bc, wc in gemms:
                flops = fc * M * C * C
                byts = bc * M * C + 2 * wc * C * C
                t_t, t_h = flops / tf_peak, byts / hbm_peak
                rows.append({"level": name, "gemm": g, "M": M, "C": C, "launches": count, "flops": flops, "bytes": byts,
                             "flop_per_byte": flops / byts, "tensor_us": t_t * 1e6, "hbm_us": t_h * 1e6,
                             "bound": "hbm" if t_h > t_t else "tensor"})
                tot["flops"] += count * flops
                tot["bytes"] += count * byts
                tot["tensor_s"] += count * t_t
                tot["hbm_s"] += count * t_h
                tot["bound_s"] += count * max(t_t, t_h)
    out = {"groups_per_call": args.groups, "peaks": {"tensor_tflops": tf_peak / 1e12, "hbm_gbs": hbm_peak / 1e9},
           "ridge_flop_per_byte": tf_peak / hbm_peak,
           "class_flops_T": tot["flops"] / 1e12, "class_bytes_GB": tot["bytes"] / 1e9,
           "tensor_only_bound_ms": tot["tensor_s"] * 1e3, "hbm_only_bound_ms": tot["hbm_s"] * 1e3,
           "mixed_bound_ms": tot["bound_s"] * 1e3}
    try:
        with open(args.bench) as f:
            k = json.loads(f.readline())["kernels"]["linear"]
        # the class also holds the input-stage GEMM (conv_in + cond_linear) and the skinny timestep linears: < 2 %
        out["measured_ms"] = k["ms_per_call"]
        out["frac_of_tensor_peak"] = tot["tensor_s"] * 1e3 / k["ms_per_call"]
        out["frac_of_mixed_roofline"] = tot["bound_s"] * 1e3 / k["ms_per_call"]
        out["measured_source"] = os.path.relpath(args.bench, ROOT)
    except (OSError, KeyError, ValueError):
        pass
    by_gemm = {}
    for r in rows:
        d = by_gemm.setdefault(r["level"], {})
        d[r["gemm"]] = d.get(r["gemm"], 0.0) + r["launches"] * max(r["tensor_us"], r["hbm_us"]) / 1e3
    out["mixed_bound_ms_by_level_and_gemm"] = by_gemm
    out["hbm_bound_launches"] = sorted({f'{r["level"]} {r["gemm"]} ({r["flop_per_byte"]:.0f} FLOP/B)' for r in rows
                                        if r["bound"] == "hbm"})
    # what fusing ff1 -> ff2 (the 4C-wide GEGLU activation never leaves the SM) would remove
    saved = sum(r["launches"] * 16 * r["M"] * r["C"] for r in rows if r["gemm"] == "ff2")
    out["ffn_fusion_saves_GB"] = saved / 1e9
    out["ffn_fusion_saves_ms_at_hbm_peak"] = saved / hbm_peak * 1e3
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

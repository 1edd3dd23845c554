#!/bin/bash
# Build the round-2 attention (and GroupNorm) A/B candidates HERE (nvcc cross-compiles), then run scripts/attn_variants.sh on the box:
#   scripts/prep_attn_variants.sh && gpurun --timeout 900 -- 'scripts/attn_variants.sh > gpurun_out/attn_variants.log 2>&1'
# Every candidate is correct by construction only on paper: attn_variants.sh runs the attention parity tests first.
cd "$(dirname "$0")/.." || exit 1
rm -rf gpurun_variants
set -e
python scripts/build_variant.py a_head
python scripts/build_variant.py b_packed -DCAP4D_ATTN_PACKED_F32X2=1
python scripts/build_variant.py c_packed_poly4 -DCAP4D_ATTN_PACKED_F32X2=1 -DCAP4D_ATTN_POLY_EVERY=4
python scripts/build_variant.py d_packed_poly8 -DCAP4D_ATTN_PACKED_F32X2=1 -DCAP4D_ATTN_POLY_EVERY=8
python scripts/build_variant.py e_regs112 -DCAP4D_ATTN_REGS_CTRL=32 -DCAP4D_ATTN_REGS_SOFTMAX=112
python scripts/build_variant.py f_packed_regs112 -DCAP4D_ATTN_PACKED_F32X2=1 -DCAP4D_ATTN_REGS_CTRL=32 -DCAP4D_ATTN_REGS_SOFTMAX=112
# GroupNorm candidates for scripts/gn_variants.sh
python scripts/build_variant.py gn_a_head
python scripts/build_variant.py gn_b_cache_hints -DCAP4D_GN_CACHE_HINTS=1
ls -la gpurun_variants

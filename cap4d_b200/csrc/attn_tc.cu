// Flash-style multi-view self-attention for sm_100a (tcgen05 + TMEM + TMA), head_dim 64.
//
// Reference semantics: cap4d/mmdm/net/attention.py:112-132 (legacy_attention: softmax(q k^T * d^-0.5) v,
// no mask) under the two rearranges at :233 ("3d": the tokens of ALL V views of a group form one
// sequence) and :237 ("spatial": one sequence per view).  Softmax attention without a mask is
// invariant to the order of the keys, so the '(n t)' interleave of the reference is not reproduced:
// a sequence is simply the contiguous token rows [s*L, (s+1)*L) of the fused QKV matrix.
//
// One CTA = one 128-row Q tile of one (sequence, head).  192 threads:
//   warp 0    TMA producer (Q once, K/V tiles through two 3-deep rings)
//   warp 1    MMA issuer: S_j = Q K_j^T (128x128, TMEM, double buffered), PV_j = P_j V_j (128x64, TMEM)
//   warps 2-5 softmax: thread <-> query row; S row TMEM->registers, online max/sum with exp2,
//             P_j -> bf16 -> shared memory in the UMMA K-major 128B-swizzle layout, O accumulated
//             in registers from the PV_j partial products (rescaled FA2-style).
// V is consumed straight from its row-major [kv][64] TMA tile as an MN-major B operand.
#include <cstdio>
#include <cstring>

#include "kernels.h"
#include "ptx.cuh"

namespace cap4d {

bool make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_el,
                    const uint32_t* box);

namespace {

constexpr int BQ = 128;   // query rows per CTA
constexpr int BKV = 128;  // keys per tile
constexpr int HD = 64;    // head dim
constexpr int TILE_BYTES = 128 * HD * 2;  // 16 KiB: one Q / K / V tile
constexpr int P_BYTES = BQ * BKV * 2;     // 32 KiB
constexpr int KS = 3, VS = 3;
constexpr int ATTN_THREADS = 192;
constexpr int TM_COLS = 512;
constexpr int TM_S0 = 0, TM_S1 = 128, TM_PV = 256;

struct AttnBars {
  uint64_t q_full;
  uint64_t k_full[KS], k_empty[KS];
  uint64_t v_full[VS], v_empty[VS];
  uint64_t s_full[2];
  uint64_t p_full[2];
  uint64_t pv_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnParams {
  int C, L, heads;
  int nkv;
  float scale_log2;
  bf16* out;
};

__global__ void __launch_bounds__(ATTN_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TILE_BYTES;
  uint8_t* sV = sK + KS * TILE_BYTES;
  uint8_t* sP = sV + VS * TILE_BYTES;
  AttnBars* bars = reinterpret_cast<AttnBars*>(sP + 2 * P_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y, seq = blockIdx.z;
  const int row_base = seq * p.L;  // first token row of this sequence
  const int q0 = qt * BQ;
  const int nkv = p.nkv;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(&bars->q_full, 1);
    for (int i = 0; i < KS; ++i) {
      mbar_init(&bars->k_full[i], 1);
      mbar_init(&bars->k_empty[i], 1);
    }
    for (int i = 0; i < VS; ++i) {
      mbar_init(&bars->v_full[i], 1);
      mbar_init(&bars->v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->s_full[i], 1);
      mbar_init(&bars->p_full[i], 128);
    }
    mbar_init(&bars->pv_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&bars->tmem_base, TM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars->q_full, TILE_BYTES);
      tma_load_2d(sQ, &tmQKV, &bars->q_full, head * HD, row_base + q0);
      for (int j = 0; j < nkv; ++j) {
        const int ks = j % KS, vs = j % VS;
        mbar_wait(&bars->k_empty[ks], ((j / KS) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars->k_full[ks], TILE_BYTES);
        tma_load_2d(sK + ks * TILE_BYTES, &tmQKV, &bars->k_full[ks], p.C + head * HD, row_base + j * BKV);
        mbar_wait(&bars->v_empty[vs], ((j / VS) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars->v_full[vs], TILE_BYTES);
        tma_load_2d(sV + vs * TILE_BYTES, &tmQKV, &bars->v_full[vs], 2 * p.C + head * HD, row_base + j * BKV);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc_qk = umma_idesc_bf16(BQ, BKV, 0);  // B = K tile, K-major (d contiguous)
      const uint32_t idesc_pv = umma_idesc_bf16(BQ, HD, 1);   // B = V tile, MN-major (d contiguous, k = kv row)
      const uint64_t qdesc = umma_smem_desc_sw128(smem_u32(sQ));
      mbar_wait(&bars->q_full, 0);
      auto issue_qk = [&](int j) {
        const int ks = j % KS;
        mbar_wait(&bars->k_full[ks], (j / KS) & 1);
        tc_fence_after();
        const uint64_t kdesc = umma_smem_desc_sw128(smem_u32(sK + ks * TILE_BYTES));
        const uint32_t d = tmem_base + ((j & 1) ? TM_S1 : TM_S0);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16(d, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0);
        umma_commit(&bars->k_empty[ks]);
        umma_commit(&bars->s_full[j & 1]);
      };
      issue_qk(0);
      for (int j = 0; j < nkv; ++j) {
        if (j + 1 < nkv) issue_qk(j + 1);  // S buffer (j+1)&1 was released by p_full of tile j-1
        const int vs = j % VS;
        mbar_wait(&bars->p_full[j & 1], (j >> 1) & 1);
        mbar_wait(&bars->v_full[vs], (j / VS) & 1);
        tc_fence_after();
        const uint64_t pdesc0 = umma_smem_desc_sw128(smem_u32(sP + (j & 1) * P_BYTES));
        const uint64_t pdesc1 = umma_smem_desc_sw128(smem_u32(sP + (j & 1) * P_BYTES + TILE_BYTES));
        const uint64_t vdesc = umma_smem_desc_sw128(smem_u32(sV + vs * TILE_BYTES));
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {
          // A: P sub-tile kk/4 (64 keys each), 32 B per k16 step; B: 16 kv rows = 2048 B per step
          const uint64_t ad = ((kk < 4) ? pdesc0 : pdesc1) + 2 * (kk & 3);
          const uint64_t bd = vdesc + static_cast<uint64_t>((kk * 16 * 128) >> 4);
          umma_bf16(tmem_base + TM_PV, ad, bd, idesc_pv, kk != 0);
        }
        umma_commit(&bars->v_empty[vs]);
        umma_commit(&bars->pv_full);
      }
    }
  } else {
    // ===================== softmax + O accumulation (warps 2..5) =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;  // row inside the Q tile
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    float o[HD];
#pragma unroll
    for (int i = 0; i < HD; ++i) o[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, alpha_pending = 1.f;
    const int valid_last = p.L - (nkv - 1) * BKV;  // valid keys in the last tile (1..128)

    for (int j = 0; j < nkv; ++j) {
      mbar_wait(&bars->s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      float s[BKV];
      {
        const uint32_t ta = tmem_base + lane_addr + ((j & 1) ? TM_S1 : TM_S0);
        uint32_t* su = reinterpret_cast<uint32_t*>(s);
        tmem_ld32(ta, su);
        tmem_ld32(ta + 32, su + 32);
        tmem_ld32(ta + 64, su + 64);
        tmem_ld32(ta + 96, su + 96);
        tmem_ld_wait();
      }
      if (j == nkv - 1 && valid_last < BKV) {
#pragma unroll
        for (int i = 0; i < BKV; ++i)
          if (i >= valid_last) s[i] = -INFINITY;
      }
      float mx = s[0];
#pragma unroll
      for (int i = 1; i < BKV; ++i) mx = fmaxf(mx, s[i]);
      const float m_new = fmaxf(m_run, mx * p.scale_log2);
      const float alpha = ex2f(m_run - m_new);
      float rowsum = 0.f;
      uint32_t pk[BKV / 2];
#pragma unroll
      for (int i = 0; i < BKV; i += 2) {
        const float p0 = ex2f(fmaf(s[i], p.scale_log2, -m_new));
        const float p1 = ex2f(fmaf(s[i + 1], p.scale_log2, -m_new));
        rowsum += p0 + p1;
        pk[i >> 1] = pack_bf16x2(p0, p1);
      }
      l_run = l_run * alpha + rowsum;
      m_run = m_new;

      // fold the previous tile's P V product into O (also frees P buffer / PV columns for reuse)
      if (j > 0) {
        mbar_wait(&bars->pv_full, (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t pv[32];
          tmem_ld32(tmem_base + lane_addr + TM_PV + h * 32, pv);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[h * 32 + i] = fmaf(o[h * 32 + i], alpha_pending, __uint_as_float(pv[i]));
        }
      }
      alpha_pending = alpha;

      // P_j -> smem, K-major SW128: row r, 16-byte chunk c of sub-tile t at (c ^ (r & 7)) * 16
      {
        uint8_t* prow = sP + (j & 1) * P_BYTES + r * 128;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint4 u = make_uint4(pk[t * 32 + c * 4], pk[t * 32 + c * 4 + 1], pk[t * 32 + c * 4 + 2],
                                 pk[t * 32 + c * 4 + 3]);
            *reinterpret_cast<uint4*>(prow + t * TILE_BYTES + ((c ^ (r & 7)) << 4)) = u;
          }
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&bars->p_full[j & 1]);
    }
    // last partial product
    mbar_wait(&bars->pv_full, (nkv - 1) & 1);
    tc_fence_after();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t pv[32];
      tmem_ld32(tmem_base + lane_addr + TM_PV + h * 32, pv);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[h * 32 + i] = fmaf(o[h * 32 + i], alpha_pending, __uint_as_float(pv[i]));
    }
    if (q0 + r < p.L) {
      const float inv = 1.0f / l_run;
      bf16* dst = p.out + static_cast<size_t>(row_base + q0 + r) * p.C + head * HD;
#pragma unroll
      for (int i = 0; i < HD; i += 8) {
        uint4 u;
        u.x = pack_bf16x2(o[i] * inv, o[i + 1] * inv);
        u.y = pack_bf16x2(o[i + 2] * inv, o[i + 3] * inv);
        u.z = pack_bf16x2(o[i + 4] * inv, o[i + 5] * inv);
        u.w = pack_bf16x2(o[i + 6] * inv, o[i + 7] * inv);
        *reinterpret_cast<uint4*>(dst + i) = u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TM_COLS);
  }
}

}  // namespace

bool make_attn_plan(AttnPlan* plan, const bf16* qkv, bf16* out, int M, int C, int L, float scale) {
  memset(plan, 0, sizeof(*plan));
  if (C % HD != 0 || L <= 0 || M % L != 0) {
    set_error("attention: C must be a multiple of 64 and M a multiple of the sequence length");
    return false;
  }
  plan->M = M;
  plan->C = C;
  plan->L = L;
  plan->n_seq = M / L;
  plan->heads = C / HD;
  plan->out = out;
  plan->scale_log2 = scale * 1.4426950408889634f;
  plan->grid = dim3((L + BQ - 1) / BQ, plan->heads, plan->n_seq);
  plan->smem_bytes = TILE_BYTES * (1 + KS + VS) + 2 * P_BYTES + sizeof(AttnBars) + 1024;
  plan->flops = 4.0 * plan->n_seq * plan->heads * static_cast<double>(L) * L * HD;
  uint64_t dims[2] = {static_cast<uint64_t>(3 * C), static_cast<uint64_t>(M)};
  uint64_t strides[2] = {1, static_cast<uint64_t>(3 * C)};
  uint32_t box[2] = {HD, 128};
  return make_tmap_bf16(&plan->tmQKV, qkv, 2, dims, strides, box);
}

cudaError_t launch_attn(const AttnPlan& plan, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  AttnParams p;
  p.C = plan.C;
  p.L = plan.L;
  p.heads = plan.heads;
  p.nkv = (plan.L + BKV - 1) / BKV;
  p.scale_log2 = plan.scale_log2;
  p.out = plan.out;
  attn_tc_kernel<<<plan.grid, ATTN_THREADS, plan.smem_bytes, stream>>>(plan.tmQKV, p);
  return cudaGetLastError();
}

}  // namespace cap4d

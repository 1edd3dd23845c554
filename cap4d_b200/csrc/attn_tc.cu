// Flash-style multi-view self-attention for sm_100a (tcgen05 + TMEM + TMA), head_dim 64.
//
// Reference semantics: cap4d/mmdm/net/attention.py:112-132 (legacy_attention: softmax(q k^T * d^-0.5) v,
// no mask) under the two rearranges at :233 ("3d": the tokens of ALL V views of a group form one
// sequence) and :237 ("spatial": one sequence per view).  Softmax attention without a mask is
// invariant to the order of the keys, so the '(n t)' interleave of the reference is not reproduced:
// a sequence is simply the contiguous token rows [s*L, (s+1)*L) of the fused QKV matrix.
//
// One CTA = two 128-row Q tiles (A, B) of one (sequence, head).  640 threads = 5 warpgroups:
//   WG0  warp 0 TMA producer (Q once, K/V tiles through two 4-deep rings), warps 1 / 3 MMA issuers of Q tile
//        A / B, warp 2 TMEM owner; registers trimmed to 40/thread (setmaxnreg) and handed to the softmax warpgroups.
//   WG1-2 / WG3-4  softmax of Q tile A / B: TWO threads per query row (same TMEM lane, warps q and q + 4), each
//        owning 64 of a tile's 128 keys and 32 of the row's 64 output columns - four softmax warps per scheduler
//        instead of two (the two-warp version sat in fixed-latency waits with the MUFU pipe 2/3 busy).
// Everything between the two GEMMs stays in tensor memory (512 columns: S_A S_B | O_A O_B | P_A P_B):
//   S_X(j) = Q_X K_j^T (SS MMA, fp32)  ->  registers (S_X is released to the tensor core at once, so QK of
//   tile j+1 overlaps the exponentials of tile j)  ->  P_X(j) = exp2(S - m) as bf16x2 back into TMEM
//   (tcgen05.st)  ->  O_X += P_X(j) V_j (TS MMA: A operand from TMEM, V straight from its row-major TMA
//   tile as an MN-major B operand).  O_X accumulates in TMEM over all KV tiles and is rescaled lazily
//   (only when a row max grows by more than 2^8), so the steady-state softmax is: load S, max (exchanged
//   between the row's two threads through shared memory), 64 exp2 per thread, pack, store P.
#include <atomic>
#include <cstdio>
#include <cstring>

#include "kernels.h"
#include "ptx.cuh"

namespace cap4d {

bool make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_el,
                    const uint32_t* box);

namespace {

constexpr int BQ = 128;   // query rows per softmax warpgroup (one UMMA M tile)
#ifndef CAP4D_ATTN_NQT
#define CAP4D_ATTN_NQT 2  // Q tiles per CTA.  1 (A/B builds): half the TMEM, threads and shared memory, two CTAs per SM
#endif
constexpr int NQT = CAP4D_ATTN_NQT;    // Q tiles per CTA
static_assert(NQT == 1 || NQT == 2, "one or two Q tiles per CTA");
constexpr int BKV = 128;  // keys per tile
constexpr int HD = 64;    // head dim
constexpr int TILE_BYTES = 128 * HD * 2;  // 16 KiB: one Q / K / V tile
#ifndef CAP4D_ATTN_KS
#define CAP4D_ATTN_KS (NQT == 2 ? 4 : 3)
#endif
#ifndef CAP4D_ATTN_VS
#define CAP4D_ATTN_VS (NQT == 2 ? 4 : 2)
#endif
constexpr int KS = CAP4D_ATTN_KS, VS = CAP4D_ATTN_VS;  // K / V ring depths (NQT == 1: two CTAs share the SM's 227 KB)
constexpr int ATTN_THREADS = 128 + NQT * 256;  // WG0: TMA / MMA / TMEM owner; WG1-2: softmax of Q tile A; WG3-4: Q tile B
constexpr int CTAS_PER_SM = (NQT == 2) ? 1 : 2;
constexpr int LAUNCH_REGS = (NQT == 2) ? 96 : 80;  // what ptxas gives every thread under the launch bounds
constexpr int SPLIT = 2;                  // threads per query row: each owns BKV / SPLIT keys of every tile
constexpr int HK = BKV / SPLIT;           // keys per softmax thread and tile
constexpr int TM_COLS = NQT * 256;
constexpr int TM_S = 0;                    // S_X (fp32 128x128)   at   0 + 128 x
constexpr int TM_O = NQT * 128;            // O_X (fp32 128x64)    behind the S tiles, + 64 x
constexpr int TM_P = NQT * 128 + NQT * 64; // P_X (bf16 128x128 = 64 columns) behind the O tiles, + 64 x
// pool = 640 * 96 at launch: 128 * (96 - REGS_CTRL) >= 512 * (REGS_SOFTMAX - 96) or the kernel deadlocks.
// Overridable for A/B builds (scripts/attn_variants.sh): -DCAP4D_ATTN_REGS_CTRL=32 -DCAP4D_ATTN_REGS_SOFTMAX=112
#ifndef CAP4D_ATTN_PACKED_F32X2
#define CAP4D_ATTN_PACKED_F32X2 1  // 0 = the scalar FFMA / FADD softmax of round 1 (kept for A/B builds)
#endif
#ifndef CAP4D_ATTN_POLY_EVERY
#define CAP4D_ATTN_POLY_EVERY 0  // needs CAP4D_ATTN_PACKED_F32X2: every N-th pair of scores uses exp2_poly_pair
#endif
#ifndef CAP4D_ATTN_LD64
#define CAP4D_ATTN_LD64 0   // A/B builds: one tcgen05.ld x64 per tile instead of two x32
#endif
#ifndef CAP4D_ATTN_ST_SPLIT
#define CAP4D_ATTN_ST_SPLIT 0   // A/B builds: P leaves in two x16 stores, the first one behind the first 32 exponentials
#endif
#ifndef CAP4D_ATTN_B_DELAY_NS
#define CAP4D_ATTN_B_DELAY_NS 1500  // head start of Q tile A's softmax warps over tile B's (sweep: profiles/r02_attn_delay.log)
#endif
#ifndef CAP4D_ATTN_MAX4
#define CAP4D_ATTN_MAX4 1   // four row-max chains instead of two (+0.7-0.9 %, profiles/r02_attn_micro2.log); 0 = round-1 code
#endif
#ifndef CAP4D_ATTN_REGS_CTRL
#define CAP4D_ATTN_REGS_CTRL (NQT == 2 ? 40 : 32)
#endif
#ifndef CAP4D_ATTN_REGS_SOFTMAX
#define CAP4D_ATTN_REGS_SOFTMAX 104
#endif
// Diagnostic builds (WRONG results, scripts/attn_diag.sh): what does the kernel cost without the row max / exchange,
// without the exponentials, without both (= the TMEM / MMA / barrier skeleton)?  Round 2, level-0 shape, cycles
// per KV tile pair: shipped 2850, no max 2450, no exp 2250, neither 2000 (profiles/r02_attn_diag.log).
#ifndef CAP4D_ATTN_DIAG_NOMAX
#define CAP4D_ATTN_DIAG_NOMAX 0
#endif
#ifndef CAP4D_ATTN_DIAG_NOEXP
#define CAP4D_ATTN_DIAG_NOEXP 0
#endif
constexpr int REGS_CTRL = CAP4D_ATTN_REGS_CTRL, REGS_SOFTMAX = CAP4D_ATTN_REGS_SOFTMAX;
static_assert(REGS_CTRL % 8 == 0 && REGS_SOFTMAX % 8 == 0 && REGS_CTRL >= 24 && REGS_SOFTMAX <= 256 &&
                  128 * (LAUNCH_REGS - REGS_CTRL) >= NQT * 256 * (REGS_SOFTMAX - LAUNCH_REGS),
              "setmaxnreg budget: the control warpgroup must release what the softmax warpgroups take");
constexpr float RESCALE_LOG2 = 8.0f;  // O / l are only rescaled when the row max grew by more than 2^8

struct AttnBars {
  uint64_t q_full;
  uint64_t k_full[KS], k_empty[KS];
  uint64_t v_full[VS], v_empty[VS];
  uint64_t s_full[NQT];   // MMA -> softmax: S_X(j) is in TMEM
  uint64_t s_free[NQT];   // softmax -> MMA: S_X(j) has been read into registers
  uint64_t p_full[NQT];   // softmax -> MMA: P_X(j) is in TMEM (and O_X has been rescaled if needed)
  uint64_t pv_full[NQT];  // MMA -> softmax: O_X += P_X(j) V_j is complete
  uint32_t tmem_base;
  // the two threads of a query row exchange their half-row maxima (double-buffered over tiles) and, at the end,
  // their half-row sums
  float xchg[2][NQT][SPLIT][BQ];
  float lsum[NQT][SPLIT][BQ];
};

__device__ __forceinline__ float ex2f(float x) {
#if CAP4D_ATTN_DIAG_NOEXP
  return x;
#else
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#endif
}

template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

struct AttnParams {
  int C, L, heads;
  int nkv;
  float scale_log2;
  bf16* out;
  long long* trace;  // TRACE builds only: [2 softmax WGs + 1 MMA][TRACE_J][8] clock64 stamps of CTA (0,0,0)
};

constexpr int TRACE_J = 16;
#define ATTN_STAMP(slot)                                                                            \
  do {                                                                                              \
    if (TRACE && trace != nullptr && j < TRACE_J) trace[(static_cast<size_t>(x) * TRACE_J + j) * 8 + (slot)] = clock64(); \
  } while (0)

// One KV tile of the online softmax for HALF a query row: the row's 128 scores are split between two threads
// (same TMEM lane, warps q and q + 4 of the tile's two warpgroups), which doubles the softmax warps per
// scheduler (4 instead of 2) - ncu showed the two-warp version waiting on fixed-latency dependencies with the
// MUFU pipe only 2/3 busy.  The halves agree on the row max through shared memory (one 64-thread named
// barrier per tile); everything else is independent: each thread exponentiates its 64 scores, stores its 32
// packed P columns, keeps its own partial row sum and rescales its 32 columns of O_X when the stabiliser moves.
// MASK: only the first `valid` keys of the tile count.  m_used is the stabiliser the running sums and O_X are
// expressed in; it only follows the true row max when that grew by more than 2^RESCALE_LOG2.
template <bool MASK, bool TRACE>
__device__ __forceinline__ void softmax_tile(AttnBars* bars, int x, int half, int r, int j, uint32_t s_addr,
                                             uint32_t o_addr, uint32_t p_addr, float scale_log2, int valid, float& m_used,
                                             float& l_run, long long* trace) {
  ATTN_STAMP(0);
  mbar_wait(&bars->s_full[x], j & 1);
  tc_fence_after();
  ATTN_STAMP(1);
  float s[HK];
  {
    uint32_t* su = reinterpret_cast<uint32_t*>(s);
#if CAP4D_ATTN_LD64
    tmem_ld64(s_addr, su);
#else
    tmem_ld32(s_addr, su);
    tmem_ld32(s_addr + 32, su + 32);
#endif
    tmem_ld_wait();
  }
  tc_fence_before();
  mbar_arrive(&bars->s_free[x]);  // the tensor core may overwrite S_X with tile j+1 once all 256 threads arrived
  ATTN_STAMP(2);
  if (MASK) {
#pragma unroll
    for (int i = 0; i < HK; ++i)
      if (half * HK + i >= valid) s[i] = -INFINITY;
  }
#if CAP4D_ATTN_DIAG_NOMAX
  bool pv_waited = (j == 0);
  if (j == 0) m_used = 0.f;
#else
#if CAP4D_ATTN_MAX4
  float mx0 = fmaxf(s[0], s[1]), mx1 = fmaxf(s[2], s[3]), mx2 = fmaxf(s[4], s[5]), mx3 = fmaxf(s[6], s[7]);
#pragma unroll
  for (int i = 8; i < HK; i += 8) {
    mx0 = fmaxf(mx0, fmaxf(s[i], s[i + 1]));
    mx1 = fmaxf(mx1, fmaxf(s[i + 2], s[i + 3]));
    mx2 = fmaxf(mx2, fmaxf(s[i + 4], s[i + 5]));
    mx3 = fmaxf(mx3, fmaxf(s[i + 6], s[i + 7]));
  }
  const float m_half = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
#else
  float mx0 = fmaxf(s[0], s[1]), mx1 = fmaxf(s[2], s[3]);
#pragma unroll
  for (int i = 4; i < HK; i += 4) {
    mx0 = fmaxf(mx0, fmaxf(s[i], s[i + 1]));
    mx1 = fmaxf(mx1, fmaxf(s[i + 2], s[i + 3]));
  }
  const float m_half = fmaxf(mx0, mx1);
#endif
  float* xc = &bars->xchg[j & 1][x][0][0];
  xc[half * BQ + r] = m_half;
  named_bar_sync(1 + x * 4 + (r >> 5), 64);  // the two warps that share these 32 rows
  const float m_new = fmaxf(m_used, fmaxf(m_half, xc[(half ^ 1) * BQ + r]) * scale_log2);
  bool pv_waited = (j == 0);
  if (j == 0) {
    m_used = m_new;  // O_X is still empty: nothing to rescale
  } else {
    const bool grow = (m_new - m_used) > RESCALE_LOG2;
    if (__any_sync(0xffffffffu, grow)) {  // same answer in the partner warp: it sees the same 32 row maxima
      // rare: O_X must be rescaled, which needs O_X += P_X(j-1) V_(j-1) to be complete
      mbar_wait(&bars->pv_full[x], (j - 1) & 1);
      tc_fence_after();
      pv_waited = true;
      const float f = grow ? ex2f(m_used - m_new) : 1.0f;
      uint32_t ov[32];
      tmem_ld32(o_addr, ov);  // this thread's 32 of the row's 64 output columns
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * f);
      tmem_st32(o_addr, ov);
      l_run *= f;
      if (grow) m_used = m_new;
    }
  }
#endif
  ATTN_STAMP(3);
  uint32_t pk[HK / 2];
#if CAP4D_ATTN_PACKED_F32X2
  // scale-subtract and row sums as packed f32x2 (FFMA2 / FADD2: 2.75 instead of 3.75 issue slots per score;
  // +4 % on the level-0 / level-1 shapes, profiles/r02_attn_variants.log)
  {
    const f32x2 sc2 = pack2(scale_log2, scale_log2), nm2 = pack2(-m_used, -m_used);
    f32x2 rs2 = pack2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < HK; i += 2) {
      float a0, a1;
      unpack2(fma2(pack2(s[i], s[i + 1]), sc2, nm2), a0, a1);
      const float p0 = ex2f(a0), p1 = ex2f(a1);
      rs2 = add2(rs2, pack2(p0, p1));
      pk[i >> 1] = pack_bf16x2(p0, p1);
#if CAP4D_ATTN_ST_SPLIT
      if (i == HK / 2 - 2) {  // the first half of this thread's P columns is complete
        if (!pv_waited) {
          mbar_wait(&bars->pv_full[x], (j - 1) & 1);
          tc_fence_after();
          pv_waited = true;
        }
        tmem_st16(p_addr, pk);
      }
#endif
    }
    float rs0, rs1;
    unpack2(rs2, rs0, rs1);
    l_run += rs0 + rs1;
  }
#else
  float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
  for (int i = 0; i < HK; i += 2) {
    const float p0 = ex2f(fmaf(s[i], scale_log2, -m_used));
    const float p1 = ex2f(fmaf(s[i + 1], scale_log2, -m_used));
    rs0 += p0;
    rs1 += p1;
    pk[i >> 1] = pack_bf16x2(p0, p1);
  }
  l_run += rs0 + rs1;
#endif
  ATTN_STAMP(4);
  if (!pv_waited) {
    // P_X may only be overwritten once the PV MMA of tile j-1 has consumed it (long done by now)
    mbar_wait(&bars->pv_full[x], (j - 1) & 1);
    tc_fence_after();
  }
  // P_X(j) -> TMEM as the A operand of the PV MMA: lane = query row, column c = keys (2c, 2c+1) as bf16x2
#if CAP4D_ATTN_ST_SPLIT && CAP4D_ATTN_PACKED_F32X2
  tmem_st16(p_addr + HK / 4, pk + HK / 4);
#else
  tmem_st32(p_addr, pk);
#endif
  tmem_st_wait();
  ATTN_STAMP(5);
  tc_fence_before();
  mbar_arrive(&bars->p_full[x]);
  ATTN_STAMP(6);
}

template <bool TRACE>
__global__ void __launch_bounds__(ATTN_THREADS, CTAS_PER_SM)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                      // [NQT] tiles
  uint8_t* sK = sQ + NQT * TILE_BYTES;
  uint8_t* sV = sK + KS * TILE_BYTES;
  AttnBars* bars = reinterpret_cast<AttnBars*>(sV + VS * TILE_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int wg = warp >> 2;
  const int qt = blockIdx.x, head = blockIdx.y, seq = blockIdx.z;
  const int row_base = seq * p.L;  // first token row of this sequence
  const int q0 = qt * (NQT * BQ);
  const int nkv = p.nkv;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(&bars->q_full, 1);
    for (int i = 0; i < KS; ++i) {
      mbar_init(&bars->k_full[i], 1);
      mbar_init(&bars->k_empty[i], NQT);  // one commit per MMA issuer
    }
    for (int i = 0; i < VS; ++i) {
      mbar_init(&bars->v_full[i], 1);
      mbar_init(&bars->v_empty[i], NQT);
    }
    for (int i = 0; i < NQT; ++i) {
      mbar_init(&bars->s_full[i], 1);
      mbar_init(&bars->s_free[i], 128 * SPLIT);
      mbar_init(&bars->p_full[i], 128 * SPLIT);
      mbar_init(&bars->pv_full[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(&bars->tmem_base, TM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (wg == 0) {
    setmaxnreg_dec<REGS_CTRL>();
    if (warp == 0 && lane == 0) {
      // ===================== TMA producer =====================
      mbar_arrive_expect_tx(&bars->q_full, NQT * TILE_BYTES);
      for (int x = 0; x < NQT; ++x)
        tma_load_2d(sQ + x * TILE_BYTES, &tmQKV, &bars->q_full, head * HD, row_base + q0 + x * BQ);
      for (int j = 0; j < nkv; ++j) {
        const int ks = j % KS, vs = j % VS;
        mbar_wait(&bars->k_empty[ks], ((j / KS) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars->k_full[ks], TILE_BYTES);
        tma_load_2d(sK + ks * TILE_BYTES, &tmQKV, &bars->k_full[ks], p.C + head * HD, row_base + j * BKV);
        mbar_wait(&bars->v_empty[vs], ((j / VS) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars->v_full[vs], TILE_BYTES);
        tma_load_2d(sV + vs * TILE_BYTES, &tmQKV, &bars->v_full[vs], 2 * p.C + head * HD, row_base + j * BKV);
      }
    } else if ((warp == 1 || (warp == 3 && NQT == 2)) && lane == 0) {
      // ===================== MMA issuers: warp 1 -> Q tile A, warp 3 -> Q tile B =====================
      // (one issuing thread per Q tile: the mbarrier round trips of one tile's chain do not delay the other's)
      const int x = warp >> 1;
      const uint32_t idesc_qk = umma_idesc_bf16(BQ, BKV, 0);  // B = K tile, K-major (d contiguous)
      const uint32_t idesc_pv = umma_idesc_bf16(BQ, HD, 1);   // A = P in TMEM; B = V tile, MN-major
      const uint64_t qdesc = umma_smem_desc_sw128(smem_u32(sQ + x * TILE_BYTES));
      const uint32_t s_tmem = tmem_base + TM_S + x * BKV;
      const uint32_t o_tmem = tmem_base + TM_O + x * HD;
      const uint32_t p_tmem = tmem_base + TM_P + x * (BKV / 2);
      long long* trace = (TRACE && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && x == 0) ? p.trace : nullptr;
      mbar_wait(&bars->q_full, 0);
      // S_X(j) = Q_X K_j^T; K stage j is released when both issuers have committed
      auto issue_qk = [&](int j) {
        const int ks = j % KS;
        mbar_wait(&bars->k_full[ks], (j / KS) & 1);
        if (j > 0) mbar_wait(&bars->s_free[x], (j - 1) & 1);  // S_X(j-1) is in registers
        tc_fence_after();
        const uint64_t kdesc = umma_smem_desc_sw128(smem_u32(sK + ks * TILE_BYTES));
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16(s_tmem, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0);
        umma_commit(&bars->s_full[x]);
        umma_commit(&bars->k_empty[ks]);
      };
      issue_qk(0);
      for (int j = 0; j < nkv; ++j) {
        if (TRACE && trace != nullptr && j < TRACE_J) trace[(static_cast<size_t>(2) * TRACE_J + j) * 8 + 0] = clock64();
        if (j + 1 < nkv) issue_qk(j + 1);
        if (TRACE && trace != nullptr && j < TRACE_J) trace[(static_cast<size_t>(2) * TRACE_J + j) * 8 + 1] = clock64();
        const int vs = j % VS;
        mbar_wait(&bars->v_full[vs], (j / VS) & 1);
        mbar_wait(&bars->p_full[x], j & 1);
        tc_fence_after();
        const uint64_t vdesc = umma_smem_desc_sw128(smem_u32(sV + vs * TILE_BYTES));
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {
          // A: 16 keys = 8 TMEM columns per step; B: 16 kv rows = 2048 B per step
          const uint64_t bd = vdesc + static_cast<uint64_t>((kk * 16 * 128) >> 4);
          umma_bf16_ts(o_tmem, p_tmem + kk * 8, bd, idesc_pv, (j | kk) != 0);
        }
        umma_commit(&bars->pv_full[x]);
        umma_commit(&bars->v_empty[vs]);
        if (TRACE && trace != nullptr && j < TRACE_J) trace[(static_cast<size_t>(2) * TRACE_J + j) * 8 + 2] = clock64();
      }
    }
  } else {
    // ===================== softmax: WG1, WG2 -> Q tile 0 (key halves 0, 1); WG3, WG4 -> Q tile 1 =====================
    setmaxnreg_inc<REGS_SOFTMAX>();
    const int x = (wg - 1) >> 1;
    const int half = (wg - 1) & 1;
    const int q = warp & 3;
    const int r = q * 32 + lane;  // row inside the Q tile
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_addr + TM_S + x * BKV + half * HK;
    const uint32_t o_addr = tmem_base + lane_addr + TM_O + x * HD + half * (HD / SPLIT);
    const uint32_t p_addr = tmem_base + lane_addr + TM_P + x * (BKV / 2) + half * (HK / 2);
    float m_used = -INFINITY, l_run = 0.f;
    // Tile B's softmax warps start about half a tile late: the two tiles then tend to alternate on the MUFU pipe
    // (one exponentiates while the other loads / takes its max / stores P) instead of running in lockstep.
    // Measured -3 % on the large shapes; enforcing the alternation with named barriers costs 10 % instead.
    if (NQT == 2 && x == 1 && CAP4D_ATTN_B_DELAY_NS > 0) __nanosleep(CAP4D_ATTN_B_DELAY_NS);
    const int valid_last = p.L - (nkv - 1) * BKV;  // valid keys in the last tile (1..128)

    // the key mask costs 2 ALU ops per score, so it is compiled only into the (peeled) last tile
    const bool tail = valid_last < BKV;
    const int n_main = tail ? nkv - 1 : nkv;
    long long* trace = nullptr;
    if (TRACE && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (threadIdx.x & 127) == 0 && half == 0)
      trace = p.trace;
    for (int j = 0; j < n_main; ++j)
      softmax_tile<false, TRACE>(bars, x, half, r, j, s_addr, o_addr, p_addr, p.scale_log2, BKV, m_used, l_run, trace);
    if (tail)
      softmax_tile<true, TRACE>(bars, x, half, r, nkv - 1, s_addr, o_addr, p_addr, p.scale_log2, valid_last, m_used,
                                l_run, trace);
    // the row sum is the sum of the two halves' partial sums (same stabiliser in both)
    bars->lsum[x][half][r] = l_run;
    named_bar_sync(1 + x * 4 + q, 64);
    const float inv = 1.0f / (l_run + bars->lsum[x][half ^ 1][r]);
    // O_X is complete once the last PV MMA has landed
    mbar_wait(&bars->pv_full[x], (nkv - 1) & 1);
    tc_fence_after();
    const int qrow = q0 + x * BQ + r;
    bf16* dst = p.out + static_cast<size_t>(row_base + qrow) * p.C + head * HD + half * (HD / SPLIT);
    uint32_t ov[32];
    tmem_ld32(o_addr, ov);
    tmem_ld_wait();
    if (qrow < p.L) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(ov[i]) * inv, __uint_as_float(ov[i + 1]) * inv);
        u.y = pack_bf16x2(__uint_as_float(ov[i + 2]) * inv, __uint_as_float(ov[i + 3]) * inv);
        u.z = pack_bf16x2(__uint_as_float(ov[i + 4]) * inv, __uint_as_float(ov[i + 5]) * inv);
        u.w = pack_bf16x2(__uint_as_float(ov[i + 6]) * inv, __uint_as_float(ov[i + 7]) * inv);
        *reinterpret_cast<uint4*>(dst + i) = u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
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
  plan->grid = dim3((L + NQT * BQ - 1) / (NQT * BQ), plan->heads, plan->n_seq);
  plan->smem_bytes = TILE_BYTES * (NQT + KS + VS) + sizeof(AttnBars) + 1024;
  plan->flops = 4.0 * plan->n_seq * plan->heads * static_cast<double>(L) * L * HD;
  uint64_t dims[2] = {static_cast<uint64_t>(3 * C), static_cast<uint64_t>(M)};
  uint64_t strides[2] = {1, static_cast<uint64_t>(3 * C)};
  uint32_t box[2] = {HD, 128};
  return make_tmap_bf16(&plan->tmQKV, qkv, 2, dims, strides, box);
}

cudaError_t launch_attn(const AttnPlan& plan, cudaStream_t stream, long long* trace) {
  // the attribute is per device: one flag per device ordinal
  static std::atomic<bool> attr_set[64];
  int dev = 0;
  cudaGetDevice(&dev);
  const bool known = dev >= 0 && dev < 64;
  if (!known || !attr_set[dev].load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(attn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    if (known) attr_set[dev].store(true, std::memory_order_release);
  }
  AttnParams p;
  p.C = plan.C;
  p.L = plan.L;
  p.heads = plan.heads;
  p.nkv = (plan.L + BKV - 1) / BKV;
  p.scale_log2 = plan.scale_log2;
  p.out = plan.out;
  p.trace = trace;
  if (trace != nullptr)
    attn_tc_kernel<true><<<plan.grid, ATTN_THREADS, plan.smem_bytes, stream>>>(plan.tmQKV, p);
  else
    attn_tc_kernel<false><<<plan.grid, ATTN_THREADS, plan.smem_bytes, stream>>>(plan.tmQKV, p);
  return cudaGetLastError();
}

}  // namespace cap4d

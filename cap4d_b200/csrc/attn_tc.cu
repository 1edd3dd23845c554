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
//        A / B, warp 2 TMEM owner; registers trimmed to 32/thread (setmaxnreg) and handed to the softmax warpgroups.
//   WG1-2 / WG3-4  softmax of Q tile A / B, 112 registers per thread: eight warps per tile, each owning 16 query
//        rows (a 16-lane half of its TMEM quadrant) and all 128 keys of a KV tile, FOUR threads per row
//        (softmax_tile_rq) - four softmax warps per scheduler and no communication between warps.
// Everything between the two GEMMs stays in tensor memory (512 columns: S_A S_B | O_A O_B | P_A P_B):
//   S_X(j) = Q_X K_j^T (SS MMA, fp32)  ->  registers (S_X is released to the tensor core at once, so QK of
//   tile j+1 overlaps the exponentials of tile j)  ->  P_X(j) = exp2(S - m) as bf16x2 back into TMEM
//   (tcgen05.st)  ->  O_X += P_X(j) V_j (TS MMA: A operand from TMEM, V straight from its row-major TMA
//   tile as an MN-major B operand).  O_X accumulates in TMEM over all KV tiles and is rescaled lazily
//   (only when a row max grows by more than 2^8), so the steady-state softmax is: load S, row max (two shuffles),
//   64 exp2 per thread, pack, store P.
// What bounds it (profiles/r02b_attn_*.log): three streams of ~2000 cycles per pair of KV
// tiles - the MUFU (2 x 128 x 128 exp2 at 16 per clock), the MMA / TMA / barrier skeleton (1.41 ms of the 2.05 ms
// level-0 launch with the softmax arithmetic deleted) and the softmax warps' own serial latency - overlap imperfectly.
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
#ifndef CAP4D_ATTN_ROWQUAD
#define CAP4D_ATTN_ROWQUAD 1   // a query row's 128 scores live in FOUR threads of ONE warp (16-lane TMEM shapes), see softmax_tile_rq; 0 = the 32-lane softmax of round 2a (two threads per row in two warps)
#endif
#ifndef CAP4D_ATTN_SCHED_FENCE
#define CAP4D_ATTN_SCHED_FENCE 1   // row-quad variant: keep the s_free arrive ahead of the exponentials (sched_fence())
#endif
#ifndef CAP4D_ATTN_TURNS
#define CAP4D_ATTN_TURNS 0   // row-quad variant: the two Q tiles take turns on each scheduler's MUFU (exp(j) of A, exp(j) of B, exp(j+1) of A, ...)
#endif
#ifndef CAP4D_ATTN_EARLY_TMA
#define CAP4D_ATTN_EARLY_TMA 0   // Q, K(0), V(0) requested before the TMEM allocation / first block barrier
#endif
#ifndef CAP4D_ATTN_LAZYMAX
#define CAP4D_ATTN_LAZYMAX 0   // row-quad variant: exact row max at KV tile 0 only, growth detected through the row sums (softmax_tile_rq)
#endif
#ifndef CAP4D_ATTN_REGS_CTRL
#define CAP4D_ATTN_REGS_CTRL ((NQT == 2 && !CAP4D_ATTN_ROWQUAD) ? 40 : 32)
#endif
#ifndef CAP4D_ATTN_REGS_SOFTMAX
#define CAP4D_ATTN_REGS_SOFTMAX ((NQT == 2 && CAP4D_ATTN_ROWQUAD) ? 112 : 104)   // row-quad: 112 / 32 measured +1-4 % over 104 / 40
#endif
// Diagnostic builds (WRONG results, scripts/attn_diag.sh): what does the kernel cost without the row max / exchange,
// without the exponentials, without both (= the TMEM / MMA / barrier skeleton)?  Round 2, level-0 shape, cycles
// per KV tile pair: shipped 2850, no max 2450, no exp 2250, neither 2000 (profiles/r02_attn_diag.log).
#ifndef CAP4D_ATTN_DIAG_NOMAX
#define CAP4D_ATTN_DIAG_NOMAX 0
#endif
#ifndef CAP4D_ATTN_DIAG_NOFMA
#define CAP4D_ATTN_DIAG_NOFMA 0   // row-quad variant only: exponentiate the raw scores (no scale / stabiliser FFMA2)
#endif
#ifndef CAP4D_ATTN_DIAG_NOQK
#define CAP4D_ATTN_DIAG_NOQK 0   // skeleton builds: the QK / PV MMAs are not issued (their commits are)
#endif
#ifndef CAP4D_ATTN_DIAG_NOPV
#define CAP4D_ATTN_DIAG_NOPV 0
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
  uint64_t always;        // phase 0 completed at start-up and never used again: see sched_fence()
  uint64_t turn[NQT][4];  // CAP4D_ATTN_TURNS: tile x's two softmax warps on scheduler q may start their exponentials
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
  int C, L, heads, M;
  int nkv;
  const bf16* qkv;   // the fused QKV matrix the tensor map describes (rq_regrow reads it directly)
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

// Row-quad variant (CAP4D_ATTN_ROWQUAD): the tile's eight softmax warps each own 16 query rows (lanes 32q + 16h ..
// + 15 of the tile's TMEM quadrant q) and ALL 128 keys, read with the 16-lane shapes of tcgen05.ld: thread t holds,
// for rows t/4 and t/4 + 8, the scores of keys 8k + 2(t%4) + {0,1} (k = 0..15) - an mma.sync-style fragment.  A
// row reduction is two shuffles inside the warp: no shared-memory exchange, no named barrier, no coupling between
// warps (the 32-lane version spent 350-400 of a tile's 2800 cycles there, scripts/attn_trace.py).  Packing a
// thread's fp32 pairs to bf16x2 gives exactly the .16x128b store fragment of P, and a warp owns all 64 output
// columns of its rows, so rescaling O_X is warp-local as well.
//
// Lazy stabiliser (CAP4D_ATTN_LAZYMAX): any stabiliser m gives the same softmax as long as nothing overflows -
// fp32 sums and bf16 probabilities keep their RELATIVE precision at any magnitude.  So only KV tile 0 takes the
// exact row max; every later tile exponentiates straight against the stabiliser it inherited (no FMNMX, no
// shuffle, no branch before the exponentials: +6-9 % in the diagnostic build, profiles/r02b_attn_rq_diag.log) and
// looks at its own row sums afterwards: a partial sum >= 2^12 (or inf / NaN) sends the warp to rq_regrow, which
//   * shifts the stabiliser by a whole number n of octaves (P, the tile's sums, l and O_X are multiplied by
//     2^-n: exact), when the tile's row sum is finite, or
//   * recomputes the tile's scores from global memory and redoes the tile with the exact row max, when a score
//     overflowed the exponential (a jump of more than 2^100 over everything the row has seen: unheard of in
//     a trained model, but the kernel must not be wrong there).
constexpr float GROW_SUM = 4096.f;        // a thread's 32-score partial row sum that triggers rq_regrow
constexpr float FINITE_SUM = 1.2676506e30f;  // 2^100: above this (or NaN) the tile is recomputed

// ptxas schedules a basic block by critical path: the s_free arrive has no consumers, so without a block boundary
// behind it ptxas sinks it below the whole exponential phase (seen in the SASS: QK(j+1) then starts a tile late).
// A wait on a barrier whose phase 0 is complete for good is a branch ptxas cannot see through; it costs one
// try_wait (~30 cycles) and never blocks.
__device__ __forceinline__ void sched_fence(AttnBars* bars) { mbar_wait(&bars->always, 0); }

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// O_X rows of this warp (16 lanes x 64 columns) times fA (rows t/4) / fB (rows t/4 + 8)
__device__ __forceinline__ void rq_rescale_o(uint32_t o_addr, float fA, float fB) {
  uint32_t ov[32];
  tmem_ld_16x256b_x8(o_addr, ov);
  tmem_ld_wait();
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    ov[4 * k] = __float_as_uint(__uint_as_float(ov[4 * k]) * fA);
    ov[4 * k + 1] = __float_as_uint(__uint_as_float(ov[4 * k + 1]) * fA);
    ov[4 * k + 2] = __float_as_uint(__uint_as_float(ov[4 * k + 2]) * fB);
    ov[4 * k + 3] = __float_as_uint(__uint_as_float(ov[4 * k + 3]) * fB);
  }
  tmem_st_16x256b_x8(o_addr, ov);  // completes before the caller's tcgen05.wait::st / p_full arrive
}

// exact stabiliser: row maxima of the scores in registers; O_X / l follow when the stabiliser moves by more than
// 2^RESCALE_LOG2 (nothing to rescale at tile 0)
__device__ __forceinline__ void rq_exact_stabiliser(const float (&s)[BKV / 2], AttnBars* bars, int x, int j,
                                                    uint32_t o_addr, float scale_log2, float (&m_used)[2],
                                                    float (&l_run)[2], bool& pv_waited) {
  float a0 = fmaxf(s[0], s[1]), b0 = fmaxf(s[2], s[3]), a1 = fmaxf(s[4], s[5]), b1 = fmaxf(s[6], s[7]);
#pragma unroll
  for (int k = 2; k < BKV / 8; k += 2) {
    a0 = fmaxf(a0, fmaxf(s[4 * k], s[4 * k + 1]));
    b0 = fmaxf(b0, fmaxf(s[4 * k + 2], s[4 * k + 3]));
    a1 = fmaxf(a1, fmaxf(s[4 * k + 4], s[4 * k + 5]));
    b1 = fmaxf(b1, fmaxf(s[4 * k + 6], s[4 * k + 7]));
  }
  const float mA = quad_max(fmaxf(a0, a1)), mB = quad_max(fmaxf(b0, b1));
  const float nA = fmaxf(m_used[0], mA * scale_log2), nB = fmaxf(m_used[1], mB * scale_log2);
  if (j == 0) {
    m_used[0] = nA;  // O_X is still empty: nothing to rescale
    m_used[1] = nB;
    return;
  }
  const bool gA = (nA - m_used[0]) > RESCALE_LOG2, gB = (nB - m_used[1]) > RESCALE_LOG2;
  if (__any_sync(0xffffffffu, gA || gB)) {
    if (!pv_waited) {  // O_X += P_X(j-1) V_(j-1) must be complete
      mbar_wait(&bars->pv_full[x], (j - 1) & 1);
      tc_fence_after();
      pv_waited = true;
    }
    const float fA = gA ? ex2f(m_used[0] - nA) : 1.0f, fB = gB ? ex2f(m_used[1] - nB) : 1.0f;
    rq_rescale_o(o_addr, fA, fB);
    l_run[0] *= fA;
    l_run[1] *= fB;
    if (gA) m_used[0] = nA;
    if (gB) m_used[1] = nB;
  }
}

// P = exp2(s * scale - m) as bf16x2 in the .16x128b store fragment, and this thread's partial row sums
template <typename Mid>
__device__ __forceinline__ void rq_exp(const float (&s)[BKV / 2], float scale_log2, const float (&m_used)[2],
                                       uint32_t (&pk)[BKV / 4], float& rsA, float& rsB, Mid mid) {
  const f32x2 sc2 = pack2(scale_log2, scale_log2);
  const f32x2 nmA = pack2(-m_used[0], -m_used[0]), nmB = pack2(-m_used[1], -m_used[1]);
  f32x2 accA = pack2(0.f, 0.f), accB = pack2(0.f, 0.f);
#pragma unroll
  for (int k = 0; k < BKV / 8; ++k) {
    float e0, e1, e2, e3;
#if CAP4D_ATTN_DIAG_NOFMA
    e0 = s[4 * k], e1 = s[4 * k + 1], e2 = s[4 * k + 2], e3 = s[4 * k + 3];
#else
    unpack2(fma2(pack2(s[4 * k], s[4 * k + 1]), sc2, nmA), e0, e1);
    unpack2(fma2(pack2(s[4 * k + 2], s[4 * k + 3]), sc2, nmB), e2, e3);
#endif
    const float p0 = ex2f(e0), p1 = ex2f(e1), p2 = ex2f(e2), p3 = ex2f(e3);
    accA = add2(accA, pack2(p0, p1));
    accB = add2(accB, pack2(p2, p3));
    pk[2 * k] = pack_bf16x2(p0, p1);      // row A, P column 4k + c
    pk[2 * k + 1] = pack_bf16x2(p2, p3);  // row B
    if (k == BKV / 16 - 1) {  // half of the exponentials are through
      float h0, h1;
      unpack2(add2(accA, accB), h0, h1);
      mid(h0 + h1);
    }
  }
  float r0, r1;
  unpack2(accA, r0, r1);
  rsA = r0 + r1;
  unpack2(accB, r0, r1);
  rsB = r0 + r1;
}

#if CAP4D_ATTN_LAZYMAX
struct RegrowCtx {       // what the recomputation needs to find the tile's Q rows and keys in global memory
  const bf16* q_rowA;    // this thread's first query row (head offset applied, clamped into the matrix)
  const bf16* q_rowB;
  const bf16* k_tile;    // K row of the tile's first key (head offset applied)
  int ld;                // row pitch of the fused QKV matrix (3 C)
};
struct RegrowIO {
  uint32_t pk[BKV / 4];
  float m_used[2], l_run[2], rs[2];
  int pv_waited;
};

// Rare path of the lazy stabiliser (see the comment above softmax_tile_rq); out of line so that its registers and
// local arrays do not weigh on the tile loop.
__device__ __noinline__ RegrowIO rq_regrow(RegrowIO io, RegrowCtx ctx, AttnBars* bars, int x, int c, int j,
                                           uint32_t o_addr, float scale_log2, int valid) {
  const float SA = quad_sum(io.rs[0]), SB = quad_sum(io.rs[1]);
  const bool finite = (SA < FINITE_SUM) && (SB < FINITE_SUM);  // false for inf and NaN
  if (!io.pv_waited) {  // both branches touch O_X: O_X += P_X(j-1) V_(j-1) must be complete
    mbar_wait(&bars->pv_full[x], (j - 1) & 1);
    tc_fence_after();
    io.pv_waited = 1;
  }
  if (__all_sync(0xffffffffu, finite)) {
    // whole octaves: the row sum of the tile comes back into [0.5, 1)
    const int nA = SA >= GROW_SUM ? static_cast<int>((__float_as_uint(SA) >> 23) & 0xff) - 126 : 0;
    const int nB = SB >= GROW_SUM ? static_cast<int>((__float_as_uint(SB) >> 23) & 0xff) - 126 : 0;
    const float fA = __uint_as_float(static_cast<uint32_t>(127 - nA) << 23);
    const float fB = __uint_as_float(static_cast<uint32_t>(127 - nB) << 23);
    rq_rescale_o(o_addr, fA, fB);
#pragma unroll
    for (int k = 0; k < BKV / 8; ++k) {
      const uint32_t a = io.pk[2 * k], b = io.pk[2 * k + 1];
      io.pk[2 * k] = pack_bf16x2(__uint_as_float(a << 16) * fA, __uint_as_float(a & 0xffff0000u) * fA);
      io.pk[2 * k + 1] = pack_bf16x2(__uint_as_float(b << 16) * fB, __uint_as_float(b & 0xffff0000u) * fB);
    }
    io.rs[0] *= fA;
    io.rs[1] *= fB;
    io.l_run[0] *= fA;
    io.l_run[1] *= fB;
    io.m_used[0] += static_cast<float>(nA);
    io.m_used[1] += static_cast<float>(nB);
    return io;
  }
  // a score overflowed the exponential: the tile's scores again, from global memory (fp32 dot products of the same
  // bf16 values the tensor core multiplied), then the exact-max tile
  float s[BKV / 2];
  {
    float qa[HD], qb[HD];
#pragma unroll
    for (int d = 0; d < HD; d += 2) {
      const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(ctx.q_rowA + d));
      const float2 fb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(ctx.q_rowB + d));
      qa[d] = fa.x, qa[d + 1] = fa.y, qb[d] = fb.x, qb[d + 1] = fb.y;
    }
#pragma unroll 1
    for (int i = 0; i < BKV / 4; ++i) {
      const int k = i >> 1, e = i & 1, key = 8 * k + 2 * c + e;
      float dA = -INFINITY, dB = -INFINITY;
      if (key < valid) {
        const bf16* kp = ctx.k_tile + static_cast<size_t>(key) * ctx.ld;
        dA = dB = 0.f;
#pragma unroll 8
        for (int d = 0; d < HD; d += 2) {
          const float2 kv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(kp + d));
          dA = fmaf(qa[d], kv.x, dA), dA = fmaf(qa[d + 1], kv.y, dA);
          dB = fmaf(qb[d], kv.x, dB), dB = fmaf(qb[d + 1], kv.y, dB);
        }
      }
      s[4 * k + e] = dA;
      s[4 * k + 2 + e] = dB;
    }
  }
  bool pvw = true;
  rq_exact_stabiliser(s, bars, x, j, o_addr, scale_log2, io.m_used, io.l_run, pvw);
  rq_exp(s, scale_log2, io.m_used, io.pk, io.rs[0], io.rs[1], [](float) {});
  return io;
}
#else
struct RegrowCtx {};
#endif

// S_X(j) -> registers, and S_X back to the tensor core.  (Loading S_X(j+1) between the store of P_X(j) and the wait
// for it - the two TMEM round trips are ~300 cycles each - was measured: carrying the scores across the loop edge
// costs ~250 B of spills per tile and 10 %, profiles/r02b_attn_pipe.log.)
__device__ __forceinline__ void rq_load_issue(AttnBars* bars, int x, int j, uint32_t s_addr, float (&s)[BKV / 2]) {
  mbar_wait(&bars->s_full[x], j & 1);
  tc_fence_after();
  uint32_t* su = reinterpret_cast<uint32_t*>(s);
  tmem_ld_16x256b_x8(s_addr, su);
  tmem_ld_16x256b_x8(s_addr + 64, su + 32);
}
__device__ __forceinline__ void rq_load_finish(AttnBars* bars, int x) {
  tmem_ld_wait();
  tc_fence_before();
  mbar_arrive(&bars->s_free[x]);  // the tensor core may overwrite S_X with tile j+1 once all 256 threads arrived
}
__device__ __forceinline__ void rq_store_finish(AttnBars* bars, int x) {
  tmem_st_wait();
  tc_fence_before();
  mbar_arrive(&bars->p_full[x]);
}

// One KV tile of the online softmax for this warp's 16 rows; S_X(j) in registers: s[4k + {0,1}]: row A (t/4),
// keys 8k + 2c + {0,1};  s[4k + {2,3}]: row B (t/4 + 8).
template <bool MASK, bool FIRST, bool TRACE>
__device__ __forceinline__ void softmax_tile_rq(AttnBars* bars, int x, int c, int j, uint32_t s_addr, uint32_t o_addr,
                                                uint32_t p_addr, float scale_log2, int valid, float (&m_used)[2],
                                                float (&l_run)[2], const RegrowCtx& ctx, long long* trace) {
  ATTN_STAMP(0);
  float s[BKV / 2];
  rq_load_issue(bars, x, j, s_addr, s);
  rq_load_finish(bars, x);
#if CAP4D_ATTN_SCHED_FENCE
  sched_fence(bars);
#endif
  ATTN_STAMP(2);
  if (MASK) {
#pragma unroll
    for (int k = 0; k < BKV / 8; ++k) {
      const int col = 8 * k + 2 * c;
      if (col >= valid) s[4 * k] = s[4 * k + 2] = -INFINITY;
      if (col + 1 >= valid) s[4 * k + 1] = s[4 * k + 3] = -INFINITY;
    }
  }
  bool pv_waited = FIRST;  // FIRST <=> j == 0
#if CAP4D_ATTN_DIAG_NOMAX
  if (FIRST) m_used[0] = m_used[1] = 0.f;
#else
  if (FIRST || !CAP4D_ATTN_LAZYMAX) rq_exact_stabiliser(s, bars, x, j, o_addr, scale_log2, m_used, l_run, pv_waited);
#endif
#if CAP4D_ATTN_TURNS
  {
    const int tq = (threadIdx.x >> 5) & 3;
    if (x == 0) {
      if (!FIRST) mbar_wait(&bars->turn[0][tq], (j - 1) & 1);  // tile B's warps of this scheduler are through exp(j-1)
    } else {
      mbar_wait(&bars->turn[NQT - 1][tq], j & 1);              // tile A's are through exp(j)
    }
  }
#endif
  ATTN_STAMP(3);
  uint32_t pk[BKV / 4];
  float rsA, rsB;
  // TURNS: hand the MUFU to the other tile - after all exponentials (1) or after half of them (2: the handover then
  // overlaps the second half).  The predicate depends on the exponentials, so ptxas cannot hoist the arrive.
  rq_exp(s, scale_log2, m_used, pk, rsA, rsB, [&](float half_sum) {
#if CAP4D_ATTN_TURNS == 2
    if ((threadIdx.x & 31) == 0 && !(half_sum < -1.0f)) mbar_arrive(&bars->turn[(NQT - 1) - x][(threadIdx.x >> 5) & 3]);
#endif
  });
#if CAP4D_ATTN_TURNS == 1
  if ((threadIdx.x & 31) == 0 && !(rsA + rsB < -1.0f)) mbar_arrive(&bars->turn[(NQT - 1) - x][(threadIdx.x >> 5) & 3]);
#endif
#if CAP4D_ATTN_LAZYMAX && !CAP4D_ATTN_DIAG_NOMAX
  if (!FIRST) {
    const bool big = !(rsA < GROW_SUM) || !(rsB < GROW_SUM);  // catches inf and NaN as well
    if (__any_sync(0xffffffffu, big)) {
      RegrowIO io;
#pragma unroll
      for (int i = 0; i < BKV / 4; ++i) io.pk[i] = pk[i];
      io.m_used[0] = m_used[0], io.m_used[1] = m_used[1], io.l_run[0] = l_run[0], io.l_run[1] = l_run[1];
      io.rs[0] = rsA, io.rs[1] = rsB, io.pv_waited = pv_waited;
      io = rq_regrow(io, ctx, bars, x, c, j, o_addr, scale_log2, valid);
#pragma unroll
      for (int i = 0; i < BKV / 4; ++i) pk[i] = io.pk[i];
      m_used[0] = io.m_used[0], m_used[1] = io.m_used[1], l_run[0] = io.l_run[0], l_run[1] = io.l_run[1];
      rsA = io.rs[0], rsB = io.rs[1], pv_waited = io.pv_waited != 0;
    }
  }
#endif
  l_run[0] += rsA;
  l_run[1] += rsB;
  ATTN_STAMP(4);
  if (!pv_waited) {
    // P_X may only be overwritten once the PV MMA of tile j-1 has consumed it (long done by now)
    mbar_wait(&bars->pv_full[x], (j - 1) & 1);
    tc_fence_after();
  }
  tmem_st_16x128b_x16(p_addr, pk);
  rq_store_finish(bars, x);
  ATTN_STAMP(5);
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
    mbar_init(&bars->always, 1);
    for (int i = 0; i < NQT; ++i)
      for (int k = 0; k < 4; ++k) mbar_init(&bars->turn[i][k], 2);
    fence_mbar_init();
    mbar_arrive(&bars->always);
#if CAP4D_ATTN_EARLY_TMA
    // Q and the first K / V tiles are requested before the TMEM allocation and the block-wide barrier below: their
    // latency is most of a CTA's start-up, and nothing they touch depends on either
    mbar_arrive_expect_tx(&bars->q_full, NQT * TILE_BYTES);
    for (int x = 0; x < NQT; ++x)
      tma_load_2d(sQ + x * TILE_BYTES, &tmQKV, &bars->q_full, head * HD, row_base + q0 + x * BQ);
    mbar_arrive_expect_tx(&bars->k_full[0], TILE_BYTES);
    tma_load_2d(sK, &tmQKV, &bars->k_full[0], p.C + head * HD, row_base);
    mbar_arrive_expect_tx(&bars->v_full[0], TILE_BYTES);
    tma_load_2d(sV, &tmQKV, &bars->v_full[0], 2 * p.C + head * HD, row_base);
#endif
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
#if !CAP4D_ATTN_EARLY_TMA
      mbar_arrive_expect_tx(&bars->q_full, NQT * TILE_BYTES);
      for (int x = 0; x < NQT; ++x)
        tma_load_2d(sQ + x * TILE_BYTES, &tmQKV, &bars->q_full, head * HD, row_base + q0 + x * BQ);
#endif
      for (int j = CAP4D_ATTN_EARLY_TMA ? 1 : 0; j < nkv; ++j) {
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
#if !CAP4D_ATTN_DIAG_NOQK
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16(s_tmem, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0);
#endif
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
#if !CAP4D_ATTN_DIAG_NOPV
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {
          // A: 16 keys = 8 TMEM columns per step; B: 16 kv rows = 2048 B per step
          const uint64_t bd = vdesc + static_cast<uint64_t>((kk * 16 * 128) >> 4);
          umma_bf16_ts(o_tmem, p_tmem + kk * 8, bd, idesc_pv, (j | kk) != 0);
        }
#endif
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
#if CAP4D_ATTN_ROWQUAD
    const int c = lane & 3;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32 + half * 16) << 16;
    const uint32_t s_addr = tmem_base + lane_addr + TM_S + x * BKV;
    const uint32_t o_addr = tmem_base + lane_addr + TM_O + x * HD;
    const uint32_t p_addr = tmem_base + lane_addr + TM_P + x * (BKV / 2);
    float m_used[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    if (NQT == 2 && x == 1 && CAP4D_ATTN_B_DELAY_NS > 0) __nanosleep(CAP4D_ATTN_B_DELAY_NS);
    const int valid_last = p.L - (nkv - 1) * BKV;  // valid keys in the last tile (1..128)
    const bool tail = valid_last < BKV;
    const int n_main = tail ? nkv - 1 : nkv;
    long long* trace = nullptr;
    if (TRACE && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (threadIdx.x & 127) == 0 && half == 0)
      trace = p.trace;
    RegrowCtx ctx;
#if CAP4D_ATTN_LAZYMAX
    {
      const int rA = q0 + x * BQ + q * 32 + half * 16 + (lane >> 2);
      const long long last = static_cast<long long>(p.M) - 1;
      const long long gA = min(static_cast<long long>(row_base + rA), last), gB = min(static_cast<long long>(row_base + rA + 8), last);
      ctx.ld = 3 * p.C;
      ctx.q_rowA = p.qkv + gA * ctx.ld + head * HD;
      ctx.q_rowB = p.qkv + gB * ctx.ld + head * HD;
      ctx.k_tile = p.qkv + static_cast<long long>(row_base) * ctx.ld + p.C + head * HD;
    }
    const bf16* k_seq = ctx.k_tile;
#endif
    // tile 0 (always the exact row max); it is also the masked tile when the sequence is shorter than one tile
    if (n_main > 0)
      softmax_tile_rq<false, true, TRACE>(bars, x, c, 0, s_addr, o_addr, p_addr, p.scale_log2, BKV, m_used, l_run, ctx, trace);
    else
      softmax_tile_rq<true, true, TRACE>(bars, x, c, 0, s_addr, o_addr, p_addr, p.scale_log2, valid_last, m_used, l_run, ctx, trace);
    for (int j = 1; j < n_main; ++j) {
#if CAP4D_ATTN_LAZYMAX
      ctx.k_tile = k_seq + static_cast<long long>(j) * BKV * ctx.ld;
#endif
      softmax_tile_rq<false, false, TRACE>(bars, x, c, j, s_addr, o_addr, p_addr, p.scale_log2, BKV, m_used, l_run, ctx, trace);
    }
    if (tail && nkv > 1) {
#if CAP4D_ATTN_LAZYMAX
      ctx.k_tile = k_seq + static_cast<long long>(nkv - 1) * BKV * ctx.ld;
#endif
      softmax_tile_rq<true, false, TRACE>(bars, x, c, nkv - 1, s_addr, o_addr, p_addr, p.scale_log2, valid_last, m_used, l_run, ctx, trace);
    }
    // row sums: the four threads of a row hold partial sums in the same stabiliser
    float lA = l_run[0], lB = l_run[1];
    lA += __shfl_xor_sync(0xffffffffu, lA, 1);
    lB += __shfl_xor_sync(0xffffffffu, lB, 1);
    lA += __shfl_xor_sync(0xffffffffu, lA, 2);
    lB += __shfl_xor_sync(0xffffffffu, lB, 2);
    const float invA = 1.0f / lA, invB = 1.0f / lB;
    mbar_wait(&bars->pv_full[x], (nkv - 1) & 1);  // O_X is complete once the last PV MMA has landed
    tc_fence_after();
    uint32_t ov[32];
    tmem_ld_16x256b_x8(o_addr, ov);
    tmem_ld_wait();
    const int qrowA = q0 + x * BQ + q * 32 + half * 16 + (lane >> 2), qrowB = qrowA + 8;
    bf16* dstA = p.out + static_cast<size_t>(row_base + qrowA) * p.C + head * HD + 2 * c;
    bf16* dstB = dstA + static_cast<size_t>(8) * p.C;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (qrowA < p.L)
        *reinterpret_cast<uint32_t*>(dstA + 8 * k) =
            pack_bf16x2(__uint_as_float(ov[4 * k]) * invA, __uint_as_float(ov[4 * k + 1]) * invA);
      if (qrowB < p.L)
        *reinterpret_cast<uint32_t*>(dstB + 8 * k) =
            pack_bf16x2(__uint_as_float(ov[4 * k + 2]) * invB, __uint_as_float(ov[4 * k + 3]) * invB);
    }
#else
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
#endif
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
  plan->qkv = qkv;
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
  p.M = plan.M;
  p.qkv = plan.qkv;
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

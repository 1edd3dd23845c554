// Persistent, warp-specialised tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   out[M, N] = epilogue( A[M, K] * W[N, K]^T ),  bf16 operands, fp32 accumulation in TMEM.
//
// This one kernel serves every dense contraction of the MMDM U-Net
// (reference: controlnet/ldm/modules/diffusionmodules/openaimodel.py:256-276 ResBlock convs,
//  :92-161 Up/Downsample convs, cap4d/mmdm/net/attention.py:68-95,168-178,356-371 linears):
//   * plain GEMM: A is a row-major [M][K] matrix (TMA 2-D map);
//   * 3x3 convolution: A is an NHWC activation tensor (TMA 4-D map, C innermost).  An M tile is
//     128 consecutive output pixels = a (box_n x box_h x W) box; filter tap (dy,dx) is the same box
//     shifted by (dy,dx) and TMA's out-of-bounds zero fill supplies the padding, so the im2col
//     matrix is never materialised.  Stride-2 convs read 4 parity planes through the same table;
//   * an optional second plain A segment appends K (the ResBlock's 1x1 skip conv is accumulated
//     into the same TMEM tile as its second 3x3 conv).
//
// CTA = 384 threads = 3 warpgroups: WG0 = warp 0 TMA producer, warp 1 MMA issuer (+TMEM owner), warps 2-3
// idle; WG1-2 = warps 4-11 epilogue.  setmaxnreg moves WG0's registers to the epilogue warpgroups (56 / 216).
// CTA tile (msub*128) x BN; smem ring of `stages` {A msub x 128x64, B BNx64} bf16 tiles (128B swizzle);
// two TMEM accumulator stages when they fit (msub*BN <= 256) so the epilogue of tile i overlaps the
// main loop of tile i+1.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "kernels.h"
#include "ptx.cuh"

namespace cap4d {

namespace {

constexpr int BM = 128;                     // rows of one UMMA (TMEM lanes)
constexpr int BK = 64;
constexpr int A_SUB_BYTES = BM * BK * 2;    // 16 KiB: one 128-row A sub-tile
constexpr int MAX_STAGES = 8;
constexpr int GEMM_THREADS = 384;           // warp 0 TMA, warp 1 MMA, warps 2-3 idle, warps 4..11 epilogue
constexpr int EPI_WARP0 = 4;                // first epilogue warp
constexpr int REGS_CTRL = 56, REGS_EPI = 216;  // 128*56 + 256*216 = 62464 <= 65536
constexpr int TMEM_COLS = 512;

struct SmemTail {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

// EPI == 1 (residual GEMMs): one 32 x 32 fp32 transpose buffer per epilogue warp behind the barriers
constexpr int EPI_BUF_OFFSET = (static_cast<int>(sizeof(SmemTail)) + 127) & ~127;
constexpr int EPI_BUF_BYTES = 8 * 32 * 32 * 4;

template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

// 256-bit global accesses (sm_100): a thread owns a whole accumulator row, so its 32 columns are 4
// (fp32) or 2 (bf16) 32-byte pieces of ONE 128 B line; wider pieces halve the number of L1 wavefronts the
// row-per-thread pattern costs.
__device__ __forceinline__ void st_global_v8(void* dst, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void ld_global_nc_v8(const void* src, float* v) {
  asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(src));
}

__device__ __forceinline__ void epilogue_store_f32(float* dst, const float* acc) {
#pragma unroll
  for (int j = 0; j < 32; j += 8) st_global_v8(dst + j, reinterpret_cast<const uint32_t*>(acc + j));
}

__device__ __forceinline__ void epilogue_store_bf16(bf16* dst, const float* acc) {
#pragma unroll
  for (int j = 0; j < 32; j += 16) {
    uint32_t u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) u[i] = pack_bf16x2(acc[j + 2 * i], acc[j + 2 * i + 1]);
    st_global_v8(dst + j, u);
  }
}

__device__ __forceinline__ void load_vec32(float (&dst)[32], const float* __restrict__ src) {
#pragma unroll
  for (int j = 0; j < 32; j += 8) ld_global_nc_v8(src + j, dst + j);
}

// acc[0..31] += 32 consecutive floats at shared address `addr` (every lane reads the same words: broadcast)
__device__ __forceinline__ void add_vec32_shared(float* acc, uint32_t addr) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 t = ld_shared_v4(addr + j * 4);
    acc[j] += t.x, acc[j + 1] += t.y, acc[j + 2] += t.z, acc[j + 3] += t.w;
  }
}

__device__ __forceinline__ void add_vec32(float* acc, const float* __restrict__ src) {
  float b[32];
#pragma unroll
  for (int j = 0; j < 32; j += 8) ld_global_nc_v8(src + j, b + j);
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] += b[j];
}

// folded upsample: low-res pixel row (n, y, x) -> row of pixel (n, 2y+py, 2x+px) in the 2H x 2W output
__device__ __forceinline__ int up_row(int row, int H, int W, int py, int px) {
  const int hw = H * W;
  const int n = row / hw, rem = row - n * hw;
  const int y = rem / W, x = rem - y * W;
  return (n * 2 * H + 2 * y + py) * 2 * W + 2 * x + px;
}

// A CTA tile is (msub * 128) x BN: msub in {1, 2} 128-row sub-tiles share one B (weight) tile per k-block,
// which raises the FLOPs per byte streamed from L2 - the resource this kernel is bound by - from
// 2*128*BN*64 / (16K + 128 BN) to 2*256*BN*64 / (32K + 128 BN).
// EPI = 1, 2: plain GEMM + fp32 residual (no conv operand, no GEGLU, no per-image bias): their own instantiations so
// that the residual look-ahead buffers below do not disturb the register allocation of the generic epilogue.
// 2 = the HBM-bound shapes (short K: to_out / proj_out, FF2 at C = 320), coalescing epilogue through shared memory;
// 1 = the tensor-bound ones (row-per-thread accesses, no shared-memory buffer: one more pipeline stage).
template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled tiles need 1024 B alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_stage_bytes = p.msub * A_SUB_BYTES;
  const int b_stage_bytes = p.BN * BK * 2;
  uint8_t* sA = smem;
  uint8_t* sB = smem + p.stages * a_stage_bytes;
  SmemTail* tail = reinterpret_cast<SmemTail*>(sB + p.stages * b_stage_bytes);
  float* epi_buf = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(tail) + EPI_BUF_OFFSET);  // EPI == 1 only

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.tiles_m * p.tiles_n;
  const int acc_stride = (p.n_acc == 2) ? (TMEM_COLS / 2) : 0;  // TMEM columns between accumulator stages

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&tail->full[s], 1);
      mbar_init(&tail->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tail->tmem_full[a], 1);
      mbar_init(&tail->tmem_empty[a], 256);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tail->tmem_base, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  if (warp < EPI_WARP0) {  // ---- WG0: control warps (the two role blocks below belong to this branch)
  setmaxnreg_dec<REGS_CTRL>();
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = a_stage_bytes + b_stage_bytes;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.tiles_n;
        const int n_tile = tile - m_tile * p.tiles_n;
        // per-tile coordinates of the (up to two) 128-row sub-tiles; nothing in the k loop divides
        int row0[2], n0[2], y0[2], x0[2];
        for (int sub = 0; sub < 2; ++sub) {
          row0[sub] = (m_tile * p.msub + sub) * BM;  // first output row (pixel) of the sub-tile
          n0[sub] = y0[sub] = x0[sub] = 0;
          if (p.a_conv) {
            const int hw = p.H * p.W;
            n0[sub] = row0[sub] / hw;
            y0[sub] = (row0[sub] - n0[sub] * hw) / p.W;
            x0[sub] = row0[sub] - n0[sub] * hw - y0[sub] * p.W;  // != 0 only when an image row is wider than a tile
          }
        }
        const int b_row = n_tile * p.BN;
        int tap = 0, c0 = 0;  // conv: filter tap and channel offset of the current k-block
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&tail->empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&tail->full[stage], tx_bytes);
          uint8_t* dstA = sA + stage * a_stage_bytes;
          if (kb < p.seg0_kb) {
            if (p.a_conv) {
              const int dx = p.tap_dx[tap], dy = p.tap_dy[tap], dn = p.tap_dn[tap];
              tma_load_4d(dstA, &tmA, &tail->full[stage], c0, x0[0] + dx, y0[0] + dy, n0[0] + dn);
              if (p.msub == 2)
                tma_load_4d(dstA + A_SUB_BYTES, &tmA, &tail->full[stage], c0, x0[1] + dx, y0[1] + dy, n0[1] + dn);
              c0 += BK;
              if (c0 == p.cin_kb * BK) {
                c0 = 0;
                ++tap;
              }
            } else {
              tma_load_2d(dstA, &tmA, &tail->full[stage], kb * BK, row0[0]);
              if (p.msub == 2) tma_load_2d(dstA + A_SUB_BYTES, &tmA, &tail->full[stage], kb * BK, row0[1]);
            }
          } else {
            tma_load_2d(dstA, &tmA2, &tail->full[stage], (kb - p.seg0_kb) * BK, row0[0]);
            if (p.msub == 2) tma_load_2d(dstA + A_SUB_BYTES, &tmA2, &tail->full[stage], (kb - p.seg0_kb) * BK, row0[1]);
          }
          tma_load_2d(sB + stage * b_stage_bytes, &tmB, &tail->full[stage], kb * BK, b_row);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(BM, p.BN, 0, p.a_f16, p.b_f16);
      const uint64_t adesc0 = umma_smem_desc_sw128(smem_u32(sA));  // stage / sub-tile / k offsets are added
      const uint64_t bdesc0 = umma_smem_desc_sw128(smem_u32(sB));  // to the 14-bit (>>4) address field
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int acc = (p.n_acc == 2) ? (it & 1) : 0;
        const uint32_t acc_phase = (p.n_acc == 2) ? ((it >> 1) & 1) : (it & 1);
        mbar_wait(&tail->tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * acc_stride;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&tail->full[stage], phase);
          tc_fence_after();
          const uint64_t bdesc = bdesc0 + static_cast<uint64_t>((stage * b_stage_bytes) >> 4);
          const uint64_t adesc = adesc0 + static_cast<uint64_t>((stage * a_stage_bytes) >> 4);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in the (>>4) address field
            umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          if (p.msub == 2) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16(d_tmem + p.BN, adesc + (A_SUB_BYTES >> 4) + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&tail->empty[stage]);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tail->tmem_full[acc]);
      }
    }
  }
  } else {  // ---- WG1-2: epilogue warps
    setmaxnreg_inc<REGS_EPI>();
    // ===================== epilogue (warps 4..11) =====================
    // Thread <-> accumulator row (TMEM lane); two warps share each TMEM lane quarter and split the
    // tile's column chunks, so every SM sub-partition has two epilogue warps to hide the
    // tcgen05.ld / global-load latencies.  Each thread reads / writes whole 128 B (fp32) or 64 B (bf16)
    // row segments; the next chunk's tcgen05.ld is in flight while the current one is written out.
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - EPI_WARP0) >> 2;  // which half of the chunks this warp handles
    const bool geglu = (p.out_mode & 15) == OUT_GEGLU_BF16;
    const bool dbg_noepi = (p.out_mode & 32) != 0;
    const int tcols = geglu ? 64 : 32;         // TMEM columns per chunk
    const int cps = p.BN / tcols;              // chunks per sub-tile
    const int nchunks = p.msub * cps;
    const int c_begin = half == 0 ? 0 : (nchunks + 1) / 2;
    const int c_end = half == 0 ? (nchunks + 1) / 2 : nchunks;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int m_tile = tile / p.tiles_n;
      const int n_tile = tile - m_tile * p.tiles_n;
      const int acc = (p.n_acc == 2) ? (it & 1) : 0;
      const uint32_t acc_phase = (p.n_acc == 2) ? ((it >> 1) & 1) : (it & 1);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * acc_stride;
      auto chunk_col = [&](int cc) { return (cc / cps) * p.BN + (cc % cps) * tcols; };  // TMEM column of chunk cc
      if constexpr (EPI == 2) {
        // Plain GEMM + fp32 residual (to_out, FF2, proj_out: K is short, the kernel IS its epilogue).  A thread
        // owns an accumulator ROW, and row-per-thread global accesses cost one L1 wavefront per thread and
        // instruction (ncu, round 2: LSU wavefronts 62 % of peak - the busiest unit of the launch - with DRAM at
        // 48 %).  So the accumulator chunk (32 rows x 32 columns per warp) goes through a swizzled 4 KB
        // shared-memory buffer into the coalesced layout - eight threads per row, four rows per instruction:
        // 4 lines per request instead of 32 - in which the residual is read and the result written.  The residual
        // does not depend on the accumulator: chunk cc+1's is requested before chunk cc is read from TMEM (the
        // first one before the tile's MMAs have finished): two buffers, used alternately.
        const uint32_t tb = smem_u32(epi_buf) + (warp - EPI_WARP0) * 4096;  // this warp's 32 x 32 fp32 buffer
        const int lr = lane >> 3, lc = lane & 7;  // coalesced layout: element i = row 4 i + lr, columns 4 lc .. 4 lc + 3
        float4 r0[8], r1[8];
        // (sub, c) of the chunk in hand and of the next one, advanced without dividing
        int sub = c_begin / cps, c = c_begin - sub * cps;
        const int row_warp = m_tile * p.msub * BM + q * 32;
        const float* res_col = p.residual + n_tile * p.BN + lc * 4;
        const bool has_res = p.residual != nullptr;  // proj_in takes this epilogue without one
        auto res_load = [&](int sb, int ch, float4(&r)[8]) {
          const int rb = row_warp + sb * BM + lr;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has_res && rb + 4 * i < p.M)
              r[i] = __ldg(reinterpret_cast<const float4*>(res_col + static_cast<size_t>(rb + 4 * i) * p.ldr + ch * 32));
          }
        };
        if (c_begin < c_end) res_load(sub, c, r0);
        mbar_wait(&tail->tmem_full[acc], acc_phase);
        tc_fence_after();
        uint32_t v[32];
        if (c_begin < c_end) tmem_ld32(taddr + sub * p.BN + c * 32, v);
        auto process = [&](int cc, float4(&cur)[8], float4(&nxt)[8]) {
          int sub_n = sub, c_n = c + 1;
          if (c_n == cps) {
            c_n = 0;
            ++sub_n;
          }
          const bool more = cc + 1 < c_end;
          if (more) res_load(sub_n, c_n, nxt);
          tmem_ld_wait();
          // own row -> buffer; 16 B pieces XOR-swizzled by the row so that both directions are conflict-free
#pragma unroll
          for (int k = 0; k < 8; ++k)
            st_shared_v4(tb + lane * 128 + ((k ^ (lane & 7)) << 4), v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
          if (more) tmem_ld32(taddr + sub_n * p.BN + c_n * 32, v);
          __syncwarp();
          const int col0 = n_tile * p.BN + c * 32 + lc * 4;
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.bias != nullptr) bv = __ldg(reinterpret_cast<const float4*>(p.bias + col0));
          const int rb = row_warp + sub * BM + lr;
          float4 acc4[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rl = 4 * i + lr;
            acc4[i] = ld_shared_v4(tb + rl * 128 + ((lc ^ (rl & 7)) << 4));
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 a = acc4[i];
            if (rb + 4 * i < p.M) {
              // (acc + residual) + bias, the order of every other epilogue: all tile configurations stay bit-identical
              a.x = (a.x + cur[i].x) + bv.x, a.y = (a.y + cur[i].y) + bv.y;
              a.z = (a.z + cur[i].z) + bv.z, a.w = (a.w + cur[i].w) + bv.w;
              const size_t off = static_cast<size_t>(rb + 4 * i) * p.ldo + col0;
              if ((p.out_mode & 15) == OUT_F32) {
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off) = a;
              } else {
                *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.out) + off) =
                    make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
              }
            }
          }
          __syncwarp();  // the buffer is rewritten by the next chunk
          sub = sub_n;
          c = c_n;
        };
        for (int cc = c_begin; cc < c_end; cc += 2) {
          process(cc, r0, r1);
          if (cc + 1 < c_end) process(cc + 1, r1, r0);
        }
      } else if constexpr (EPI == 1) {
        // Plain GEMM + fp32 residual (to_out, FF2, proj_out: K is short, the kernel IS its epilogue, and ncu
        // shows it waiting on the residual rows: ~4 KB in flight per warp).  The residual does not depend on
        // the accumulator, so chunk cc+1's row segment is requested before chunk cc is even read from TMEM
        // (the first one before the tile's MMAs have finished): two buffers, used alternately.
        float r0[32], r1[32];
        // (sub, c) of the chunk in hand and of the next one, advanced without dividing
        int sub = c_begin / cps, c = c_begin - sub * cps;
        const int row_lane = m_tile * p.msub * BM + q * 32 + lane;
        const float* res_col = p.residual + n_tile * p.BN;
        auto res_load = [&](int sb, int ch, float(&r)[32]) {
          const int ri = row_lane + sb * BM;
          if (ri < p.M) load_vec32(r, res_col + static_cast<size_t>(ri) * p.ldr + ch * 32);
        };
        if (c_begin < c_end) res_load(sub, c, r0);
        mbar_wait(&tail->tmem_full[acc], acc_phase);
        tc_fence_after();
        uint32_t v[32];
        if (c_begin < c_end) tmem_ld32(taddr + sub * p.BN + c * 32, v);
        auto process = [&](int cc, float(&cur)[32], float(&nxt)[32]) {
          int sub_n = sub, c_n = c + 1;
          if (c_n == cps) {
            c_n = 0;
            ++sub_n;
          }
          const bool more = cc + 1 < c_end;
          if (more) res_load(sub_n, c_n, nxt);
          tmem_ld_wait();
          float a[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) a[j] = __uint_as_float(v[j]);
          if (more) tmem_ld32(taddr + sub_n * p.BN + c_n * 32, v);
          const int row = row_lane + sub * BM;
          if (row < p.M) {
            const int col0 = n_tile * p.BN + c * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j) a[j] += cur[j];
            if (p.bias != nullptr) add_vec32(a, p.bias + col0);
            if ((p.out_mode & 15) == OUT_F32) {
              epilogue_store_f32(reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + col0, a);
            } else {
              epilogue_store_bf16(reinterpret_cast<bf16*>(p.out) + static_cast<size_t>(row) * p.ldo + col0, a);
            }
          }
          sub = sub_n;
          c = c_n;
        };
        for (int cc = c_begin; cc < c_end; cc += 2) {
          process(cc, r0, r1);
          if (cc + 1 < c_end) process(cc + 1, r1, r0);
        }
      } else {
      // The tile's BN bias values are the same for every row.  Read from global memory inside the chunk loop they
      // were the top stall of the GEGLU GEMM (ncu: 37 % of all warp samples waiting on the bias adds); each epilogue
      // warp keeps its own copy of the slice in shared memory instead - no block-wide barrier (the staged version of
      // round 1 needed one and lost) - and requests the NEXT tile's slice before it starts on this tile's chunks.
      const bool bias_smem = p.bias_smem != 0 && p.bias != nullptr;
      const uint32_t wb = smem_u32(epi_buf) + (warp - EPI_WARP0) * 1024;  // 256 floats per warp
      float4 nb0 = make_float4(0.f, 0.f, 0.f, 0.f), nb1 = nb0;
      auto bias_request = [&](int n_t) {
        const float* bsrc = p.bias + n_t * p.BN;
        if (lane * 4 < p.BN) nb0 = __ldg(reinterpret_cast<const float4*>(bsrc + lane * 4));
        if (lane * 4 + 128 < p.BN) nb1 = __ldg(reinterpret_cast<const float4*>(bsrc + lane * 4 + 128));
      };
      auto bias_publish = [&]() {
        __syncwarp();  // every lane is done reading the previous slice
        st_shared_v4(wb + lane * 16, __float_as_uint(nb0.x), __float_as_uint(nb0.y), __float_as_uint(nb0.z), __float_as_uint(nb0.w));
        st_shared_v4(wb + 512 + lane * 16, __float_as_uint(nb1.x), __float_as_uint(nb1.y), __float_as_uint(nb1.z), __float_as_uint(nb1.w));
        __syncwarp();
      };
      if (bias_smem) {
        if (it == 0) {
          bias_request(n_tile);
          bias_publish();
        }
        const int next = tile + static_cast<int>(gridDim.x);
        if (next < total_tiles) bias_request(next - (next / p.tiles_n) * p.tiles_n);
      }
      mbar_wait(&tail->tmem_full[acc], acc_phase);
      tc_fence_after();
      uint32_t v[32], vg[32];
      if (c_begin < c_end) {
        tmem_ld32(taddr + chunk_col(c_begin), v);
        if (geglu) tmem_ld32(taddr + chunk_col(c_begin) + 32, vg);
      }
      for (int cc = c_begin; cc < c_end; ++cc) {
        const int sub = cc / cps, c = cc - sub * cps;
        const int row_in = (m_tile * p.msub + sub) * BM + q * 32 + lane;
        const bool row_ok = row_in < p.M && !dbg_noepi;
        const int row = (p.up_py < 0) ? row_in : up_row(row_in, p.H, p.W, p.up_py, p.up_px);
        tmem_ld_wait();
        float a[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) a[j] = __uint_as_float(v[j]);
        if (geglu) {
          // weight rows are interleaved in blocks of 32: [x(32) | gate(32)] -> 32 output columns
          float ag[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) ag[j] = __uint_as_float(vg[j]);
          if (cc + 1 < c_end) {
            tmem_ld32(taddr + chunk_col(cc + 1), v);
            tmem_ld32(taddr + chunk_col(cc + 1) + 32, vg);
          }
          if (row_ok) {
            const int col0 = n_tile * p.BN + c * 64;
            if (bias_smem) {
              add_vec32_shared(a, wb + c * 256);
              add_vec32_shared(ag, wb + c * 256 + 128);
            } else if (p.bias != nullptr) {
              add_vec32(a, p.bias + col0);
              add_vec32(ag, p.bias + col0 + 32);
            }
#pragma unroll
            for (int j = 0; j < 32; j += 2) geglu_pair(a[j], a[j + 1], ag[j], ag[j + 1]);
            epilogue_store_bf16(reinterpret_cast<bf16*>(p.out) + static_cast<size_t>(row) * p.ldo + (col0 >> 1), a);
          }
        } else {
          if (cc + 1 < c_end) tmem_ld32(taddr + chunk_col(cc + 1), v);
          if (row_ok) {
            const int col0 = n_tile * p.BN + c * 32;
            if (p.residual != nullptr) add_vec32(a, p.residual + static_cast<size_t>(row) * p.ldr + col0);
            if (p.rowbias != nullptr)
              add_vec32(a, p.rowbias + static_cast<size_t>(row / p.rowbias_div) * p.rowbias_ld + col0);
            if (bias_smem) add_vec32_shared(a, wb + c * 128);
            else if (p.bias != nullptr) add_vec32(a, p.bias + col0);
            if ((p.out_mode & 15) == OUT_F32) {
              epilogue_store_f32(reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + col0, a);
            } else {
              epilogue_store_bf16(reinterpret_cast<bf16*>(p.out) + static_cast<size_t>(row) * p.ldo + col0, a);
            }
          }
        }
      }
      if (bias_smem && tile + static_cast<int>(gridDim.x) < total_tiles) bias_publish();
      }  // generic epilogue
      tc_fence_before();
      mbar_arrive(&tail->tmem_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2).  A cluster of two CTAs computes a 256 x BN tile with ONE UMMA per
// k-step: CTA r stages rows [128 r, 128 r + 128) of A and rows [r BN/2, (r+1) BN/2) of the weight tile
// in its own shared memory, the leader's MMA thread issues tcgen05.mma.cta_group::2 (M = 256), and each
// CTA's TMEM receives its own 128 rows of the accumulator.  Per CTA and k-block the shared memory then
// sees 16 KiB + BN*64 B written and read instead of 16 KiB + BN*128 B - the single-CTA kernel is bound
// by exactly that traffic (see pick_tile) - so the tensor pipe can run at (or near) its peak rate.
// Barriers: both CTAs' TMA loads signal the LEADER's full[stage]; the leader's commits are multicast to
// both CTAs' empty[stage] / tmem_full[acc]; both CTAs' epilogue threads arrive on the leader's
// tmem_empty[acc].
// ---------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_2sm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                   const __grid_constant__ CUtensorMap tmB, const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_stage_bytes = A_SUB_BYTES;
  const int b_stage_bytes = (p.BN >> 1) * BK * 2;  // this CTA's half of the weight tile
  uint8_t* sA = smem;
  uint8_t* sB = smem + p.stages * a_stage_bytes;
  SmemTail* tail = reinterpret_cast<SmemTail*>(sB + p.stages * b_stage_bytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader_cta = rank == 0;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int pair_tiles = p.tiles_m * p.tiles_n;  // tiles_m counts 256-row pair tiles here
  const int acc_stride = (p.n_acc == 2) ? (TMEM_COLS / 2) : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&tail->full[s], 1);   // leader producer's arrive.expect_tx (+ the tx bytes of both CTAs)
      mbar_init(&tail->empty[s], 1);  // leader MMA commit (multicast)
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tail->tmem_full[a], 1);     // leader MMA commit (multicast)
      mbar_init(&tail->tmem_empty[a], 512);  // epilogue threads of BOTH CTAs (only the leader's copy is used)
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(&tail->tmem_base, TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers and TMEM allocations exist before anything crosses the pair
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  if (warp < EPI_WARP0) {  // ---- WG0: control warps (the two role blocks below belong to this branch)
  setmaxnreg_dec<REGS_CTRL>();
  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = 2u * (a_stage_bytes + b_stage_bytes);  // both CTAs' loads land on the leader's barrier
      const int b_half = p.BN >> 1;
      for (int pt = cluster_id; pt < pair_tiles; pt += n_clusters) {
        const int m_tile = pt / p.tiles_n;
        const int n_tile = pt - m_tile * p.tiles_n;
        const int row0 = m_tile * 2 * BM + static_cast<int>(rank) * BM;  // first output row of this CTA's half
        int n0 = 0, y0 = 0, x0 = 0;
        if (p.a_conv) {
          const int hw = p.H * p.W;
          n0 = row0 / hw;
          y0 = (row0 - n0 * hw) / p.W;
          x0 = row0 - n0 * hw - y0 * p.W;
        }
        const int b_row = n_tile * p.BN + static_cast<int>(rank) * b_half;
        int tap = 0, c0 = 0;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&tail->empty[stage], phase ^ 1);
          if (leader_cta) mbar_arrive_expect_tx(&tail->full[stage], tx_bytes);
          uint8_t* dstA = sA + stage * a_stage_bytes;
          if (kb < p.seg0_kb) {
            if (p.a_conv) {
              tma_load_4d_2sm(dstA, &tmA, &tail->full[stage], c0, x0 + p.tap_dx[tap], y0 + p.tap_dy[tap],
                              n0 + p.tap_dn[tap]);
              c0 += BK;
              if (c0 == p.cin_kb * BK) {
                c0 = 0;
                ++tap;
              }
            } else {
              tma_load_2d_2sm(dstA, &tmA, &tail->full[stage], kb * BK, row0);
            }
          } else {
            tma_load_2d_2sm(dstA, &tmA2, &tail->full[stage], (kb - p.seg0_kb) * BK, row0);
          }
          tma_load_2d_2sm(sB + stage * b_stage_bytes, &tmB, &tail->full[stage], kb * BK, b_row);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && leader_cta) {
      const uint32_t idesc = umma_idesc_bf16(2 * BM, p.BN, 0, p.a_f16, p.b_f16);
      const uint64_t adesc0 = umma_smem_desc_sw128(smem_u32(sA));
      const uint64_t bdesc0 = umma_smem_desc_sw128(smem_u32(sB));
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int pt = cluster_id; pt < pair_tiles; pt += n_clusters, ++it) {
        const int acc = (p.n_acc == 2) ? (it & 1) : 0;
        const uint32_t acc_phase = (p.n_acc == 2) ? ((it >> 1) & 1) : (it & 1);
        mbar_wait(&tail->tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * acc_stride;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&tail->full[stage], phase);
          tc_fence_after();
          const uint64_t adesc = adesc0 + static_cast<uint64_t>((stage * a_stage_bytes) >> 4);
          const uint64_t bdesc = bdesc0 + static_cast<uint64_t>((stage * b_stage_bytes) >> 4);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_bf16_2sm(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit_2sm(&tail->empty[stage], 0x3);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2sm(&tail->tmem_full[acc], 0x3);
      }
    }
  }
  } else {  // ---- WG1-2: epilogue warps
    setmaxnreg_inc<REGS_EPI>();
    // ===================== epilogue (warps 4..11, both CTAs: own 128 rows) =====================
    const int q = warp & 3;
    const int half = (warp - EPI_WARP0) >> 2;
    const bool geglu = (p.out_mode & 15) == OUT_GEGLU_BF16;
    const int tcols = geglu ? 64 : 32;
    const int nchunks = p.BN / tcols;
    const int c_begin = half == 0 ? 0 : (nchunks + 1) / 2;
    const int c_end = half == 0 ? (nchunks + 1) / 2 : nchunks;
    int it = 0;
    for (int pt = cluster_id; pt < pair_tiles; pt += n_clusters, ++it) {
      const int m_tile = pt / p.tiles_n;
      const int n_tile = pt - m_tile * p.tiles_n;
      const int acc = (p.n_acc == 2) ? (it & 1) : 0;
      const uint32_t acc_phase = (p.n_acc == 2) ? ((it >> 1) & 1) : (it & 1);
      const int row_in = m_tile * 2 * BM + static_cast<int>(rank) * BM + q * 32 + lane;
      const bool row_ok = row_in < p.M;
      const int row = (p.up_py < 0) ? row_in : up_row(row_in, p.H, p.W, p.up_py, p.up_px);
      mbar_wait(&tail->tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * acc_stride;
      uint32_t v[32], vg[32];
      if (c_begin < c_end) {
        tmem_ld32(taddr + c_begin * tcols, v);
        if (geglu) tmem_ld32(taddr + c_begin * tcols + 32, vg);
      }
      for (int c = c_begin; c < c_end; ++c) {
        tmem_ld_wait();
        float a[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) a[j] = __uint_as_float(v[j]);
        if (geglu) {
          float ag[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) ag[j] = __uint_as_float(vg[j]);
          if (c + 1 < c_end) {
            tmem_ld32(taddr + (c + 1) * tcols, v);
            tmem_ld32(taddr + (c + 1) * tcols + 32, vg);
          }
          if (row_ok) {
            const int col0 = n_tile * p.BN + c * 64;
            if (p.bias != nullptr) {
              add_vec32(a, p.bias + col0);
              add_vec32(ag, p.bias + col0 + 32);
            }
#pragma unroll
            for (int j = 0; j < 32; j += 2) geglu_pair(a[j], a[j + 1], ag[j], ag[j + 1]);
            epilogue_store_bf16(reinterpret_cast<bf16*>(p.out) + static_cast<size_t>(row) * p.ldo + (col0 >> 1), a);
          }
        } else {
          if (c + 1 < c_end) tmem_ld32(taddr + (c + 1) * tcols, v);
          if (row_ok) {
            const int col0 = n_tile * p.BN + c * 32;
            if (p.residual != nullptr) add_vec32(a, p.residual + static_cast<size_t>(row) * p.ldr + col0);
            if (p.rowbias != nullptr)
              add_vec32(a, p.rowbias + static_cast<size_t>(row / p.rowbias_div) * p.rowbias_ld + col0);
            if (p.bias != nullptr) add_vec32(a, p.bias + col0);
            if ((p.out_mode & 15) == OUT_F32) {
              epilogue_store_f32(reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + col0, a);
            } else {
              epilogue_store_bf16(reinterpret_cast<bf16*>(p.out) + static_cast<size_t>(row) * p.ldo + col0, a);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive_leader(&tail->tmem_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody leaves (or frees TMEM) while the pair may still touch its smem / barriers / TMEM
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

}  // namespace

// bf16 tensor map with 128B swizzle; dims/strides innermost first; strides in elements.
bool make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_el,
                    const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    return false;
  }
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_el[i] * 2;  // bytes
  }
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu box %u %u base %p",
             static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
             box[0], rank > 1 ? box[1] : 0, base);
    set_error(buf);
    return false;
  }
  return true;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

namespace {

struct TileCfg {
  int msub, bn, n_acc, two_cta;
};

// Pick (msub, BN).  Measured on B200 (scripts/bench_shapes.py, profiles/): the main loop of this
// kernel is bound by SHARED-MEMORY bandwidth (128 B/clk/SM), not by L2 or the tensor pipe: per
// k-block the TMA writes msub*16 KiB (A) + BN*128 B (B) into smem and the UMMAs read msub*16 KiB (A)
// + msub*BN*128 B (B is re-read by every 128-row sub-tile).  Cost = waves * k-blocks *
// max(tensor cycles, smem bytes / 128) + the epilogue where it cannot overlap (single TMEM stage).
TileCfg pick_tile(int M, int N, int num_kb, bool geglu) {
  static const int cands[] = {256, 224, 192, 160, 128, 96, 64, 32};
  const int sms = sm_count();
  TileCfg best{0, 0, 0, 0};
  if (const char* f = getenv("CAP4D_GEMM_FORCE")) {  // "msub,bn[,two_cta]" for experiments
    int ms = 0, bn = 0, tc = 0;
    const int nf = sscanf(f, "%d,%d,%d", &ms, &bn, &tc);
    if (nf >= 2 && (ms == 1 || ms == 2) && bn > 0 && N % bn == 0 && bn % 32 == 0 && bn <= 256 &&
        (!geglu || bn % 64 == 0)) {
      if (nf == 3 && tc == 1 && bn % 64 == 0) return TileCfg{1, bn, 2, 1};
      return TileCfg{ms, bn, (ms * bn <= 256) ? 2 : 1, 0};
    }
  }
  const bool allow_2cta = getenv("CAP4D_GEMM_NO_2CTA") == nullptr;
  double best_cost = 1e30;
  for (int msub = 1; msub <= 2; ++msub) {
    const int tiles_m = (M + msub * BM - 1) / (msub * BM);
    for (int bn : cands) {
      if (N % bn != 0) continue;
      if (geglu && (bn % 64 != 0)) continue;
      const int n_acc = (msub * bn <= 256) ? 2 : 1;
      const long tiles = static_cast<long>(tiles_m) * (N / bn);
      const long waves = (tiles + sms - 1) / sms;
      const double mma = 2.0 * msub * bn;
      const double smem = (2.0 * msub * 16384.0 + (1.0 + msub) * bn * 128.0) / 128.0;
      const double epi = 8.0 * msub * bn;  // cycles to drain one tile's accumulator (8 epilogue warps)
      double per_tile = num_kb * std::max(mma, smem) + 600.0;
      per_tile = (n_acc == 2) ? std::max(per_tile, epi) : per_tile + epi;
      const double cost = waves * per_tile;
      if (cost < best_cost * 0.999) {
        best_cost = cost;
        best = TileCfg{msub, bn, n_acc, 0};
      }
    }
  }
  // CTA pairs: 256 x BN per cluster, 74 clusters; per CTA and k-block 2*(16 KiB + BN*64 B) of smem traffic.
  // Measured (scripts/bench_shapes.py with CAP4D_GEMM_FORCE=1,256,1): the pair kernel reaches ~1400 TFLOP/s
  // at BN = 256, 1.45x the time this model predicts (the tensor pipe tops out near 60 % of its nominal rate
  // under the power cap), so it only wins where BN = 256 divides N and the pair tiles fill the 74 clusters.
  if (allow_2cta) {
    const int slots = std::max(1, sms / 2);
    const int tiles_m = (M + 2 * BM - 1) / (2 * BM);
    for (int bn : cands) {
      if (N % bn != 0 || bn % 64 != 0) continue;  // each CTA's half must stay a multiple of the 8-row swizzle group
      const long tiles = static_cast<long>(tiles_m) * (N / bn);
      const long waves = (tiles + slots - 1) / slots;
      const double mma = 2.0 * bn;
      const double smem = (2.0 * 16384.0 + 2.0 * bn * 64.0) / 128.0;
      const double epi = 8.0 * bn;
      const double per_tile = std::max(1.45 * num_kb * std::max(mma, smem) + 800.0, epi);
      const double cost = waves * per_tile;
      if (cost < best_cost * 0.99) {
        best_cost = cost;
        best = TileCfg{1, bn, 2, 1};
      }
    }
  }
  return best;
}

// Derive everything that depends on the tile configuration: tile counts, smem ring, grid, weight tensor map.
// which problems take the EPI = 1 instantiation (see gemm_tc_kernel)
int env_int_or(const char* name, int dflt) {
  const char* e = getenv(name);
  return e != nullptr ? atoi(e) : dflt;
}

// plain linear layer: no conv operand, no GEGLU, no per-image bias, no phase scatter
bool plain_linear(const GemmParams& p) {
  return !p.a_conv && (p.out_mode & 15) != OUT_GEGLU_BF16 && (p.out_mode & 32) == 0 && p.rowbias == nullptr && p.up_py < 0;
}
// 0: generic epilogue, 1: residual look-ahead, 2: (residual look-ahead +) coalescing epilogue.  2 is for the
// HBM-bound shapes - FLOPs per algorithmic byte below the ridge: to_out / proj_out at every level, FF2 and proj_in
// at C = 320 (measured: 4.2 -> 4.6-5.0 TB/s on the K = C shapes, profiles/r02b_residual_gemm.log); the K = 4C shapes
// at C >= 640 are tensor-bound and would lose a pipeline stage to the 32 KB buffer.  CAP4D_GEMM_EPI=0|1|2 forces
// (1 needs a residual), CAP4D_GEMM_NO_LOOKAHEAD is the same as 0.
int epilogue_kind(const GemmParams& p, int N, int Ktot) {
  static const bool off = getenv("CAP4D_GEMM_NO_LOOKAHEAD") != nullptr;
  static const char* force = getenv("CAP4D_GEMM_EPI");
  if (off || !plain_linear(p)) return 0;
  const bool res = p.residual != nullptr, f32 = (p.out_mode & 15) == OUT_F32;
  if (force != nullptr && force[0] >= '0' && force[0] <= '2') return (force[0] == '1' && !res) ? 0 : force[0] - '0';
  if (!res && !f32) return 0;  // QKV: bf16 rows are 64 B per chunk, see the measurement in DESIGN.md
  const double bytes_per_row = 2.0 * Ktot + (f32 ? 4.0 : 2.0) * N + (res ? 4.0 * N : 0.0);
  const double intensity = 2.0 * N * Ktot / bytes_per_row;
  return intensity < 200.0 ? 2 : (res ? 1 : 0);
}

bool apply_cfg(GemmPlan* plan, const TileCfg& cfg, const bf16* Wt, int N, int Ktot) {
  GemmParams& p = plan->p;
  const int bn = cfg.bn;
  p.N = N;
  p.BN = bn;
  p.msub = cfg.msub;
  p.n_acc = cfg.n_acc;
  plan->two_cta = cfg.two_cta;
  const int rows_per_tile = cfg.two_cta ? 2 * BM : cfg.msub * BM;
  p.tiles_m = (p.M + rows_per_tile - 1) / rows_per_tile;
  p.tiles_n = N / bn;
  p.num_kb = Ktot / BK;
  const int stage_bytes = cfg.two_cta ? (A_SUB_BYTES + (bn / 2) * BK * 2) : (cfg.msub * A_SUB_BYTES + bn * BK * 2);
  p.epi = cfg.two_cta ? 0 : epilogue_kind(p, N, Ktot);
  // generic epilogue: per-warp bias slices in shared memory (CAP4D_GEMM_BIAS_SMEM: 0 off, 1 GEGLU only, 2 every bias)
  static const int bias_mode = env_int_or("CAP4D_GEMM_BIAS_SMEM", 2);
  const bool is_geglu = (p.out_mode & 15) == OUT_GEGLU_BF16;
  p.bias_smem = (!cfg.two_cta && p.epi == 0 && p.bias != nullptr && (p.out_mode & 32) == 0 &&
                 (bias_mode == 2 || (bias_mode == 1 && is_geglu))) ? 1 : 0;
  const int epi_extra = p.epi == 2 ? EPI_BUF_OFFSET + EPI_BUF_BYTES : (p.bias_smem ? EPI_BUF_OFFSET + 8 * 1024 : 0);
  int stages = (220 * 1024 - static_cast<int>(sizeof(SmemTail)) - 1024 - epi_extra) / stage_bytes;
  stages = std::min(stages, MAX_STAGES);
  stages = std::min(stages, std::max(2, p.num_kb));
  p.stages = stages;
  plan->smem_bytes = static_cast<size_t>(stages) * stage_bytes + sizeof(SmemTail) + 1024 + epi_extra;
  plan->grid = cfg.two_cta ? 2 * std::min(p.tiles_m * p.tiles_n, std::max(1, sm_count() / 2))
                           : std::min(p.tiles_m * p.tiles_n, sm_count());
  // weights: [N][Ktot] row-major
  uint64_t dims[2] = {static_cast<uint64_t>(Ktot), static_cast<uint64_t>(N)};
  uint64_t strides[2] = {1, static_cast<uint64_t>(plan->ldw > 0 ? plan->ldw : Ktot)};
  uint32_t box[2] = {BK, static_cast<uint32_t>(cfg.two_cta ? bn / 2 : bn)};
  return make_tmap_bf16(&plan->tmB, Wt, 2, dims, strides, box);
}

bool finish_plan(GemmPlan* plan, const bf16* Wt, int N, int Ktot, int out_mode, void* out, int ldo,
                 const float* bias, const float* rowbias, int rowbias_div, int rowbias_ld, const float* residual,
                 int ldr) {
  GemmParams& p = plan->p;
  if (Ktot % BK != 0) {
    set_error("gemm: K must be a multiple of 64");
    return false;
  }
  TileCfg cfg = pick_tile(p.M, N, Ktot / BK, (out_mode & 15) == OUT_GEGLU_BF16);
  // The cost model ranks the HBM-bound residual layers poorly (it has no term for their epilogue traffic).  Measured
  // winners on the transformer shapes of the shipped model, 5-13 % ahead of the model's pick at 5 groups per call
  // (scripts/bench_residual_gemm.py with FORCES, profiles/r02b_residual_gemm_force.log); every configuration computes
  // bit-identical results, so this only moves time.
  if (getenv("CAP4D_GEMM_FORCE") == nullptr && residual != nullptr && rowbias == nullptr && !p.a_conv &&
      (out_mode & 15) != OUT_GEGLU_BF16 && p.M >= 16384) {
    if (N == 320 && Ktot == 320) cfg = TileCfg{2, 64, 2, 0};          // to_out / proj_out, level 0
    else if (N == 320 && Ktot == 1280) cfg = TileCfg{1, 160, 2, 0};   // FF2, level 0
    else if (N == 640 && Ktot == 2560) cfg = TileCfg{2, 128, 2, 0};   // FF2, level 1
    else if (N == 1280 && Ktot == 1280) cfg = TileCfg{1, 128, 2, 0};  // to_out / proj_out, level 2
  }
  if (cfg.bn == 0) {
    set_error("gemm: N must be a multiple of 32 (64 for GEGLU)");
    return false;
  }
  p.out_mode = out_mode;
  p.out = out;
  p.ldo = ldo;
  p.bias = bias;
  p.rowbias = rowbias;
  p.rowbias_div = rowbias_div > 0 ? rowbias_div : 1;
  p.rowbias_ld = rowbias_ld;
  p.residual = residual;
  p.ldr = ldr;
  plan->flops = 2.0 * p.M * static_cast<double>(N) * Ktot;
  return apply_cfg(plan, cfg, Wt, N, Ktot);
}

bool make_plain_a_map(CUtensorMap* map, const bf16* A, int M, int K, int lda = 0) {
  uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
  uint64_t strides[2] = {1, static_cast<uint64_t>(lda > 0 ? lda : K)};
  uint32_t box[2] = {BK, BM};
  return make_tmap_bf16(map, A, 2, dims, strides, box);
}

}  // namespace

cudaError_t launch_gemm(const GemmPlan& plan, cudaStream_t stream);

namespace {

// ---- tile autotuner (opt-in: CAP4D_GEMM_AUTOTUNE=1) ---------------------------------------------------
// The cost model above ranks the long-K conv shapes well but not the short-K linear layers, whose time
// is set by the epilogue and by how tiles fill the SMs (measured in isolation: a CTA pair or a narrower
// tile is 5-13 % faster on several of them).  With autotuning on, the first plan of every distinct problem
// times the plausible configurations on the device - the plan's own buffers are used; at plan-build time
// their contents are scratch, and every configuration computes the same values - and the winner is cached
// for the process.  It is off by default because it does not pay end to end on this part: the U-Net step
// runs into the board's power cap (SM clock ~1.7 of 1.965 GHz under load), the tuned linear layers finish
// 10 % sooner in isolation and the forward pass as a whole moves by < 0.5 % (profiles/r01d_autotune.txt).
using TuneKey = std::tuple<int, int, int, int, int, int, int, int>;
std::map<TuneKey, TileCfg>& tune_cache() {
  static std::map<TuneKey, TileCfg> c;
  return c;
}
std::mutex& tune_mutex() {
  static std::mutex m;
  return m;
}

bool autotune_enabled() {  // read per plan, so a process can switch it on for the plans it builds next
  if (getenv("CAP4D_GEMM_FORCE") != nullptr) return false;
  const char* e = getenv("CAP4D_GEMM_AUTOTUNE");
  return e != nullptr && atoi(e) != 0;
}

bool autotune(GemmPlan* plan, const bf16* Wt, int N, int Ktot) {
  if (!autotune_enabled()) return true;
  GemmParams& p = plan->p;
  const bool geglu = (p.out_mode & 15) == OUT_GEGLU_BF16;
  const int flags = (p.bias != nullptr) | ((p.rowbias != nullptr) << 1) | ((p.residual != nullptr) << 2) |
                    ((p.up_py >= 0) << 3) | (p.a_conv << 4) | ((p.num_kb != p.seg0_kb) << 5);
  const TuneKey key{p.M, N, Ktot / BK, p.out_mode & 15, flags, p.W, p.H, p.n_taps};
  std::lock_guard<std::mutex> lock(tune_mutex());
  auto it = tune_cache().find(key);
  if (it != tune_cache().end()) return apply_cfg(plan, it->second, Wt, N, Ktot);

  const TileCfg model{p.msub, p.BN, p.n_acc, plan->two_cta};
  std::vector<TileCfg> cands;
  static const int bns[] = {256, 224, 192, 160, 128, 96, 64};
  for (int bn : bns) {
    if (N % bn != 0 || (geglu && bn % 64 != 0)) continue;
    cands.push_back(TileCfg{1, bn, 2, 0});
    cands.push_back(TileCfg{2, bn, (2 * bn <= 256) ? 2 : 1, 0});
    if (bn % 64 == 0) cands.push_back(TileCfg{1, bn, 2, 1});
  }
  cudaStream_t stream = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  TileCfg best = model;
  float best_ms = 1e30f;
  bool ok = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess &&
            cudaDeviceSynchronize() == cudaSuccess;
  if (ok) {
    GemmPlan trial = *plan;
    auto time_cfg = [&](const TileCfg& cfg) -> float {
      if (!apply_cfg(&trial, cfg, Wt, N, Ktot)) return 1e30f;
      // One warm-up, then several launches back to back under one event pair: inside the U-Net a kernel
      // never runs alone, so what counts includes how its grid drains and how the next one ramps up
      // (timing single launches favoured CTA-pair configurations that were no faster end to end).
      if (launch_gemm(trial, stream) != cudaSuccess) return 1e30f;
      const int reps = 4;
      cudaEventRecord(e0, stream);
      for (int r = 0; r < reps; ++r)
        if (launch_gemm(trial, stream) != cudaSuccess) return 1e30f;
      cudaEventRecord(e1, stream);
      if (cudaEventSynchronize(e1) != cudaSuccess) return 1e30f;
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      return ms / reps;
    };
    best_ms = time_cfg(model);
    for (const TileCfg& c : cands) {
      if (c.msub == model.msub && c.bn == model.bn && c.two_cta == model.two_cta) continue;
      const float ms = time_cfg(c);
      if (ms < best_ms * 0.98f) {  // only a clear win displaces the model's choice
        best_ms = ms;
        best = c;
      }
    }
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) best = model;
  } else {
    cudaGetLastError();  // no device / no context: keep the model's choice
  }
  if (e0 != nullptr) cudaEventDestroy(e0);
  if (e1 != nullptr) cudaEventDestroy(e1);
  if (stream != nullptr) cudaStreamDestroy(stream);
  tune_cache()[key] = best;
  return apply_cfg(plan, best, Wt, N, Ktot);
}

}  // namespace

bool make_gemm_plan(GemmPlan* plan, const bf16* A, int M, int K, const bf16* A2, int K2, const bf16* Wt, int N,
                    int out_mode, void* out, int ldo, const float* bias, const float* rowbias, int rowbias_div,
                    int rowbias_ld, const float* residual, int ldr, int lda, int ldw) {
  memset(plan, 0, sizeof(*plan));
  GemmParams& p = plan->p;
  if ((lda % 8) != 0 || (ldw % 8) != 0 || (ldw > 0 && A2 != nullptr)) {
    set_error("gemm: lda / ldw must be multiples of 8 (and ldw excludes a second A segment)");
    return false;
  }
  plan->ldw = ldw;
  if (K % BK != 0 || (A2 != nullptr && K2 % BK != 0)) {
    set_error("gemm: K must be a multiple of 64");
    return false;
  }
  p.M = M;
  p.a_conv = 0;
  p.seg0_kb = K / BK;
  p.cin_kb = 1;
  p.n_taps = 1;
  if (!make_plain_a_map(&plan->tmA, A, M, K, lda)) return false;
  if (A2 != nullptr) {
    if (!make_plain_a_map(&plan->tmA2, A2, M, K2)) return false;
  } else {
    plan->tmA2 = plan->tmA;
    K2 = 0;
  }
  p.up_py = p.up_px = -1;  // before finish_plan: apply_cfg sizes shared memory by the epilogue the launch will pick
  if (!finish_plan(plan, Wt, N, K + K2, out_mode, out, ldo, bias, rowbias, rowbias_div, rowbias_ld, residual, ldr))
    return false;
  return autotune(plan, Wt, N, K + K2);
}

bool make_conv_plan(GemmPlan* plan, const bf16* A, const ConvGeom& g, int Cin, const bf16* A2, int K2,
                    const bf16* Wt, int N, int out_mode, void* out, int ldo, const float* bias, const float* rowbias,
                    int rowbias_div, int rowbias_ld, const float* residual, int ldr) {
  memset(plan, 0, sizeof(*plan));
  GemmParams& p = plan->p;
  const int H = g.H, W = g.W;
  auto is_pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
  if (!is_pow2(W) || !is_pow2(H)) {
    set_error("conv: output H and W must be powers of two");
    return false;
  }
  if (Cin % BK != 0 || (A2 != nullptr && K2 % BK != 0)) {
    set_error("conv: Cin must be a multiple of 64");
    return false;
  }
  const bool up = g.up_phase >= 0;
  if ((!up && g.taps != 9) || (up && (g.taps != 4 || g.stride != 1 || g.up_phase > 3)) ||
      (g.stride != 1 && g.stride != 2)) {
    set_error("conv: only 3x3 stride 1/2 and the folded upsample phases are implemented");
    return false;
  }
  const int ntaps = up ? 4 : 9;
  p.M = g.n_img * H * W;
  p.a_conv = 1;
  p.cin_kb = Cin / BK;
  p.seg0_kb = ntaps * p.cin_kb;
  p.W = W;
  p.H = H;
  // an M sub-tile is 128 consecutive pixels: box_n images x box_h rows x box_w pixels (a piece of ONE row
  // when the image is wider than 128)
  const int box_w = std::min(W, BM);
  p.box_h = std::max(1, std::min(H, BM / W));
  p.box_n = BM / (box_w * p.box_h);
  p.n_taps = ntaps;
  int planes = 1;
  if (up) {
    const int py = g.up_phase >> 1, px = g.up_phase & 1;
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) {
        p.tap_dy[a * 2 + b] = a - 1 + py;  // py = 0: rows y-1, y;  py = 1: rows y, y+1
        p.tap_dx[a * 2 + b] = b - 1 + px;
        p.tap_dn[a * 2 + b] = 0;
      }
  }
  for (int ky = 0; ky < 3 && !up; ++ky) {
    for (int kx = 0; kx < 3; ++kx) {
      const int t = ky * 3 + kx;
      if (g.stride == 1) {
        p.tap_dx[t] = kx - 1;
        p.tap_dy[t] = ky - 1;
        p.tap_dn[t] = 0;
      } else {
        if (g.asym_pad) {
          // input pixel (2*oy + ky, 2*ox + kx): parity plane (ky&1, kx&1), shifted by +1 for ky/kx == 2; the zero
          // row/column F.pad adds below / right of the image is the TMA out-of-bounds fill
          const int py = ky & 1, px = kx & 1;
          p.tap_dy[t] = ky >> 1;
          p.tap_dx[t] = kx >> 1;
          p.tap_dn[t] = (py * 2 + px) * g.n_img;
        } else {
          // input pixel (2*oy + ky - 1, 2*ox + kx - 1): parity plane ((ky-1)&1, (kx-1)&1), shifted by -1 for ky/kx == 0
          const int py = (ky + 1) & 1, px = (kx + 1) & 1;  // ky=0 -> odd, 1 -> even, 2 -> odd
          p.tap_dy[t] = (ky == 0) ? -1 : 0;
          p.tap_dx[t] = (kx == 0) ? -1 : 0;
          p.tap_dn[t] = (py * 2 + px) * g.n_img;
        }
        planes = 4;
      }
    }
  }
  {
    uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                        static_cast<uint64_t>(g.n_img) * planes};
    uint64_t strides[4] = {1, static_cast<uint64_t>(Cin), static_cast<uint64_t>(Cin) * W,
                           static_cast<uint64_t>(Cin) * W * H};
    uint32_t box[4] = {BK, static_cast<uint32_t>(box_w), static_cast<uint32_t>(p.box_h),
                       static_cast<uint32_t>(p.box_n)};
    if (!make_tmap_bf16(&plan->tmA, A, 4, dims, strides, box)) return false;
  }
  if (A2 != nullptr) {
    if (!make_plain_a_map(&plan->tmA2, A2, p.M, K2)) return false;
  } else {
    plan->tmA2 = plan->tmA;
    K2 = 0;
  }
  if (!finish_plan(plan, Wt, N, ntaps * Cin + K2, out_mode, out, ldo, bias, rowbias, rowbias_div, rowbias_ld,
                   residual, ldr))
    return false;
  p.up_py = up ? (g.up_phase >> 1) : -1;
  p.up_px = up ? (g.up_phase & 1) : -1;
  if (up && (residual != nullptr || rowbias != nullptr)) {
    set_error("conv: the folded upsample phases support bias only");
    return false;
  }
  return autotune(plan, Wt, N, ntaps * Cin + K2);
}

cudaError_t launch_gemm(const GemmPlan& plan, cudaStream_t stream) {
  // the attribute is per device: one flag per device ordinal (several executors of one process may sit on
  // different GPUs, e.g. the reference's device_model_map)
  static std::atomic<bool> attr_set[64];
  int dev = 0;
  cudaGetDevice(&dev);
  const bool known = dev >= 0 && dev < 64;
  if (!known || !attr_set[dev].load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(gemm_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(gemm_tc_2sm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    if (known) attr_set[dev].store(true, std::memory_order_release);
  }
  if (plan.two_cta)
    gemm_tc_2sm_kernel<<<plan.grid, GEMM_THREADS, plan.smem_bytes, stream>>>(plan.tmA, plan.tmA2, plan.tmB, plan.p);
  else if (plan.p.epi == 2)
    gemm_tc_kernel<2><<<plan.grid, GEMM_THREADS, plan.smem_bytes, stream>>>(plan.tmA, plan.tmA2, plan.tmB, plan.p);
  else if (plan.p.epi == 1)
    gemm_tc_kernel<1><<<plan.grid, GEMM_THREADS, plan.smem_bytes, stream>>>(plan.tmA, plan.tmA2, plan.tmB, plan.p);
  else
    gemm_tc_kernel<0><<<plan.grid, GEMM_THREADS, plan.smem_bytes, stream>>>(plan.tmA, plan.tmA2, plan.tmB, plan.p);
  return cudaGetLastError();
}

}  // namespace cap4d

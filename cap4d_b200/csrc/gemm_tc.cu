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
// CTA = 192 threads: warp 0 TMA producer, warp 1 MMA issuer (+TMEM owner), warps 2-5 epilogue.
// smem ring of `stages` {A 128x64, B BNx64} bf16 tiles (128B swizzle); two TMEM accumulator
// stages (2 x 256 columns) so the epilogue of tile i overlaps the main loop of tile i+1.
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "kernels.h"
#include "ptx.cuh"

namespace cap4d {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int MAX_STAGES = 8;
constexpr int GEMM_THREADS = 192;
constexpr int TMEM_COLS = 512;
constexpr int ACC_STRIDE = 256;  // TMEM columns between the two accumulator stages

struct SmemTail {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void epilogue_store_f32(float* dst, const float* acc) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    *reinterpret_cast<float4*>(dst + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
  }
}

__device__ __forceinline__ void epilogue_store_bf16(bf16* dst, const float* acc) {
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    uint4 u;
    u.x = pack_bf16x2(acc[j], acc[j + 1]);
    u.y = pack_bf16x2(acc[j + 2], acc[j + 3]);
    u.z = pack_bf16x2(acc[j + 4], acc[j + 5]);
    u.w = pack_bf16x2(acc[j + 6], acc[j + 7]);
    *reinterpret_cast<uint4*>(dst + j) = u;
  }
}

__device__ __forceinline__ void add_vec32(float* acc, const float* __restrict__ src) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    float4 b = __ldg(reinterpret_cast<const float4*>(src + j));
    acc[j] += b.x;
    acc[j + 1] += b.y;
    acc[j + 2] += b.z;
    acc[j + 3] += b.w;
  }
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled tiles need 1024 B alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_stage_bytes = p.BN * BK * 2;
  uint8_t* sA = smem;
  uint8_t* sB = smem + p.stages * A_STAGE_BYTES;
  SmemTail* tail = reinterpret_cast<SmemTail*>(sB + p.stages * b_stage_bytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.tiles_m * p.tiles_n;

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
      mbar_init(&tail->tmem_empty[a], 128);
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

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = A_STAGE_BYTES + b_stage_bytes;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.tiles_n;
        const int n_tile = tile - m_tile * p.tiles_n;
        int n0 = 0, y0 = 0;
        if (p.a_conv) {
          const int pix0 = m_tile * BM;
          const int hw = p.H * p.W;
          n0 = pix0 / hw;
          y0 = (pix0 - n0 * hw) / p.W;
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&tail->empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&tail->full[stage], tx_bytes);
          void* dstA = sA + stage * A_STAGE_BYTES;
          if (kb < p.seg0_kb) {
            if (p.a_conv) {
              const int tap = kb / p.cin_kb;
              const int c0 = (kb - tap * p.cin_kb) * BK;
              tma_load_4d(dstA, &tmA, &tail->full[stage], c0, p.tap_dx[tap], y0 + p.tap_dy[tap],
                          n0 + p.tap_dn[tap]);
            } else {
              tma_load_2d(dstA, &tmA, &tail->full[stage], kb * BK, m_tile * BM);
            }
          } else {
            tma_load_2d(dstA, &tmA2, &tail->full[stage], (kb - p.seg0_kb) * BK, m_tile * BM);
          }
          tma_load_2d(sB + stage * b_stage_bytes, &tmB, &tail->full[stage], kb * BK, n_tile * p.BN);
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
      const uint32_t idesc = umma_idesc_bf16(BM, p.BN, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tail->tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&tail->full[stage], phase);
          tc_fence_after();
          const uint64_t adesc = umma_smem_desc_sw128(smem_u32(sA + stage * A_STAGE_BYTES));
          const uint64_t bdesc = umma_smem_desc_sw128(smem_u32(sB + stage * b_stage_bytes));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in the (>>4) address field
            umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
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
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int m_tile = tile / p.tiles_n;
      const int n_tile = tile - m_tile * p.tiles_n;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tail->tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row = m_tile * BM + q * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * ACC_STRIDE;
      const float* rb = nullptr;
      if (p.rowbias != nullptr && row_ok) rb = p.rowbias + static_cast<size_t>(row / p.rowbias_div) * p.rowbias_ld;
      const float* res = nullptr;
      if (p.residual != nullptr && row_ok) res = p.residual + static_cast<size_t>(row) * p.ldr;

      if (p.out_mode == OUT_GEGLU_BF16) {
        // weight rows are interleaved in blocks of 32: [x(32) | gate(32)] -> out 32 columns
        const int npairs = p.BN / 64;
        for (int c = 0; c < npairs; ++c) {
          uint32_t vx[32], vg[32];
          tmem_ld32(taddr + c * 64, vx);
          tmem_ld32(taddr + c * 64 + 32, vg);
          tmem_ld_wait();
          if (row_ok) {
            float* ax = reinterpret_cast<float*>(vx);
            float* ag = reinterpret_cast<float*>(vg);
            const int col0 = n_tile * p.BN + c * 64;
            if (p.bias != nullptr) {
              add_vec32(ax, p.bias + col0);
              add_vec32(ag, p.bias + col0 + 32);
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) ax[j] = ax[j] * gelu_erf_f(ag[j]);
            bf16* dst = reinterpret_cast<bf16*>(p.out) + static_cast<size_t>(row) * p.ldo + (col0 >> 1);
            epilogue_store_bf16(dst, ax);
          }
        }
      } else {
        const int nchunks = p.BN / 32;
        for (int c = 0; c < nchunks; ++c) {
          uint32_t v[32];
          tmem_ld32(taddr + c * 32, v);
          tmem_ld_wait();
          if (row_ok) {
            float* a = reinterpret_cast<float*>(v);
            const int col0 = n_tile * p.BN + c * 32;
            if (p.bias != nullptr) add_vec32(a, p.bias + col0);
            if (rb != nullptr) add_vec32(a, rb + col0);
            if (res != nullptr) add_vec32(a, res + col0);
            if (p.out_mode == OUT_F32) {
              epilogue_store_f32(reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + col0, a);
            } else {
              epilogue_store_bf16(reinterpret_cast<bf16*>(p.out) + static_cast<size_t>(row) * p.ldo + col0, a);
            }
          }
        }
      }
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

int pick_bn(int M, int N, bool geglu) {
  static const int cands[] = {256, 224, 192, 160, 128, 96, 64, 32};
  const int tiles_m = (M + BM - 1) / BM;
  const int sms = sm_count();
  int best = 0;
  double best_cost = 1e30;
  for (int bn : cands) {
    if (N % bn != 0) continue;
    if (geglu && (bn % 64 != 0)) continue;
    const long tiles = static_cast<long>(tiles_m) * (N / bn);
    const long waves = (tiles + sms - 1) / sms;
    // per-tile main-loop time ~ bn (MMA cycles per k-step); small tiles pay a fixed cost and are
    // shared-memory-bandwidth bound (A tile re-read per N tile)
    const double per_tile = std::max(bn, 128) + 40.0;
    const double cost = waves * per_tile;
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

bool finish_plan(GemmPlan* plan, const bf16* Wt, int N, int Ktot, int out_mode, void* out, int ldo,
                 const float* bias, const float* rowbias, int rowbias_div, int rowbias_ld, const float* residual,
                 int ldr) {
  GemmParams& p = plan->p;
  if (Ktot % BK != 0) {
    set_error("gemm: K must be a multiple of 64");
    return false;
  }
  const int bn = pick_bn(p.M, N, out_mode == OUT_GEGLU_BF16);
  if (bn == 0) {
    set_error("gemm: N must be a multiple of 32 (64 for GEGLU)");
    return false;
  }
  p.N = N;
  p.BN = bn;
  p.tiles_m = (p.M + BM - 1) / BM;
  p.tiles_n = N / bn;
  p.num_kb = Ktot / BK;
  const int stage_bytes = A_STAGE_BYTES + bn * BK * 2;
  int stages = (220 * 1024 - static_cast<int>(sizeof(SmemTail)) - 1024) / stage_bytes;
  stages = std::min(stages, MAX_STAGES);
  stages = std::min(stages, std::max(2, p.num_kb));
  p.stages = stages;
  plan->smem_bytes = static_cast<size_t>(stages) * stage_bytes + sizeof(SmemTail) + 1024;
  p.out_mode = out_mode;
  p.out = out;
  p.ldo = ldo;
  p.bias = bias;
  p.rowbias = rowbias;
  p.rowbias_div = rowbias_div > 0 ? rowbias_div : 1;
  p.rowbias_ld = rowbias_ld;
  p.residual = residual;
  p.ldr = ldr;
  const int total = p.tiles_m * p.tiles_n;
  plan->grid = std::min(total, sm_count());
  plan->flops = 2.0 * p.M * static_cast<double>(N) * Ktot;
  // weights: [N][Ktot] row-major
  uint64_t dims[2] = {static_cast<uint64_t>(Ktot), static_cast<uint64_t>(N)};
  uint64_t strides[2] = {1, static_cast<uint64_t>(Ktot)};
  uint32_t box[2] = {BK, static_cast<uint32_t>(bn)};
  return make_tmap_bf16(&plan->tmB, Wt, 2, dims, strides, box);
}

bool make_plain_a_map(CUtensorMap* map, const bf16* A, int M, int K) {
  uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
  uint64_t strides[2] = {1, static_cast<uint64_t>(K)};
  uint32_t box[2] = {BK, BM};
  return make_tmap_bf16(map, A, 2, dims, strides, box);
}

}  // namespace

bool make_gemm_plan(GemmPlan* plan, const bf16* A, int M, int K, const bf16* A2, int K2, const bf16* Wt, int N,
                    int out_mode, void* out, int ldo, const float* bias, const float* rowbias, int rowbias_div,
                    int rowbias_ld, const float* residual, int ldr) {
  memset(plan, 0, sizeof(*plan));
  GemmParams& p = plan->p;
  if (K % BK != 0 || (A2 != nullptr && K2 % BK != 0)) {
    set_error("gemm: K must be a multiple of 64");
    return false;
  }
  p.M = M;
  p.a_conv = 0;
  p.seg0_kb = K / BK;
  p.cin_kb = 1;
  p.n_taps = 1;
  if (!make_plain_a_map(&plan->tmA, A, M, K)) return false;
  if (A2 != nullptr) {
    if (!make_plain_a_map(&plan->tmA2, A2, M, K2)) return false;
  } else {
    plan->tmA2 = plan->tmA;
    K2 = 0;
  }
  return finish_plan(plan, Wt, N, K + K2, out_mode, out, ldo, bias, rowbias, rowbias_div, rowbias_ld, residual, ldr);
}

bool make_conv_plan(GemmPlan* plan, const bf16* A, const ConvGeom& g, int Cin, const bf16* A2, int K2,
                    const bf16* Wt, int N, int out_mode, void* out, int ldo, const float* bias, const float* rowbias,
                    int rowbias_div, int rowbias_ld, const float* residual, int ldr) {
  memset(plan, 0, sizeof(*plan));
  GemmParams& p = plan->p;
  const int H = g.H, W = g.W;
  auto is_pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
  if (!is_pow2(W) || !is_pow2(H) || W > BM) {
    set_error("conv: output H and W must be powers of two with W <= 128");
    return false;
  }
  if (Cin % BK != 0 || (A2 != nullptr && K2 % BK != 0)) {
    set_error("conv: Cin must be a multiple of 64");
    return false;
  }
  if (g.taps != 9 || (g.stride != 1 && g.stride != 2)) {
    set_error("conv: only 3x3 stride 1/2 is implemented");
    return false;
  }
  p.M = g.n_img * H * W;
  p.a_conv = 1;
  p.cin_kb = Cin / BK;
  p.seg0_kb = 9 * p.cin_kb;
  p.W = W;
  p.H = H;
  p.box_h = std::min(H, BM / W);
  p.box_n = BM / (W * p.box_h);
  p.n_taps = 9;
  int planes = 1;
  for (int ky = 0; ky < 3; ++ky) {
    for (int kx = 0; kx < 3; ++kx) {
      const int t = ky * 3 + kx;
      if (g.stride == 1) {
        p.tap_dx[t] = kx - 1;
        p.tap_dy[t] = ky - 1;
        p.tap_dn[t] = 0;
      } else {
        // input pixel (2*oy + ky - 1, 2*ox + kx - 1): parity plane ((ky-1)&1, (kx-1)&1), shifted by -1 for ky/kx == 0
        const int py = (ky + 1) & 1, px = (kx + 1) & 1;  // ky=0 -> odd, 1 -> even, 2 -> odd
        p.tap_dy[t] = (ky == 0) ? -1 : 0;
        p.tap_dx[t] = (kx == 0) ? -1 : 0;
        p.tap_dn[t] = (py * 2 + px) * g.n_img;
        planes = 4;
      }
    }
  }
  {
    uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                        static_cast<uint64_t>(g.n_img) * planes};
    uint64_t strides[4] = {1, static_cast<uint64_t>(Cin), static_cast<uint64_t>(Cin) * W,
                           static_cast<uint64_t>(Cin) * W * H};
    uint32_t box[4] = {BK, static_cast<uint32_t>(W), static_cast<uint32_t>(p.box_h),
                       static_cast<uint32_t>(p.box_n)};
    if (!make_tmap_bf16(&plan->tmA, A, 4, dims, strides, box)) return false;
  }
  if (A2 != nullptr) {
    if (!make_plain_a_map(&plan->tmA2, A2, p.M, K2)) return false;
  } else {
    plan->tmA2 = plan->tmA;
    K2 = 0;
  }
  return finish_plan(plan, Wt, N, 9 * Cin + K2, out_mode, out, ldo, bias, rowbias, rowbias_div, rowbias_ld,
                     residual, ldr);
}

cudaError_t launch_gemm(const GemmPlan& plan, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  gemm_tc_kernel<<<plan.grid, GEMM_THREADS, plan.smem_bytes, stream>>>(plan.tmA, plan.tmA2, plan.tmB, plan.p);
  return cudaGetLastError();
}

}  // namespace cap4d

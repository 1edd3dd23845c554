// Fused GroupNorm(32)+SiLU and LayerNorm for NHWC fp32 activations -> bf16 MMA operands.
//
// Reference: GroupNorm32 / LayerNorm32 (controlnet/ldm/modules/diffusionmodules/util.py:217-223,
// fp32 statistics), used as `normalization(ch)` + SiLU in ResBlocks and `out`
// (openaimodel.py:180-184, 222-231, 770-774; eps 1e-5) and as `Normalize` in the transformer
// (cap4d/mmdm/net/attention.py:107-109; eps 1e-6, no SiLU); LayerNorm32 at attention.py:311-326.
//
// These kernels are HBM-bound: fp32 in (GroupNorm: read twice, the second time mostly from L2), bf16 out.
// (Measured alternative, rejected: keeping each block's slice in shared memory between the two passes
// via 1-D bulk copies reads HBM once but, with only 2-4 slices per SM, overlaps loads and stores worse -
// 2.2-2.8 TB/s against 3.4-3.7 TB/s for this version on the production tensors.)
// The up-path ResBlocks normalise the channel concatenation of two tensors whose 32 groups
// straddle the seam (openaimodel.py:766 via mmdm_unet.py:115); both sources are read in place and
// the concatenation only ever exists as the bf16 output.
#include <cstdlib>

#include "kernels.h"
#include "ptx.cuh"

namespace cap4d {

namespace {

constexpr int GN_GROUPS = 32;
constexpr int GN_MAX_QI = 4;  // quads per thread along C (C <= 4 * 256 * 4 = 4096)
constexpr int GN_THREADS = 256;
constexpr int GN_MAX_TX = 256;

struct GnGeom {
  int C, C1, quads, TX, TY, nqi, cpg;
  int rows_per_chunk, n_chunks;
};

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e != nullptr ? atoi(e) : dflt;
}

// Chunk geometry of the persistent kernel below.  `resident` = the blocks the cooperative launch keeps on the GPU
// (occupancy x SMs, queried from the runtime for the current device).  Every block pays fixed costs per work item
// (partial reduction, one atomic, the image barrier), so chunks are as large as keeping the GPU full allows (about
// `waves` items per resident block); large images get more chunks so that the images in flight stay L2-sized.
// Measured (scripts/gn_sweep.sh, round 2): small chunks that would make the second pass a guaranteed L2 hit cost
// more in per-item overhead than the DRAM re-read they save, with 256-thread blocks and with one 1024-thread
// block per SM alike - the defaults are the best of the sweep at 16 and 80 images.  The same holds for keeping the
// chunk in REGISTERS between the phases (one HBM read, 40 KB items, two in flight per block): correct, 22.9 GB of
// DRAM traffic per call instead of 33.5, and 1.45-2.5x SLOWER (264 / 739 / 1265 us against 182 / 342 / 506 us on the
// three level-0 shapes) - an item then lasts ~2 us of transfer against ~5-9 us of per-item round trips (ticket,
// partials + fence, arrive, flag poll, done counter).  The fix for the re-read is statistics from the producer's
// epilogue, not a different GroupNorm kernel.
__host__ GnGeom gn_geometry(int C1, int C2, int hw, int n_img, int resident) {
  static const int l2_mb = env_int("CAP4D_GN_L2_MB", 160), waves = env_int("CAP4D_GN_WAVES", 2);
  GnGeom g;
  g.C = C1 + C2;
  g.C1 = C1;
  g.quads = g.C / 4;
  g.cpg = g.C / GN_GROUPS;
  // per-thread quad count must be one of the instantiated template values {1, 2, 4}
  g.TX = 0;
  g.nqi = 0;
  for (int nq = 1; nq <= 4; nq *= 2) {
    if (g.quads % nq == 0 && g.quads / nq <= GN_MAX_TX) {
      g.nqi = nq;
      g.TX = g.quads / nq;
      break;
    }
  }
  if (g.TX == 0) {  // unsupported (C > 4096 or odd quad count): reported by the launcher
    g.TX = 1;
    g.nqi = GN_MAX_QI + 1;
  }
  g.TY = GN_THREADS / g.TX;
  const int ty_smem = static_cast<int>(48 * 1024 / (static_cast<size_t>(2) * g.C * sizeof(float)));  // s_ch is [TY][2][C]
  if (g.TY > ty_smem) g.TY = ty_smem;
  if (g.TY < 1) g.TY = 1;
  if (g.TY > hw) g.TY = hw;
  if (resident < 1) resident = 1;
  const double image_mb = static_cast<double>(hw) * g.C * 4 / (1024.0 * 1024.0);
  int chunks = static_cast<int>(resident * image_mb / l2_mb) + 1;            // (1) L2 residency
  if (2.0 * image_mb > l2_mb) chunks = 1;  // an image that cannot share L2 with a second one (VAE decoder, 512^2 x 128:
                                           // 134 MB) is re-read from HBM whatever the chunking: fewest, largest chunks
  const int fill = (waves * resident + n_img - 1) / n_img;                   // (2) enough items
  if (chunks < fill) chunks = fill;
  if (chunks > GN_MAX_CHUNKS) chunks = GN_MAX_CHUNKS;
  if (chunks > resident) chunks = resident;  // never more chunks per image than blocks (see the kernel's ticket loop)
  int rpc = (hw + chunks - 1) / chunks;
  if (rpc < 4 * g.TY) rpc = 4 * g.TY;
  if (rpc > hw) rpc = hw;
  g.rows_per_chunk = rpc;
  g.n_chunks = (hw + rpc - 1) / rpc;
  return g;
}

// row2 = row + x2_shift: the second source may hold more images than the first (see launch_groupnorm)
__device__ __forceinline__ float4 ld_quad(const float* __restrict__ x1, const float* __restrict__ x2, int C1, int C2,
                                          size_t row, size_t row2, int c) {
  // c is a multiple of 4 and C1 is a multiple of 4, so a quad never straddles the seam
  if (c < C1) return __ldg(reinterpret_cast<const float4*>(x1 + row * C1 + c));
  return __ldg(reinterpret_cast<const float4*>(x2 + row2 * C2 + (c - C1)));
}

// ---- GroupNorm (+SiLU), one persistent cooperative kernel ----------------------------------------
// Work item = (image n, chunk of rows), items in image-major order, drawn by the resident blocks from a ticket
// counter.
// Block (TX, TY): thread (tx, ty) owns channel quads tx + i*TX (i < NQI) and rows ty, ty+TY, ... of the chunk,
// UNROLL rows (= UNROLL*NQI independent 16 B loads) per iteration.
//   phase 1  per-chunk (sum, sumsq) of the 32 groups -> partial[n][chunk][g][2], then arrive on the image's
//            counter; the LAST chunk to arrive reduces the image's partials in chunk order (fixed order:
//            deterministic), publishes mean / rstd of the 32 groups and raises the image's ready flag.
//   phase 2  wait for the flag, normalise the same rows again - they were read a few microseconds ago and
//            come from L2 (the working set is two chunks per resident block, sized for L2 by gn_geometry),
//            so HBM sees 4 B in + 2 B out per element instead of 8 + 2.
// The two phases are software-pipelined over a block's items: phase 1 of item k+1 runs BEFORE the wait of
// item k, so the barrier latency hides behind useful loads and a block never idles at a barrier while it
// still has statistics to contribute.  Waiting on other blocks is legitimate here because the launch is
// cooperative (cudaLaunchCooperativeKernel: the runtime refuses the launch unless every block is resident)
// and an image never has more chunks than there are blocks (see the ticket loop at the end of the kernel).
// All counters return to zero by the end of the launch (graph replays need no reset).
struct GnSync {
  unsigned int arrived, ready, done;
};
struct GnTicket {
  unsigned int next, left;
};

template <int NQI, int UNROLL>
__global__ void __launch_bounds__(GN_THREADS, (NQI == 1) ? 4 : 2)
gn_fused_kernel(const float* __restrict__ x1, const float* __restrict__ x2, int C1, int C2, int hw, int cpg,
                int rows_per_chunk, int n_chunks, int n_img, float* __restrict__ partial, float2* __restrict__ stats,
                GnSync* __restrict__ sync, GnTicket* __restrict__ ticket, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                int apply_silu, bf16* __restrict__ out, bf16* __restrict__ raw_out, int x2_G, int x2_V, int x2_R,
                int pipelined, int out_f16) {
  extern __shared__ float s_ch[];  // [TY][2][C]: per-row-lane channel partials (no atomics: deterministic)
  __shared__ double s_part[8][GN_GROUPS][2];
  __shared__ float s_mean[GN_GROUPS], s_rstd[GN_GROUPS];
  __shared__ int s_last;
  const int C = C1 + C2;
  const int tx = threadIdx.x, ty = threadIdx.y, TX = blockDim.x, TY = blockDim.y;
  const int tid = ty * TX + tx, nthreads = TX * TY;
  const int n_items = n_img * n_chunks;

  auto rows_of = [&](int item, int* n, size_t* img_row, size_t* img_row2, int* r0, int* r1) {
    *n = item / n_chunks;
    const int chunk = item - *n * n_chunks;
    *r0 = chunk * rows_per_chunk;
    *r1 = min(hw, *r0 + rows_per_chunk);
    *img_row = static_cast<size_t>(*n) * hw;
    *img_row2 = (x2_G > 0) ? static_cast<size_t>((*n / x2_G) * x2_V + x2_R + *n % x2_G) * hw : *img_row;
  };

  // ---------------- phase 1: statistics of one chunk, arrive, (last arriver) publish the image ----------------
  auto phase1 = [&](int item) {
    int n, r0, r1;
    size_t img_row, img_row2;
    rows_of(item, &n, &img_row, &img_row2, &r0, &r1);
    const int chunk = item - n * n_chunks;
    {
      float sum[NQI][4], sq[NQI][4];
#pragma unroll
      for (int qi = 0; qi < NQI; ++qi)
#pragma unroll
        for (int k = 0; k < 4; ++k) sum[qi][k] = sq[qi][k] = 0.f;
      int r = r0 + ty;
      for (; r + (UNROLL - 1) * TY < r1; r += UNROLL * TY) {
        float4 v[UNROLL][NQI];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
          for (int qi = 0; qi < NQI; ++qi)
            v[u][qi] = ld_quad(x1, x2, C1, C2, img_row + r + u * TY, img_row2 + r + u * TY, (tx + qi * TX) * 4);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
          for (int qi = 0; qi < NQI; ++qi) {
            sum[qi][0] += v[u][qi].x; sq[qi][0] = fmaf(v[u][qi].x, v[u][qi].x, sq[qi][0]);
            sum[qi][1] += v[u][qi].y; sq[qi][1] = fmaf(v[u][qi].y, v[u][qi].y, sq[qi][1]);
            sum[qi][2] += v[u][qi].z; sq[qi][2] = fmaf(v[u][qi].z, v[u][qi].z, sq[qi][2]);
            sum[qi][3] += v[u][qi].w; sq[qi][3] = fmaf(v[u][qi].w, v[u][qi].w, sq[qi][3]);
          }
      }
      for (; r < r1; r += TY) {
#pragma unroll
        for (int qi = 0; qi < NQI; ++qi) {
          const float4 v = ld_quad(x1, x2, C1, C2, img_row + r, img_row2 + r, (tx + qi * TX) * 4);
          sum[qi][0] += v.x; sq[qi][0] = fmaf(v.x, v.x, sq[qi][0]);
          sum[qi][1] += v.y; sq[qi][1] = fmaf(v.y, v.y, sq[qi][1]);
          sum[qi][2] += v.z; sq[qi][2] = fmaf(v.z, v.z, sq[qi][2]);
          sum[qi][3] += v.w; sq[qi][3] = fmaf(v.w, v.w, sq[qi][3]);
        }
      }
      float* my = s_ch + static_cast<size_t>(ty) * 2 * C;
#pragma unroll
      for (int qi = 0; qi < NQI; ++qi) {
        const int c = (tx + qi * TX) * 4;
        *reinterpret_cast<float4*>(my + c) = make_float4(sum[qi][0], sum[qi][1], sum[qi][2], sum[qi][3]);
        *reinterpret_cast<float4*>(my + C + c) = make_float4(sq[qi][0], sq[qi][1], sq[qi][2], sq[qi][3]);
      }
    }
    __syncthreads();
    // row-lanes -> lane 0 (2C sums of TY values, fixed order), then channels -> groups (64 sums of cpg values)
    for (int i = tid; i < 2 * C; i += nthreads) {
      float acc = s_ch[i];
      for (int y = 1; y < TY; ++y) acc += s_ch[static_cast<size_t>(y) * 2 * C + i];
      s_ch[i] = acc;
    }
    __syncthreads();
    if (tid < 2 * GN_GROUPS) {
      const int g = tid >> 1, which = tid & 1;
      const float* src = s_ch + which * C + g * cpg;
      float acc = 0.f;
      for (int c = 0; c < cpg; ++c) acc += src[c];
      __stcg(partial + ((static_cast<size_t>(n) * GN_MAX_CHUNKS + chunk) * GN_GROUPS + g) * 2 + which, acc);
      __threadfence();
    }
    __syncthreads();
    if (tid == 0) {
      const unsigned int old = atomicAdd(&sync[n].arrived, 1u);
      s_last = (old == static_cast<unsigned int>(n_chunks) - 1u);
      if (s_last) sync[n].arrived = 0u;  // every chunk of the image has arrived: nobody adds to it again in this launch
    }
    __syncthreads();
    if (s_last) {
      __threadfence();  // acquire the other chunks' partials
      const int nparts = min(8, max(1, nthreads / 32));
      const int part = tid >> 5, g = tid & 31;
      if (part < nparts) {
        double s = 0.0, q = 0.0;
        for (int ch = part; ch < n_chunks; ch += nparts) {
          const float2 v = __ldcg(reinterpret_cast<const float2*>(
              partial + ((static_cast<size_t>(n) * GN_MAX_CHUNKS + ch) * GN_GROUPS + g) * 2));
          s += v.x;
          q += v.y;
        }
        s_part[part][g][0] = s;
        s_part[part][g][1] = q;
      }
      __syncthreads();
      if (tid < GN_GROUPS) {
        double s = 0.0, q = 0.0;
        for (int pt = 0; pt < nparts; ++pt) {
          s += s_part[pt][tid][0];
          q += s_part[pt][tid][1];
        }
        const double cnt = static_cast<double>(cpg) * hw;
        const double mean = s / cnt;
        double var = q / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        __stcg(stats + static_cast<size_t>(n) * GN_GROUPS + tid,
               make_float2(static_cast<float>(mean), static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)))));
        __threadfence();
      }
      __syncthreads();
      if (tid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(&sync[n].ready), "r"(1u) : "memory");
    }
  };

  // ---------------- phase 2: wait for the image's statistics, normalise the chunk ----------------
  auto phase2 = [&](int item) {
    int n, r0, r1;
    size_t img_row, img_row2;
    rows_of(item, &n, &img_row, &img_row2, &r0, &r1);
    if (tid == 0) {
      const long long t0 = clock64();
      unsigned int seen;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(&sync[n].ready) : "memory");
        if (seen == 0u && clock64() - t0 > CAP4D_WATCHDOG_CYCLES) __trap();
      } while (seen == 0u);
    }
    __syncthreads();
    if (tid < GN_GROUPS) {
      const float2 st = __ldcg(stats + static_cast<size_t>(n) * GN_GROUPS + tid);
      s_mean[tid] = st.x;
      s_rstd[tid] = st.y;
    }
    __syncthreads();
    // per-thread affine: y = x * a + b with a = rstd*gamma, b = beta - mean*rstd*gamma
    float a[NQI][4], b[NQI][4];
#pragma unroll
    for (int qi = 0; qi < NQI; ++qi) {
      const int c = (tx + qi * TX) * 4;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int g = (c + k) / cpg;
        const float ga = __ldg(gamma + c + k), be = __ldg(beta + c + k);
        a[qi][k] = s_rstd[g] * ga;
        b[qi][k] = be - s_mean[g] * s_rstd[g] * ga;
      }
    }
    auto emit = [&](size_t row, int qi, const float4& v) {
      const int c = (tx + qi * TX) * 4;
      float y0 = fmaf(v.x, a[qi][0], b[qi][0]);
      float y1 = fmaf(v.y, a[qi][1], b[qi][1]);
      float y2 = fmaf(v.z, a[qi][2], b[qi][2]);
      float y3 = fmaf(v.w, a[qi][3], b[qi][3]);
      if (out_f16 == 2) {
        // fp32-accuracy mode: exact SiLU, fp32 out (split into bf16 triples by the next kernel, precise.cu)
        if (apply_silu) {
          y0 = y0 / (1.0f + expf(-y0));
          y1 = y1 / (1.0f + expf(-y1));
          y2 = y2 / (1.0f + expf(-y2));
          y3 = y3 / (1.0f + expf(-y3));
        }
        __stcs(reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + row * C + c), make_float4(y0, y1, y2, y3));
        if (raw_out != nullptr) __stcs(reinterpret_cast<float4*>(reinterpret_cast<float*>(raw_out) + row * C + c), v);
        return;
      }
      if (apply_silu) {
        y0 = silu_f(y0);
        y1 = silu_f(y1);
        y2 = silu_f(y2);
        y3 = silu_f(y3);
      }
      __stcs(reinterpret_cast<uint2*>(out + row * C + c), make_uint2(pack16x2(y0, y1, out_f16), pack16x2(y2, y3, out_f16)));
      if (raw_out != nullptr)
        __stcs(reinterpret_cast<uint2*>(raw_out + row * C + c), make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w)));
    };
    int r = r0 + ty;
    for (; r + (UNROLL - 1) * TY < r1; r += UNROLL * TY) {
      float4 v[UNROLL][NQI];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
#pragma unroll
        for (int qi = 0; qi < NQI; ++qi)
          v[u][qi] = ld_quad(x1, x2, C1, C2, img_row + r + u * TY, img_row2 + r + u * TY, (tx + qi * TX) * 4);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
#pragma unroll
        for (int qi = 0; qi < NQI; ++qi) emit(img_row + r + u * TY, qi, v[u][qi]);
    }
    for (; r < r1; r += TY) {
#pragma unroll
      for (int qi = 0; qi < NQI; ++qi)
        emit(img_row + r, qi, ld_quad(x1, x2, C1, C2, img_row + r, img_row2 + r, (tx + qi * TX) * 4));
    }
    // the last chunk of the image to finish lowers the flag again for the next launch
    __syncthreads();
    if (tid == 0) {
      const unsigned int old = atomicAdd(&sync[n].done, 1u);
      if (old == static_cast<unsigned int>(n_chunks) - 1u) {
        sync[n].done = 0u;
        sync[n].ready = 0u;
        __threadfence();
      }
    }
  };

  // Items are handed out in order by a global ticket counter (dynamic balance: a block that drew slow chunks does
  // not hold a fixed share back).  Deadlock freedom with tickets: the drawn items always form a prefix of the item
  // order; a block only waits while holding two drawn items, so if every block waited on the same unfinished image
  // that image would have 2 x grid > n_chunks drawn chunks - impossible with n_chunks <= grid.
  __shared__ int s_item;
  auto draw = [&]() {
    __syncthreads();  // everybody is done with the previous value of s_item
    if (tid == 0) s_item = static_cast<int>(atomicAdd(&ticket->next, 1u));
    __syncthreads();
    return s_item;
  };
  // bit 1 of `pipelined`: tickets count DOWN through the items (last image first).  The producer of this tensor wrote
  // it front to back, so its tail is what is still in L2 when this kernel starts, and the consumer (a conv that walks
  // its tiles front to back) then finds the images this kernel wrote last.  Items stay image-contiguous in ticket
  // order, which is all the deadlock argument needs.
  const bool reverse = (pipelined & 2) != 0;
  if (pipelined & 1) {
    int prev = -1;
    for (;;) {
      const int t = draw();
      const bool has = t < n_items;
      const int item = reverse ? n_items - 1 - t : t;
      if (has) phase1(item);
      if (prev >= 0) phase2(prev);
      if (!has) break;
      prev = item;
    }
  } else {
    for (int t = draw(); t < n_items; t = draw()) {
      const int item = reverse ? n_items - 1 - t : t;
      phase1(item);
      phase2(item);
    }
  }
  // the last block to leave rewinds the ticket counter for the next launch
  __syncthreads();
  if (tid == 0) {
    const unsigned int left = atomicAdd(&ticket->left, 1u);
    if (left == gridDim.x - 1u) {
      ticket->next = 0u;
      ticket->left = 0u;
      __threadfence();
    }
  }
}

// ---- LayerNorm: one warp per row; the row lives in registers between the two passes --------------
template <int NQ>  // quads per lane: C <= 128 * NQ
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, int M, int C, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, bf16* __restrict__ out, int out_f16, int reverse) {
  const int warps_per_block = blockDim.x >> 5;
  // reverse: last rows first - the tail of the tensor is what the producing GEMM left in L2 (same idea as the
  // GroupNorm kernel's ticket order), and the consuming GEMM starts at the rows written last
  const int blk = reverse ? static_cast<int>(gridDim.x) - 1 - static_cast<int>(blockIdx.x) : static_cast<int>(blockIdx.x);
  const int row = blk * warps_per_block + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const int quads = C >> 2;
  const float* xr = x + static_cast<size_t>(row) * C;
  float4 v[NQ];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    const int qd = i * 32 + lane;
    if (qd < quads) v[i] = __ldg(reinterpret_cast<const float4*>(xr) + qd);
  }
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    const int qd = i * 32 + lane;
    if (qd < quads) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    const int qd = i * 32 + lane;
    if (qd < quads) {
      const float d0 = v[i].x - mean, d1 = v[i].y - mean, d2 = v[i].z - mean, d3 = v[i].w - mean;
      q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / C + eps);
  bf16* orow = out + static_cast<size_t>(row) * C;
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    const int qd = i * 32 + lane;
    if (qd < quads) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + qd);
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + qd);
      const float y0 = (v[i].x - mean) * rstd * g.x + b.x;
      const float y1 = (v[i].y - mean) * rstd * g.y + b.y;
      const float y2 = (v[i].z - mean) * rstd * g.z + b.z;
      const float y3 = (v[i].w - mean) * rstd * g.w + b.w;
      if (out_f16 == 2)
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + static_cast<size_t>(row) * C + qd * 4) = make_float4(y0, y1, y2, y3);
      else
        *reinterpret_cast<uint2*>(orow + qd * 4) = make_uint2(pack16x2(y0, y1, out_f16), pack16x2(y2, y3, out_f16));
    }
  }
}

// blocks of gn_fused_kernel<NQI, UNROLL> the current device keeps resident for this shared-memory size
template <int NQI, int UNROLL>
int gn_resident_blocks(size_t smem_bytes, int threads) {
  int dev = 0, per_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gn_fused_kernel<NQI, UNROLL>, threads, smem_bytes) !=
      cudaSuccess)
    return 0;
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return per_sm * sms;
}

template <int NQI>
cudaError_t launch_gn_t(int C1, int C2, const float* x1, const float* x2, int n_img, int hw, const float* gamma,
                        const float* beta, float eps, int apply_silu, bf16* out, bf16* raw_out, float* partial,
                        cudaStream_t stream, int x2_G, int x2_V, int x2_R, int n_layout, int out_f16) {
  constexpr int UNROLL = (NQI >= 4) ? 2 : 4;
  const int C = C1 + C2;
  // the block shape does not depend on the residency, the residency depends on the block's shared memory
  GnGeom g = gn_geometry(C1, C2, hw, n_img, 1);
  const size_t smem = static_cast<size_t>(2) * C * g.TY * sizeof(float);
  const int resident = gn_resident_blocks<NQI, UNROLL>(smem, g.TX * g.TY);
  if (resident < 1) {
    set_error("groupnorm: the kernel does not fit an SM / occupancy query failed");
    return cudaErrorInvalidConfiguration;
  }
  g = gn_geometry(C1, C2, hw, n_img, resident);
  const int n_items = n_img * g.n_chunks;
  // n_chunks <= grid is what makes the ticket loop deadlock-free (see the kernel)
  int grid = resident;
  if (grid > n_items) grid = n_items;
  if (grid < g.n_chunks) {
    set_error("groupnorm: more chunks per image than resident blocks");
    return cudaErrorInvalidConfiguration;
  }
  dim3 block(g.TX, g.TY);
  // scratch: partials | per-image group statistics | per-image counters (zero between launches)
  float2* stats = reinterpret_cast<float2*>(partial + static_cast<size_t>(n_layout) * GN_MAX_CHUNKS * GN_GROUPS * 2);
  GnSync* sync = reinterpret_cast<GnSync*>(stats + static_cast<size_t>(n_layout) * GN_GROUPS);
  GnTicket* ticket = reinterpret_cast<GnTicket*>(sync + n_layout);
  int cpg = g.cpg, rpc = g.rows_per_chunk, nch = g.n_chunks;
  // CAP4D_GN_REVERSE=0: first image first (A/B: 8.04 -> 7.85 ms per U-Net call with the reversed order, two same-box pairs)
  static int pipelined = env_int("CAP4D_GN_PIPELINE", 1) | (env_int("CAP4D_GN_REVERSE", 1) << 1);
  void* args[] = {&x1, &x2, &C1, &C2, &hw, &cpg, &rpc, &nch, &n_img, &partial, &stats, &sync, &ticket, &gamma, &beta, &eps,
                  &apply_silu, &out, &raw_out, &x2_G, &x2_V, &x2_R, &pipelined, &out_f16};
  // CAP4D_GN_COOPERATIVE=0 (experiments only): an ordinary launch of the same grid - every block still fits the GPU at
  // once, but nothing then guarantees it against other work sharing the device
  static const int cooperative = env_int("CAP4D_GN_COOPERATIVE", 1);
  if (!cooperative) {
    gn_fused_kernel<NQI, UNROLL><<<dim3(grid), block, smem, stream>>>(x1, x2, C1, C2, hw, cpg, rpc, nch, n_img, partial, stats,
                                                                      sync, ticket, gamma, beta, eps, apply_silu, out,
                                                                      raw_out, x2_G, x2_V, x2_R, pipelined, out_f16);
    return cudaGetLastError();
  }
  return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(gn_fused_kernel<NQI, UNROLL>), dim3(grid), block,
                                     args, smem, stream);
}

}  // namespace

// per-chunk partials, per-image statistics, then the per-image counters (which must be zero before the first
// launch, see groupnorm_sync_offset; every launch leaves them zeroed again)
size_t groupnorm_sync_offset(int n_img) {
  return static_cast<size_t>(n_img) * GN_MAX_CHUNKS * GN_GROUPS * 2 * sizeof(float) +
         static_cast<size_t>(n_img) * GN_GROUPS * sizeof(float2);
}
size_t groupnorm_partial_bytes(int n_img) {
  return groupnorm_sync_offset(n_img) + static_cast<size_t>(n_img) * sizeof(GnSync) + sizeof(GnTicket);
}

cudaError_t launch_groupnorm(const float* x1, int C1, const float* x2, int C2, int n_img, int hw, const float* gamma,
                             const float* beta, float eps, int apply_silu, bf16* out, bf16* raw_out, float* partial,
                             cudaStream_t stream, int x2_G, int x2_V, int x2_R, int n_img_layout, int out_f16) {
  const int C = C1 + C2;
  const int n_layout = n_img_layout > 0 ? n_img_layout : n_img;
  if (C % GN_GROUPS != 0 || C1 % 4 != 0 || C2 % 4 != 0) {
    set_error("groupnorm: channels must be a multiple of 32 (each source a multiple of 4)");
    return cudaErrorInvalidValue;
  }
  if (n_img <= 0 || hw <= 0) return cudaSuccess;
  const GnGeom g = gn_geometry(C1, C2, hw, n_img, 1);
  if (g.nqi > GN_MAX_QI || static_cast<size_t>(2) * C * g.TY * sizeof(float) > 48 * 1024) {
    set_error("groupnorm: too many channels for this kernel");
    return cudaErrorInvalidValue;
  }
#define CAP4D_GN_CASE(N) \
  return launch_gn_t<N>(C1, C2, x1, x2, n_img, hw, gamma, beta, eps, apply_silu, out, raw_out, partial, stream, x2_G, \
                        x2_V, x2_R, n_layout, out_f16)
  if (g.nqi == 1) CAP4D_GN_CASE(1);
  if (g.nqi == 2) CAP4D_GN_CASE(2);
  CAP4D_GN_CASE(4);
#undef CAP4D_GN_CASE
}

cudaError_t launch_layernorm(const float* x, int M, int C, const float* gamma, const float* beta, float eps,
                             bf16* out, cudaStream_t stream, int out_f16) {
  if (C % 4 != 0 || C > 4 * 32 * 16) {
    set_error("layernorm: C must be a multiple of 4 and <= 2048");
    return cudaErrorInvalidValue;
  }
  const int warps = 8;
  const int grid = (M + warps - 1) / warps;
  const int quads = C / 4;
  static const int reverse = env_int("CAP4D_LN_REVERSE", 1);  // A/B: 2.36 -> 2.27 ms per U-Net call, two same-box pairs
  if (quads <= 32 * 4)
    layernorm_kernel<4><<<grid, warps * 32, 0, stream>>>(x, M, C, gamma, beta, eps, out, out_f16, reverse);
  else if (quads <= 32 * 8)
    layernorm_kernel<8><<<grid, warps * 32, 0, stream>>>(x, M, C, gamma, beta, eps, out, out_f16, reverse);
  else
    layernorm_kernel<16><<<grid, warps * 32, 0, stream>>>(x, M, C, gamma, beta, eps, out, out_f16, reverse);
  return cudaGetLastError();
}

}  // namespace cap4d

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

struct GnGeom {
  int C, C1, quads, TX, TY, nqi, cpg;
  int rows_per_chunk, n_chunks;
};

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e != nullptr ? atoi(e) : dflt;
}

// Chunks per image.  Every block pays the fixed latency of the image barrier, so chunks should be as
// large as keeping the GPU full allows (one wave of resident blocks over all images); large images get
// more chunks so that the images in flight (resident blocks / chunks per image) stay L2-sized.  The
// defaults are the best of a measured sweep (scripts/gn_sweep.sh) at 16 and 80 images.
__host__ GnGeom gn_geometry(int C1, int C2, int hw, int n_img) {
  static const int l2_mb = env_int("CAP4D_GN_L2_MB", 160), waves = env_int("CAP4D_GN_WAVES", 1);
  GnGeom g;
  g.C = C1 + C2;
  g.C1 = C1;
  g.quads = g.C / 4;
  g.cpg = g.C / GN_GROUPS;
  // per-thread quad count must be one of the instantiated template values {1, 2, 4}
  g.TX = 0;
  g.nqi = 0;
  for (int nq = 1; nq <= 4; nq *= 2) {
    if (g.quads % nq == 0 && g.quads / nq <= 256) {
      g.nqi = nq;
      g.TX = g.quads / nq;
      break;
    }
  }
  if (g.TX == 0) {  // unsupported (C > 4096 or odd quad count): reported by the launcher
    g.TX = 1;
    g.nqi = GN_MAX_QI + 1;
  }
  g.TY = 256 / g.TX;
  if (g.TY < 1) g.TY = 1;
  if (g.TY > hw) g.TY = hw;
  const int resident = 148 * (g.nqi == 1 ? 4 : 3);  // blocks the GPU holds at once (register-limited)
  const double image_mb = static_cast<double>(hw) * g.C * 4 / (1024.0 * 1024.0);
  int chunks = static_cast<int>(resident * image_mb / l2_mb) + 1;            // (1) L2 residency
  const int fill = (waves * resident + n_img - 1) / n_img;                   // (2) enough blocks
  if (chunks < fill) chunks = fill;
  if (chunks > GN_MAX_CHUNKS) chunks = GN_MAX_CHUNKS;
  int rpc = (hw + chunks - 1) / chunks;
  if (rpc < 4 * g.TY) rpc = 4 * g.TY;
  if (rpc > hw) rpc = hw;
  g.rows_per_chunk = rpc;
  g.n_chunks = (hw + rpc - 1) / rpc;
  return g;
}

// A/B builds only (-DCAP4D_GN_CACHE_HINTS=1, not the default): L2 eviction priorities for the two-pass GroupNorm -
// first-pass reads are kept (evict_last: the same rows are read again after the image barrier), second-pass reads and
// the bf16 outputs are marked evict_first so that they do not push the waiting inputs out.
#ifndef CAP4D_GN_CACHE_HINTS
#define CAP4D_GN_CACHE_HINTS 0
#endif
#if CAP4D_GN_CACHE_HINTS
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ld_hint(const float* ptr, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(ptr), "l"(pol));
  return v;
}
__device__ __forceinline__ void st_hint(bf16* ptr, uint2 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.b32 [%0], {%1, %2}, %3;" ::"l"(ptr), "r"(v.x), "r"(v.y), "l"(pol) : "memory");
}
#endif

// row2 = row + x2_shift: the second source may hold more images than the first (see launch_groupnorm)
__device__ __forceinline__ float4 ld_quad(const float* __restrict__ x1, const float* __restrict__ x2, int C1, int C2,
                                          size_t row, size_t row2, int c, uint64_t pol = 0) {
  // c is a multiple of 4 and C1 is a multiple of 4, so a quad never straddles the seam
#if CAP4D_GN_CACHE_HINTS
  if (c < C1) return ld_hint(x1 + row * C1 + c, pol);
  return ld_hint(x2 + row2 * C2 + (c - C1), pol);
#else
  if (c < C1) return __ldg(reinterpret_cast<const float4*>(x1 + row * C1 + c));
  return __ldg(reinterpret_cast<const float4*>(x2 + row2 * C2 + (c - C1)));
#endif
}

// ---- GroupNorm (+SiLU), one kernel ---------------------------------------------------------------
// grid (n_chunks, n_img) = image-major block order, block (TX, TY): thread (tx, ty) owns channel quads
// tx + i*TX (i < NQI) and rows ty, ty+TY, ... of its chunk, UNROLL rows (= UNROLL*NQI independent 16 B
// loads) per iteration.
//   phase 1  per-chunk (sum, sumsq) of the 32 groups -> partial[n][chunk][g][2] (fixed order: deterministic)
//   barrier  the chunks of ONE image wait for each other on a global counter.  Blocks are dispatched in
//            linear order, so every block an image waits for is already resident or ahead of it in the
//            queue, and the blocks of earlier images can always finish: no deadlock.
//   phase 2  reduce the partials, normalise the same rows again - they were read a few microseconds ago
//            and come from L2 (the working set is the handful of images in flight, not the whole tensor),
//            so HBM sees 4 B in + 2 B out per element instead of 8 + 2.
struct GnSync {
  unsigned int arrived, done;
};

template <int NQI, int UNROLL>
__global__ void __launch_bounds__(256)
gn_fused_kernel(const float* __restrict__ x1, const float* __restrict__ x2, int C1, int C2, int hw, int cpg,
                int rows_per_chunk, int n_chunks, float* __restrict__ partial, GnSync* __restrict__ sync,
                const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int apply_silu,
                bf16* __restrict__ out, bf16* __restrict__ raw_out, int x2_G, int x2_V, int x2_R) {
  extern __shared__ float s_ch[];  // [TY][2][C]: per-row-lane channel partials (no atomics: deterministic)
  __shared__ double s_part[8][GN_GROUPS][2];
  __shared__ float s_mean[GN_GROUPS], s_rstd[GN_GROUPS];
  const int C = C1 + C2;
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int tx = threadIdx.x, ty = threadIdx.y, TX = blockDim.x, TY = blockDim.y;
  const int tid = ty * TX + tx, nthreads = TX * TY;
  const int r0 = chunk * rows_per_chunk;
  const int r1 = min(hw, r0 + rows_per_chunk);
  const size_t img_row = static_cast<size_t>(n) * hw;
  const size_t img_row2 = (x2_G > 0) ? static_cast<size_t>((n / x2_G) * x2_V + x2_R + n % x2_G) * hw : img_row;

#if CAP4D_GN_CACHE_HINTS
  const uint64_t pol1 = l2_policy_evict_last(), pol2 = l2_policy_evict_first();
#else
  const uint64_t pol1 = 0, pol2 = 0;
#endif
  // ---------------- phase 1: statistics of this chunk ----------------
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
        for (int qi = 0; qi < NQI; ++qi) v[u][qi] = ld_quad(x1, x2, C1, C2, img_row + r + u * TY, img_row2 + r + u * TY, (tx + qi * TX) * 4, pol1);
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
        const float4 v = ld_quad(x1, x2, C1, C2, img_row + r, img_row2 + r, (tx + qi * TX) * 4, pol1);
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
  if (tid < GN_GROUPS) {
    float s = 0.f, q = 0.f;
    for (int y = 0; y < TY; ++y) {
      const float* src = s_ch + static_cast<size_t>(y) * 2 * C;
      for (int c = tid * cpg; c < (tid + 1) * cpg; ++c) {
        s += src[c];
        q += src[C + c];
      }
    }
    float* dst = partial + ((static_cast<size_t>(n) * GN_MAX_CHUNKS + chunk) * GN_GROUPS + tid) * 2;
    __stcg(reinterpret_cast<float2*>(dst), make_float2(s, q));
    __threadfence();
  }
  __syncthreads();

  // ---------------- barrier over the chunks of image n ----------------
  if (tid == 0) {
    atomicAdd(&sync[n].arrived, 1u);
    long long t0 = clock64();
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(&sync[n].arrived) : "memory");
      if (seen < static_cast<unsigned int>(n_chunks) && clock64() - t0 > CAP4D_WATCHDOG_CYCLES) __trap();
    } while (seen < static_cast<unsigned int>(n_chunks));
  }
  __syncthreads();

  // ---------------- phase 2: statistics of the image, then normalise ----------------
  const int nparts = min(8, max(1, nthreads / 32));
  {
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
    s_mean[tid] = static_cast<float>(mean);
    s_rstd[tid] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
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
    if (apply_silu) {
      y0 = silu_f(y0);
      y1 = silu_f(y1);
      y2 = silu_f(y2);
      y3 = silu_f(y3);
    }
#if CAP4D_GN_CACHE_HINTS
    st_hint(out + row * C + c, make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3)), pol2);
    if (raw_out != nullptr)
      st_hint(raw_out + row * C + c, make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w)), pol2);
#else
    *reinterpret_cast<uint2*>(out + row * C + c) = make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
    if (raw_out != nullptr)
      *reinterpret_cast<uint2*>(raw_out + row * C + c) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
#endif
  };
  int r = r0 + ty;
  for (; r + (UNROLL - 1) * TY < r1; r += UNROLL * TY) {
    float4 v[UNROLL][NQI];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
#pragma unroll
      for (int qi = 0; qi < NQI; ++qi) v[u][qi] = ld_quad(x1, x2, C1, C2, img_row + r + u * TY, img_row2 + r + u * TY, (tx + qi * TX) * 4, pol2);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
#pragma unroll
      for (int qi = 0; qi < NQI; ++qi) emit(img_row + r + u * TY, qi, v[u][qi]);
  }
  for (; r < r1; r += TY) {
#pragma unroll
    for (int qi = 0; qi < NQI; ++qi) emit(img_row + r, qi, ld_quad(x1, x2, C1, C2, img_row + r, img_row2 + r, (tx + qi * TX) * 4, pol2));
  }
  // ---------------- the last chunk of the image resets its counters for the next launch ----------------
  __syncthreads();
  if (tid == 0) {
    const unsigned int old = atomicAdd(&sync[n].done, 1u);
    if (old == static_cast<unsigned int>(n_chunks) - 1u) {
      sync[n].arrived = 0u;
      sync[n].done = 0u;
      __threadfence();
    }
  }
}

// ---- LayerNorm: one warp per row; the row lives in registers between the two passes --------------
template <int NQ>  // quads per lane: C <= 128 * NQ
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, int M, int C, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, bf16* __restrict__ out) {
  const int warps_per_block = blockDim.x >> 5;
  const int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
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
      *reinterpret_cast<uint2*>(orow + qd * 4) = make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
    }
  }
}

template <int NQI>
cudaError_t launch_gn_t(const GnGeom& g, const float* x1, int C1, const float* x2, int C2, int n_img, int hw,
                        const float* gamma, const float* beta, float eps, int apply_silu, bf16* out, bf16* raw_out,
                        float* partial, cudaStream_t stream, int x2_G, int x2_V, int x2_R, int n_layout) {
  constexpr int UNROLL = (NQI >= 4) ? 2 : 4;
  const int C = C1 + C2;
  dim3 grid(g.n_chunks, n_img), block(g.TX, g.TY);
  // the barrier counters sit behind the partials of the image count the scratch was laid out (and zeroed) for
  GnSync* sync = reinterpret_cast<GnSync*>(partial + static_cast<size_t>(n_layout) * GN_MAX_CHUNKS * GN_GROUPS * 2);
  gn_fused_kernel<NQI, UNROLL><<<grid, block, static_cast<size_t>(2) * C * g.TY * sizeof(float), stream>>>(
      x1, x2, C1, C2, hw, g.cpg, g.rows_per_chunk, g.n_chunks, partial, sync, gamma, beta, eps, apply_silu, out,
      raw_out, x2_G, x2_V, x2_R);
  return cudaGetLastError();
}

}  // namespace

// per-chunk partials followed by the per-image barrier counters (which must start zeroed, see
// groupnorm_sync_offset; the kernel leaves them zeroed again)
size_t groupnorm_partial_bytes(int n_img) {  // partials | GnSync[n_img]
  return static_cast<size_t>(n_img) * GN_MAX_CHUNKS * GN_GROUPS * 2 * sizeof(float) +
         static_cast<size_t>(n_img) * sizeof(GnSync);
}
size_t groupnorm_sync_offset(int n_img) { return static_cast<size_t>(n_img) * GN_MAX_CHUNKS * GN_GROUPS * 2 * sizeof(float); }

cudaError_t launch_groupnorm(const float* x1, int C1, const float* x2, int C2, int n_img, int hw, const float* gamma,
                             const float* beta, float eps, int apply_silu, bf16* out, bf16* raw_out, float* partial,
                             cudaStream_t stream, int x2_G, int x2_V, int x2_R, int n_img_layout) {
  const int C = C1 + C2;
  const int n_layout = n_img_layout > 0 ? n_img_layout : n_img;
  if (C % GN_GROUPS != 0 || C1 % 4 != 0 || C2 % 4 != 0) {
    set_error("groupnorm: channels must be a multiple of 32 (each source a multiple of 4)");
    return cudaErrorInvalidValue;
  }
  GnGeom g = gn_geometry(C1, C2, hw, n_img);
  if (g.nqi > GN_MAX_QI || static_cast<size_t>(2) * C * g.TY * sizeof(float) > 48 * 1024) {
    set_error("groupnorm: too many channels for this kernel");
    return cudaErrorInvalidValue;
  }
#define CAP4D_GN_CASE(N) \
  return launch_gn_t<N>(g, x1, C1, x2, C2, n_img, hw, gamma, beta, eps, apply_silu, out, raw_out, partial, stream, \
                        x2_G, x2_V, x2_R, n_layout)
  if (g.nqi == 1) CAP4D_GN_CASE(1);
  if (g.nqi == 2) CAP4D_GN_CASE(2);
  CAP4D_GN_CASE(4);
#undef CAP4D_GN_CASE
}

cudaError_t launch_layernorm(const float* x, int M, int C, const float* gamma, const float* beta, float eps,
                             bf16* out, cudaStream_t stream) {
  if (C % 4 != 0 || C > 4 * 32 * 16) {
    set_error("layernorm: C must be a multiple of 4 and <= 2048");
    return cudaErrorInvalidValue;
  }
  const int warps = 8;
  const int grid = (M + warps - 1) / warps;
  const int quads = C / 4;
  if (quads <= 32 * 4)
    layernorm_kernel<4><<<grid, warps * 32, 0, stream>>>(x, M, C, gamma, beta, eps, out);
  else if (quads <= 32 * 8)
    layernorm_kernel<8><<<grid, warps * 32, 0, stream>>>(x, M, C, gamma, beta, eps, out);
  else
    layernorm_kernel<16><<<grid, warps * 32, 0, stream>>>(x, M, C, gamma, beta, eps, out);
  return cudaGetLastError();
}

}  // namespace cap4d

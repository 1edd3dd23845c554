// Fused GroupNorm(32)+SiLU and LayerNorm for NHWC fp32 activations -> bf16 MMA operands.
//
// Reference: GroupNorm32 / LayerNorm32 (controlnet/ldm/modules/diffusionmodules/util.py:217-223,
// fp32 statistics), used as `normalization(ch)` + SiLU in ResBlocks and `out`
// (openaimodel.py:180-184, 222-231, 770-774; eps 1e-5) and as `Normalize` in the transformer
// (cap4d/mmdm/net/attention.py:107-109; eps 1e-6, no SiLU); LayerNorm32 at attention.py:311-326.
//
// These kernels are HBM-bound: fp32 in (read twice, the second time mostly from L2), bf16 out.
// The up-path ResBlocks normalise the channel concatenation of two tensors whose 32 groups
// straddle the seam (openaimodel.py:766 via mmdm_unet.py:115); both sources are read in place and
// the concatenation only ever exists as the bf16 output.
#include "kernels.h"
#include "ptx.cuh"

namespace cap4d {

namespace {

constexpr int GN_GROUPS = 32;
constexpr int GN_MAX_QI = 8;  // quads per thread along C (C <= 4 * 256 * 8)

struct GnGeom {
  int C, C1, quads, TX, TY, nqi, cpg;
  int rows_per_chunk, n_chunks;
};

__host__ GnGeom gn_geometry(int C1, int C2, int hw) {
  GnGeom g;
  g.C = C1 + C2;
  g.C1 = C1;
  g.quads = g.C / 4;
  g.cpg = g.C / GN_GROUPS;
  int tx = 1;
  for (int d = 1; d <= 256 && d <= g.quads; ++d)
    if (g.quads % d == 0) tx = d;
  g.TX = tx;
  g.nqi = g.quads / tx;
  g.TY = 256 / tx;
  if (g.TY < 1) g.TY = 1;
  if (g.TY > hw) g.TY = hw;
  int rpc = (hw + GN_MAX_CHUNKS - 1) / GN_MAX_CHUNKS;
  if (rpc < 4 * g.TY) rpc = 4 * g.TY;
  if (rpc > hw) rpc = hw;
  g.rows_per_chunk = rpc;
  g.n_chunks = (hw + rpc - 1) / rpc;
  return g;
}

__device__ __forceinline__ float4 ld_quad(const float* __restrict__ x1, const float* __restrict__ x2, int C1, int C2,
                                          size_t row, int c) {
  // c is a multiple of 4 and C1 is a multiple of 4, so a quad never straddles the seam
  if (c < C1) return __ldg(reinterpret_cast<const float4*>(x1 + row * C1 + c));
  return __ldg(reinterpret_cast<const float4*>(x2 + row * C2 + (c - C1)));
}

// grid (n_chunks, n_img), block (TX, TY)
__global__ void gn_stats_kernel(const float* __restrict__ x1, const float* __restrict__ x2, int C1, int C2, int hw,
                                int nqi, int cpg, int rows_per_chunk, float* __restrict__ partial) {
  extern __shared__ float s_ch[];  // [TY][2][C]: per-row-lane channel partials (no atomics: deterministic)
  const int C = C1 + C2;
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int tx = threadIdx.x, ty = threadIdx.y, TX = blockDim.x, TY = blockDim.y;
  const int tid = ty * TX + tx;
  const int r0 = chunk * rows_per_chunk;
  const int r1 = min(hw, r0 + rows_per_chunk);
  float sum[GN_MAX_QI][4], sq[GN_MAX_QI][4];
#pragma unroll
  for (int qi = 0; qi < GN_MAX_QI; ++qi)
#pragma unroll
    for (int k = 0; k < 4; ++k) sum[qi][k] = sq[qi][k] = 0.f;
  for (int r = r0 + ty; r < r1; r += TY) {
    const size_t row = static_cast<size_t>(n) * hw + r;
#pragma unroll
    for (int qi = 0; qi < GN_MAX_QI; ++qi) {
      if (qi < nqi) {
        const float4 v = ld_quad(x1, x2, C1, C2, row, (tx + qi * TX) * 4);
        sum[qi][0] += v.x; sq[qi][0] += v.x * v.x;
        sum[qi][1] += v.y; sq[qi][1] += v.y * v.y;
        sum[qi][2] += v.z; sq[qi][2] += v.z * v.z;
        sum[qi][3] += v.w; sq[qi][3] += v.w * v.w;
      }
    }
  }
  float* my = s_ch + static_cast<size_t>(ty) * 2 * C;
#pragma unroll
  for (int qi = 0; qi < GN_MAX_QI; ++qi) {
    if (qi < nqi) {
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
    dst[0] = s;
    dst[1] = q;
  }
}

// grid (n_chunks, n_img), block (TX, TY)
__global__ void gn_apply_kernel(const float* __restrict__ x1, const float* __restrict__ x2, int C1, int C2, int hw,
                                int nqi, int cpg, int rows_per_chunk, int n_chunks,
                                const float* __restrict__ partial, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float eps, int apply_silu, bf16* __restrict__ out,
                                bf16* __restrict__ raw_out) {
  __shared__ float s_mean[GN_GROUPS], s_rstd[GN_GROUPS];
  const int C = C1 + C2;
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int tx = threadIdx.x, ty = threadIdx.y, TX = blockDim.x, TY = blockDim.y;
  const int tid = ty * TX + tx;
  if (tid < GN_GROUPS) {
    // fixed-order reduction of the per-chunk partials: deterministic
    double s = 0.0, q = 0.0;
    for (int ch = 0; ch < n_chunks; ++ch) {
      const float* src = partial + ((static_cast<size_t>(n) * GN_MAX_CHUNKS + ch) * GN_GROUPS + tid) * 2;
      s += src[0];
      q += src[1];
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
  float a[GN_MAX_QI][4], b[GN_MAX_QI][4];
#pragma unroll
  for (int qi = 0; qi < GN_MAX_QI; ++qi) {
    if (qi < nqi) {
      const int c = (tx + qi * TX) * 4;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int g = (c + k) / cpg;
        const float ga = __ldg(gamma + c + k), be = __ldg(beta + c + k);
        a[qi][k] = s_rstd[g] * ga;
        b[qi][k] = be - s_mean[g] * s_rstd[g] * ga;
      }
    }
  }
  const int r0 = chunk * rows_per_chunk;
  const int r1 = min(hw, r0 + rows_per_chunk);
  for (int r = r0 + ty; r < r1; r += TY) {
    const size_t row = static_cast<size_t>(n) * hw + r;
#pragma unroll
    for (int qi = 0; qi < GN_MAX_QI; ++qi) {
      if (qi < nqi) {
        const int c = (tx + qi * TX) * 4;
        const float4 v = ld_quad(x1, x2, C1, C2, row, c);
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
        uint2 u = make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
        *reinterpret_cast<uint2*>(out + row * C + c) = u;
        if (raw_out != nullptr) {
          uint2 w = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
          *reinterpret_cast<uint2*>(raw_out + row * C + c) = w;
        }
      }
    }
  }
}

constexpr int LN_MAX_Q = 16;  // C <= 4 * 32 * 16 = 2048

// one warp per row; the row lives in registers between the two passes
__global__ void layernorm_kernel(const float* __restrict__ x, int M, int C, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float eps, bf16* __restrict__ out) {
  const int warps_per_block = blockDim.x >> 5;
  const int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const int quads = C >> 2;
  const float* xr = x + static_cast<size_t>(row) * C;
  float4 v[LN_MAX_Q];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_Q; ++i) {
    const int qd = i * 32 + lane;
    if (qd < quads) {
      v[i] = __ldg(reinterpret_cast<const float4*>(xr) + qd);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_Q; ++i) {
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
  for (int i = 0; i < LN_MAX_Q; ++i) {
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

}  // namespace

size_t groupnorm_partial_bytes(int n_img) {
  return static_cast<size_t>(n_img) * GN_MAX_CHUNKS * GN_GROUPS * 2 * sizeof(float);
}

cudaError_t launch_groupnorm(const float* x1, int C1, const float* x2, int C2, int n_img, int hw, const float* gamma,
                             const float* beta, float eps, int apply_silu, bf16* out, bf16* raw_out, float* partial,
                             cudaStream_t stream) {
  const int C = C1 + C2;
  if (C % GN_GROUPS != 0 || C1 % 4 != 0 || C2 % 4 != 0) {
    set_error("groupnorm: channels must be a multiple of 32 (each source a multiple of 4)");
    return cudaErrorInvalidValue;
  }
  GnGeom g = gn_geometry(C1, C2, hw);
  if (g.nqi > GN_MAX_QI || static_cast<size_t>(2) * C * g.TY * sizeof(float) > 48 * 1024) {
    set_error("groupnorm: too many channels for this kernel");
    return cudaErrorInvalidValue;
  }
  dim3 grid(g.n_chunks, n_img), block(g.TX, g.TY);
  gn_stats_kernel<<<grid, block, static_cast<size_t>(2) * C * g.TY * sizeof(float), stream>>>(x1, x2, C1, C2, hw, g.nqi, g.cpg,
                                                                  g.rows_per_chunk, partial);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  gn_apply_kernel<<<grid, block, 0, stream>>>(x1, x2, C1, C2, hw, g.nqi, g.cpg, g.rows_per_chunk, g.n_chunks,
                                              partial, gamma, beta, eps, apply_silu, out, raw_out);
  return cudaGetLastError();
}

cudaError_t launch_layernorm(const float* x, int M, int C, const float* gamma, const float* beta, float eps,
                             bf16* out, cudaStream_t stream) {
  if (C % 4 != 0 || C > 4 * 32 * LN_MAX_Q) {
    set_error("layernorm: C must be a multiple of 4 and <= 2048");
    return cudaErrorInvalidValue;
  }
  const int warps = 8;
  layernorm_kernel<<<(M + warps - 1) / warps, warps * 32, 0, stream>>>(x, M, C, gamma, beta, eps, out);
  return cudaGetLastError();
}

}  // namespace cap4d

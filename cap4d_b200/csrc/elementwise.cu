// Data-movement and small fp32 kernels around the tensor-core path: input pack (ref-mask mix +
// im2col of the 4-channel latent + pose conditioning), output mix, up/down-sampling layout
// transforms, the timestep-embedding MLPs, weight repacking and the fused CFG + DDIM update.
#include <cuda_fp16.h>

#include "kernels.h"
#include "ptx.cuh"

namespace cap4d {

namespace {

// ---- input pack: mmdm_unet.py:77-95 -----------------------------------------------------------
// out[(n,y,x)][k]: k < 9*cin   -> tap (ky,kx) = k / cin, channel ci = k % cin of the masked latent
//                  k < 9*cin+cc -> pos_enc[n,y,x,k-9*cin]
//                  else 0
// One thread per (pixel, 8 consecutive k): the pixel's coordinates are decoded once per 8 outputs and the row leaves
// in 16-byte pieces (one thread per 2-byte element, with its divisions, made this the slowest of the small kernels:
// 283 us per 5-group call).
__device__ __forceinline__ float input_pack_value(const float* __restrict__ x, const float* __restrict__ z,
                                                  const float* __restrict__ mask, const float* __restrict__ pos, int n,
                                                  int yy, int xx, size_t pix, int k, int cin, int H, int W, int cc) {
  if (k < 9 * cin) {
    const int tap = k / cin, ci = k - tap * cin;
    const int sy = yy + tap / 3 - 1, sx = xx + tap % 3 - 1;
    if (sy < 0 || sy >= H || sx < 0 || sx >= W) return 0.f;
    const size_t sp = static_cast<size_t>(sy) * W + sx;
    const float m = mask[static_cast<size_t>(n) * H * W + sp];
    const size_t idx = (static_cast<size_t>(n) * cin + ci) * H * W + sp;
    // x = z_input * ref_mask + x * logical_not(ref_mask)
    return z[idx] * m + x[idx] * (m == 0.f ? 1.f : 0.f);
  }
  if (k < 9 * cin + cc) return pos[pix * cc + (k - 9 * cin)];
  return 0.f;
}

__global__ void input_pack_kernel(const float* __restrict__ x, const float* __restrict__ z,
                                  const float* __restrict__ mask, const float* __restrict__ pos, int n_img, int cin,
                                  int H, int W, int cc, int kpad, bf16* __restrict__ out, int out_f16) {
  const int kg = kpad >> 3;  // groups of 8 per row (kpad is a multiple of 64)
  const size_t total = static_cast<size_t>(n_img) * H * W * kg;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int k0 = static_cast<int>(i % kg) * 8;
    const size_t pix = i / kg;
    const int xx = static_cast<int>(pix % W);
    const int yy = static_cast<int>((pix / W) % H);
    const int n = static_cast<int>(pix / (static_cast<size_t>(W) * H));
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = input_pack_value(x, z, mask, pos, n, yy, xx, pix, k0 + j, cin, H, W, cc);
    if (out_f16 == 2) {
      float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + pix * kpad + k0);
      dst[0] = make_float4(v[0], v[1], v[2], v[3]);
      dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      *reinterpret_cast<uint4*>(out + pix * kpad + k0) =
          make_uint4(pack16x2(v[0], v[1], out_f16), pack16x2(v[2], v[3], out_f16), pack16x2(v[4], v[5], out_f16),
                     pack16x2(v[6], v[7], out_f16));
    }
  }
}

// ---- output mix: mmdm_unet.py:77,122-125 ------------------------------------------------------
// G > 0: h holds only the generated views (B*G images, view v >= R of group b at image b*G + v - R); the
// reference views' outputs do not depend on h at all (ref_mask == 1 there).
__global__ void output_mix_kernel(const float* __restrict__ h, int ldh, const float* __restrict__ x,
                                  const float* __restrict__ z, const float* __restrict__ mask, int n_img, int cout,
                                  int H, int W, int G, int V, int R, float* __restrict__ out,
                                  int* __restrict__ violations) {
  const size_t total = static_cast<size_t>(n_img) * cout * H * W;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t sp = i % (static_cast<size_t>(H) * W);
    const int c = static_cast<int>((i / (static_cast<size_t>(H) * W)) % cout);
    const size_t n = i / (static_cast<size_t>(H) * W * cout);
    const float m = mask[n * H * W + sp];
    float hv = 0.f;
    if (G <= 0) {
      hv = h[(n * H * W + sp) * ldh + c];
    } else {
      const int b = static_cast<int>(n) / V, v = static_cast<int>(n) % V;
      if (v >= R) {
        hv = h[((static_cast<size_t>(b) * G + (v - R)) * H * W + sp) * ldh + c];
      } else if (m == 0.f) {
        // the caller promised ref_mask == 1 on the first R views (cap4d_b200_unet_set_ref_views) and the network
        // never computed this view: fail loudly (NaN in the output, counted for ..._ref_view_violations)
        hv = __int_as_float(0x7fc00000);
        if (violations != nullptr && c == 0 && sp == 0) atomicAdd(violations, 1);
      }
    }
    out[i] = (x[i] - z[i]) * m + hv * (m == 0.f ? 1.f : 0.f);
  }
}

// ---- keep the generated views: dst image (b, g) = src image (b, R + g), per_img floats each -----------
__global__ void gather_views_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int B, int V, int R,
                                    size_t quads_per_img) {
  const int G = V - R;
  const size_t total = static_cast<size_t>(B) * G * quads_per_img;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t img = i / quads_per_img, off = i - img * quads_per_img;
    const size_t b = img / G, g = img - b * G;
    dst[i] = __ldg(src + (b * V + R + g) * quads_per_img + off);
  }
}

// ---- fp32 NHWC -> bf16 parity planes for the stride-2 conv (openaimodel.py:150-153) ------------
__global__ void parity_split_kernel(const float* __restrict__ x, int n_img, int H, int W, int C,
                                    bf16* __restrict__ out, int out_f32) {
  const int quads = C >> 2;
  const size_t total = static_cast<size_t>(n_img) * H * W * quads;
  const int H2 = H >> 1, W2 = W >> 1;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int qd = static_cast<int>(i % quads);
    const size_t pix = i / quads;
    const int xx = static_cast<int>(pix % W);
    const int yy = static_cast<int>((pix / W) % H);
    const size_t n = pix / (static_cast<size_t>(W) * H);
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    const int plane = (yy & 1) * 2 + (xx & 1);
    const size_t o = (((static_cast<size_t>(plane) * n_img + n) * H2 + (yy >> 1)) * W2 + (xx >> 1)) * C + qd * 4;
    if (out_f32)
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + o) = v;
    else
      *reinterpret_cast<uint2*>(out + o) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

// ---- timestep embedding: util.py:154-174 -------------------------------------------------------
__global__ void timestep_embedding_kernel(const long long* __restrict__ t, int n_img, int dim, float* __restrict__ out) {
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_img * half) return;
  const int n = i / half, k = i - n * half;
  // freqs = exp(-ln(10000) * k / half) in fp32, like torch
  const float f = expf(-9.210340371976184f * static_cast<float>(k) / static_cast<float>(half));
  const float a = static_cast<float>(t[n]) * f;
  out[static_cast<size_t>(n) * dim + k] = cosf(a);
  out[static_cast<size_t>(n) * dim + half + k] = sinf(a);
  if ((dim & 1) && k == 0) out[static_cast<size_t>(n) * dim + dim - 1] = 0.f;
}

// ---- distinct timesteps: the embedding MLP is a pure function of t, and inside the sampler every view of
// a step carries the same t (sampler.py:126), so it is evaluated once per DISTINCT value and copied.
// rep[r] = first row with the same timestep as row r; uniq = {count, rows with rep[r] == r ...}
__global__ void timestep_unique_kernel(const long long* __restrict__ t, int n, int* __restrict__ rep,
                                       int* __restrict__ uniq) {
  for (int r = threadIdx.x; r < n; r += blockDim.x) {
    int first = r;
    for (int q = 0; q < r; ++q)
      if (t[q] == t[r]) {
        first = q;
        break;
      }
    rep[r] = first;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int cnt = 0;
    for (int r = 0; r < n; ++r)
      if (rep[r] == r) uniq[1 + cnt++] = r;
    uniq[0] = cnt;
  }
}

__global__ void broadcast_rows_kernel(float* __restrict__ out, int n_rows, int n_cols, const int* __restrict__ rep) {
  const size_t total = static_cast<size_t>(n_rows) * n_cols;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / n_cols);
    const int src = rep[r];
    if (src != r) out[i] = out[static_cast<size_t>(src) * n_cols + (i - static_cast<size_t>(r) * n_cols)];
  }
}

// ---- skinny fp32 linear: out[n][j] = act(sum_k W[j][k] in[n][k] + b[j]) for the rows listed in uniq -------
// one warp per output column j; lanes split K; 16 rows of `in` per pass
__global__ void skinny_linear_kernel(const float* __restrict__ in, const int* __restrict__ uniq, int K,
                                     const float* __restrict__ Wm, const float* __restrict__ bias, int n_out,
                                     int silu_out, float* __restrict__ out) {
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= n_out) return;
  const int n_rows = uniq[0];
  const float* wr = Wm + static_cast<size_t>(j) * K;
  for (int r0 = 0; r0 < n_rows; r0 += 16) {
    float acc[16];
    int row[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      acc[r] = 0.f;
      row[r] = (r0 + r < n_rows) ? uniq[1 + r0 + r] : -1;
    }
    for (int k = lane * 4; k < K; k += 128) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(wr + k));
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        if (row[r] >= 0) {
          const float4 x = __ldg(reinterpret_cast<const float4*>(in + static_cast<size_t>(row[r]) * K + k));
          acc[r] += w.x * x.x + w.y * x.y + w.z * x.z + w.w * x.w;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      float v = acc[r];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && row[r] >= 0) {
        v += bias[j];
        if (silu_out) v = v / (1.0f + expf(-v));
        out[static_cast<size_t>(row[r]) * n_out + j] = v;
      }
    }
  }
}

// ---- VAE decode helpers (controlnet/ldm/models/autoencoder.py:87-91, diffusionmodules/model.py) -------
// Input stage: z / scale_factor (ddpm.py:829) -> post_quant_conv (1x1, 4 -> 4) -> im2col rows for conv_in
// (3x3, pad 1): out[n*h*w + y*w + x][tap*zc + c], K padded with zeros to kpad.  Padding taps are zero AFTER
// post_quant_conv (conv_in pads its own input).
__global__ void vae_input_pack_kernel(const float* __restrict__ z, int n_img, int zc, int H, int W,
                                      const float* __restrict__ wpq, const float* __restrict__ bpq, float inv_scale,
                                      int kpad, bf16* __restrict__ out) {
  const size_t total = static_cast<size_t>(n_img) * H * W * 9;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int tap = static_cast<int>(i % 9);
    const size_t pix = i / 9;
    const int x = static_cast<int>(pix % W), y = static_cast<int>((pix / W) % H);
    const size_t n = pix / (static_cast<size_t>(W) * H);
    const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
    bf16* dst = out + pix * kpad + tap * zc;
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) {
      for (int c = 0; c < zc; ++c) dst[c] = __float2bfloat16(0.f);
    } else {
      for (int co = 0; co < zc; ++co) {
        float v = bpq[co];
        for (int ci = 0; ci < zc; ++ci)
          v = fmaf(wpq[co * zc + ci], z[((n * zc + ci) * H + yy) * W + xx] * inv_scale, v);
        dst[co] = __float2bfloat16(v);
      }
    }
    if (tap == 8)
      for (int k = 9 * zc; k < kpad; ++k) out[pix * kpad + k] = __float2bfloat16(0.f);
  }
}

// h fp32 NHWC [n*H*W][ldh] (first cout columns) -> images fp32 NCHW [n][cout][H][W]
__global__ void vae_output_kernel(const float* __restrict__ h, int ldh, int n_img, int cout, int H, int W,
                                  float* __restrict__ out) {
  const size_t total = static_cast<size_t>(n_img) * cout * H * W;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t sp = i % (static_cast<size_t>(H) * W);
    const int c = static_cast<int>((i / (static_cast<size_t>(H) * W)) % cout);
    const size_t n = i / (static_cast<size_t>(H) * W * cout);
    out[i] = h[(n * H * W + sp) * ldh + c];
  }
}

// h fp32 NHWC [n*H*W][ldh] (first 3 columns = R, G, B in [-1, 1]) -> uint8 HWC with the channels reversed,
// exactly the array the reference hands to cv2.imwrite (cap4d/inference/utils.py:134-137):
// ((x + 1) / 2).clip(0, 1) * 255 truncated to uint8, channels [2, 1, 0].
__global__ void vae_output_u8_kernel(const float* __restrict__ h, int ldh, size_t n_pix, unsigned char* __restrict__ out) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n_pix;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = __fdiv_rn(__fadd_rn(h[i * ldh + (2 - c)], 1.0f), 2.0f);
      v = fminf(fmaxf(v, 0.0f), 1.0f);
      out[i * 3 + c] = static_cast<unsigned char>(__fmul_rn(v, 255.0f));
    }
  }
}

// softmax over the rows of fp32 scores [rows][L] (times scale) -> bf16 probabilities; one block per row
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, int L, float scale,
                                                           bf16* __restrict__ p) {
  __shared__ float red[8];
  const float* row = s + static_cast<size_t>(blockIdx.x) * L;
  bf16* prow = p + static_cast<size_t>(blockIdx.x) * L;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float m = -INFINITY;
  for (int i = threadIdx.x * 4; i < L; i += blockDim.x * 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(row + i));
    m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int i = threadIdx.x * 4; i < L; i += blockDim.x * 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(row + i));
    sum += expf((v.x - m) * scale) + expf((v.y - m) * scale) + expf((v.z - m) * scale) + expf((v.w - m) * scale);
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) sum += red[w];
  const float inv = 1.0f / sum;
  for (int i = threadIdx.x * 4; i < L; i += blockDim.x * 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(row + i));
    const uint2 u = make_uint2(pack_bf16x2(expf((v.x - m) * scale) * inv, expf((v.y - m) * scale) * inv),
                               pack_bf16x2(expf((v.z - m) * scale) * inv, expf((v.w - m) * scale) * inv));
    *reinterpret_cast<uint2*>(prow + i) = u;
  }
}

// dst[c][r] = src[r][c] for r < rows, c < cols (src row stride ld): 32x32 tiles through shared memory
__global__ void transpose_bf16_kernel(const bf16* __restrict__ src, int ld, int rows, int cols, bf16* __restrict__ dst) {
  __shared__ bf16 tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = src[static_cast<size_t>(r) * ld + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[static_cast<size_t>(c) * rows + r] = tile[threadIdx.x][j];
  }
}

// ---- weight repacks ---------------------------------------------------------------------------
// storage of a repacked weight (kernels.h: set_weight_pack_format): 0 bf16, 1 the fp16 bit pattern in the same two
// bytes, 2 fp32 (out then points at floats: the fp32-accuracy mode splits them afterwards, precise.cu)
__device__ __forceinline__ void store_w(bf16* out, size_t idx, float v, int fmt) {
  if (fmt == 2) {
    reinterpret_cast<float*>(out)[idx] = v;
  } else if (fmt == 1) {
    const __half h = __float2half_rn(v);
    out[idx] = *reinterpret_cast<const bf16*>(&h);
  } else {
    out[idx] = __float2bfloat16(v);
  }
}

__global__ void pack_conv_weight_kernel(const float* __restrict__ w, int O, int I, int KH, int KW,
                                        bf16* __restrict__ out, int ldk, int k_offset, int f16) {
  const size_t total = static_cast<size_t>(O) * I * KH * KW;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    // destination-major enumeration: (o, tap, i)
    const int i = static_cast<int>(idx % I);
    const int tap = static_cast<int>((idx / I) % (KH * KW));
    const size_t o = idx / (static_cast<size_t>(I) * KH * KW);
    const float v = w[(o * I + i) * KH * KW + tap];
    store_w(out, o * ldk + k_offset + static_cast<size_t>(tap) * I + i, v, f16);
  }
}

__global__ void pack_matrix_kernel(const float* __restrict__ w, int rows, int cols, bf16* __restrict__ out, int ldk,
                                   int k_offset, int row_offset, int f16) {
  const size_t total = static_cast<size_t>(rows) * cols;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t r = idx / cols, c = idx % cols;
    store_w(out, (row_offset + r) * ldk + k_offset + c, w[idx], f16);
  }
}

// ---- folded upsample-conv weights (openaimodel.py:111-119) ---------------------------------------
// phase py: output row 2y+py reads low-res rows {y-1 (ky=0), y (ky=1,2)} if py == 0, {y (ky=0,1), y+1 (ky=2)} if py == 1
__global__ void pack_upconv_weight_kernel(const float* __restrict__ w, int O, int I, bf16* __restrict__ out, int f16) {
  const size_t total = static_cast<size_t>(4) * O * 4 * I;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(idx % I);
    const int tap = static_cast<int>((idx / I) % 4);
    const size_t o = (idx / (static_cast<size_t>(4) * I)) % O;
    const int phase = static_cast<int>(idx / (static_cast<size_t>(4) * I * O));
    const int py = phase >> 1, px = phase & 1, a = tap >> 1, b = tap & 1;
    // 3x3 taps merged into low-res tap a (rows) / b (columns)
    const int ky0 = (py == 0) ? (a == 0 ? 0 : 1) : (a == 0 ? 0 : 2);
    const int ky1 = (py == 0) ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : 2);
    const int kx0 = (px == 0) ? (b == 0 ? 0 : 1) : (b == 0 ? 0 : 2);
    const int kx1 = (px == 0) ? (b == 0 ? 0 : 2) : (b == 0 ? 1 : 2);
    float acc = 0.f;
    for (int ky = ky0; ky <= ky1; ++ky)
      for (int kx = kx0; kx <= kx1; ++kx) acc += w[((o * I + i) * 3 + ky) * 3 + kx];
    store_w(out, idx, acc, f16);
  }
}

__global__ void cast_bf16_kernel(const float* __restrict__ x, size_t n4, bf16* __restrict__ out) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    *reinterpret_cast<uint2*>(out + i * 4) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

// ---- CFG + DDIM update with scatter: cap4d/mmdm/sampler.py:205-231 -----------------------------
// eps layout: [2*n_groups][V][chw]; batch b < n_groups = unconditional half of group b,
// b + n_groups = conditional half.  Only views R..V-1 are generated.
__device__ __forceinline__ float ddim_one(float x, float eu, float ec, float cfg, float x_coef, float e_coef) {
  const float e = __fadd_rn(eu, __fmul_rn(cfg, __fsub_rn(ec, eu)));
  return __fadd_rn(__fmul_rn(x, x_coef), __fmul_rn(e, e_coef));
}

__global__ void cfg_ddim_kernel(float* __restrict__ latents, const float* __restrict__ eps,
                                const long long* __restrict__ gen_idx, int n_groups, int V, int R, int chw, float cfg,
                                float x_coef, float e_coef) {
  const int G = V - R;
  const int q4 = chw >> 2;
  const size_t total = static_cast<size_t>(n_groups) * G * q4;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int e = static_cast<int>(i % q4);
    const int v = static_cast<int>((i / q4) % G);
    const int g = static_cast<int>(i / (static_cast<size_t>(q4) * G));
    const float4 eu = __ldg(reinterpret_cast<const float4*>(eps + (static_cast<size_t>(g) * V + R + v) * chw) + e);
    const float4 ec =
        __ldg(reinterpret_cast<const float4*>(eps + (static_cast<size_t>(g + n_groups) * V + R + v) * chw) + e);
    const long long dst = gen_idx[g * G + v];
    float4* xp = reinterpret_cast<float4*>(latents + static_cast<size_t>(dst) * chw) + e;
    float4 xv = *xp;
    // model_output = uncond + cfg * (cond - uncond);  x = x * x_coef + e_t * e_coef
    // (separately rounded mul/add like the reference's eager ops: no FMA contraction)
    xv.x = ddim_one(xv.x, eu.x, ec.x, cfg, x_coef, e_coef);
    xv.y = ddim_one(xv.y, eu.y, ec.y, cfg, x_coef, e_coef);
    xv.z = ddim_one(xv.z, eu.z, ec.z, cfg, x_coef, e_coef);
    xv.w = ddim_one(xv.w, eu.w, ec.w, cfg, x_coef, e_coef);
    *xp = xv;
  }
}

inline int grid_for(size_t total, int block) {
  size_t g = (total + block - 1) / block;
  const size_t cap = static_cast<size_t>(sm_count()) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace

cudaError_t launch_input_pack(const float* x, const float* z_input, const float* ref_mask, const float* pos_enc,
                              int n_img, int cin, int H, int W, int ccond, int kpad, bf16* out,
                              cudaStream_t stream, int out_f16) {
  if (kpad % 8 != 0) {
    set_error("input_pack: the padded K must be a multiple of 8");
    return cudaErrorInvalidValue;
  }
  const size_t total = static_cast<size_t>(n_img) * H * W * (kpad / 8);
  input_pack_kernel<<<grid_for(total, 256), 256, 0, stream>>>(x, z_input, ref_mask, pos_enc, n_img, cin, H, W, ccond,
                                                              kpad, out, out_f16);
  return cudaGetLastError();
}

cudaError_t launch_output_mix(const float* h, int ldh, const float* x, const float* z_input, const float* ref_mask,
                              int n_img, int cout, int H, int W, int G, int V, int R, float* out,
                              cudaStream_t stream, int* violations) {
  const size_t total = static_cast<size_t>(n_img) * cout * H * W;
  output_mix_kernel<<<grid_for(total, 256), 256, 0, stream>>>(h, ldh, x, z_input, ref_mask, n_img, cout, H, W, G, V, R,
                                                              out, violations);
  return cudaGetLastError();
}

cudaError_t launch_gather_views(const float* src, float* dst, int B, int V, int R, size_t per_img, cudaStream_t stream) {
  if (per_img % 4 != 0 || R < 0 || R >= V) {
    set_error("gather_views: per-image size must be a multiple of 4 floats and 0 <= R < V");
    return cudaErrorInvalidValue;
  }
  const size_t total = static_cast<size_t>(B) * (V - R) * (per_img / 4);
  gather_views_kernel<<<grid_for(total, 256), 256, 0, stream>>>(reinterpret_cast<const float4*>(src),
                                                                reinterpret_cast<float4*>(dst), B, V, R, per_img / 4);
  return cudaGetLastError();
}

cudaError_t launch_parity_split_bf16(const float* x, int n_img, int H, int W, int C, bf16* out,
                                     cudaStream_t stream, int out_f32) {
  const size_t total = static_cast<size_t>(n_img) * H * W * (C / 4);
  parity_split_kernel<<<grid_for(total, 256), 256, 0, stream>>>(x, n_img, H, W, C, out, out_f32);
  return cudaGetLastError();
}

size_t time_embed_scratch_bytes(int n_img, int model_ch, int emb_ch) {  // temb | h1 | silu(emb) | rep | uniq
  return static_cast<size_t>(n_img) * (model_ch + 2 * emb_ch) * sizeof(float) + (2 * static_cast<size_t>(n_img) + 4) * sizeof(int);
}

cudaError_t launch_time_embed(const long long* t, int n_img, int model_ch, int emb_ch, const float* w1,
                              const float* b1, const float* w2, const float* b2, const float* wall,
                              const float* ball, int n_all, float* scratch, float* out, cudaStream_t stream) {
  if (model_ch % 4 != 0 || emb_ch % 4 != 0) {
    set_error("time_embed: channel counts must be multiples of 4");
    return cudaErrorInvalidValue;
  }
  float* temb = scratch;
  float* h1 = temb + static_cast<size_t>(n_img) * model_ch;
  float* se = h1 + static_cast<size_t>(n_img) * emb_ch;
  int* rep = reinterpret_cast<int*>(se + static_cast<size_t>(n_img) * emb_ch);
  int* uniq = rep + n_img;
  const int half = model_ch / 2;
  timestep_unique_kernel<<<1, 256, 0, stream>>>(t, n_img, rep, uniq);
  timestep_embedding_kernel<<<(n_img * half + 127) / 128, 128, 0, stream>>>(t, n_img, model_ch, temb);
  const int wpb = 8;
  // time_embed: Linear -> SiLU -> Linear (openaimodel.py:528-533); every consumer applies SiLU first
  // (ResBlock.emb_layers, openaimodel.py:203-209), so silu(emb) is what is kept.
  skinny_linear_kernel<<<(emb_ch + wpb - 1) / wpb, wpb * 32, 0, stream>>>(temb, uniq, model_ch, w1, b1, emb_ch, 1, h1);
  skinny_linear_kernel<<<(emb_ch + wpb - 1) / wpb, wpb * 32, 0, stream>>>(h1, uniq, emb_ch, w2, b2, emb_ch, 1, se);
  skinny_linear_kernel<<<(n_all + wpb - 1) / wpb, wpb * 32, 0, stream>>>(se, uniq, emb_ch, wall, ball, n_all, 0, out);
  broadcast_rows_kernel<<<grid_for(static_cast<size_t>(n_img) * n_all, 256), 256, 0, stream>>>(out, n_img, n_all, rep);
  return cudaGetLastError();
}

static thread_local int g_pack_fmt = 0;
void set_weight_pack_format(int fmt) { g_pack_fmt = fmt; }
int weight_pack_format() { return g_pack_fmt; }
void set_weight_pack_f16(bool f16) { g_pack_fmt = f16 ? 1 : 0; }

cudaError_t launch_pack_conv_weight(const float* w_oihw, int O, int I, int KH, int KW, bf16* out, int ldk,
                                    int k_offset, cudaStream_t stream) {
  const size_t total = static_cast<size_t>(O) * I * KH * KW;
  pack_conv_weight_kernel<<<grid_for(total, 256), 256, 0, stream>>>(w_oihw, O, I, KH, KW, out, ldk, k_offset, weight_pack_format());
  return cudaGetLastError();
}

cudaError_t launch_pack_upconv_weight(const float* w_oihw, int O, int I, bf16* out, cudaStream_t stream) {
  const size_t total = static_cast<size_t>(16) * O * I;
  pack_upconv_weight_kernel<<<grid_for(total, 256), 256, 0, stream>>>(w_oihw, O, I, out, weight_pack_format());
  return cudaGetLastError();
}

cudaError_t launch_cast_bf16(const float* x, size_t n, bf16* out, cudaStream_t stream) {
  if (n % 4 != 0) {
    set_error("cast_bf16: element count must be a multiple of 4");
    return cudaErrorInvalidValue;
  }
  cast_bf16_kernel<<<grid_for(n / 4, 256), 256, 0, stream>>>(x, n / 4, out);
  return cudaGetLastError();
}

cudaError_t launch_pack_matrix(const float* w, int rows, int cols, bf16* out, int ldk, int k_offset, int row_offset,
                               cudaStream_t stream) {
  const size_t total = static_cast<size_t>(rows) * cols;
  pack_matrix_kernel<<<grid_for(total, 256), 256, 0, stream>>>(w, rows, cols, out, ldk, k_offset, row_offset,
                                                               weight_pack_format());
  return cudaGetLastError();
}

cudaError_t launch_cfg_ddim_update(float* latents, const float* eps, const long long* gen_idx, int n_groups, int V,
                                   int R, int chw, float cfg, float x_coef, float e_coef, cudaStream_t stream) {
  if (chw % 4 != 0) {
    set_error("cfg_ddim_update: latent size must be a multiple of 4");
    return cudaErrorInvalidValue;
  }
  const size_t total = static_cast<size_t>(n_groups) * (V - R) * (chw / 4);
  cfg_ddim_kernel<<<grid_for(total, 256), 256, 0, stream>>>(latents, eps, gen_idx, n_groups, V, R, chw, cfg, x_coef,
                                                            e_coef);
  return cudaGetLastError();
}

cudaError_t launch_vae_input_pack(const float* z, int n_img, int zc, int H, int W, const float* wpq, const float* bpq,
                                  float inv_scale, int kpad, bf16* out, cudaStream_t stream) {
  if (9 * zc > kpad) {
    set_error("vae_input_pack: 9 * z_channels must fit the padded K");
    return cudaErrorInvalidValue;
  }
  const size_t total = static_cast<size_t>(n_img) * H * W * 9;
  vae_input_pack_kernel<<<grid_for(total, 256), 256, 0, stream>>>(z, n_img, zc, H, W, wpq, bpq, inv_scale, kpad, out);
  return cudaGetLastError();
}

cudaError_t launch_vae_output(const float* h, int ldh, int n_img, int cout, int H, int W, float* out,
                              cudaStream_t stream) {
  const size_t total = static_cast<size_t>(n_img) * cout * H * W;
  vae_output_kernel<<<grid_for(total, 256), 256, 0, stream>>>(h, ldh, n_img, cout, H, W, out);
  return cudaGetLastError();
}

cudaError_t launch_vae_output_u8(const float* h, int ldh, size_t n_pix, unsigned char* out, cudaStream_t stream) {
  vae_output_u8_kernel<<<grid_for(n_pix, 256), 256, 0, stream>>>(h, ldh, n_pix, out);
  return cudaGetLastError();
}

cudaError_t launch_softmax_rows(const float* s, int rows, int L, float scale, bf16* p, cudaStream_t stream) {
  if (L % 4 != 0) {
    set_error("softmax_rows: L must be a multiple of 4");
    return cudaErrorInvalidValue;
  }
  softmax_rows_kernel<<<rows, 256, 0, stream>>>(s, L, scale, p);
  return cudaGetLastError();
}

cudaError_t launch_transpose_bf16(const bf16* src, int ld, int rows, int cols, bf16* dst, cudaStream_t stream) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  transpose_bf16_kernel<<<grid, block, 0, stream>>>(src, ld, rows, cols, dst);
  return cudaGetLastError();
}

}  // namespace cap4d

// Device-resident data plane of the stochastic-I/O sampler: the work cap4d/mmdm/sampler.py:141-213 does with
// CPU-staged tensors per group (dict_sample, torch.cat of reference / generated views and of the unconditional /
// conditional halves, .to(device), the all_e_t scatter and the DDIM update) as three kernels over stores that
// stay in HBM for the whole sample() call.
//
//   gather   builds the U-Net inputs of n view groups, [uncond n | cond n] x V views, straight from the stores
//            through the step's index tables; the unconditional branch of CAP4DConditioning is all zeros
//            (cap4d/mmdm/conditioning/cap4dcond.py:78-88: pos_enc = 0, z_input * 0, same ref_mask), so its stores
//            may be NULL and are then neither kept nor read.
//   update   CFG combine + DDIM update scattered into the latent store (sampler.py:205-231).
//   pack / unpack   the per-step latent exchange between ranks (one all-gather in between, done by the host).
//
// Everything that changes from call to call (timestep, DDIM factors, which groups) is read from a small struct
// in DEVICE memory, so one captured CUDA graph per batch shape serves every call of every step.
#include "../../include/cap4d_b200.h"
#include "kernels.h"

namespace cap4d {
namespace {

inline int plane_grid(size_t total, int block) {
  size_t g = (total + block - 1) / block;
  const size_t cap = static_cast<size_t>(sm_count()) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

__device__ __forceinline__ float4 ld4_or_zero(const float* base, size_t row, size_t row_elems, size_t q) {
  if (base == nullptr) return make_float4(0.f, 0.f, 0.f, 0.f);
  return __ldg(reinterpret_cast<const float4*>(base + row * row_elems) + q);
}

// One virtual row per output image: [x (chw) | z_input (chw) | ref_mask (hw) | pos_enc (hw * Cc)], in float4 units.
__global__ void sampler_gather_kernel(cap4d_b200_sampler_stores st, const long long* __restrict__ ref_idx,
                                      const long long* __restrict__ gen_idx,
                                      const cap4d_b200_sampler_call* __restrict__ call, int n, int V, int R, int chw,
                                      int hw, int pos_elems, float* __restrict__ x_out, float* __restrict__ z_out,
                                      float* __restrict__ m_out, float* __restrict__ p_out,
                                      long long* __restrict__ t_out) {
  const int G = V - R;
  const int img = blockIdx.y;  // (half * n + gi) * V + v
  const int v = img % V;
  const int b2 = img / V;
  const int half = b2 / n;     // 0: unconditional, 1: conditional (sampler.py:185: cat([uncond, cond], dim=0))
  const int gi = b2 - half * n;
  const int g = call->groups[gi];
  const bool is_ref = v < R;
  const size_t src = static_cast<size_t>(is_ref ? ref_idx[static_cast<size_t>(g) * R + v]
                                                : gen_idx[static_cast<size_t>(g) * G + (v - R)]);
  // x: the reference views carry the CONDITIONAL z_input in both halves (sampler.py:189-190), the generated
  // views the current latents
  const float* xs = is_ref ? st.ref_z : st.latents;
  const float* zs = half ? (is_ref ? st.ref_z : st.gen_z) : (is_ref ? st.ref_z_u : st.gen_z_u);
  const float* ms = half ? (is_ref ? st.ref_mask : st.gen_mask) : (is_ref ? st.ref_mask_u : st.gen_mask_u);
  if (ms == nullptr) ms = is_ref ? st.ref_mask : st.gen_mask;  // the unconditional branch keeps ref_mask
  const float* ps = half ? (is_ref ? st.ref_pos : st.gen_pos) : (is_ref ? st.ref_pos_u : st.gen_pos_u);
  const int q_x = chw >> 2, q_m = hw >> 2, q_p = pos_elems >> 2;
  const int total = 2 * q_x + q_m + q_p;
  float4* xo = reinterpret_cast<float4*>(x_out + static_cast<size_t>(img) * chw);
  float4* zo = reinterpret_cast<float4*>(z_out + static_cast<size_t>(img) * chw);
  float4* mo = reinterpret_cast<float4*>(m_out + static_cast<size_t>(img) * hw);
  float4* po = reinterpret_cast<float4*>(p_out + static_cast<size_t>(img) * pos_elems);
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
    if (q < q_x) {
      xo[q] = ld4_or_zero(xs, src, chw, q);
    } else if (q < 2 * q_x) {
      zo[q - q_x] = ld4_or_zero(zs, src, chw, q - q_x);
    } else if (q < 2 * q_x + q_m) {
      mo[q - 2 * q_x] = ld4_or_zero(ms, src, hw, q - 2 * q_x);
    } else {
      po[q - 2 * q_x - q_m] = ld4_or_zero(ps, src, pos_elems, q - 2 * q_x - q_m);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) t_out[img] = call->timestep;  // sampler.py:124: one timestep for all views
}

__device__ __forceinline__ float ddim_one(float x, float eu, float ec, float cfg, float x_coef, float e_coef) {
  // model_output = uncond + cfg * (cond - uncond);  x = x * x_coef + e_t * e_coef
  // (separately rounded mul/add like the reference's eager ops: no FMA contraction)
  const float e = __fadd_rn(eu, __fmul_rn(cfg, __fsub_rn(ec, eu)));
  return __fadd_rn(__fmul_rn(x, x_coef), __fmul_rn(e, e_coef));
}

__global__ void sampler_update_kernel(float* __restrict__ latents, const float* __restrict__ eps,
                                      const long long* __restrict__ gen_idx,
                                      const cap4d_b200_sampler_call* __restrict__ call, int n, int V, int R, int chw,
                                      float cfg) {
  const int G = V - R;
  const int q4 = chw >> 2;
  const float x_coef = call->x_coef, e_coef = call->e_coef;
  const size_t total = static_cast<size_t>(n) * G * q4;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int e = static_cast<int>(i % q4);
    const int v = static_cast<int>((i / q4) % G);
    const int gi = static_cast<int>(i / (static_cast<size_t>(q4) * G));
    const float4 eu = __ldg(reinterpret_cast<const float4*>(eps + (static_cast<size_t>(gi) * V + R + v) * chw) + e);
    const float4 ec = __ldg(reinterpret_cast<const float4*>(eps + (static_cast<size_t>(gi + n) * V + R + v) * chw) + e);
    const long long dst = gen_idx[static_cast<size_t>(call->groups[gi]) * G + v];
    float4* xp = reinterpret_cast<float4*>(latents + static_cast<size_t>(dst) * chw) + e;
    float4 xv = *xp;
    xv.x = ddim_one(xv.x, eu.x, ec.x, cfg, x_coef, e_coef);
    xv.y = ddim_one(xv.y, eu.y, ec.y, cfg, x_coef, e_coef);
    xv.z = ddim_one(xv.z, eu.z, ec.z, cfg, x_coef, e_coef);
    xv.w = ddim_one(xv.w, eu.w, ec.w, cfg, x_coef, e_coef);
    *xp = xv;
  }
}

// rank r owns groups r, r + world, ... (sampler.py:151-158); slot k of rank r is group r + k * world
__global__ void sampler_pack_kernel(const float* __restrict__ latents, const long long* __restrict__ gen_idx, int n_its,
                                    int G, int chw, int rank, int world, float* __restrict__ send) {
  const int q4 = chw >> 2;
  const int mine = (n_its - rank + world - 1) / world;
  const size_t total = static_cast<size_t>(mine) * G * q4;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int e = static_cast<int>(i % q4);
    const size_t row = i / q4;  // k * G + j
    const int k = static_cast<int>(row / G), j = static_cast<int>(row % G);
    const long long src = gen_idx[static_cast<size_t>(rank + k * world) * G + j];
    reinterpret_cast<float4*>(send + row * chw)[e] =
        __ldg(reinterpret_cast<const float4*>(latents + static_cast<size_t>(src) * chw) + e);
  }
}

__global__ void sampler_unpack_kernel(float* __restrict__ latents, const float* __restrict__ recv,
                                      const long long* __restrict__ gen_idx, int n_its, int G, int chw, int rank,
                                      int world, int per_rank) {
  const int q4 = chw >> 2;
  const size_t total = static_cast<size_t>(n_its) * G * q4;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int e = static_cast<int>(i % q4);
    const size_t row = i / q4;  // g * G + j over ALL groups of the step
    const int g = static_cast<int>(row / G), j = static_cast<int>(row % G);
    const int owner = g % world;
    if (owner == rank) continue;  // updated in place by this rank
    const int k = g / world;
    const long long dst = gen_idx[row];
    const float4 val =
        __ldg(reinterpret_cast<const float4*>(recv + (static_cast<size_t>(owner) * per_rank * G + static_cast<size_t>(k) * G + j) * chw) + e);
    reinterpret_cast<float4*>(latents + static_cast<size_t>(dst) * chw)[e] = val;
  }
}

}  // namespace
}  // namespace cap4d

using namespace cap4d;

extern "C" {

int cap4d_b200_sampler_gather(const cap4d_b200_sampler_stores* stores, const int64_t* ref_idx, const int64_t* gen_idx,
                              const cap4d_b200_sampler_call* call, int n_groups, int V, int R, int C, int H, int W,
                              int Cc, float* x_in, float* z_in, float* mask_in, float* pos_in, int64_t* t_in,
                              void* stream) {
  if (stores == nullptr || gen_idx == nullptr || call == nullptr || x_in == nullptr || z_in == nullptr ||
      mask_in == nullptr || pos_in == nullptr || t_in == nullptr || (R > 0 && ref_idx == nullptr)) {
    set_error("sampler_gather: null argument");
    return 1;
  }
  if (n_groups < 1 || n_groups > CAP4D_B200_MAX_GROUPS_PER_CALL || R < 0 || R >= V) {
    set_error("sampler_gather: need 1 <= n_groups <= CAP4D_B200_MAX_GROUPS_PER_CALL and 0 <= R < V");
    return 1;
  }
  const int hw = H * W, chw = C * hw, pos_elems = hw * Cc;
  if (hw % 4 != 0 || pos_elems % 4 != 0) {
    set_error("sampler_gather: H*W and H*W*Cc must be multiples of 4");
    return 1;
  }
  if (stores->latents == nullptr || stores->gen_mask == nullptr || stores->gen_pos == nullptr ||
      (R > 0 && (stores->ref_z == nullptr || stores->ref_mask == nullptr || stores->ref_pos == nullptr))) {
    set_error("sampler_gather: the conditional stores and the latent store are required");
    return 1;
  }
  const int total = (2 * chw + hw + pos_elems) / 4;
  dim3 grid((total + 255) / 256, 2 * n_groups * V);
  if (grid.x > 64) grid.x = 64;
  sampler_gather_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      *stores, reinterpret_cast<const long long*>(ref_idx), reinterpret_cast<const long long*>(gen_idx), call, n_groups,
      V, R, chw, hw, pos_elems, x_in, z_in, mask_in, pos_in, reinterpret_cast<long long*>(t_in));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error(cudaGetErrorString(e));
    return 7;
  }
  return 0;
}

int cap4d_b200_sampler_update(float* latents, const float* eps, const int64_t* gen_idx,
                              const cap4d_b200_sampler_call* call, int n_groups, int V, int R, int chw, float cfg_scale,
                              void* stream) {
  if (latents == nullptr || eps == nullptr || gen_idx == nullptr || call == nullptr || chw % 4 != 0 || n_groups < 1 ||
      n_groups > CAP4D_B200_MAX_GROUPS_PER_CALL) {
    set_error("sampler_update: null argument, latent size not a multiple of 4 or too many groups");
    return 1;
  }
  const size_t total = static_cast<size_t>(n_groups) * (V - R) * (chw / 4);
  sampler_update_kernel<<<plane_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      latents, eps, reinterpret_cast<const long long*>(gen_idx), call, n_groups, V, R, chw, cfg_scale);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error(cudaGetErrorString(e));
    return 7;
  }
  return 0;
}

int cap4d_b200_sampler_pack(const float* latents, const int64_t* gen_idx, int n_its, int G, int chw, int rank,
                            int world, float* send, void* stream) {
  if (latents == nullptr || gen_idx == nullptr || send == nullptr || chw % 4 != 0 || world < 1 || rank < 0 ||
      rank >= world) {
    set_error("sampler_pack: bad argument");
    return 1;
  }
  const int mine = (n_its - rank + world - 1) / world;
  if (mine <= 0) return 0;
  const size_t total = static_cast<size_t>(mine) * G * (chw / 4);
  sampler_pack_kernel<<<plane_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      latents, reinterpret_cast<const long long*>(gen_idx), n_its, G, chw, rank, world, send);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error(cudaGetErrorString(e));
    return 7;
  }
  return 0;
}

int cap4d_b200_sampler_unpack(float* latents, const float* recv, const int64_t* gen_idx, int n_its, int G, int chw,
                              int rank, int world, void* stream) {
  if (latents == nullptr || recv == nullptr || gen_idx == nullptr || chw % 4 != 0 || world < 1 || rank < 0 ||
      rank >= world) {
    set_error("sampler_unpack: bad argument");
    return 1;
  }
  const int per_rank = (n_its + world - 1) / world;
  const size_t total = static_cast<size_t>(n_its) * G * (chw / 4);
  sampler_unpack_kernel<<<plane_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      latents, recv, reinterpret_cast<const long long*>(gen_idx), n_its, G, chw, rank, world, per_rank);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error(cudaGetErrorString(e));
    return 7;
  }
  return 0;
}

}  // extern "C"

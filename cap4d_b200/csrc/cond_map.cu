// Conditioning-map generation: the step on the INPUT side of the hot path.
//
// Replaces, per view, CAP4DConditioning.forward(unconditional=False)
// (cap4d/mmdm/conditioning/cap4dcond.py:91-133) including PropRenderer.render
// (cap4d/mmdm/conditioning/mesh2img.py:334-379), i.e. pytorch3d 0.7.8 `rasterize_meshes`
// (cameras=None: blur_radius 0, faces_per_pixel 1, perspective_correct False, clip_barycentric_coords True,
// cull_backfaces False, no z clipping) + `interpolate_face_attributes`, in ONE kernel:
//
//   rasterise the view's mesh at (S*sr)^2 -> barycentric interpolation of the template positions and of the
//   expression offsets -> sinusoidal encoding -> render mask -> sr x sr area average -> concat ray map,
//   reference mask, crop mask -> pos_enc[n][S][S][Ctot]   (the tensor MMDMUnetModel.forward consumes).
//
// The super-resolved images (6 + 45 channels at (S*sr)^2 per view) never reach HBM; algorithmic bytes per view are
// the vertex arrays in and S*S*Ctot*4 out.  This is byte/compare work, not a GEMM: one CTA owns an 8x8 tile of
// output pixels (one thread per super-resolved sample), culls the faces against the tile into a shared-memory list
// in passes of LIST_CAP faces, and every sample walks the list.  Compiled with -fmad=false: coverage decisions
// (pix_to_face) are made with exactly the fp32 operations of the CPU restatement.
//
// Also here: load_camera_rays (cap4d/datasets/utils.py:161-186) + the rotation into the reference camera frame
// (cap4d/inference/data/inference_data.py:89-100), evaluated in fp64 like the numpy original.
#include <math.h>

#include "../../include/cap4d_b200.h"
#include "kernels.h"

namespace cap4d {
namespace {

constexpr float kEpsilon = 1e-8f;  // pytorch3d rasterization_utils.cuh
constexpr int TILE = 8;            // output pixels per tile side
constexpr int LIST_CAP = 768;      // faces culled per pass (bounds the shared-memory list)

struct FaceRec {  // one culled face: its three vertices (NDC x, y; depth z) and its index
  float x0, y0, z0, x1, y1, z1, x2, y2, z2;
  int f;
};

struct CondParams {
  const float* verts;          // [n][Nv][3]
  const float* offsets;        // [n][Nv][3] or null
  const int* faces;            // [F][3]
  const float* props;          // [Nv][3]
  const unsigned char* fmask;  // [F]
  const float* ray_map;        // [n][3][S][S] or null
  const float* ref_mask;       // [n][S][S]
  const float* crop_mask;      // [n][S][S] or null
  float* out;                  // [n][S][S][Ctot]
  int* pix_to_face;            // [n][S*sr][S*sr] or null (debug / parity output)
  int n, Nv, F, S, sr, n_freq, Ctot;
  float pos_mult, std_expr;
};

// EdgeFunctionForward(p, v0, v1) of pytorch3d (geometry_utils.cuh)
__device__ __forceinline__ float edge_fn(float px, float py, float ax, float ay, float bx, float by) {
  return (px - ax) * (by - ay) - (py - ay) * (bx - ax);
}

// PixToNonSquareNdc for a square image: pixel centre i of S -> NDC
__device__ __forceinline__ float pix_to_ndc(int i, int S) { return -1.0f + (2.0f * i + 1.0f) / S; }

__global__ void __launch_bounds__(TILE* TILE * 16) cond_pos_enc_kernel(CondParams p) {
  __shared__ __align__(16) unsigned char smem_raw[LIST_CAP * sizeof(FaceRec)];
  __shared__ float4 list_bbox[LIST_CAP];  // xmin, xmax, ymin, ymax of the culled faces: all a rejected sample reads
  __shared__ int list_n;
  FaceRec* list = reinterpret_cast<FaceRec*>(smem_raw);
  float* tile_out = reinterpret_cast<float*>(smem_raw);  // reused after the face loop: [TILE*TILE][Ctot]

  const int sr = p.sr, sr2 = sr * sr;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int tiles_x = (p.S + TILE - 1) / TILE;
  const int tile_y = blockIdx.x / tiles_x, tile_x = blockIdx.x % tiles_x;
  const int view = blockIdx.y;
  const int pix = tid / sr2, sub = tid % sr2;
  const int oy = tile_y * TILE + pix / TILE, ox = tile_x * TILE + pix % TILE;  // output pixel
  const int SS = p.S * sr;
  const int yi = oy * sr + sub / sr, xi = ox * sr + sub % sr;  // super-resolved sample
  const bool in_img = oy < p.S && ox < p.S;
  // pytorch3d: +X points left and +Y up, so row/column i samples NDC(S - 1 - i)
  const float px = pix_to_ndc(SS - 1 - xi, SS), py = pix_to_ndc(SS - 1 - yi, SS);
  // NDC extent of the tile's sample centres (conservative cull)
  const int ty0 = tile_y * TILE * sr, tx0 = tile_x * TILE * sr;
  const int ty1 = min(ty0 + TILE * sr, SS) - 1, tx1 = min(tx0 + TILE * sr, SS) - 1;
  const float t_xlo = pix_to_ndc(SS - 1 - tx1, SS), t_xhi = pix_to_ndc(SS - 1 - tx0, SS);
  const float t_ylo = pix_to_ndc(SS - 1 - ty1, SS), t_yhi = pix_to_ndc(SS - 1 - ty0, SS);

  const float* V = p.verts + static_cast<size_t>(view) * p.Nv * 3;

  float best_z = INFINITY, b0 = 0.f, b1 = 0.f, b2 = 0.f;
  int best_f = -1;

  for (int c0 = 0; c0 < p.F; c0 += LIST_CAP) {
    if (tid == 0) list_n = 0;
    __syncthreads();
    const int c1 = min(c0 + LIST_CAP, p.F);
    for (int f = c0 + tid; f < c1; f += nthr) {
      const int i0 = __ldg(p.faces + 3 * f), i1 = __ldg(p.faces + 3 * f + 1), i2 = __ldg(p.faces + 3 * f + 2);
      FaceRec r;
      r.x0 = __ldg(V + 3 * i0); r.y0 = __ldg(V + 3 * i0 + 1); r.z0 = __ldg(V + 3 * i0 + 2);
      r.x1 = __ldg(V + 3 * i1); r.y1 = __ldg(V + 3 * i1 + 1); r.z1 = __ldg(V + 3 * i1 + 2);
      r.x2 = __ldg(V + 3 * i2); r.y2 = __ldg(V + 3 * i2 + 1); r.z2 = __ldg(V + 3 * i2 + 2);
      r.f = f;
      const float xmin = fminf(r.x0, fminf(r.x1, r.x2)), xmax = fmaxf(r.x0, fmaxf(r.x1, r.x2));
      const float ymin = fminf(r.y0, fminf(r.y1, r.y2)), ymax = fmaxf(r.y0, fmaxf(r.y1, r.y2));
      const float zmin = fminf(r.z0, fminf(r.z1, r.z2));
      // per-face rejections of CheckPixelInsideFace / CheckPointOutsideBoundingBox (rasterize_meshes.cu)
      const float area = edge_fn(r.x0, r.y0, r.x1, r.y1, r.x2, r.y2);  // EdgeFunctionForward(v0, v1, v2)
      const bool zero_area = (area <= kEpsilon && area >= -kEpsilon);
      const bool z_invalid = zmin < kEpsilon;
      const bool off_tile = (t_xlo > xmax || t_xhi < xmin || t_ylo > ymax || t_yhi < ymin);
      if (!(zero_area || z_invalid || off_tile)) {
        const int at = atomicAdd(&list_n, 1);
        list[at] = r;
        list_bbox[at] = make_float4(xmin, xmax, ymin, ymax);
      }
    }
    __syncthreads();
    const int ln = list_n;
    if (in_img) {
      for (int j = 0; j < ln; ++j) {
        const float4 bb = list_bbox[j];  // broadcast read
        if (px > bb.y || px < bb.x || py > bb.w || py < bb.z) continue;
        const FaceRec r = list[j];
        // BarycentricCoordsForward
        const float area = edge_fn(r.x2, r.y2, r.x0, r.y0, r.x1, r.y1) + kEpsilon;
        const float w0 = edge_fn(px, py, r.x1, r.y1, r.x2, r.y2) / area;
        const float w1 = edge_fn(px, py, r.x2, r.y2, r.x0, r.y0) / area;
        const float w2 = edge_fn(px, py, r.x0, r.y0, r.x1, r.y1) / area;
        if (!(w0 > 0.0f && w1 > 0.0f && w2 > 0.0f)) continue;  // blur_radius 0: only strictly inside samples
        // BarycentricClipForward
        float c_0 = fmaxf(w0, 0.0f), c_1 = fmaxf(w1, 0.0f), c_2 = fmaxf(w2, 0.0f);
        const float wsum = fmaxf(c_0 + c_1 + c_2, 1e-5f);
        c_0 /= wsum; c_1 /= wsum; c_2 /= wsum;
        const float pz = c_0 * r.z0 + c_1 * r.z1 + c_2 * r.z2;
        if (pz < 0.0f) continue;
        if (pz < best_z || (pz == best_z && r.f < best_f)) {  // nearest face; ties -> smaller face index
          best_z = pz; best_f = r.f; b0 = c_0; b1 = c_1; b2 = c_2;
        }
      }
    }
    __syncthreads();
  }

  if (p.pix_to_face != nullptr && in_img)
    p.pix_to_face[(static_cast<size_t>(view) * SS + yi) * SS + xi] = best_f;

  // ---- shade: interpolate_face_attributes -> positional encoding -> mask ------------------------
  float prop[3] = {0.f, 0.f, 0.f}, offs[3] = {0.f, 0.f, 0.f};
  float mask = 0.f;
  if (best_f >= 0) {
    const int i0 = __ldg(p.faces + 3 * best_f), i1 = __ldg(p.faces + 3 * best_f + 1), i2 = __ldg(p.faces + 3 * best_f + 2);
#pragma unroll
    for (int c = 0; c < 3; ++c)
      prop[c] = b0 * __ldg(p.props + 3 * i0 + c) + b1 * __ldg(p.props + 3 * i1 + c) + b2 * __ldg(p.props + 3 * i2 + c);
    if (p.offsets != nullptr) {
      const float* O = p.offsets + static_cast<size_t>(view) * p.Nv * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c)
        offs[c] = b0 * (__ldg(O + 3 * i0 + c) / p.std_expr) + b1 * (__ldg(O + 3 * i1 + c) / p.std_expr) +
                  b2 * (__ldg(O + 3 * i2 + c) / p.std_expr);
    }
    mask = p.fmask[best_f] ? 1.0f : 0.0f;
  }

  // area average over the sr x sr samples of an output pixel: the samples are consecutive lanes; every lane adds
  // them in raster order (the order of adaptive_avg_pool2d), lane `sub == 0` keeps the result
  const unsigned lane = tid & 31u;
  const unsigned base = lane - sub;
  const float inv_cnt = 1.0f / static_cast<float>(sr2);  // sr2 is 1, 4 or 16: multiplying is exact division
  auto pool = [&](float v) {
    float s = 0.f;
    for (int j = 0; j < sr2; ++j) s += __shfl_sync(0xffffffffu, v, base + j);
    return s * inv_cnt;
  };
  float* my_out = tile_out + pix * p.Ctot;
  const int nf = p.n_freq;
  for (int c = 0; c < 3; ++c) {
    const float v = prop[c] * p.pos_mult;
    float freq = 1.0f;
    for (int k = 0; k < nf; ++k) {
      const float a = v * freq;
      float sv = 0.f, cv = 0.f;
      if (mask != 0.f) {  // two thirds of the samples are background or masked: their contribution is 0
        sv = sinf(a);
        cv = cosf(a);
      }
      const float s = pool(sv), co = pool(cv);
      if (sub == 0) {
        my_out[c * 2 * nf + k] = s;
        my_out[c * 2 * nf + nf + k] = co;
      }
      freq *= 2.0f;
    }
  }
  int ch = 6 * nf;
  if (p.offsets != nullptr) {
    for (int c = 0; c < 3; ++c) {
      const float o = pool(offs[c] * mask);
      if (sub == 0) my_out[ch + c] = o;
    }
    ch += 3;
  }
  if (sub == 0 && in_img) {
    const size_t plane = static_cast<size_t>(p.S) * p.S;
    const size_t at = static_cast<size_t>(oy) * p.S + ox;
    if (p.ray_map != nullptr) {
      for (int c = 0; c < 3; ++c) my_out[ch + c] = __ldg(p.ray_map + (static_cast<size_t>(view) * 3 + c) * plane + at);
      ch += 3;
    }
    my_out[ch++] = __ldg(p.ref_mask + static_cast<size_t>(view) * plane + at);
    if (p.crop_mask != nullptr) my_out[ch++] = __ldg(p.crop_mask + static_cast<size_t>(view) * plane + at);
  }
  __syncthreads();
  // coalesced write of the tile: each tile row is (valid columns) * Ctot consecutive floats of the output
  const int row_len = TILE * p.Ctot;
  const int valid = min(TILE, p.S - tile_x * TILE) * p.Ctot;
  const int rows = min(TILE, p.S - tile_y * TILE);
  const bool vec = (valid == row_len) && ((p.S * p.Ctot) % 4 == 0);  // rows start on 16-byte boundaries
  for (int ry = 0; ry < rows; ++ry) {
    float* dst = p.out + ((static_cast<size_t>(view) * p.S + tile_y * TILE + ry) * p.S + tile_x * TILE) * p.Ctot;
    const float* src = tile_out + ry * row_len;
    if (vec) {
      for (int e = tid; e < row_len / 4; e += nthr)
        reinterpret_cast<float4*>(dst)[e] = reinterpret_cast<const float4*>(src)[e];
    } else {
      for (int e = tid; e < valid; e += nthr) dst[e] = src[e];
    }
  }
}

// cam[n][22] (fp64): new_fx, new_fy, new_cx, new_cy, inv(extr[:3,:3]) row-major (9), ref_extr[:3,:3] (9)
__global__ void ray_map_kernel(const double* __restrict__ cam, float* __restrict__ out, int n, int S) {
  const size_t total = static_cast<size_t>(n) * S * S;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int view = static_cast<int>(i / (static_cast<size_t>(S) * S));
  const int r = static_cast<int>(i % (static_cast<size_t>(S) * S));
  const int v = r / S, u = r % S;
  const double* c = cam + static_cast<size_t>(view) * 22;
  double d0 = (static_cast<double>(u) - c[2]) / c[0];
  double d1 = (static_cast<double>(v) - c[3]) / c[1];
  double d2 = 1.0;
  const double nrm = sqrt(d0 * d0 + d1 * d1 + d2 * d2) + 1e-8;
  d0 /= nrm; d1 /= nrm; d2 /= nrm;
  const double* A = c + 4;
  const double e0 = A[0] * d0 + A[1] * d1 + A[2] * d2;
  const double e1 = A[3] * d0 + A[4] * d1 + A[5] * d2;
  const double e2 = A[6] * d0 + A[7] * d1 + A[8] * d2;
  const double* R = c + 13;
  const size_t plane = static_cast<size_t>(S) * S;
  float* o = out + static_cast<size_t>(view) * 3 * plane + r;
  o[0] = static_cast<float>(R[0] * e0 + R[1] * e1 + R[2] * e2);
  o[plane] = static_cast<float>(R[3] * e0 + R[4] * e1 + R[5] * e2);
  o[2 * plane] = static_cast<float>(R[6] * e0 + R[7] * e1 + R[8] * e2);
}

}  // namespace
}  // namespace cap4d

using namespace cap4d;

extern "C" {

int cap4d_b200_cond_pos_enc(const float* verts_2d, const float* offsets_3d, const int32_t* faces, const float* props,
                            const uint8_t* face_mask, const float* ray_map, const float* ref_mask,
                            const float* crop_mask, float* pos_enc, int32_t* pix_to_face, int n_views, int n_verts,
                            int n_faces, int image_size, int super_resolution, int positional_channels,
                            float positional_multiplier, float std_expr_deformation, void* stream) {
  if (n_views == 0) return 0;  // an empty batch has nothing to read or write
  if (verts_2d == nullptr || faces == nullptr || props == nullptr || face_mask == nullptr || ref_mask == nullptr ||
      pos_enc == nullptr) {
    set_error("cond_pos_enc: verts_2d, faces, props, face_mask, ref_mask and pos_enc are required");
    return 20;
  }
  if (n_views < 0 || n_verts <= 0 || n_faces <= 0 || image_size <= 0) {
    set_error("cond_pos_enc: bad sizes");
    return 20;
  }
  if (super_resolution != 1 && super_resolution != 2 && super_resolution != 4) {
    set_error("cond_pos_enc: super_resolution must be 1, 2 or 4");
    return 20;
  }
  if (positional_channels <= 0 || positional_channels % 6 != 0) {
    // cap4dcond.py:63 (channels % 3 == 0) and :16 (channels_per_dim % 2 == 0)
    set_error("cond_pos_enc: positional_channels must be a positive multiple of 6");
    return 20;
  }
  CondParams p;
  p.verts = verts_2d; p.offsets = offsets_3d; p.faces = faces; p.props = props; p.fmask = face_mask;
  p.ray_map = ray_map; p.ref_mask = ref_mask; p.crop_mask = crop_mask; p.out = pos_enc; p.pix_to_face = pix_to_face;
  p.n = n_views; p.Nv = n_verts; p.F = n_faces; p.S = image_size; p.sr = super_resolution;
  p.n_freq = positional_channels / 6;
  p.Ctot = positional_channels + (offsets_3d ? 3 : 0) + (ray_map ? 3 : 0) + 1 + (crop_mask ? 1 : 0);
  p.pos_mult = positional_multiplier; p.std_expr = std_expr_deformation;
  if (static_cast<size_t>(TILE) * TILE * p.Ctot * sizeof(float) > LIST_CAP * sizeof(FaceRec)) {
    set_error("cond_pos_enc: too many channels for the tile buffer");
    return 20;
  }
  const int tiles = (image_size + TILE - 1) / TILE;
  for (int v0 = 0; v0 < n_views; v0 += 65535) {  // gridDim.y limit
    CondParams q = p;
    const int nv = (n_views - v0 < 65535) ? (n_views - v0) : 65535;
    const size_t plane = static_cast<size_t>(image_size) * image_size;
    q.verts += static_cast<size_t>(v0) * n_verts * 3;
    if (q.offsets) q.offsets += static_cast<size_t>(v0) * n_verts * 3;
    if (q.ray_map) q.ray_map += static_cast<size_t>(v0) * 3 * plane;
    q.ref_mask += static_cast<size_t>(v0) * plane;
    if (q.crop_mask) q.crop_mask += static_cast<size_t>(v0) * plane;
    q.out += static_cast<size_t>(v0) * plane * p.Ctot;
    if (q.pix_to_face) q.pix_to_face += static_cast<size_t>(v0) * plane * super_resolution * super_resolution;
    q.n = nv;
    dim3 grid(tiles * tiles, nv);
    cond_pos_enc_kernel<<<grid, TILE * TILE * super_resolution * super_resolution, 0,
                          static_cast<cudaStream_t>(stream)>>>(q);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error(std::string("cond_pos_enc: ") + cudaGetErrorString(e));
    return 21;
  }
  return 0;
}

int cap4d_b200_cond_ray_map(const double* cam, float* ray_map, int n_views, int image_size, void* stream) {
  if (cam == nullptr || ray_map == nullptr || n_views < 0 || image_size <= 0) {
    set_error("cond_ray_map: bad arguments");
    return 20;
  }
  if (n_views == 0) return 0;
  const size_t total = static_cast<size_t>(n_views) * image_size * image_size;
  ray_map_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      cam, ray_map, n_views, image_size);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error(std::string("cond_ray_map: ") + cudaGetErrorString(e));
    return 21;
  }
  return 0;
}

}  // extern "C"

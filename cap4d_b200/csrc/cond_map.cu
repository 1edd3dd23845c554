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
// output pixels; its threads first take FACES (a FLAME triangle covers about one sample at this resolution) and
// scatter (depth, face) keys into a shared-memory z-buffer with atomicMin, then take SAMPLES (one thread per
// super-resolved sample) for shading and the area average.  Compiled with -fmad=false: coverage decisions
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
constexpr int MAX_CH = 64;         // output channels the shared-memory tile buffer holds
constexpr int BIG_CAP = 256;       // faces per tile evaluated sample-parallel (the rest: by the thread that found them)
constexpr int SCAN_U = 4;          // loads of face tile ranges in flight per thread during the scan
constexpr int BIG_SAMPLES = 48;    // a face whose box spans more samples of the tile than this is "big"

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
  unsigned* bins;              // [n][F] workspace: the tile range every face can touch (face_bins_kernel)
  int n, Nv, F, S, sr, n_freq, Ctot;
  float pos_mult, std_expr;
};

// EdgeFunctionForward(p, v0, v1) of pytorch3d (geometry_utils.cuh)
__device__ __forceinline__ float edge_fn(float px, float py, float ax, float ay, float bx, float by) {
  return (px - ax) * (by - ay) - (py - ay) * (bx - ax);
}

// PixToNonSquareNdc for a square image: pixel centre i of S -> NDC
__device__ __forceinline__ float pix_to_ndc(int i, int S) { return -1.0f + (2.0f * i + 1.0f) / S; }

// One face against one sample centre (CheckPixelInsideFace for blur_radius 0): true when the sample is strictly
// inside; pz and the clipped barycentrics are what the reference's priority queue would hold.
__device__ __forceinline__ bool eval_face(const FaceRec& r, float px, float py, float& pz, float& c_0, float& c_1,
                                          float& c_2) {
  // BarycentricCoordsForward
  const float area = edge_fn(r.x2, r.y2, r.x0, r.y0, r.x1, r.y1) + kEpsilon;
  const float e0 = edge_fn(px, py, r.x1, r.y1, r.x2, r.y2);
  const float e1 = edge_fn(px, py, r.x2, r.y2, r.x0, r.y0);
  const float e2 = edge_fn(px, py, r.x0, r.y0, r.x1, r.y1);
  // e / area > 0 needs e and area of the same sign (area == 0 can only be +0): most candidates leave here, before
  // the six divisions.  (The quotient itself is still what decides: only |e / area| < 2^-149 could differ.)
  if (area >= 0.0f ? !(e0 > 0.0f && e1 > 0.0f && e2 > 0.0f) : !(e0 < 0.0f && e1 < 0.0f && e2 < 0.0f)) return false;
  const float w0 = e0 / area, w1 = e1 / area, w2 = e2 / area;
  if (!(w0 > 0.0f && w1 > 0.0f && w2 > 0.0f)) return false;  // blur_radius 0: only strictly inside samples
  // BarycentricClipForward
  c_0 = fmaxf(w0, 0.0f); c_1 = fmaxf(w1, 0.0f); c_2 = fmaxf(w2, 0.0f);
  const float wsum = fmaxf(c_0 + c_1 + c_2, 1e-5f);
  c_0 /= wsum; c_1 /= wsum; c_2 /= wsum;
  pz = c_0 * r.z0 + c_1 * r.z1 + c_2 * r.z2;
  return !(pz < 0.0f);
}

__device__ __forceinline__ FaceRec load_face(const int* __restrict__ faces, const float* __restrict__ V, int f) {
  const int i0 = __ldg(faces + 3 * f), i1 = __ldg(faces + 3 * f + 1), i2 = __ldg(faces + 3 * f + 2);
  FaceRec r;
  r.x0 = __ldg(V + 3 * i0); r.y0 = __ldg(V + 3 * i0 + 1); r.z0 = __ldg(V + 3 * i0 + 2);
  r.x1 = __ldg(V + 3 * i1); r.y1 = __ldg(V + 3 * i1 + 1); r.z1 = __ldg(V + 3 * i1 + 2);
  r.x2 = __ldg(V + 3 * i2); r.y2 = __ldg(V + 3 * i2 + 1); r.z2 = __ldg(V + 3 * i2 + 2);
  r.f = f;
  return r;
}

// depth key of a hit: pz >= 0, so its bit pattern orders like the value; ties go to the smaller face index
__device__ __forceinline__ unsigned long long depth_key(float pz, int f) {
  return (static_cast<unsigned long long>(__float_as_uint(pz)) << 32) | static_cast<unsigned>(f);
}

constexpr unsigned long long kNoFace = ~0ull;

// Sample rows/columns whose centre can lie inside [lo, hi] (NDC): centre(i) = 1 - (2 i + 1) / SS, padded by one
// sample for the rounding of this index arithmetic; the exact fp32 box test of the reference is applied per sample.
__device__ __forceinline__ void sample_window(float lo, float hi, int SS, int& a, int& b) {
  const float half = 0.5f * static_cast<float>(SS), lim = static_cast<float>(SS) + 2.0f;
  a = static_cast<int>(floorf(fminf(fmaxf((1.0f - hi) * half - 0.5f, -2.0f), lim))) - 1;
  b = static_cast<int>(ceilf(fminf(fmaxf((1.0f - lo) * half - 0.5f, -2.0f), lim))) + 1;
}

// Tile range of a face, one byte per bound so that one subtraction tests all four: tx_lo | (127 - tx_hi) << 8 |
// ty_lo << 16 | (127 - ty_hi) << 24, every byte <= 127.  A tile (tx, ty) is inside iff every byte of
// T = tx | (127 - tx) << 8 | ty << 16 | (127 - ty) << 24 is >= the code's byte, i.e. iff
// ((T | 0x80808080) - code) keeps bit 7 of every byte (no byte can borrow from its neighbour).
constexpr unsigned kBinReject = 0x0000007fu;  // tx_lo = 127 > any tile index (tiles per side <= 127)

// Pre-pass, one thread per (view, face): the per-face rejections of the rasteriser and the range of tiles the
// face's bounding box can touch.  The tile CTAs then scan 4 coalesced bytes per face instead of gathering its
// vertices (64 tiles x 10k faces per view).
__global__ void face_bins_kernel(CondParams p, int tw) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(p.n) * p.F) return;
  const int view = static_cast<int>(i / p.F), f = static_cast<int>(i % p.F);
  const FaceRec r = load_face(p.faces, p.verts + static_cast<size_t>(view) * p.Nv * 3, f);
  const float xmin = fminf(r.x0, fminf(r.x1, r.x2)), xmax = fmaxf(r.x0, fmaxf(r.x1, r.x2));
  const float ymin = fminf(r.y0, fminf(r.y1, r.y2)), ymax = fmaxf(r.y0, fmaxf(r.y1, r.y2));
  const float zmin = fminf(r.z0, fminf(r.z1, r.z2));
  const float area = edge_fn(r.x0, r.y0, r.x1, r.y1, r.x2, r.y2);
  unsigned code = kBinReject;
  if (!((area <= kEpsilon && area >= -kEpsilon) || zmin < kEpsilon || !(xmin <= xmax) || !(ymin <= ymax))) {
    const int SS = p.S * p.sr;
    int xa, xb, ya, yb;
    sample_window(xmin, xmax, SS, xa, xb);
    sample_window(ymin, ymax, SS, ya, yb);
    xa = max(xa, 0); ya = max(ya, 0); xb = min(xb, SS - 1); yb = min(yb, SS - 1);
    if (xa <= xb && ya <= yb)
      code = static_cast<unsigned>(xa / tw) | static_cast<unsigned>(127 - xb / tw) << 8 |
             static_cast<unsigned>(ya / tw) << 16 | static_cast<unsigned>(127 - yb / tw) << 24;
  }
  p.bins[i] = code;
}

template <int SR>
__global__ void __launch_bounds__(TILE* TILE* SR* SR, 16 / (SR * SR)) cond_pos_enc_kernel(CondParams p) {
  // FACE-parallel rasterisation: the faces of a FLAME-sized mesh cover about one sample each, so instead of every
  // sample walking a face list, every thread takes faces, visits the few sample centres inside the face's bounding
  // box and scatters (depth, face) keys with a shared-memory atomicMin.  Faces whose box covers many samples of the
  // tile go to `big` and are evaluated sample-parallel.
  __shared__ unsigned long long zkey[TILE * TILE * SR * SR];
  __shared__ __align__(16) float tile_out[TILE * TILE * MAX_CH];
  __shared__ int big[BIG_CAP];
  __shared__ int big_n;
  __shared__ float ndc_x[TILE * SR], ndc_y[TILE * SR];  // NDC centres of the tile's sample columns / rows

  constexpr int sr = SR, sr2 = SR * SR, nthr = TILE * TILE * SR * SR;
  const int tid = threadIdx.x;
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
  // sample rows/columns of the tile
  const int tw = TILE * sr;
  const int ty0 = tile_y * tw, tx0 = tile_x * tw;
  const int ty1 = min(ty0 + tw, SS) - 1, tx1 = min(tx0 + tw, SS) - 1;
  const int my_slot = (yi - ty0) * tw + (xi - tx0);

  const float* V = p.verts + static_cast<size_t>(view) * p.Nv * 3;

  zkey[tid] = kNoFace;  // blockDim.x == tw * tw
  if (tid == 0) big_n = 0;
  if (tid < tw) {
    ndc_x[tid] = pix_to_ndc(SS - 1 - (tx0 + tid), SS);
    ndc_y[tid] = pix_to_ndc(SS - 1 - (ty0 + tid), SS);
  }
  __syncthreads();

  // scan the view's per-face tile ranges: SCAN_U independent coalesced loads in flight per thread (a plain loop is one
  // L2 round trip per face); consecutive faces stay on consecutive lanes, so the faces of a mesh patch that hit this
  // tile are evaluated side by side
  const unsigned* bins = p.bins + static_cast<size_t>(view) * p.F;
  const unsigned t_or = (static_cast<unsigned>(tile_x) | static_cast<unsigned>(127 - tile_x) << 8 |
                         static_cast<unsigned>(tile_y) << 16 | static_cast<unsigned>(127 - tile_y) << 24) | 0x80808080u;
  for (int base = 0; base < p.F; base += nthr * SCAN_U) {
    unsigned c[SCAN_U];
#pragma unroll
    for (int u = 0; u < SCAN_U; ++u) {
      const int i = base + u * nthr + tid;
      c[u] = (i < p.F) ? __ldg(bins + i) : kBinReject;
    }
    unsigned hits = 0;
#pragma unroll
    for (int u = 0; u < SCAN_U; ++u) hits |= ((((t_or - c[u]) & 0x80808080u) == 0x80808080u) ? 1u : 0u) << u;
    while (hits) {
      const int u = __ffs(hits) - 1;
      hits &= hits - 1;
      const int f = base + u * nthr + tid;
      const FaceRec r = load_face(p.faces, V, f);
      const float xmin = fminf(r.x0, fminf(r.x1, r.x2)), xmax = fmaxf(r.x0, fmaxf(r.x1, r.x2));
      const float ymin = fminf(r.y0, fminf(r.y1, r.y2)), ymax = fmaxf(r.y0, fmaxf(r.y1, r.y2));
      int xa, xb, ya, yb;
      sample_window(xmin, xmax, SS, xa, xb);
      sample_window(ymin, ymax, SS, ya, yb);
      xa = max(xa, tx0) - tx0; xb = min(xb, tx1) - tx0; ya = max(ya, ty0) - ty0; yb = min(yb, ty1) - ty0;
      // columns / rows of the tile whose centre passes the reference's exact fp32 bounding-box test
      unsigned cols = 0, rows = 0;
      for (int x = xa; x <= xb; ++x) cols |= (ndc_x[x] > xmax || ndc_x[x] < xmin) ? 0u : (1u << x);
      for (int y = ya; y <= yb; ++y) rows |= (ndc_y[y] > ymax || ndc_y[y] < ymin) ? 0u : (1u << y);
      if (cols == 0 || rows == 0) continue;
      if (__popc(cols) * __popc(rows) > BIG_SAMPLES) {
        const int at = atomicAdd(&big_n, 1);
        if (at < BIG_CAP) {
          big[at] = f;
          continue;
        }  // list full: fall through and do it here
      }
      for (unsigned rm = rows; rm; rm &= rm - 1) {
        const int y = __ffs(rm) - 1;
        for (unsigned cm = cols; cm; cm &= cm - 1) {
          const int x = __ffs(cm) - 1;
          float pz, c_0, c_1, c_2;
          if (eval_face(r, ndc_x[x], ndc_y[y], pz, c_0, c_1, c_2)) atomicMin(&zkey[y * tw + x], depth_key(pz, f));
        }
      }
    }
  }
  __syncthreads();
  const int nbig = min(big_n, BIG_CAP);
  for (int j = 0; j < nbig; ++j) {  // large faces: every thread tests its own sample
    const FaceRec r = load_face(p.faces, V, big[j]);
    const float xmin = fminf(r.x0, fminf(r.x1, r.x2)), xmax = fmaxf(r.x0, fmaxf(r.x1, r.x2));
    const float ymin = fminf(r.y0, fminf(r.y1, r.y2)), ymax = fmaxf(r.y0, fmaxf(r.y1, r.y2));
    if (!in_img || px > xmax || px < xmin || py > ymax || py < ymin) continue;
    float pz, c_0, c_1, c_2;
    if (eval_face(r, px, py, pz, c_0, c_1, c_2)) atomicMin(&zkey[my_slot], depth_key(pz, r.f));
  }
  __syncthreads();

  // the winner of this thread's sample; its barycentrics are recomputed with the same operations
  float b0 = 0.f, b1 = 0.f, b2 = 0.f;
  int best_f = -1;
  if (in_img) {
    const unsigned long long key = zkey[my_slot];
    if (key != kNoFace) {
      best_f = static_cast<int>(key & 0xffffffffull);
      float pz;
      eval_face(load_face(p.faces, V, best_f), px, py, pz, b0, b1, b2);
    }
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
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = prop[c] * p.pos_mult;
    float freq = 1.0f;
    for (int k = 0; k < nf; ++k) {
      const float a = v * freq;
      float sv = 0.f, cv = 0.f;
      if (mask != 0.f) sincosf(a, &sv, &cv);  // two thirds of the samples are background or masked: they add 0
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
#pragma unroll
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
                            float positional_multiplier, float std_expr_deformation, void* workspace,
                            size_t workspace_bytes, void* stream) {
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
  const size_t f_pad = static_cast<size_t>(n_faces);
  if (workspace == nullptr || (reinterpret_cast<uintptr_t>(workspace) & 3) != 0 ||
      workspace_bytes < static_cast<size_t>(n_views) * f_pad * sizeof(unsigned)) {
    set_error("cond_pos_enc: workspace missing, misaligned or too small (cap4d_b200_cond_workspace_bytes)");
    return 20;
  }
  if ((image_size + TILE - 1) / TILE > 127) {
    set_error("cond_pos_enc: image_size above 1016 is not supported");
    return 20;
  }
  p.bins = static_cast<unsigned*>(workspace);
  if (p.Ctot > MAX_CH) {
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
    q.bins += static_cast<size_t>(v0) * f_pad;
    q.n = nv;
    dim3 grid(tiles * tiles, nv);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t n_bins = static_cast<size_t>(nv) * f_pad;
    face_bins_kernel<<<static_cast<unsigned>((n_bins + 255) / 256), 256, 0, st>>>(q, TILE * super_resolution);
    if (super_resolution == 1) cond_pos_enc_kernel<1><<<grid, TILE * TILE, 0, st>>>(q);
    else if (super_resolution == 2) cond_pos_enc_kernel<2><<<grid, TILE * TILE * 4, 0, st>>>(q);
    else cond_pos_enc_kernel<4><<<grid, TILE * TILE * 16, 0, st>>>(q);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error(std::string("cond_pos_enc: ") + cudaGetErrorString(e));
    return 21;
  }
  return 0;
}

int cap4d_b200_cond_workspace_bytes(int n_views, int n_faces, size_t* bytes) {
  if (bytes == nullptr || n_views < 0 || n_faces < 0) {
    set_error("cond_workspace_bytes: bad arguments");
    return 20;
  }
  *bytes = static_cast<size_t>(n_views) * static_cast<size_t>(n_faces) * sizeof(unsigned);
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

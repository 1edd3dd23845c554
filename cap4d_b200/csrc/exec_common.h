// Pieces shared by the executors (unet_exec.cu, vae_exec.cu): launch-plan ops, the workspace arena,
// raw weight tensors.  Internal linkage (one copy per translation unit).
#pragma once
#include <algorithm>
#include <cstdint>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "kernels.h"

namespace cap4d {

#define CUDA_OK(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                       \
      return false;                                                                        \
    }                                                                                      \
  } while (0)

namespace {

enum OpClass { CLS_CONV = 0, CLS_LINEAR = 1, CLS_ATTN = 2, CLS_GN = 3, CLS_LN = 4, CLS_OTHER = 5 };

struct Op {
  int cls;
  int launches;
  double flops;       // algorithmic (the reference op's FLOPs)
  double exec_flops = -1;  // executed by the tensor core when that differs (folded upsample conv); -1: same as flops
  double bytes;
  std::function<cudaError_t(cudaStream_t)> run;
};

struct RawTensor {
  float* d = nullptr;
  size_t numel = 0;
  std::vector<int64_t> shape;
};

// small device kernels used only at weight-pack time
__global__ void vec_add_kernel(const float* a, const float* b, float* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + (b != nullptr ? b[i] : 0.f);
}
// first-fit offset allocator over the workspace; identical decisions in the sizing and the real pass
struct Arena {
  std::map<size_t, size_t> free_;  // offset -> size
  size_t top = 0, peak = 0;
  static size_t align(size_t b) { return (b + 1023) & ~static_cast<size_t>(1023); }
  size_t alloc(size_t bytes) {
    bytes = align(bytes ? bytes : 1);
    for (auto it = free_.begin(); it != free_.end(); ++it) {
      if (it->second >= bytes) {
        size_t off = it->first, rem = it->second - bytes;
        free_.erase(it);
        if (rem) free_[off + bytes] = rem;
        return off;
      }
    }
    size_t off = top;
    top += bytes;
    peak = std::max(peak, top);
    return off;
  }
  void release(size_t off, size_t bytes) {
    bytes = align(bytes ? bytes : 1);
    auto it = free_.emplace(off, bytes).first;
    auto nx = std::next(it);
    if (nx != free_.end() && it->first + it->second == nx->first) {
      it->second += nx->second;
      free_.erase(nx);
    }
    if (it != free_.begin()) {
      auto pv = std::prev(it);
      if (pv->first + pv->second == it->first) {
        pv->second += it->second;
        free_.erase(it);
        it = pv;
      }
    }
    if (it->first + it->second == top) {
      top = it->first;
      free_.erase(it);
    }
  }
};

struct Buf {  // a workspace tensor
  size_t off = 0, bytes = 0;
  int M = 0, C = 0;
  bool valid = false;
};

}  // namespace

}  // namespace cap4d

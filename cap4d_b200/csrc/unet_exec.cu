// MMDM U-Net executor: topology, weight repacking, static launch plan.
//
// Mirrors MMDMUnetModel (cap4d/mmdm/net/mmdm_unet.py:14-126) on top of UNetModel's block topology
// (controlnet/ldm/modules/diffusionmodules/openaimodel.py:544-774) and SpatioTemporalTransformer
// (cap4d/mmdm/net/attention.py:330-387).  Activations are NHWC (token-major) fp32 on the residual
// stream and bf16 as MMA operands; a forward is a fixed list of kernel launches on one stream with
// every intermediate carved out of a caller-provided workspace (no allocation after planning).
#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include <cuda_fp16.h>

#include "../../include/cap4d_b200.h"
#include "exec_common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace cap4d {

// ---------------------------------------------------------------------------------------------
// error text
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
const char* get_error() { return g_error.c_str(); }

namespace {

// GEGLU interleave (attention.py:68-75): rows [0,inner) = x, [inner,2*inner) = gate ->
// packed row (r/32)*64 + half*32 + r%32
__global__ void pack_geglu_kernel(const float* w, const float* b, int inner, int K, bf16* wout, float* bout, int f16) {
  const size_t total = static_cast<size_t>(2) * inner * K;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t row = i / K, c = i % K;
    const int half = row >= static_cast<size_t>(inner);
    const size_t r = row - static_cast<size_t>(half) * inner;
    const size_t prow = (r / 32) * 64 + half * 32 + (r % 32);
    if (f16) {
      const __half hv = __float2half_rn(w[i]);
      wout[prow * K + c] = *reinterpret_cast<const bf16*>(&hv);
    } else {
      wout[prow * K + c] = __float2bfloat16(w[i]);
    }
    if (c == 0) bout[prow] = b[row];
  }
}

// max |x| into *out (non-negative floats order like their bit patterns)
__global__ void absmax_kernel(const float* x, size_t n, float* out) {
  float m = 0.f;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float v = fabsf(x[i]);
    m = (v > m || v != v) ? v : m;  // NaN propagates
  }
  if (m != m) m = __int_as_float(0x7f800000);
  atomicMax(reinterpret_cast<unsigned int*>(out), __float_as_uint(m));
}

struct ResW {
  std::string prefix;
  int cin = 0, cout = 0;
  bool skip = false;
  float *gn1_g = nullptr, *gn1_b = nullptr, *gn2_g = nullptr, *gn2_b = nullptr;
  bf16 *w1 = nullptr, *w2 = nullptr;
  float *b1 = nullptr, *b2 = nullptr;
  int emb_off = 0;
};
struct TfW {
  std::string prefix;
  int C = 0;
  bool is3d = false;
  float *gn_g = nullptr, *gn_b = nullptr, *ln1_g = nullptr, *ln1_b = nullptr, *ln3_g = nullptr, *ln3_b = nullptr;
  bf16 *wpi = nullptr, *wqkv = nullptr, *wo = nullptr, *wff1 = nullptr, *wff2 = nullptr, *wpo = nullptr;
  float *bpi = nullptr, *bo = nullptr, *bff1 = nullptr, *bff2 = nullptr, *bpo = nullptr;
};
struct ConvW {
  std::string prefix;
  int cin = 0, cout = 0;
  bf16* w = nullptr;
  float* b = nullptr;
};

enum LayerKind { L_RES, L_TF, L_DOWN, L_UP };
struct Layer {
  LayerKind kind;
  int idx;
};
typedef std::vector<Layer> Block;

struct IoPtrs {
  const float *x = nullptr, *z = nullptr, *mask = nullptr, *pos = nullptr;
  const long long* t = nullptr;
  float* out = nullptr;
};

struct Unet {
  cap4d_b200_unet_config cfg;
  int emb_ch = 0, kpad_in = 0, n_all = 0;
  std::vector<ResW> res;
  std::vector<TfW> tf;
  std::vector<ConvW> down, up;
  std::vector<Block> input_blocks, output_blocks;  // input_blocks[0] is conv_in (handled separately)
  Block middle;
  std::vector<int> input_block_chans;
  std::map<std::string, RawTensor> raw;
  std::vector<void*> owned;  // packed device buffers
  bool finalized = false;
  // packed globals
  bf16* w_in = nullptr;
  float* b_in = nullptr;
  float *te_w1 = nullptr, *te_b1 = nullptr, *te_w2 = nullptr, *te_b2 = nullptr, *wall = nullptr, *ball = nullptr;
  float *out_gn_g = nullptr, *out_gn_b = nullptr;
  bf16* w_out = nullptr;
  float* b_out = nullptr;
  // plan
  std::vector<Op> ops;
  IoPtrs io;
  int pB = 0, pV = 0, pH = 0, pW = 0, pR = 0;
  // Caller's promise (cap4d_b200_unet_set_ref_views): the first ref_views views of every group are reference
  // views (ref_mask == 1), so their outputs are x - z_input whatever the network computes for them.
  int ref_views = 0;
  int* d_violations = nullptr;  // broken n_ref_views promises seen by the output mix (device counter)
  // Debug taps (cap4d_b200_unet_enable_taps): the fp32 NHWC activation after every block of the topology is copied
  // into a slot of the workspace, for per-block error budgets against the fp32 reference arithmetic (tests/test_gpu_parity_budget.py).
  bool taps_on = false;
  struct TapInfo {
    std::string name;
    const float* ptr;
    int64_t rows;
    int C, n_img;
  };
  std::vector<TapInfo> taps;
  void* p_ws = nullptr;
  size_t p_ws_bytes = 0;
  struct CachedPlan {
    int B, V, H, W, R;
    void* ws;
    size_t ws_bytes;
    std::vector<Op> ops;
  };
  std::vector<CachedPlan> cache;
  std::vector<cudaEvent_t> events;
  std::vector<std::vector<cudaEvent_t>> deferred, event_pool;
  std::vector<std::vector<int>> deferred_cls;  // op classes of the plan each deferred run used

  ~Unet() {
    for (void* p : owned) cudaFree(p);
    for (auto& kv : raw)
      if (kv.second.d) cudaFree(kv.second.d);
    for (cudaEvent_t e : events) cudaEventDestroy(e);
    for (auto& v : deferred)
      for (cudaEvent_t e : v) cudaEventDestroy(e);
    for (auto& v : event_pool)
      for (cudaEvent_t e : v) cudaEventDestroy(e);
  }

  // ------------------------------------------------------------------ topology
  bool attn_at(int ds) const {
    for (int i = 0; i < cfg.n_attention_resolutions; ++i)
      if (cfg.attention_resolutions[i] == ds) return true;
    return false;
  }

  int add_res(const std::string& prefix, int cin, int cout) {
    ResW r;
    r.prefix = prefix;
    r.cin = cin;
    r.cout = cout;
    r.skip = cin != cout;
    r.emb_off = n_all;
    n_all += cout;
    res.push_back(r);
    return static_cast<int>(res.size()) - 1;
  }
  int add_tf(const std::string& prefix, int C, int mult) {
    TfW t;
    t.prefix = prefix;
    t.C = C;
    t.is3d = mult >= 2;  // mmdm_unet.py:49-55
    tf.push_back(t);
    return static_cast<int>(tf.size()) - 1;
  }

  bool build_topology() {
    const int mc = cfg.model_channels;
    emb_ch = 4 * mc;
    kpad_in = ((9 * cfg.in_channels + cfg.condition_channels + 63) / 64) * 64;
    if (cfg.num_head_channels != 64) {
      set_error("only num_head_channels == 64 is implemented");
      return false;
    }
    if (cfg.n_levels < 1 || cfg.n_levels > CAP4D_B200_MAX_LEVELS || mc % 64 != 0) {
      set_error("model_channels must be a multiple of 64 and 1 <= n_levels <= 8");
      return false;
    }
    char buf[128];
    input_blocks.clear();
    input_blocks.push_back(Block());  // conv_in
    input_block_chans.assign(1, mc);
    int ch = mc, ds = 1;
    for (int level = 0; level < cfg.n_levels; ++level) {
      const int mult = cfg.channel_mult[level];
      for (int nr = 0; nr < cfg.num_res_blocks; ++nr) {
        Block b;
        snprintf(buf, sizeof(buf), "input_blocks.%d.", static_cast<int>(input_blocks.size()));
        b.push_back({L_RES, add_res(std::string(buf) + "0.", ch, mult * mc)});
        ch = mult * mc;
        if (attn_at(ds)) b.push_back({L_TF, add_tf(std::string(buf) + "1.", ch, mult)});
        input_blocks.push_back(b);
        input_block_chans.push_back(ch);
      }
      if (level != cfg.n_levels - 1) {
        Block b;
        snprintf(buf, sizeof(buf), "input_blocks.%d.0.op.", static_cast<int>(input_blocks.size()));
        ConvW d;
        d.prefix = buf;
        d.cin = d.cout = ch;
        down.push_back(d);
        b.push_back({L_DOWN, static_cast<int>(down.size()) - 1});
        input_blocks.push_back(b);
        input_block_chans.push_back(ch);
        ds *= 2;
      }
    }
    middle.clear();
    middle.push_back({L_RES, add_res("middle_block.0.", ch, ch)});
    middle.push_back({L_TF, add_tf("middle_block.1.", ch, cfg.channel_mult[cfg.n_levels - 1])});
    middle.push_back({L_RES, add_res("middle_block.2.", ch, ch)});
    output_blocks.clear();
    std::vector<int> chans = input_block_chans;
    for (int level = cfg.n_levels - 1; level >= 0; --level) {
      const int mult = cfg.channel_mult[level];
      for (int i = 0; i <= cfg.num_res_blocks; ++i) {
        const int ich = chans.back();
        chans.pop_back();
        Block b;
        snprintf(buf, sizeof(buf), "output_blocks.%d.", static_cast<int>(output_blocks.size()));
        b.push_back({L_RES, add_res(std::string(buf) + "0.", ch + ich, mc * mult)});
        ch = mc * mult;
        int sub = 1;
        if (attn_at(ds)) {
          b.push_back({L_TF, add_tf(std::string(buf) + "1.", ch, mult)});
          sub = 2;
        }
        if (level && i == cfg.num_res_blocks) {
          ConvW u;
          u.prefix = std::string(buf) + std::to_string(sub) + ".conv.";
          u.cin = u.cout = ch;
          up.push_back(u);
          b.push_back({L_UP, static_cast<int>(up.size()) - 1});
          ds /= 2;
        }
        output_blocks.push_back(b);
      }
    }
    if (ch != mc) {
      set_error("topology: final channel count != model_channels");
      return false;
    }
    return true;
  }

  // ------------------------------------------------------------------ parameter census
  // every state_dict entry of MMDMUnetModel for this topology, in a fixed order
  std::vector<std::pair<std::string, std::vector<int64_t>>> expected_params() const {
    std::vector<std::pair<std::string, std::vector<int64_t>>> v;
    const int64_t mc = cfg.model_channels, e = emb_ch;
    auto add = [&](const std::string& n, std::vector<int64_t> shp) { v.emplace_back(n, std::move(shp)); };
    add("time_embed.0.weight", {e, mc});
    add("time_embed.0.bias", {e});
    add("time_embed.2.weight", {e, e});
    add("time_embed.2.bias", {e});
    add("input_blocks.0.0.weight", {mc, cfg.in_channels, 3, 3});
    add("input_blocks.0.0.bias", {mc});
    for (const ResW& r : res) {
      const std::string& p = r.prefix;
      add(p + "in_layers.0.weight", {r.cin});
      add(p + "in_layers.0.bias", {r.cin});
      add(p + "in_layers.2.weight", {r.cout, r.cin, 3, 3});
      add(p + "in_layers.2.bias", {r.cout});
      add(p + "emb_layers.1.weight", {r.cout, e});
      add(p + "emb_layers.1.bias", {r.cout});
      add(p + "out_layers.0.weight", {r.cout});
      add(p + "out_layers.0.bias", {r.cout});
      add(p + "out_layers.3.weight", {r.cout, r.cout, 3, 3});
      add(p + "out_layers.3.bias", {r.cout});
      if (r.skip) {
        add(p + "skip_connection.weight", {r.cout, r.cin, 1, 1});
        add(p + "skip_connection.bias", {r.cout});
      }
    }
    for (const TfW& w : tf) {
      const std::string& p = w.prefix;
      const std::string tb = p + "transformer_blocks.0.";
      const int64_t C = w.C;
      add(p + "norm.weight", {C});
      add(p + "norm.bias", {C});
      add(p + "proj_in.weight", {C, C});
      add(p + "proj_in.bias", {C});
      add(tb + "attn1.to_q.weight", {C, C});
      add(tb + "attn1.to_k.weight", {C, C});
      add(tb + "attn1.to_v.weight", {C, C});
      add(tb + "attn1.to_out.0.weight", {C, C});
      add(tb + "attn1.to_out.0.bias", {C});
      add(tb + "norm1.weight", {C});
      add(tb + "norm1.bias", {C});
      add(tb + "norm3.weight", {C});
      add(tb + "norm3.bias", {C});
      add(tb + "ff.net.0.proj.weight", {8 * C, C});
      add(tb + "ff.net.0.proj.bias", {8 * C});
      add(tb + "ff.net.2.weight", {C, 4 * C});
      add(tb + "ff.net.2.bias", {C});
      add(p + "proj_out.weight", {C, C});
      add(p + "proj_out.bias", {C});
    }
    for (int pass = 0; pass < 2; ++pass)
      for (const ConvW& c : (pass == 0 ? down : up)) {
        add(c.prefix + "weight", {c.cout, c.cin, 3, 3});
        add(c.prefix + "bias", {c.cout});
      }
    add("out.0.weight", {mc});
    add("out.0.bias", {mc});
    add("out.2.weight", {cfg.out_channels, mc, 3, 3});
    add("out.2.bias", {cfg.out_channels});
    add("cond_linear.weight", {mc, cfg.condition_channels});
    add("cond_linear.bias", {mc});
    return v;
  }

  // ------------------------------------------------------------------ weights
  template <typename T>
  T* dev_alloc(size_t n, bool zero = false) {
    void* p = nullptr;
    if (cudaMalloc(&p, std::max<size_t>(n * sizeof(T), 16)) != cudaSuccess) return nullptr;
    if (zero) cudaMemset(p, 0, std::max<size_t>(n * sizeof(T), 16));
    owned.push_back(p);
    return static_cast<T*>(p);
  }

  bool get(const std::string& name, size_t numel, const RawTensor** out) {
    auto it = raw.find(name);
    if (it == raw.end()) {
      set_error("missing weight: " + name);
      return false;
    }
    if (it->second.numel != numel) {
      set_error("weight " + name + ": expected " + std::to_string(numel) + " elements, got " +
                std::to_string(it->second.numel));
      return false;
    }
    *out = &it->second;
    return true;
  }
  // fp32 parameter used as is (kept alive in `raw`)
  bool fp(const std::string& name, size_t numel, float** out) {
    const RawTensor* t;
    if (!get(name, numel, &t)) return false;
    *out = t->d;
    return true;
  }
  bool pack_linear(const std::string& name, int rows, int cols, bf16** out) {
    const RawTensor* t;
    if (!get(name, static_cast<size_t>(rows) * cols, &t)) return false;
    *out = walloc(static_cast<size_t>(rows) * cols);
    if (!*out) {
      set_error("cudaMalloc failed for " + name);
      return false;
    }
    CUDA_OK(launch_pack_matrix(t->d, rows, cols, *out, cols, 0, 0, 0));
    consumed.push_back(name);
    return split_weight(out, rows, cols, {{0, cols, cols}});
  }
  std::vector<std::string> consumed;

  // Operand formats.  tcgen05 kind::f16 takes bf16 or fp16 operands (one format for both operands of an MMA: a
  // mixed bf16 x fp16 descriptor faults on sm_100a).  fp16 has three more mantissa bits, and operand rounding IS the
  // error budget of this path (bf16 everywhere: 7.7e-3 .. 1.03e-2 of the 1e-2 tolerance over seeds / timesteps,
  // half of it from the weights alone, scripts/weight_rounding_error.py).  So every GEMM whose activation operand
  // has a range bounded by construction - the output of a GroupNorm / LayerNorm (conv1, conv2, proj_in, QKV, FF1,
  // the output conv) or the packed network input - runs fp16 x fp16; everything fed by an unnormalised tensor (raw
  // residual stream into the skip / down / up convs, attention Q K V P and output, GEGLU output, FF2 / proj_out
  // inputs) stays bf16 x bf16.  finalize() refuses weights that fp16 cannot hold.  CAP4D_OPERANDS=bf16: all bf16.
  bool f16_ok = true;
  // fp32-accuracy mode (cap4d_b200_unet_set_precision, before finalize; see precise.cu): every MMA operand is the
  // exact three-way bf16 split of an fp32 tensor laid out as six K segments, the pointwise maths is exact, attention
  // runs in fp32.  kx() = how many times wider an operand is than the tensor it represents.
  bool precise = false;
  int kx() const { return precise ? 6 : 1; }
  void pack_as(bool f16) { set_weight_pack_format(precise ? 2 : ((f16 && f16_ok) ? 1 : 0)); }
  // weight buffer of n elements in the current pack format
  bf16* walloc(size_t n, bool zero = false) { return dev_alloc<bf16>(precise ? 2 * n : n, zero); }
  std::vector<void*> split_sources;  // fp32 repacked weights, freed once their splits exist
  // precise mode: *w holds fp32 [rows][K] in the GEMM's K order -> bf16 [rows][6 K], weight-side segment order.
  // parts: (first column, column count, segment width) - the K ranges that face different operand tensors
  // conv51: the layout of a 3x3 conv's weights (precise.cu, split51): [rows][5 K] small-term segments followed by
  // [rows][K] leading segments, for the two launches of conv51()
  bool split_weight(bf16** w, int rows, int K, const std::vector<std::array<int, 3>>& parts, bool conv51 = false) {
    if (!precise) return true;
    bf16* out = dev_alloc<bf16>(static_cast<size_t>(rows) * 6 * K);
    if (!out) {
      set_error("cudaMalloc failed (split weights)");
      return false;
    }
    const float* src = reinterpret_cast<const float*>(*w);
    bf16* hi = out + static_cast<size_t>(rows) * 5 * K;
    for (const auto& pt : parts) {
      if (conv51)
        CUDA_OK(launch_split51(src + pt[0], rows, pt[1], K, pt[2], out + static_cast<size_t>(5) * pt[0],
                               static_cast<size_t>(5) * K, hi + pt[0], K, 1, 0));
      else
        CUDA_OK(launch_split6(src + pt[0], rows, pt[1], K, pt[2], out + static_cast<size_t>(6) * pt[0],
                              static_cast<size_t>(6) * K, 1, 0));
    }
    split_sources.push_back(*w);
    *w = out;
    return true;
  }

  // every matrix / conv weight must be representable in fp16 (|w| <= 65504); checked on the device
  bool weights_fit_f16() {
    float* d_max = dev_alloc<float>(1, true);
    if (!d_max) {
      set_error("cudaMalloc failed");
      return false;
    }
    for (const auto& kv : raw)
      if (kv.second.shape.size() >= 2 && kv.second.numel > 0)
        absmax_kernel<<<static_cast<unsigned>(std::min<size_t>(1024, (kv.second.numel + 255) / 256)), 256>>>(
            kv.second.d, kv.second.numel, d_max);
    float h = 0.f;
    CUDA_OK(cudaMemcpy(&h, d_max, sizeof(float), cudaMemcpyDeviceToHost));
    if (!(h <= 65504.f)) {
      set_error("a weight exceeds the fp16 range (max |w| = " + std::to_string(h) + "); set CAP4D_OPERANDS=bf16");
      return false;
    }
    return true;
  }

  bool finalize() {
    if (finalized) return true;
    if (const char* e = getenv("CAP4D_OPERANDS")) f16_ok = std::string(e) != "bf16";
    struct PackGuard {  // the pack kernels read the target format from a thread-local switch
      ~PackGuard() { set_weight_pack_f16(false); }
    } pack_guard;
    if (precise) f16_ok = false;  // the split operands are bf16 triples
    if (f16_ok && !weights_fit_f16()) return false;
    const int mc = cfg.model_channels;
    const RawTensor* t;
    // ---- conv_in + cond_linear fused into one [mc][kpad] matrix (mmdm_unet.py:92-107)
    {
      w_in = walloc(static_cast<size_t>(mc) * kpad_in, true);
      b_in = dev_alloc<float>(mc);
      pack_as(true);
      if (!get("input_blocks.0.0.weight", static_cast<size_t>(mc) * cfg.in_channels * 9, &t)) return false;
      CUDA_OK(launch_pack_conv_weight(t->d, mc, cfg.in_channels, 3, 3, w_in, kpad_in, 0, 0));
      consumed.push_back("input_blocks.0.0.weight");
      if (!get("cond_linear.weight", static_cast<size_t>(mc) * cfg.condition_channels, &t)) return false;
      CUDA_OK(launch_pack_matrix(t->d, mc, cfg.condition_channels, w_in, kpad_in, 9 * cfg.in_channels, 0, 0));
      consumed.push_back("cond_linear.weight");
      float *b0, *b1;
      if (!fp("input_blocks.0.0.bias", mc, &b0) || !fp("cond_linear.bias", mc, &b1)) return false;
      vec_add_kernel<<<(mc + 255) / 256, 256>>>(b0, b1, b_in, mc);
      if (!split_weight(&w_in, mc, kpad_in, {{0, kpad_in, kpad_in}})) return false;
    }
    // ---- time embedding (fp32)
    if (!fp("time_embed.0.weight", static_cast<size_t>(emb_ch) * mc, &te_w1) || !fp("time_embed.0.bias", emb_ch, &te_b1) ||
        !fp("time_embed.2.weight", static_cast<size_t>(emb_ch) * emb_ch, &te_w2) ||
        !fp("time_embed.2.bias", emb_ch, &te_b2))
      return false;
    wall = dev_alloc<float>(static_cast<size_t>(n_all) * emb_ch);
    ball = dev_alloc<float>(n_all);
    if (!wall || !ball) {
      set_error("cudaMalloc failed (emb layers)");
      return false;
    }
    // ---- ResBlocks
    for (ResW& r : res) {
      const std::string& p = r.prefix;
      if (!fp(p + "in_layers.0.weight", r.cin, &r.gn1_g) || !fp(p + "in_layers.0.bias", r.cin, &r.gn1_b) ||
          !fp(p + "out_layers.0.weight", r.cout, &r.gn2_g) || !fp(p + "out_layers.0.bias", r.cout, &r.gn2_b) ||
          !fp(p + "in_layers.2.bias", r.cout, &r.b1))
        return false;
      if (!get(p + "in_layers.2.weight", static_cast<size_t>(r.cout) * r.cin * 9, &t)) return false;
      r.w1 = walloc(static_cast<size_t>(r.cout) * 9 * r.cin);
      if (!r.w1) {
        set_error("cudaMalloc failed");
        return false;
      }
      pack_as(true);
      CUDA_OK(launch_pack_conv_weight(t->d, r.cout, r.cin, 3, 3, r.w1, 9 * r.cin, 0, 0));
      consumed.push_back(p + "in_layers.2.weight");
      if (!split_weight(&r.w1, r.cout, 9 * r.cin, {{0, 9 * r.cin, r.cin}}, true)) return false;
      const int k2 = 9 * r.cout + (r.skip ? r.cin : 0);
      if (!get(p + "out_layers.3.weight", static_cast<size_t>(r.cout) * r.cout * 9, &t)) return false;
      r.w2 = walloc(static_cast<size_t>(r.cout) * k2);
      r.b2 = dev_alloc<float>(r.cout);
      if (!r.w2 || !r.b2) {
        set_error("cudaMalloc failed");
        return false;
      }
      pack_as(!r.skip);  // the fused 1x1 skip conv reads the raw residual stream: that MMA stays bf16
      CUDA_OK(launch_pack_conv_weight(t->d, r.cout, r.cout, 3, 3, r.w2, k2, 0, 0));
      consumed.push_back(p + "out_layers.3.weight");
      float *bo, *bs = nullptr;
      if (!fp(p + "out_layers.3.bias", r.cout, &bo)) return false;
      if (r.skip) {
        // 1x1 skip conv (openaimodel.py:235-242) appended along K of the second 3x3 conv
        if (!get(p + "skip_connection.weight", static_cast<size_t>(r.cout) * r.cin, &t)) return false;
        CUDA_OK(launch_pack_matrix(t->d, r.cout, r.cin, r.w2, k2, 9 * r.cout, 0, 0));
        consumed.push_back(p + "skip_connection.weight");
        if (!fp(p + "skip_connection.bias", r.cout, &bs)) return false;
      }
      vec_add_kernel<<<(r.cout + 255) / 256, 256>>>(bo, bs, r.b2, r.cout);
      {
        std::vector<std::array<int, 3>> parts = {{0, 9 * r.cout, r.cout}};
        if (r.skip) parts.push_back({9 * r.cout, r.cin, r.cin});
        if (!split_weight(&r.w2, r.cout, k2, parts, true)) return false;
      }
      // emb_layers (openaimodel.py:203-209) -> rows of the shared [n_all][emb_ch] matrix
      if (!get(p + "emb_layers.1.weight", static_cast<size_t>(r.cout) * emb_ch, &t)) return false;
      CUDA_OK(cudaMemcpy(wall + static_cast<size_t>(r.emb_off) * emb_ch, t->d, t->numel * sizeof(float),
                         cudaMemcpyDeviceToDevice));
      consumed.push_back(p + "emb_layers.1.weight");
      if (!get(p + "emb_layers.1.bias", r.cout, &t)) return false;
      CUDA_OK(cudaMemcpy(ball + r.emb_off, t->d, t->numel * sizeof(float), cudaMemcpyDeviceToDevice));
    }
    // ---- transformers
    for (TfW& w : tf) {
      const std::string& p = w.prefix;
      const std::string tb = p + "transformer_blocks.0.";
      const int C = w.C;
      if (!fp(p + "norm.weight", C, &w.gn_g) || !fp(p + "norm.bias", C, &w.gn_b) ||
          !fp(p + "proj_in.bias", C, &w.bpi) || !fp(p + "proj_out.bias", C, &w.bpo) ||
          !fp(tb + "norm1.weight", C, &w.ln1_g) || !fp(tb + "norm1.bias", C, &w.ln1_b) ||
          !fp(tb + "norm3.weight", C, &w.ln3_g) || !fp(tb + "norm3.bias", C, &w.ln3_b) ||
          !fp(tb + "attn1.to_out.0.bias", C, &w.bo) || !fp(tb + "ff.net.2.bias", C, &w.bff2))
        return false;
      pack_as(true);   // proj_in reads a GroupNorm output
      if (!pack_linear(p + "proj_in.weight", C, C, &w.wpi)) return false;
      pack_as(false);  // proj_out / to_out / FF2 read unnormalised activations
      if (!pack_linear(p + "proj_out.weight", C, C, &w.wpo) || !pack_linear(tb + "attn1.to_out.0.weight", C, C, &w.wo) ||
          !pack_linear(tb + "ff.net.2.weight", C, 4 * C, &w.wff2))
        return false;
      pack_as(true);   // QKV and FF1 read LayerNorm outputs
      // fused QKV [3C][C] (attention.py:168-170, no bias)
      w.wqkv = walloc(static_cast<size_t>(3) * C * C);
      if (!w.wqkv) {
        set_error("cudaMalloc failed");
        return false;
      }
      const char* names[3] = {"attn1.to_q.weight", "attn1.to_k.weight", "attn1.to_v.weight"};
      for (int i = 0; i < 3; ++i) {
        if (!get(tb + names[i], static_cast<size_t>(C) * C, &t)) return false;
        CUDA_OK(launch_pack_matrix(t->d, C, C, w.wqkv, C, 0, i * C, 0));
        consumed.push_back(tb + names[i]);
      }
      if (!split_weight(&w.wqkv, 3 * C, C, {{0, C, C}})) return false;
      // GEGLU projection [8C][C] interleaved
      const RawTensor* tbias;
      if (!get(tb + "ff.net.0.proj.weight", static_cast<size_t>(8) * C * C, &t) ||
          !get(tb + "ff.net.0.proj.bias", static_cast<size_t>(8) * C, &tbias))
        return false;
      w.wff1 = walloc(static_cast<size_t>(8) * C * C);
      w.bff1 = dev_alloc<float>(static_cast<size_t>(8) * C);
      if (!w.wff1 || !w.bff1) {
        set_error("cudaMalloc failed");
        return false;
      }
      if (precise) {
        // the reference's row order [x | gate]: GEGLU is a separate exact kernel in this mode (precise.cu)
        CUDA_OK(launch_pack_matrix(t->d, 8 * C, C, w.wff1, C, 0, 0, 0));
        CUDA_OK(cudaMemcpy(w.bff1, tbias->d, static_cast<size_t>(8) * C * sizeof(float), cudaMemcpyDeviceToDevice));
        if (!split_weight(&w.wff1, 8 * C, C, {{0, C, C}})) return false;
      } else {
        pack_geglu_kernel<<<sm_count() * 8, 256>>>(t->d, tbias->d, 4 * C, C, w.wff1, w.bff1, f16_ok ? 1 : 0);
      }
      consumed.push_back(tb + "ff.net.0.proj.weight");
    }
    // ---- down / up convs (inputs: the raw residual stream -> bf16)
    pack_as(false);
    for (int pass = 0; pass < 2; ++pass) {
      for (ConvW& c : (pass == 0 ? down : up)) {
        if (!get(c.prefix + "weight", static_cast<size_t>(c.cout) * c.cin * 9, &t)) return false;
        // up: nearest-2x + conv3x3 is folded into four 2x2-tap phase convs on the low-res grid: [4][Cout][4*Cin]
        c.w = walloc(static_cast<size_t>(c.cout) * (pass == 0 ? 9 : 16) * c.cin);
        if (!c.w) {
          set_error("cudaMalloc failed");
          return false;
        }
        if (pass == 1) {
          CUDA_OK(launch_pack_upconv_weight(t->d, c.cout, c.cin, c.w, 0));
        } else
        CUDA_OK(launch_pack_conv_weight(t->d, c.cout, c.cin, 3, 3, c.w, 9 * c.cin, 0, 0));
        consumed.push_back(c.prefix + "weight");
        if (!fp(c.prefix + "bias", c.cout, &c.b)) return false;
        if (pass == 0) {
          if (!split_weight(&c.w, c.cout, 9 * c.cin, {{0, 9 * c.cin, c.cin}}, true)) return false;
        } else if (!split_weight(&c.w, 4 * c.cout, 4 * c.cin, {{0, 4 * c.cin, c.cin}})) {  // [phase][cout][4 taps][cin]
          return false;
        }
      }
    }
    // ---- out: GN -> SiLU -> conv3x3 mc -> out_channels, N padded to 32 (openaimodel.py:770-774)
    {
      if (!fp("out.0.weight", mc, &out_gn_g) || !fp("out.0.bias", mc, &out_gn_b)) return false;
      if (cfg.out_channels > 32) {
        set_error("out_channels > 32 is not implemented");
        return false;
      }
      pack_as(true);
      w_out = walloc(static_cast<size_t>(32) * 9 * mc, true);
      b_out = dev_alloc<float>(32, true);
      if (!get("out.2.weight", static_cast<size_t>(cfg.out_channels) * mc * 9, &t)) return false;
      CUDA_OK(launch_pack_conv_weight(t->d, cfg.out_channels, mc, 3, 3, w_out, 9 * mc, 0, 0));
      consumed.push_back("out.2.weight");
      if (!get("out.2.bias", cfg.out_channels, &t)) return false;
      CUDA_OK(cudaMemcpy(b_out, t->d, cfg.out_channels * sizeof(float), cudaMemcpyDeviceToDevice));
      if (!split_weight(&w_out, 32, 9 * mc, {{0, 9 * mc, mc}}, true)) return false;
    }
    d_violations = dev_alloc<int>(1, true);
    CUDA_OK(cudaDeviceSynchronize());
    for (void* ptr : split_sources) {
      owned.erase(std::remove(owned.begin(), owned.end(), ptr), owned.end());
      cudaFree(ptr);
    }
    split_sources.clear();
    // the big fp32 originals are no longer needed
    for (const std::string& name : consumed) {
      auto it = raw.find(name);
      if (it != raw.end() && it->second.d) {
        cudaFree(it->second.d);
        raw.erase(it);
      }
    }
    consumed.clear();
    finalized = true;
    return true;
  }

  // ------------------------------------------------------------------ planning
  struct PlanCtx {
    Arena arena;
    bool dry = true;
    uint8_t* base = nullptr;
    int n_img = 0, V = 0;      // n_img: images the layers being planned work on (B*V, or B*G once compact)
    int B = 0, R = 0, G = 0;   // R > 0: drop the R leading views of every group after the last cross-view layer
    bool compact = false;      // the current activations hold only the generated views
    float* emb_all = nullptr;  // rows of the images being planned (all views, or the gathered generated views)
    float* emb_gen = nullptr;
    float* gn_partial = nullptr;
    std::vector<Op>* ops = nullptr;
    template <typename T>
    T* ptr(const Buf& b) const {
      return reinterpret_cast<T*>(base + b.off);
    }
    Buf alloc(int M, int C, size_t elem) {
      Buf b;
      b.M = M;
      b.C = C;
      b.bytes = static_cast<size_t>(M) * C * elem;
      b.off = arena.alloc(b.bytes);
      b.valid = true;
      return b;
    }
    void release(Buf& b) {
      if (b.valid) arena.release(b.off, b.bytes);
      b.valid = false;
    }
  };

  // f16: both operands of this GEMM are stored as fp16 (see "Operand formats" above)
  bool add_gemm_op(PlanCtx& c, int cls, const GemmPlan& plan, bool f16, double exec_flops = -1) {
    Op op;
    op.cls = cls;
    op.launches = 1;
    op.flops = plan.flops;
    op.exec_flops = exec_flops;
    op.bytes = 0;
    GemmPlan copy = plan;
    copy.p.a_f16 = copy.p.b_f16 = (f16 && f16_ok) ? 1 : 0;
    op.run = [copy](cudaStream_t s) { return launch_gemm(copy, s); };
    c.ops->push_back(op);
    return true;
  }

  // fp32 workspace tensor [M][C] -> its six-segment bf16 operand dst [M][6C] (activation-side order); precise mode
  // conv51: dst is laid out for conv51() - [M][5C] small-term segments followed by [M][C] leading segments
  void op_split(PlanCtx& c, const Buf& src, const Buf& dst, bool conv51 = false) {
    if (c.dry) return;
    const float* ps = c.ptr<float>(src);
    bf16* pd = c.ptr<bf16>(dst);
    const size_t M = src.M;
    const int C = src.C;
    Op op;
    op.cls = CLS_OTHER;
    op.launches = 1;
    op.flops = 0;
    op.bytes = static_cast<double>(M) * C * 16;
    if (conv51)
      op.run = [=](cudaStream_t s) {
        return launch_split51(ps, M, C, C, C, pd, static_cast<size_t>(5) * C, pd + M * 5 * C, C, 0, s);
      };
    else
      op.run = [=](cudaStream_t s) { return launch_split6(ps, M, C, C, C, pd, static_cast<size_t>(6) * C, 0, s); };
    c.ops->push_back(op);
  }

  // precise mode: a 3x3 conv over conv51-laid-out operands = the small-terms launch (bias, timestep bias, residual)
  // followed by the leading-term launch, which adds the first result as its fp32 residual.  Cin / K2: channels of ONE
  // segment of the conv operand `a` / of the appended 1x1 operand `a2`.
  bool conv51(PlanCtx& c, const ConvGeom& g, const Buf& a, int Cin, const Buf* a2, int K2, const bf16* w, int N,
              const Buf& out, const float* bias, const float* rowbias, int rb_div, int rb_ld, const float* residual) {
    Buf tmp = c.alloc(out.M, N, 4);
    if (!c.dry) {
      const size_t Ma = a.M;
      const bf16* a_lo = c.ptr<bf16>(a);
      const bf16* a_hi = a_lo + Ma * 5 * Cin;
      const bf16* a2_lo = a2 ? c.ptr<bf16>(*a2) : nullptr;
      const bf16* a2_hi = a2 ? a2_lo + static_cast<size_t>(a2->M) * 5 * K2 : nullptr;
      const int Kt = 9 * Cin + (a2 ? K2 : 0);
      const bf16* w_hi = w + static_cast<size_t>(N) * 5 * Kt;
      GemmPlan p;
      if (!make_conv_plan(&p, a_lo, g, 5 * Cin, a2_lo, a2 ? 5 * K2 : 0, w, N, OUT_F32, c.ptr<float>(tmp), N, bias, rowbias,
                          rb_div, rb_ld, residual, N))
        return false;
      add_gemm_op(c, CLS_CONV, p, false);
      if (!make_conv_plan(&p, a_hi, g, Cin, a2_hi, a2 ? K2 : 0, w_hi, N, OUT_F32, c.ptr<float>(out), N, nullptr, nullptr, 1,
                          0, c.ptr<float>(tmp), N))
        return false;
      add_gemm_op(c, CLS_CONV, p, false);
    }
    c.release(tmp);
    return true;
  }

  // out (and raw_out) are operand buffers: [M][kx() * C] 16-bit
  bool op_gn(PlanCtx& c, const Buf& x1, const Buf* x2, int hw, const float* g, const float* b, float eps, int silu,
             const Buf& out, const Buf* raw_out, bool out_f16, bool conv_operand = false) {
    const int C1 = x1.C, C2 = x2 ? x2->C : 0;
    Buf t32, r32;  // precise mode: the norm writes fp32 (exact SiLU), which is then split
    if (precise) {
      t32 = c.alloc(x1.M, C1 + C2, 4);
      if (raw_out) r32 = c.alloc(x1.M, C1 + C2, 4);
    }
    if (!c.dry) {
      const float* p1 = c.ptr<float>(x1);
      const float* p2 = x2 ? c.ptr<float>(*x2) : nullptr;
      bf16* po = precise ? reinterpret_cast<bf16*>(c.ptr<float>(t32)) : c.ptr<bf16>(out);
      bf16* pr = raw_out ? (precise ? reinterpret_cast<bf16*>(c.ptr<float>(r32)) : c.ptr<bf16>(*raw_out)) : nullptr;
      float* partial = c.gn_partial;
      const int n_img = c.n_img;
      // once compact, x1 holds the generated views only while a skip tensor x2 still holds every view
      const int x2G = (c.compact && x2 != nullptr) ? c.G : 0, x2V = c.V, x2R = c.R, n_layout = c.B * c.V;
      const int fmt = precise ? 2 : ((out_f16 && f16_ok) ? 1 : 0);
      Op op;
      op.cls = CLS_GN;
      op.launches = 1;
      op.flops = 0;
      op.bytes = static_cast<double>(x1.M) * (C1 + C2) * (4 + 2 + (raw_out ? 2 : 0));
      op.run = [=](cudaStream_t s) {
        return launch_groupnorm(p1, C1, p2, C2, n_img, hw, g, b, eps, silu, po, pr, partial, s, x2G, x2V, x2R,
                                n_layout, fmt);
      };
      c.ops->push_back(op);
    }
    if (precise) {
      op_split(c, t32, out, conv_operand);
      if (raw_out) op_split(c, r32, *raw_out, conv_operand);
      c.release(t32);
      c.release(r32);
    }
    return true;
  }

  // out: operand buffer [M][kx() * C]
  bool op_ln(PlanCtx& c, const Buf& x, const float* g, const float* b, const Buf& out) {
    Buf t32;
    if (precise) t32 = c.alloc(x.M, x.C, 4);
    if (!c.dry) {
      const float* px = c.ptr<float>(x);
      bf16* po = precise ? reinterpret_cast<bf16*>(c.ptr<float>(t32)) : c.ptr<bf16>(out);
      const int M = x.M, C = x.C;
      const int fmt = precise ? 2 : (f16_ok ? 1 : 0);
      Op op;
      op.cls = CLS_LN;
      op.launches = 1;
      op.flops = 0;
      op.bytes = static_cast<double>(M) * C * 6;
      op.run = [=](cudaStream_t s) { return launch_layernorm(px, M, C, g, b, 1e-5f, po, s, fmt); };
      c.ops->push_back(op);
    }
    if (precise) {
      op_split(c, t32, out);
      c.release(t32);
    }
    return true;
  }

  // ResBlock._forward (openaimodel.py:256-276); x2 != null: input is cat([x1, x2], channels)
  bool plan_res(PlanCtx& c, const ResW& r, const Buf& x1, const Buf* x2, int H, int W, Buf* out) {
    const int M = x1.M, hw = H * W, k = kx();
    const int cin = x1.C + (x2 ? x2->C : 0);
    if (cin != r.cin) {
      set_error("plan: channel mismatch at " + r.prefix);
      return false;
    }
    Buf a1 = c.alloc(M, k * cin, 2), xb;
    if (r.skip) xb = c.alloc(M, k * cin, 2);
    if (!op_gn(c, x1, x2, hw, r.gn1_g, r.gn1_b, 1e-5f, 1, a1, r.skip ? &xb : nullptr, true, true)) return false;
    Buf h = c.alloc(M, r.cout, 4);
    ConvGeom g{c.n_img, H, W, 9, 1};
    if (precise) {
      if (!conv51(c, g, a1, cin, nullptr, 0, r.w1, r.cout, h, r.b1, c.dry ? nullptr : c.emb_all + r.emb_off, hw, n_all,
                  nullptr))
        return false;
    } else if (!c.dry) {
      GemmPlan p;
      if (!make_conv_plan(&p, c.ptr<bf16>(a1), g, k * cin, nullptr, 0, r.w1, r.cout, OUT_F32, c.ptr<float>(h), r.cout,
                          r.b1, c.emb_all + r.emb_off, hw, n_all, nullptr, 0))
        return false;
      add_gemm_op(c, CLS_CONV, p, true);
    }
    c.release(a1);
    Buf a2 = c.alloc(M, k * r.cout, 2);
    if (!op_gn(c, h, nullptr, hw, r.gn2_g, r.gn2_b, 1e-5f, 1, a2, nullptr, !r.skip, true)) return false;
    c.release(h);
    *out = c.alloc(M, r.cout, 4);
    if (precise) {
      if (!conv51(c, g, a2, r.cout, r.skip ? &xb : nullptr, cin, r.w2, r.cout, *out, r.b2, nullptr, 1, 0,
                  (r.skip || c.dry) ? nullptr : c.ptr<float>(x1)))
        return false;
    } else if (!c.dry) {
      GemmPlan p;
      const float* residual = r.skip ? nullptr : c.ptr<float>(x1);
      if (!make_conv_plan(&p, c.ptr<bf16>(a2), g, k * r.cout, r.skip ? c.ptr<bf16>(xb) : nullptr, r.skip ? k * cin : 0,
                          r.w2, r.cout, OUT_F32, c.ptr<float>(*out), r.cout, r.b2, nullptr, 1, 0, residual, r.cout))
        return false;
      add_gemm_op(c, CLS_CONV, p, !r.skip);
    }
    c.release(a2);
    c.release(xb);
    return true;
  }

  // SpatioTemporalTransformer.forward (attention.py:375-387) + BasicTransformerBlock (:311-326)
  bool plan_tf(PlanCtx& c, const TfW& w, const Buf& x, int H, int W, Buf* out) {
    if (precise) return plan_tf_precise(c, w, x, H, W, out);
    const int M = x.M, C = w.C, hw = H * W;
    Buf a = c.alloc(M, C, 2);
    if (!op_gn(c, x, nullptr, hw, w.gn_g, w.gn_b, 1e-6f, 0, a, nullptr, true)) return false;
    Buf t0 = c.alloc(M, C, 4);
    auto gemm = [&](const Buf& A, int K, const bf16* Wt, int N, int mode, const Buf& o, int ldo, const float* bias,
                    const float* residual, bool f16) -> bool {
      if (c.dry) return true;
      GemmPlan p;
      if (!make_gemm_plan(&p, c.ptr<bf16>(A), M, K, nullptr, 0, Wt, N, mode, c.base + o.off, ldo, bias, nullptr, 1, 0,
                          residual, C))
        return false;
      return add_gemm_op(c, CLS_LINEAR, p, f16);
    };
    if (!gemm(a, C, w.wpi, C, OUT_F32, t0, C, w.bpi, nullptr, true)) return false;
    c.release(a);
    Buf n1 = c.alloc(M, C, 2);
    if (!op_ln(c, t0, w.ln1_g, w.ln1_b, n1)) return false;
    Buf qkv = c.alloc(M, 3 * C, 2);
    if (!gemm(n1, C, w.wqkv, 3 * C, OUT_BF16, qkv, 3 * C, nullptr, nullptr, true)) return false;
    c.release(n1);
    Buf o = c.alloc(M, C, 2);
    if (!c.dry) {
      AttnPlan ap;
      const int L = w.is3d ? c.V * hw : hw;  // attention.py:233 vs :237
      if (!make_attn_plan(&ap, c.ptr<bf16>(qkv), c.ptr<bf16>(o), M, C, L, 0.125f)) return false;
      Op op;
      op.cls = CLS_ATTN;
      op.launches = 1;
      op.flops = ap.flops;
      op.bytes = 0;
      op.run = [ap](cudaStream_t s) { return launch_attn(ap, s); };
      c.ops->push_back(op);
    }
    c.release(qkv);
    Buf t1 = c.alloc(M, C, 4);
    if (!gemm(o, C, w.wo, C, OUT_F32, t1, C, w.bo, c.dry ? nullptr : c.ptr<float>(t0), false)) return false;
    c.release(o);
    c.release(t0);
    Buf n3 = c.alloc(M, C, 2);
    if (!op_ln(c, t1, w.ln3_g, w.ln3_b, n3)) return false;
    Buf gg = c.alloc(M, 4 * C, 2);
    if (!gemm(n3, C, w.wff1, 8 * C, OUT_GEGLU_BF16, gg, 4 * C, w.bff1, nullptr, true)) return false;
    c.release(n3);
    Buf t2 = c.alloc(M, C, 2);
    if (!gemm(gg, 4 * C, w.wff2, C, OUT_BF16, t2, C, w.bff2, c.dry ? nullptr : c.ptr<float>(t1), false)) return false;
    c.release(gg);
    c.release(t1);
    *out = c.alloc(M, C, 4);
    if (!gemm(t2, C, w.wpo, C, OUT_F32, *out, C, w.bpo, c.dry ? nullptr : c.ptr<float>(x), false)) return false;
    c.release(t2);
    return true;
  }

  // The same block in fp32-accuracy mode: every GEMM reads six-segment operands and writes fp32, the attention core
  // and GEGLU are the exact fp32 kernels of precise.cu.
  bool plan_tf_precise(PlanCtx& c, const TfW& w, const Buf& x, int H, int W, Buf* out) {
    const int M = x.M, C = w.C, hw = H * W;
    // fp32 [M][K] operand tensor `src` x weights [N][6K] -> fp32 `o` [M][N] (+ bias, + residual [M][N])
    auto gemm6 = [&](const Buf& src6, int K, const bf16* Wt, int N, const Buf& o, const float* bias,
                     const float* residual) -> bool {
      if (c.dry) return true;
      GemmPlan p;
      if (!make_gemm_plan(&p, c.ptr<bf16>(src6), M, 6 * K, nullptr, 0, Wt, N, OUT_F32, c.ptr<float>(o), N, bias, nullptr,
                          1, 0, residual, N))
        return false;
      return add_gemm_op(c, CLS_LINEAR, p, false);
    };
    Buf a = c.alloc(M, 6 * C, 2);
    if (!op_gn(c, x, nullptr, hw, w.gn_g, w.gn_b, 1e-6f, 0, a, nullptr, true)) return false;
    Buf t0 = c.alloc(M, C, 4);
    if (!gemm6(a, C, w.wpi, C, t0, w.bpi, nullptr)) return false;
    c.release(a);
    Buf n1 = c.alloc(M, 6 * C, 2);
    if (!op_ln(c, t0, w.ln1_g, w.ln1_b, n1)) return false;
    Buf qkv = c.alloc(M, 3 * C, 4);
    if (!gemm6(n1, C, w.wqkv, 3 * C, qkv, nullptr, nullptr)) return false;
    c.release(n1);
    Buf o32 = c.alloc(M, C, 4);
    if (!c.dry) {
      const float* pq = c.ptr<float>(qkv);
      float* po = c.ptr<float>(o32);
      const int L = w.is3d ? c.V * hw : hw;  // attention.py:233 vs :237
      Op op;
      op.cls = CLS_ATTN;
      op.launches = 1;
      op.flops = 4.0 * (M / L) * (C / 64) * static_cast<double>(L) * L * 64;
      op.bytes = 0;
      op.run = [=](cudaStream_t s) { return launch_attention_f32(pq, po, M, C, L, 0.125f, s); };
      c.ops->push_back(op);
    }
    c.release(qkv);
    Buf o6 = c.alloc(M, 6 * C, 2);
    op_split(c, o32, o6);
    c.release(o32);
    Buf t1 = c.alloc(M, C, 4);
    if (!gemm6(o6, C, w.wo, C, t1, w.bo, c.dry ? nullptr : c.ptr<float>(t0))) return false;
    c.release(o6);
    c.release(t0);
    Buf n3 = c.alloc(M, 6 * C, 2);
    if (!op_ln(c, t1, w.ln3_g, w.ln3_b, n3)) return false;
    Buf u = c.alloc(M, 8 * C, 4);  // [x | gate], the reference's row order
    if (!gemm6(n3, C, w.wff1, 8 * C, u, w.bff1, nullptr)) return false;
    c.release(n3);
    Buf gl = c.alloc(M, 4 * C, 4);
    if (!c.dry) {
      const float* pu = c.ptr<float>(u);
      float* pg = c.ptr<float>(gl);
      const size_t rows = M;
      const int inner = 4 * C;
      Op op;
      op.cls = CLS_OTHER;
      op.launches = 1;
      op.flops = 0;
      op.bytes = static_cast<double>(M) * C * 48;
      op.run = [=](cudaStream_t s) { return launch_geglu_f32(pu, rows, inner, pg, s); };
      c.ops->push_back(op);
    }
    c.release(u);
    Buf g6 = c.alloc(M, 24 * C, 2);
    op_split(c, gl, g6);
    c.release(gl);
    Buf t2 = c.alloc(M, C, 4);
    if (!gemm6(g6, 4 * C, w.wff2, C, t2, w.bff2, c.dry ? nullptr : c.ptr<float>(t1))) return false;
    c.release(g6);
    c.release(t1);
    Buf t6 = c.alloc(M, 6 * C, 2);
    op_split(c, t2, t6);
    c.release(t2);
    *out = c.alloc(M, C, 4);
    if (!gemm6(t6, C, w.wpo, C, *out, w.bpo, c.dry ? nullptr : c.ptr<float>(x))) return false;
    c.release(t6);
    return true;
  }

  // Downsample (openaimodel.py:135-161): conv3x3 stride 2
  bool plan_down(PlanCtx& c, const ConvW& w, const Buf& x, int H, int W, Buf* out) {
    const int C = w.cin, k = kx();
    Buf pp = c.alloc(x.M, k * C, 2), p32;
    if (precise) p32 = c.alloc(x.M, C, 4);  // fp32 parity planes, then split per pixel
    *out = c.alloc(x.M / 4, w.cout, 4);
    if (!c.dry) {
      const float* px = c.ptr<float>(x);
      bf16* ppp = c.ptr<bf16>(pp);
      bf16* dst = precise ? reinterpret_cast<bf16*>(c.ptr<float>(p32)) : ppp;
      const int n_img = c.n_img, to_f32 = precise ? 1 : 0;
      Op op;
      op.cls = CLS_OTHER;
      op.launches = 1;
      op.flops = 0;
      op.bytes = static_cast<double>(x.M) * C * 6;
      op.run = [=](cudaStream_t s) { return launch_parity_split_bf16(px, n_img, H, W, C, dst, s, to_f32); };
      c.ops->push_back(op);
    }
    if (precise) {
      op_split(c, p32, pp, true);
      c.release(p32);
      ConvGeom g{c.n_img, H / 2, W / 2, 9, 2};
      if (!conv51(c, g, pp, C, nullptr, 0, w.w, w.cout, *out, w.b, nullptr, 1, 0, nullptr)) return false;
    } else if (!c.dry) {
      GemmPlan p;
      ConvGeom g{c.n_img, H / 2, W / 2, 9, 2};
      if (!make_conv_plan(&p, c.ptr<bf16>(pp), g, k * C, nullptr, 0, w.w, w.cout, OUT_F32, c.ptr<float>(*out), w.cout, w.b,
                          nullptr, 1, 0, nullptr, 0))
        return false;
      add_gemm_op(c, CLS_CONV, p, false);
    }
    c.release(pp);
    return true;
  }

  // Upsample (openaimodel.py:92-120): nearest 2x, conv3x3.  Folded: output pixel (2y+py, 2x+px) only sees a
  // 2x2 neighbourhood of the low-res input, so each of the four phases is a 4-tap conv with pre-summed
  // weights (4/9 of the FLOPs, and the 4x upsampled tensor is never written).
  bool plan_up(PlanCtx& c, const ConvW& w, const Buf& x, int H, int W, Buf* out) {
    const int C = w.cin, k = kx();
    Buf xb = c.alloc(x.M, k * C, 2);
    *out = c.alloc(x.M * 4, w.cout, 4);
    if (precise) {
      op_split(c, x, xb);
    } else if (!c.dry) {
      const float* px = c.ptr<float>(x);
      bf16* pb = c.ptr<bf16>(xb);
      const size_t n = static_cast<size_t>(x.M) * C;
      Op op;
      op.cls = CLS_OTHER;
      op.launches = 1;
      op.flops = 0;
      op.bytes = static_cast<double>(n) * 6;
      op.run = [=](cudaStream_t s) { return launch_cast_bf16(px, n, pb, s); };
      c.ops->push_back(op);
    }
    if (!c.dry) {
      for (int phase = 0; phase < 4; ++phase) {
        GemmPlan p;
        ConvGeom g{c.n_img, H, W, 4, 1};
        g.up_phase = phase;
        if (!make_conv_plan(&p, c.ptr<bf16>(xb), g, k * C, nullptr, 0, w.w + static_cast<size_t>(phase) * w.cout * 4 * k * C,
                            w.cout, OUT_F32, c.ptr<float>(*out), w.cout, w.b, nullptr, 1, 0, nullptr, 0))
          return false;
        const double executed = p.flops;                            // 4 taps on the low-resolution grid
        p.flops = 2.0 * x.M * static_cast<double>(w.cout) * 9 * C * k;  // algorithmic: the un-folded conv's share
        add_gemm_op(c, CLS_CONV, p, false, executed);
      }
    }
    c.release(xb);
    return true;
  }

  // After the last cross-view ("3d") transformer every remaining layer treats the views independently,
  // and the reference views' outputs are discarded by the output mix (mmdm_unet.py:122-125: x - z_input
  // where ref_mask == 1).  With R promised reference views the rest of the network therefore runs on the
  // generated views only: one gather of the activations here, skip tensors are read in place by GroupNorm.
  bool plan_compact(PlanCtx& c, Buf* cur, int hw) {
    const int C = cur->C;
    Buf dst = c.alloc(c.B * c.G * hw, C, 4);
    if (!c.dry) {
      const float* src = c.ptr<float>(*cur);
      float* d = c.ptr<float>(dst);
      const int B = c.B, V = c.V, R = c.R;
      const size_t per_img = static_cast<size_t>(hw) * C;
      Op op;
      op.cls = CLS_OTHER;
      op.launches = 1;
      op.flops = 0;
      op.bytes = static_cast<double>(dst.M) * C * 8;
      op.run = [=](cudaStream_t s) { return launch_gather_views(src, d, B, V, R, per_img, s); };
      c.ops->push_back(op);
    }
    c.release(*cur);
    *cur = dst;
    c.compact = true;
    c.n_img = c.B * c.G;
    c.emb_all = c.emb_gen;
    return true;
  }

  void add_tap(PlanCtx& c, const std::string& name, const Buf& cur) {
    if (!taps_on) return;
    Buf slot = c.alloc(cur.M, cur.C, 4);  // never released: lives until the next forward
    if (c.dry) return;
    const float* src = c.ptr<float>(cur);
    float* dst = c.ptr<float>(slot);
    const size_t bytes = cur.bytes;
    taps.push_back({name, dst, cur.M, cur.C, c.n_img});
    Op op;
    op.cls = CLS_OTHER;
    op.launches = 0;
    op.flops = 0;
    op.bytes = 2.0 * bytes;
    op.run = [=](cudaStream_t s) { return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s); };
    c.ops->push_back(op);
  }

  bool plan_block(PlanCtx& c, const Block& blk, Buf x, const Buf* skip, int* H, int* W, Buf* out,
                  int compact_after = -1) {
    // x is owned by this call (released once consumed); skip is owned by the caller
    Buf cur = x;
    for (size_t li = 0; li < blk.size(); ++li) {
      const Layer& l = blk[li];
      Buf nxt;
      switch (l.kind) {
        case L_RES:
          if (!plan_res(c, res[l.idx], cur, (li == 0) ? skip : nullptr, *H, *W, &nxt)) return false;
          break;
        case L_TF:
          if (!plan_tf(c, tf[l.idx], cur, *H, *W, &nxt)) return false;
          break;
        case L_DOWN:
          if (!plan_down(c, down[l.idx], cur, *H, *W, &nxt)) return false;
          *H /= 2;
          *W /= 2;
          break;
        case L_UP:
          if (!plan_up(c, up[l.idx], cur, *H, *W, &nxt)) return false;
          *H *= 2;
          *W *= 2;
          break;
      }
      c.release(cur);
      cur = nxt;
      if (static_cast<int>(li) == compact_after && !plan_compact(c, &cur, *H * *W)) return false;
    }
    *out = cur;
    return true;
  }

  // the (block, layer) of the up path's last cross-view transformer; false if a later stage has none
  bool last_3d_in_up_path(int* block, int* layer) const {
    *block = *layer = -1;
    for (size_t bi = 0; bi < output_blocks.size(); ++bi)
      for (size_t li = 0; li < output_blocks[bi].size(); ++li)
        if (output_blocks[bi][li].kind == L_TF && tf[output_blocks[bi][li].idx].is3d) {
          *block = static_cast<int>(bi);
          *layer = static_cast<int>(li);
        }
    return *block >= 0;
  }

  bool build_plan(PlanCtx& c, int B, int V, int H, int W, int R) {
    const int mc = cfg.model_channels;
    const int n_img = B * V;
    c.n_img = n_img;
    c.V = V;
    c.B = B;
    int cb = -1, cl = -1;
    if (R < 0 || R >= V) {
      set_error("n_ref_views must be in [0, V)");
      return false;
    }
    // the "3d" transformers regroup '(b t) n c -> b (t n) c' with t = time_steps (attention.py:233): with another
    // V the reference attends over different token sets than one sequence per group
    for (const TfW& w : tf)
      if (w.is3d && V != cfg.time_steps) {
        set_error("V = " + std::to_string(V) + " views per group, but the model's cross-view attention was built for "
                  "time_steps = " + std::to_string(cfg.time_steps));
        return false;
      }
    if (R > 0 && last_3d_in_up_path(&cb, &cl)) {
      c.R = R;
      c.G = V - R;
    }
    const int down_factor = 1 << (cfg.n_levels - 1);
    if (H % down_factor != 0 || W % down_factor != 0) {
      set_error("H and W must be divisible by 2^(n_levels-1)");
      return false;
    }
    // persistent small buffers
    Buf emb = c.alloc(n_img, n_all, 4);
    Buf te_scratch;
    te_scratch.bytes = time_embed_scratch_bytes(n_img, mc, emb_ch);
    te_scratch.off = c.arena.alloc(te_scratch.bytes);
    te_scratch.valid = true;
    Buf embg;
    if (c.R > 0) embg = c.alloc(B * c.G, n_all, 4);
    Buf gnp;
    gnp.bytes = groupnorm_partial_bytes(n_img);
    gnp.off = c.arena.alloc(gnp.bytes);
    gnp.valid = true;
    if (!c.dry) {
      c.emb_all = c.ptr<float>(emb);
      c.emb_gen = (c.R > 0) ? c.ptr<float>(embg) : nullptr;
      c.gn_partial = c.ptr<float>(gnp);
      // The GroupNorm kernel's per-image counters must be zero when a forward starts.  Every launch leaves them
      // zeroed, but a plan can be parked while its workspace is reused by the caller (or a launch can have been
      // aborted), so they are cleared on the caller's stream at the head of every forward (a few hundred bytes).
      {
        uint8_t* sync_ptr = c.base + gnp.off + groupnorm_sync_offset(n_img);
        const size_t sync_bytes = gnp.bytes - groupnorm_sync_offset(n_img);
        Op op;
        op.cls = CLS_OTHER;
        op.launches = 0;
        op.flops = 0;
        op.bytes = static_cast<double>(sync_bytes);
        op.run = [=](cudaStream_t s) { return cudaMemsetAsync(sync_ptr, 0, sync_bytes, s); };
        c.ops->push_back(op);
      }
    }
    const int M0 = n_img * H * W;
    // ---- input stage
    Buf a0 = c.alloc(M0, kx() * kpad_in, 2), a32;
    if (precise) a32 = c.alloc(M0, kpad_in, 4);
    Buf h = c.alloc(M0, mc, 4);
    if (!c.dry) {
      IoPtrs* iop = &io;
      bf16* pa0 = precise ? reinterpret_cast<bf16*>(c.ptr<float>(a32)) : c.ptr<bf16>(a0);
      const int cin = cfg.in_channels, cc = cfg.condition_channels, kp = kpad_in, in_f16 = precise ? 2 : (f16_ok ? 1 : 0);
      {
        Op op;
        op.cls = CLS_OTHER;
        op.launches = 1;
        op.flops = 0;
        op.bytes = static_cast<double>(M0) * (kp * 2 + cc * 4 + cin * 8);
        op.run = [=](cudaStream_t s) {
          return launch_input_pack(iop->x, iop->z, iop->mask, iop->pos, n_img, cin, H, W, cc, kp, pa0, s, in_f16);
        };
        c.ops->push_back(op);
      }
      {
        float* scratch = c.ptr<float>(te_scratch);
        float* embp = c.emb_all;
        const int nall = n_all, ech = emb_ch;
        const float *w1 = te_w1, *b1 = te_b1, *w2 = te_w2, *b2 = te_b2, *wa = wall, *ba = ball;
        Op op;
        op.cls = CLS_OTHER;
        op.launches = 6;
        op.flops = 2.0 * n_img * (static_cast<double>(mc) * ech + static_cast<double>(ech) * ech +
                                  static_cast<double>(ech) * nall);
        op.bytes = 4.0 * (static_cast<double>(mc) * ech + static_cast<double>(ech) * ech + static_cast<double>(ech) * nall);
        op.run = [=](cudaStream_t s) {
          return launch_time_embed(iop->t, n_img, mc, ech, w1, b1, w2, b2, wa, ba, nall, scratch, embp, s);
        };
        c.ops->push_back(op);
        if (c.R > 0) {  // the generated views' rows, for the ResBlocks planned after the compaction
          float* eg = c.emb_gen;
          const int Bv = B, Vv = V, Rv = c.R;
          Op og;
          og.cls = CLS_OTHER;
          og.launches = 1;
          og.flops = 0;
          og.bytes = static_cast<double>(B) * c.G * nall * 8;
          og.run = [=](cudaStream_t s) { return launch_gather_views(embp, eg, Bv, Vv, Rv, nall, s); };
          c.ops->push_back(og);
        }
      }
    }
    if (precise) {
      op_split(c, a32, a0);
      c.release(a32);
    }
    if (!c.dry) {
      GemmPlan p;
      if (!make_gemm_plan(&p, c.ptr<bf16>(a0), M0, kx() * kpad_in, nullptr, 0, w_in, mc, OUT_F32, c.ptr<float>(h), mc, b_in,
                          nullptr, 1, 0, nullptr, 0))
        return false;
      add_gemm_op(c, CLS_LINEAR, p, true);
    }
    c.release(a0);
    // ---- down path
    std::vector<Buf> hs;
    hs.push_back(h);
    int curH = H, curW = W;
    Buf cur = h;
    if (!c.dry) taps.clear();
    add_tap(c, "input_blocks.0", h);
    for (size_t bi = 1; bi < input_blocks.size(); ++bi) {
      Buf nxt;
      // the block input is also a skip tensor: do not let plan_block release it
      Buf keep = cur;
      keep.valid = false;
      if (!plan_block(c, input_blocks[bi], keep, nullptr, &curH, &curW, &nxt)) return false;
      hs.push_back(nxt);
      cur = nxt;
      add_tap(c, "input_blocks." + std::to_string(bi), cur);
    }
    // ---- middle
    {
      Buf keep = cur;
      keep.valid = false;
      Buf nxt;
      if (!plan_block(c, middle, keep, nullptr, &curH, &curW, &nxt)) return false;
      cur = nxt;  // owned
      add_tap(c, "middle_block", cur);
    }
    // ---- up path
    for (size_t bi = 0; bi < output_blocks.size(); ++bi) {
      Buf skip = hs.back();
      hs.pop_back();
      Buf nxt;
      const int compact_after = (c.R > 0 && static_cast<int>(bi) == cb) ? cl : -1;
      if (!plan_block(c, output_blocks[bi], cur, &skip, &curH, &curW, &nxt, compact_after)) return false;
      c.release(skip);
      cur = nxt;
      add_tap(c, "output_blocks." + std::to_string(bi), cur);
    }
    // ---- out (on the generated views only once compact)
    const int Mo = c.n_img * H * W;
    Buf a = c.alloc(Mo, kx() * mc, 2);
    if (!op_gn(c, cur, nullptr, H * W, out_gn_g, out_gn_b, 1e-5f, 1, a, nullptr, true, true)) return false;
    c.release(cur);
    Buf o32 = c.alloc(Mo, 32, 4);
    if (precise) {
      ConvGeom g{c.n_img, H, W, 9, 1};
      if (!conv51(c, g, a, mc, nullptr, 0, w_out, 32, o32, b_out, nullptr, 1, 0, nullptr)) return false;
    }
    if (!c.dry) {
      if (!precise) {
        GemmPlan p;
        ConvGeom g{c.n_img, H, W, 9, 1};
        if (!make_conv_plan(&p, c.ptr<bf16>(a), g, mc, nullptr, 0, w_out, 32, OUT_F32, c.ptr<float>(o32), 32, b_out,
                            nullptr, 1, 0, nullptr, 0))
          return false;
        add_gemm_op(c, CLS_CONV, p, true);
      }
      IoPtrs* iop = &io;
      const float* po = c.ptr<float>(o32);
      const int cout = cfg.out_channels;
      const int mixG = c.compact ? c.G : 0, mixR = c.R;
      int* viol = d_violations;
      Op op;
      op.cls = CLS_OTHER;
      op.launches = 1;
      op.flops = 0;
      op.bytes = static_cast<double>(M0) * cout * 16;
      op.run = [=](cudaStream_t s) {
        return launch_output_mix(po, 32, iop->x, iop->z, iop->mask, n_img, cout, H, W, mixG, V, mixR, iop->out, s, viol);
      };
      c.ops->push_back(op);
    }
    c.release(a);
    c.release(o32);
    return true;
  }

  bool workspace_bytes(int B, int V, int H, int W, size_t* bytes) {
    PlanCtx c;
    c.dry = true;
    std::vector<Op> dummy;
    c.ops = &dummy;
    if (!build_plan(c, B, V, H, W, ref_views)) return false;
    *bytes = c.arena.peak + 1024;
    return true;
  }

  bool ensure_plan(int B, int V, int H, int W, void* ws, size_t ws_bytes) {
    if (!finalized) {
      set_error("cap4d_b200_unet_finalize has not been called");
      return false;
    }
    const int R = ref_views;
    if (B == pB && V == pV && H == pH && W == pW && R == pR && ws == p_ws && ws_bytes == p_ws_bytes && !ops.empty())
      return true;
    // park the current plan and look for a cached one (a sampler alternates between at most a few batch
    // shapes, e.g. 4 groups per call and a 3-group remainder; each shape has its own workspace)
    if (!ops.empty()) {
      if (cache.size() >= 4) cache.erase(cache.begin());
      cache.emplace_back();
      CachedPlan& cp = cache.back();
      cp.B = pB, cp.V = pV, cp.H = pH, cp.W = pW, cp.R = pR, cp.ws = p_ws, cp.ws_bytes = p_ws_bytes;
      cp.ops.swap(ops);
    }
    for (size_t i = 0; i < cache.size(); ++i) {
      CachedPlan& cp = cache[i];
      if (cp.B == B && cp.V == V && cp.H == H && cp.W == W && cp.R == R && cp.ws == ws && cp.ws_bytes == ws_bytes) {
        ops.swap(cp.ops);
        pB = B, pV = V, pH = H, pW = W, pR = R, p_ws = ws, p_ws_bytes = ws_bytes;
        cache.erase(cache.begin() + i);
        return true;
      }
    }
    size_t need = 0;
    if (!workspace_bytes(B, V, H, W, &need)) return false;
    if (ws == nullptr || ws_bytes < need) {
      set_error("workspace too small: need " + std::to_string(need) + " bytes");
      return false;
    }
    ops.clear();
    PlanCtx c;
    c.dry = false;
    c.base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~static_cast<uintptr_t>(1023));
    c.ops = &ops;
    if (!build_plan(c, B, V, H, W, R)) {
      ops.clear();
      return false;
    }
    pR = R;
    pB = B;
    pV = V;
    pH = H;
    pW = W;
    p_ws = ws;
    p_ws_bytes = ws_bytes;
    return true;
  }

  // timed == 0: plain; 1: events around every op, synchronise and return per-class ms;
  // 2: events recorded but NOT synchronised (collected later by collect_timings)
  bool forward(const IoPtrs& ptrs, int B, int V, int H, int W, void* ws, size_t ws_bytes, cudaStream_t stream,
               int timed, float* class_ms) {
    if (!ensure_plan(B, V, H, W, ws, ws_bytes)) return false;
    io = ptrs;
    std::vector<cudaEvent_t>* evs = nullptr;
    if (timed == 1) {
      evs = &events;
    } else if (timed == 2) {
      if (!event_pool.empty()) {
        deferred.push_back(std::move(event_pool.back()));
        event_pool.pop_back();
      } else {
        deferred.emplace_back();
      }
      evs = &deferred.back();
      deferred_cls.emplace_back();
      for (const Op& op : ops) deferred_cls.back().push_back(op.cls);
    }
    if (evs != nullptr) {
      while (evs->size() < ops.size() + 1) {
        cudaEvent_t ev;
        CUDA_OK(cudaEventCreate(&ev));
        evs->push_back(ev);
      }
      CUDA_OK(cudaEventRecord((*evs)[0], stream));
    }
    for (size_t i = 0; i < ops.size(); ++i) {
      cudaError_t e = ops[i].run(stream);
      if (e != cudaSuccess) {
        set_error("launch of op " + std::to_string(i) + " failed: " + cudaGetErrorString(e) + " / " + get_error());
        return false;
      }
      if (evs != nullptr) CUDA_OK(cudaEventRecord((*evs)[i + 1], stream));
    }
    if (timed == 1) {
      CUDA_OK(cudaStreamSynchronize(stream));
      for (int k = 0; k < CAP4D_B200_N_CLASSES; ++k) class_ms[k] = 0.f;
      if (!accumulate(events, class_ms)) return false;
    }
    return true;
  }

  bool accumulate(const std::vector<cudaEvent_t>& evs, float* class_ms) {
    for (size_t i = 0; i < ops.size(); ++i) {
      float ms = 0.f;
      CUDA_OK(cudaEventElapsedTime(&ms, evs[i], evs[i + 1]));
      class_ms[ops[i].cls] += ms;
    }
    return true;
  }

  bool collect_timings(float* class_ms, int* n_runs) {
    for (int k = 0; k < CAP4D_B200_N_CLASSES; ++k) class_ms[k] = 0.f;
    *n_runs = 0;
    for (size_t r = 0; r < deferred.size(); ++r) {
      const std::vector<cudaEvent_t>& evs = deferred[r];
      const std::vector<int>& cls = deferred_cls[r];
      if (evs.size() < cls.size() + 1) continue;
      CUDA_OK(cudaEventSynchronize(evs[cls.size()]));
      for (size_t i = 0; i < cls.size(); ++i) {
        float ms = 0.f;
        CUDA_OK(cudaEventElapsedTime(&ms, evs[i], evs[i + 1]));
        class_ms[cls[i]] += ms;
      }
      ++*n_runs;
    }
    for (auto& evs : deferred) event_pool.push_back(std::move(evs));
    deferred.clear();
    deferred_cls.clear();
    return true;
  }
};

}  // namespace
}  // namespace cap4d

// =============================================================================================
// C ABI
// =============================================================================================
using namespace cap4d;

// timing helper for the single-kernel entry points
template <typename F>
static int run_timed(F&& launch, cudaStream_t s, float* ms_out, int iters) {
  if (iters < 1) iters = 1;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ms_out != nullptr) {
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaError_t e = launch(s);  // warm-up
    if (e != cudaSuccess) {
      set_error(std::string("launch failed: ") + cudaGetErrorString(e) + " / " + get_error());
      return 8;
    }
    cudaEventRecord(e0, s);
  }
  for (int i = 0; i < iters; ++i) {
    cudaError_t e = launch(s);
    if (e != cudaSuccess) {
      set_error(std::string("launch failed: ") + cudaGetErrorString(e) + " / " + get_error());
      return 8;
    }
  }
  if (ms_out != nullptr) {
    cudaEventRecord(e1, s);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) {
      set_error(std::string("kernel failed: ") + cudaGetErrorString(e));
      return 9;
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    *ms_out = ms / iters;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }
  return 0;
}

extern "C" {

const char* cap4d_b200_last_error(void) { return get_error(); }
const char* cap4d_b200_version(void) { return "cap4d_b200 0.1 (sm_100a)"; }

int cap4d_b200_unet_create(const cap4d_b200_unet_config* cfg, void** handle) {
  if (cfg == nullptr || handle == nullptr) {
    set_error("null argument");
    return 1;
  }
  Unet* u = new Unet();
  u->cfg = *cfg;
  if (!u->build_topology()) {
    delete u;
    return 2;
  }
  *handle = u;
  return 0;
}

int cap4d_b200_unet_load_weight(void* handle, const char* name, const float* data, const int64_t* shape, int ndim) {
  Unet* u = static_cast<Unet*>(handle);
  if (u == nullptr || name == nullptr || data == nullptr) {
    set_error("null argument");
    return 1;
  }
  if (u->finalized) {
    set_error("weights are already finalized");
    return 1;
  }
  RawTensor t;
  t.numel = 1;
  for (int i = 0; i < ndim; ++i) {
    t.shape.push_back(shape[i]);
    t.numel *= static_cast<size_t>(shape[i]);
  }
  if (cudaMalloc(reinterpret_cast<void**>(&t.d), std::max<size_t>(t.numel * sizeof(float), 16)) != cudaSuccess) {
    set_error(std::string("cudaMalloc failed for ") + name);
    return 3;
  }
  cudaError_t e = cudaMemcpy(t.d, data, t.numel * sizeof(float), cudaMemcpyDefault);
  if (e != cudaSuccess) {
    cudaFree(t.d);
    set_error(std::string("cudaMemcpy failed for ") + name + ": " + cudaGetErrorString(e));
    return 3;
  }
  auto it = u->raw.find(name);
  if (it != u->raw.end() && it->second.d) cudaFree(it->second.d);
  u->raw[name] = t;
  return 0;
}

int cap4d_b200_unet_finalize(void* handle) {
  Unet* u = static_cast<Unet*>(handle);
  if (u == nullptr) {
    set_error("null handle");
    return 1;
  }
  return u->finalize() ? 0 : 4;
}

int cap4d_b200_unet_workspace_bytes(void* handle, int B, int V, int H, int W, size_t* bytes) {
  Unet* u = static_cast<Unet*>(handle);
  if (u == nullptr || bytes == nullptr) {
    set_error("null argument");
    return 1;
  }
  return u->workspace_bytes(B, V, H, W, bytes) ? 0 : 5;
}

static int forward_impl(void* handle, const float* x, const int64_t* timesteps, const float* z_input,
                        const float* ref_mask, const float* pos_enc, float* out, int B, int V, int H, int W,
                        void* workspace, size_t workspace_bytes, void* stream, int timed, float* class_ms) {
  Unet* u = static_cast<Unet*>(handle);
  if (u == nullptr || x == nullptr || timesteps == nullptr || z_input == nullptr || ref_mask == nullptr ||
      pos_enc == nullptr || out == nullptr) {
    set_error("null argument");
    return 1;
  }
  IoPtrs io;
  io.x = x;
  io.t = reinterpret_cast<const long long*>(timesteps);
  io.z = z_input;
  io.mask = ref_mask;
  io.pos = pos_enc;
  io.out = out;
  return u->forward(io, B, V, H, W, workspace, workspace_bytes, static_cast<cudaStream_t>(stream), timed, class_ms)
             ? 0
             : 6;
}

int cap4d_b200_unet_set_ref_views(void* handle, int n_ref_views) {
  Unet* u = static_cast<Unet*>(handle);
  if (u == nullptr || n_ref_views < 0) {
    set_error("set_ref_views: null handle or negative count");
    return 1;
  }
  u->ref_views = n_ref_views;
  return 0;
}

int cap4d_b200_unet_set_precision(void* handle, int fp32_accuracy) {
  Unet* u = static_cast<Unet*>(handle);
  if (u == nullptr) {
    set_error("null handle");
    return 1;
  }
  if (u->finalized) {
    set_error("set_precision must be called before finalize: it decides how the weights are packed");
    return 1;
  }
  u->precise = fp32_accuracy != 0;
  return 0;
}

int cap4d_b200_unet_enable_taps(void* handle, int on) {
  Unet* u = static_cast<Unet*>(handle);
  if (u == nullptr) {
    set_error("null handle");
    return 1;
  }
  if (u->taps_on != (on != 0)) {  // the taps are part of the launch plan: drop every plan
    u->taps_on = on != 0;
    u->ops.clear();
    u->cache.clear();
    u->taps.clear();
    u->pB = u->pV = u->pH = u->pW = 0;
  }
  return 0;
}

int cap4d_b200_unet_num_taps(void* handle, int* n) {
  Unet* u = static_cast<Unet*>(handle);
  if (u == nullptr || n == nullptr) {
    set_error("null argument");
    return 1;
  }
  *n = static_cast<int>(u->taps.size());
  return 0;
}

int cap4d_b200_unet_tap_info(void* handle, int index, char* name, int name_capacity, const float** data, int64_t* rows,
                             int* channels, int* n_img) {
  Unet* u = static_cast<Unet*>(handle);
  if (u == nullptr || name == nullptr || data == nullptr || rows == nullptr || channels == nullptr || n_img == nullptr ||
      index < 0 || index >= static_cast<int>(u->taps.size())) {
    set_error("tap_info: bad argument");
    return 1;
  }
  const Unet::TapInfo& t = u->taps[index];
  if (static_cast<int>(t.name.size()) + 1 > name_capacity) {
    set_error("name buffer too small");
    return 1;
  }
  std::strcpy(name, t.name.c_str());
  *data = t.ptr;
  *rows = t.rows;
  *channels = t.C;
  *n_img = t.n_img;
  return 0;
}

int cap4d_b200_unet_ref_view_violations(void* handle, int* n) {
  Unet* u = static_cast<Unet*>(handle);
  if (u == nullptr || n == nullptr) {
    set_error("null argument");
    return 1;
  }
  *n = 0;
  if (u->d_violations == nullptr) return 0;
  // blocking read on the legacy default stream: it waits for the caller's (blocking) streams
  cudaError_t e = cudaMemcpy(n, u->d_violations, sizeof(int), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemset(u->d_violations, 0, sizeof(int));
  if (e != cudaSuccess) {
    set_error(cudaGetErrorString(e));
    return 6;
  }
  return 0;
}

int cap4d_b200_unet_forward(void* handle, const float* x, const int64_t* timesteps, const float* z_input,
                            const float* ref_mask, const float* pos_enc, float* out, int B, int V, int H, int W,
                            void* workspace, size_t workspace_bytes, void* stream) {
  return forward_impl(handle, x, timesteps, z_input, ref_mask, pos_enc, out, B, V, H, W, workspace, workspace_bytes,
                      stream, 0, nullptr);
}

int cap4d_b200_unet_forward_timed(void* handle, const float* x, const int64_t* timesteps, const float* z_input,
                                  const float* ref_mask, const float* pos_enc, float* out, int B, int V, int H,
                                  int W, void* workspace, size_t workspace_bytes, void* stream, float* class_ms) {
  return forward_impl(handle, x, timesteps, z_input, ref_mask, pos_enc, out, B, V, H, W, workspace, workspace_bytes,
                      stream, class_ms != nullptr ? 1 : 2, class_ms);
}

int cap4d_b200_unet_collect_timings(void* handle, float* class_ms, int* n_runs) {
  Unet* u = static_cast<Unet*>(handle);
  if (u == nullptr || class_ms == nullptr || n_runs == nullptr) {
    set_error("null argument");
    return 1;
  }
  return u->collect_timings(class_ms, n_runs) ? 0 : 6;
}

int cap4d_b200_unet_plan(void* handle, int B, int V, int H, int W, void* workspace, size_t workspace_bytes) {
  Unet* u = static_cast<Unet*>(handle);
  if (u == nullptr) {
    set_error("null handle");
    return 1;
  }
  return u->ensure_plan(B, V, H, W, workspace, workspace_bytes) ? 0 : 6;
}

int cap4d_b200_unet_num_launches(void* handle, int* n) {
  Unet* u = static_cast<Unet*>(handle);
  if (u == nullptr || n == nullptr) {
    set_error("null argument");
    return 1;
  }
  int total = 0;
  for (const Op& op : u->ops) total += op.launches;
  *n = total;
  return 0;
}

int cap4d_b200_unet_class_stats(void* handle, double* flops, double* bytes, int* launches) {
  Unet* u = static_cast<Unet*>(handle);
  if (u == nullptr) {
    set_error("null handle");
    return 1;
  }
  for (int k = 0; k < CAP4D_B200_N_CLASSES; ++k) {
    if (flops) flops[k] = 0;
    if (bytes) bytes[k] = 0;
    if (launches) launches[k] = 0;
  }
  for (const Op& op : u->ops) {
    if (flops) flops[op.cls] += op.flops;
    if (bytes) bytes[op.cls] += op.bytes;
    if (launches) launches[op.cls] += op.launches;
  }
  return 0;
}

int cap4d_b200_unet_class_exec_flops(void* handle, double* flops) {
  Unet* u = static_cast<Unet*>(handle);
  if (u == nullptr || flops == nullptr) {
    set_error("null argument");
    return 1;
  }
  for (int k = 0; k < CAP4D_B200_N_CLASSES; ++k) flops[k] = 0;
  for (const Op& op : u->ops) flops[op.cls] += (op.exec_flops >= 0 ? op.exec_flops : op.flops);
  return 0;
}

int cap4d_b200_unet_num_params(void* handle, int* n) {
  Unet* u = static_cast<Unet*>(handle);
  if (u == nullptr || n == nullptr) {
    set_error("null argument");
    return 1;
  }
  *n = static_cast<int>(u->expected_params().size());
  return 0;
}

int cap4d_b200_unet_param_info(void* handle, int index, char* name, int name_capacity, int64_t* shape, int* ndim) {
  Unet* u = static_cast<Unet*>(handle);
  if (u == nullptr || name == nullptr || shape == nullptr || ndim == nullptr) {
    set_error("null argument");
    return 1;
  }
  auto v = u->expected_params();
  if (index < 0 || index >= static_cast<int>(v.size())) {
    set_error("parameter index out of range");
    return 1;
  }
  if (static_cast<int>(v[index].first.size()) + 1 > name_capacity) {
    set_error("name buffer too small");
    return 1;
  }
  std::strcpy(name, v[index].first.c_str());
  *ndim = static_cast<int>(v[index].second.size());
  for (int i = 0; i < *ndim; ++i) shape[i] = v[index].second[i];
  return 0;
}

int cap4d_b200_unet_destroy(void* handle) {
  delete static_cast<Unet*>(handle);
  return 0;
}

int cap4d_b200_cfg_ddim_update(float* latents, const float* eps, const int64_t* gen_idx, int n_groups, int V, int R,
                               int chw, float cfg_scale, float x_coef, float e_coef, void* stream) {
  cudaError_t e = launch_cfg_ddim_update(latents, eps, reinterpret_cast<const long long*>(gen_idx), n_groups, V, R, chw,
                                         cfg_scale, x_coef, e_coef, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) {
    if (e != cudaErrorInvalidValue) set_error(cudaGetErrorString(e));
    return 7;
  }
  return 0;
}

// ---- single-kernel entry points ---------------------------------------------------------------
int cap4d_b200_gemm_bf16(const uint16_t* A, const uint16_t* Wt, int M, int N, int K, const float* bias,
                         const float* residual, void* out, int out_mode, void* stream, float* ms_out, int iters) {
  GemmPlan p;
  const int ldo = (out_mode == OUT_GEGLU_BF16) ? N / 2 : N;
  if (!make_gemm_plan(&p, reinterpret_cast<const bf16*>(A), M, K, nullptr, 0, reinterpret_cast<const bf16*>(Wt), N,
                      out_mode, out, ldo, bias, nullptr, 1, 0, residual, ldo))
    return 10;
  return run_timed([&](cudaStream_t s) { return launch_gemm(p, s); }, static_cast<cudaStream_t>(stream), ms_out, iters);
}

int cap4d_b200_gemm_mixed(const uint16_t* A, const uint16_t* Wt, int M, int N, int K, int a_f16, int b_f16,
                          const float* bias, const float* residual, void* out, int out_mode, void* stream,
                          float* ms_out, int iters) {
  GemmPlan p;
  const int ldo = (out_mode == OUT_GEGLU_BF16) ? N / 2 : N;
  if (!make_gemm_plan(&p, reinterpret_cast<const bf16*>(A), M, K, nullptr, 0, reinterpret_cast<const bf16*>(Wt), N,
                      out_mode, out, ldo, bias, nullptr, 1, 0, residual, ldo))
    return 10;
  p.p.a_f16 = a_f16 ? 1 : 0;
  p.p.b_f16 = b_f16 ? 1 : 0;
  return run_timed([&](cudaStream_t s) { return launch_gemm(p, s); }, static_cast<cudaStream_t>(stream), ms_out, iters);
}

int cap4d_b200_conv3x3_bf16(const uint16_t* A, const uint16_t* Wt, int n_img, int H_out, int W_out, int Cin, int Cout,
                            int stride, const float* bias, const float* rowbias, const float* residual, float* out,
                            void* stream, float* ms_out, int iters) {
  GemmPlan p;
  ConvGeom g{n_img, H_out, W_out, 9, stride};
  if (!make_conv_plan(&p, reinterpret_cast<const bf16*>(A), g, Cin, nullptr, 0, reinterpret_cast<const bf16*>(Wt), Cout,
                      OUT_F32, out, Cout, bias, rowbias, H_out * W_out, Cout, residual, Cout))
    return 10;
  return run_timed([&](cudaStream_t s) { return launch_gemm(p, s); }, static_cast<cudaStream_t>(stream), ms_out, iters);
}

int cap4d_b200_upsample_conv3x3_bf16(const uint16_t* A, const float* w_oihw, int n_img, int H, int W, int Cin, int Cout,
                                     const float* bias, float* out, void* stream, float* ms_out, int iters) {
  bf16* wp = nullptr;
  if (cudaMalloc(reinterpret_cast<void**>(&wp), static_cast<size_t>(16) * Cout * Cin * sizeof(bf16)) != cudaSuccess) {
    set_error("cudaMalloc failed");
    return 3;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = launch_pack_upconv_weight(w_oihw, Cout, Cin, wp, s);
  GemmPlan plans[4];
  int rc = 0;
  for (int phase = 0; phase < 4 && rc == 0; ++phase) {
    ConvGeom g{n_img, H, W, 4, 1};
    g.up_phase = phase;
    if (!make_conv_plan(&plans[phase], reinterpret_cast<const bf16*>(A), g, Cin, nullptr, 0,
                        wp + static_cast<size_t>(phase) * Cout * 4 * Cin, Cout, OUT_F32, out, Cout, bias, nullptr, 1, 0,
                        nullptr, 0))
      rc = 10;
  }
  if (rc == 0 && e == cudaSuccess) {
    rc = run_timed(
        [&](cudaStream_t st) {
          for (int phase = 0; phase < 4; ++phase) {
            cudaError_t le = launch_gemm(plans[phase], st);
            if (le != cudaSuccess) return le;
          }
          return cudaSuccess;
        },
        s, ms_out, iters);
  } else if (rc == 0) {
    set_error(cudaGetErrorString(e));
    rc = 8;
  }
  cudaStreamSynchronize(s);
  cudaFree(wp);
  return rc;
}

int cap4d_b200_attention_bf16(const uint16_t* qkv, uint16_t* out, int M, int C, int L, float scale, void* stream,
                              float* ms_out, int iters) {
  AttnPlan p;
  if (!make_attn_plan(&p, reinterpret_cast<const bf16*>(qkv), reinterpret_cast<bf16*>(out), M, C, L, scale)) return 10;
  return run_timed([&](cudaStream_t s) { return launch_attn(p, s); }, static_cast<cudaStream_t>(stream), ms_out, iters);
}

int cap4d_b200_attention_trace(const uint16_t* qkv, uint16_t* out, int M, int C, int L, float scale, void* stream,
                               long long* trace) {
  AttnPlan p;
  if (!make_attn_plan(&p, reinterpret_cast<const bf16*>(qkv), reinterpret_cast<bf16*>(out), M, C, L, scale)) return 10;
  cudaError_t e = launch_attn(p, static_cast<cudaStream_t>(stream), trace);
  if (e != cudaSuccess) {
    set_error(cudaGetErrorString(e));
    return 8;
  }
  return 0;
}

int cap4d_b200_groupnorm_bf16(const float* x1, int C1, const float* x2, int C2, int n_img, int hw,
                              const float* gamma, const float* beta, float eps, int apply_silu, uint16_t* out,
                              uint16_t* raw_out, void* stream, float* ms_out, int iters) {
  float* partial = nullptr;
  if (cudaMalloc(reinterpret_cast<void**>(&partial), groupnorm_partial_bytes(n_img)) != cudaSuccess) {
    set_error("cudaMalloc failed");
    return 3;
  }
  cudaMemset(partial, 0, groupnorm_partial_bytes(n_img));
  int rc = run_timed(
      [&](cudaStream_t s) {
        return launch_groupnorm(x1, C1, x2, C2, n_img, hw, gamma, beta, eps, apply_silu, reinterpret_cast<bf16*>(out),
                                reinterpret_cast<bf16*>(raw_out), partial, s);
      },
      static_cast<cudaStream_t>(stream), ms_out, iters);
  cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
  cudaFree(partial);
  return rc;
}

int cap4d_b200_layernorm_bf16(const float* x, int M, int C, const float* gamma, const float* beta, float eps,
                              uint16_t* out, void* stream, float* ms_out, int iters) {
  return run_timed(
      [&](cudaStream_t s) { return launch_layernorm(x, M, C, gamma, beta, eps, reinterpret_cast<bf16*>(out), s); },
      static_cast<cudaStream_t>(stream), ms_out, iters);
}

}  // extern "C"

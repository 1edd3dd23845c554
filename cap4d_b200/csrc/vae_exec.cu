// VAE decoder executor (the step on the output side of the sampler).
//
// Mirrors decode_first_stage (controlnet/ldm/models/diffusion/ddpm.py:822-830: z / scale_factor) ->
// AutoencoderKL.decode (controlnet/ldm/models/autoencoder.py:87-91: post_quant_conv, decoder) -> Decoder.forward
// (controlnet/ldm/modules/diffusionmodules/model.py:606-640) with ResnetBlock (:125-146, temb = None),
// AttnBlock (:176-203) and Upsample (:71-76).  Same kernel family as the U-Net: GroupNorm+SiLU ->
// bf16 operand, implicit-GEMM conv3x3 with the 1x1 nin_shortcut appended along K, nearest-2x + conv3x3
// folded into four phase convs.  The single mid-block attention is one head over all C channels
// (C = 512, 4096 tokens): scores, softmax and P.V are three launches per image through a materialised
// fp32 score matrix (64 MB) - 3 % of the decoder's FLOPs, not worth a second flash kernel.
#include <cstdio>
#include <cstring>

#include "../../include/cap4d_b200.h"
#include "exec_common.h"
#include "kernels.h"

namespace cap4d {

namespace {

struct VRes {
  std::string prefix;
  int cin = 0, cout = 0;
  bool skip = false;
  float *gn1_g = nullptr, *gn1_b = nullptr, *gn2_g = nullptr, *gn2_b = nullptr;
  bf16 *w1 = nullptr, *w2 = nullptr;
  float *b1 = nullptr, *b2 = nullptr;
};
struct VAttn {
  std::string prefix;
  int C = 0;
  float *gn_g = nullptr, *gn_b = nullptr;
  bf16 *wqkv = nullptr, *wo = nullptr;
  float *bqkv = nullptr, *bo = nullptr;
};
struct VConv {
  std::string prefix;
  int cin = 0, cout = 0;
  bf16* w = nullptr;
  float* b = nullptr;
};
enum VKind { V_RES, V_ATTN, V_UP, V_DOWN };
struct VLayer {
  VKind kind;
  int idx;
};

struct Vae {
  cap4d_b200_vae_config cfg;
  std::vector<VRes> res;
  std::vector<VAttn> attn;
  std::vector<VConv> up;
  std::vector<VLayer> layers;
  int c_in = 0, c_last = 0, kpad_in = 64;
  std::map<std::string, RawTensor> raw;
  std::vector<void*> owned;
  bool finalized = false;
  bf16* w_in = nullptr;
  float *b_in = nullptr, *w_pq = nullptr, *b_pq = nullptr, *out_gn_g = nullptr, *out_gn_b = nullptr, *b_out = nullptr;
  bf16* w_out = nullptr;
  // encoder (model.py:452-545 + quant_conv, autoencoder.py:82-85); optional: built when its weights were loaded
  std::vector<VRes> eres;
  std::vector<VAttn> eattn;
  std::vector<VConv> down;
  std::vector<VLayer> elayers;
  bool has_encoder = false;
  int e_c_last = 0;
  bf16 *ew_in = nullptr, *ew_out = nullptr;
  float *eb_in = nullptr, *eb_out = nullptr, *e_gn_g = nullptr, *e_gn_b = nullptr, *e_ident = nullptr, *e_zero = nullptr;
  std::vector<Op> eops;
  const float* io_img = nullptr;
  float* io_moments = nullptr;
  int eN = 0, eH = 0, eW = 0;
  void* e_ws = nullptr;
  size_t e_ws_bytes = 0;
  // plan
  std::vector<Op> ops;
  const float* io_z = nullptr;
  float* io_out = nullptr;         // fp32 NCHW images, or
  unsigned char* io_out_u8 = nullptr;  // uint8 HWC BGR images (exactly one of the two is set per call)
  float io_inv_scale = 1.f;
  int pN = 0, pH = 0, pW = 0;
  void* p_ws = nullptr;
  size_t p_ws_bytes = 0;

  ~Vae() {
    for (void* p : owned) cudaFree(p);
    for (auto& kv : raw)
      if (kv.second.d) cudaFree(kv.second.d);
  }

  // ------------------------------------------------------------------ topology (model.py:546-604)
  bool build_topology() {
    if (cfg.n_levels < 1 || cfg.n_levels > 8 || cfg.ch % 32 != 0 || cfg.z_channels < 1 || 9 * cfg.z_channels > kpad_in ||
        cfg.out_ch < 1 || cfg.out_ch > 32 || cfg.num_res_blocks < 0) {
      set_error("vae: unsupported configuration (ch must be a multiple of 32, 9*z_channels <= 64, out_ch <= 32)");
      return false;
    }
    int block_in = cfg.ch * cfg.ch_mult[cfg.n_levels - 1];
    c_in = block_in;
    auto add_res = [&](const std::string& p, int cin, int cout) {
      VRes r;
      r.prefix = p;
      r.cin = cin;
      r.cout = cout;
      r.skip = cin != cout;
      res.push_back(r);
      layers.push_back({V_RES, static_cast<int>(res.size()) - 1});
    };
    add_res("decoder.mid.block_1.", block_in, block_in);
    {
      VAttn a;
      a.prefix = "decoder.mid.attn_1.";
      a.C = block_in;
      attn.push_back(a);
      layers.push_back({V_ATTN, 0});
    }
    add_res("decoder.mid.block_2.", block_in, block_in);
    for (int lvl = cfg.n_levels - 1; lvl >= 0; --lvl) {
      const int block_out = cfg.ch * cfg.ch_mult[lvl];
      if (block_out % 64 != 0 && block_out % 32 != 0) {
        set_error("vae: channel counts must be multiples of 32");
        return false;
      }
      for (int i = 0; i <= cfg.num_res_blocks; ++i) {
        add_res("decoder.up." + std::to_string(lvl) + ".block." + std::to_string(i) + ".", block_in, block_out);
        block_in = block_out;
      }
      if (lvl != 0) {
        VConv u;
        u.prefix = "decoder.up." + std::to_string(lvl) + ".upsample.conv.";
        u.cin = u.cout = block_in;
        up.push_back(u);
        layers.push_back({V_UP, static_cast<int>(up.size()) - 1});
      }
    }
    c_last = block_in;
    for (const VRes& r : res)
      if (r.cin % 64 != 0 || r.cout % 32 != 0) {
        set_error("vae: ResnetBlock channels must be multiples of 64 (in) / 32 (out) for the conv kernel");
        return false;
      }
    // Encoder.__init__ (model.py:466-516): attn_resolutions = [] : mid-block attention only
    int eb = cfg.ch;
    auto add_eres = [&](const std::string& p, int cin, int cout) {
      VRes r;
      r.prefix = p;
      r.cin = cin;
      r.cout = cout;
      r.skip = cin != cout;
      eres.push_back(r);
      elayers.push_back({V_RES, static_cast<int>(eres.size()) - 1});
    };
    for (int lvl = 0; lvl < cfg.n_levels; ++lvl) {
      const int block_out = cfg.ch * cfg.ch_mult[lvl];
      for (int i = 0; i < cfg.num_res_blocks; ++i) {
        add_eres("encoder.down." + std::to_string(lvl) + ".block." + std::to_string(i) + ".", eb, block_out);
        eb = block_out;
      }
      if (lvl != cfg.n_levels - 1) {
        VConv d;
        d.prefix = "encoder.down." + std::to_string(lvl) + ".downsample.conv.";
        d.cin = d.cout = eb;
        down.push_back(d);
        elayers.push_back({V_DOWN, static_cast<int>(down.size()) - 1});
      }
    }
    add_eres("encoder.mid.block_1.", eb, eb);
    {
      VAttn a;
      a.prefix = "encoder.mid.attn_1.";
      a.C = eb;
      eattn.push_back(a);
      elayers.push_back({V_ATTN, 0});
    }
    add_eres("encoder.mid.block_2.", eb, eb);
    e_c_last = eb;
    return true;
  }

  // parameters of the encoder half (same naming as the reference's state_dict); optional as a whole
  std::vector<std::pair<std::string, std::vector<int64_t>>> expected_encoder_params() const {
    std::vector<std::pair<std::string, std::vector<int64_t>>> v;
    auto add = [&](const std::string& n, std::vector<int64_t> s) { v.emplace_back(n, std::move(s)); };
    const int z2 = 2 * cfg.z_channels, e2 = 2 * cfg.embed_dim;
    add("encoder.conv_in.weight", {cfg.ch, cfg.out_ch, 3, 3});
    add("encoder.conv_in.bias", {cfg.ch});
    for (const VRes& r : eres) {
      add(r.prefix + "norm1.weight", {r.cin});
      add(r.prefix + "norm1.bias", {r.cin});
      add(r.prefix + "conv1.weight", {r.cout, r.cin, 3, 3});
      add(r.prefix + "conv1.bias", {r.cout});
      add(r.prefix + "norm2.weight", {r.cout});
      add(r.prefix + "norm2.bias", {r.cout});
      add(r.prefix + "conv2.weight", {r.cout, r.cout, 3, 3});
      add(r.prefix + "conv2.bias", {r.cout});
      if (r.skip) {
        add(r.prefix + "nin_shortcut.weight", {r.cout, r.cin, 1, 1});
        add(r.prefix + "nin_shortcut.bias", {r.cout});
      }
    }
    for (const VAttn& a : eattn) {
      add(a.prefix + "norm.weight", {a.C});
      add(a.prefix + "norm.bias", {a.C});
      for (const char* n : {"q", "k", "v", "proj_out"}) {
        add(a.prefix + n + ".weight", {a.C, a.C, 1, 1});
        add(a.prefix + n + ".bias", {a.C});
      }
    }
    for (const VConv& d : down) {
      add(d.prefix + "weight", {d.cout, d.cin, 3, 3});
      add(d.prefix + "bias", {d.cout});
    }
    add("encoder.norm_out.weight", {e_c_last});
    add("encoder.norm_out.bias", {e_c_last});
    add("encoder.conv_out.weight", {z2, e_c_last, 3, 3});
    add("encoder.conv_out.bias", {z2});
    add("quant_conv.weight", {e2, z2, 1, 1});
    add("quant_conv.bias", {e2});
    return v;
  }

  std::vector<std::pair<std::string, std::vector<int64_t>>> expected_params() const {
    std::vector<std::pair<std::string, std::vector<int64_t>>> v;
    auto add = [&](const std::string& n, std::vector<int64_t> s) { v.emplace_back(n, std::move(s)); };
    const int zc = cfg.z_channels;
    add("post_quant_conv.weight", {zc, cfg.embed_dim, 1, 1});
    add("post_quant_conv.bias", {zc});
    add("decoder.conv_in.weight", {c_in, zc, 3, 3});
    add("decoder.conv_in.bias", {c_in});
    for (const VRes& r : res) {
      add(r.prefix + "norm1.weight", {r.cin});
      add(r.prefix + "norm1.bias", {r.cin});
      add(r.prefix + "conv1.weight", {r.cout, r.cin, 3, 3});
      add(r.prefix + "conv1.bias", {r.cout});
      add(r.prefix + "norm2.weight", {r.cout});
      add(r.prefix + "norm2.bias", {r.cout});
      add(r.prefix + "conv2.weight", {r.cout, r.cout, 3, 3});
      add(r.prefix + "conv2.bias", {r.cout});
      if (r.skip) {
        add(r.prefix + "nin_shortcut.weight", {r.cout, r.cin, 1, 1});
        add(r.prefix + "nin_shortcut.bias", {r.cout});
      }
    }
    for (const VAttn& a : attn) {
      add(a.prefix + "norm.weight", {a.C});
      add(a.prefix + "norm.bias", {a.C});
      for (const char* n : {"q", "k", "v", "proj_out"}) {
        add(a.prefix + n + ".weight", {a.C, a.C, 1, 1});
        add(a.prefix + n + ".bias", {a.C});
      }
    }
    for (const VConv& u : up) {
      add(u.prefix + "weight", {u.cout, u.cin, 3, 3});
      add(u.prefix + "bias", {u.cout});
    }
    add("decoder.norm_out.weight", {c_last});
    add("decoder.norm_out.bias", {c_last});
    add("decoder.conv_out.weight", {cfg.out_ch, c_last, 3, 3});
    add("decoder.conv_out.bias", {cfg.out_ch});
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
  bool fp(const std::string& name, size_t numel, float** out) {
    const RawTensor* t;
    if (!get(name, numel, &t)) return false;
    *out = t->d;
    return true;
  }
  bool pack_conv(const std::string& name, int cout, int cin, bf16* dst, int ldk) {
    const RawTensor* t;
    if (!get(name, static_cast<size_t>(cout) * cin * 9, &t)) return false;
    CUDA_OK(launch_pack_conv_weight(t->d, cout, cin, 3, 3, dst, ldk, 0, 0));
    return true;
  }

  // ResnetBlock / AttnBlock weights -> kernel layouts (shared by the decoder and the encoder)
  bool finalize_res(VRes& r) {
    const RawTensor* t;
    const std::string& p = r.prefix;
    if (!fp(p + "norm1.weight", r.cin, &r.gn1_g) || !fp(p + "norm1.bias", r.cin, &r.gn1_b) ||
        !fp(p + "norm2.weight", r.cout, &r.gn2_g) || !fp(p + "norm2.bias", r.cout, &r.gn2_b) ||
        !fp(p + "conv1.bias", r.cout, &r.b1))
      return false;
    r.w1 = dev_alloc<bf16>(static_cast<size_t>(r.cout) * 9 * r.cin);
    const int k2 = 9 * r.cout + (r.skip ? r.cin : 0);
    r.w2 = dev_alloc<bf16>(static_cast<size_t>(r.cout) * k2);
    r.b2 = dev_alloc<float>(r.cout);
    if (!r.w1 || !r.w2 || !r.b2) {
      set_error("cudaMalloc failed");
      return false;
    }
    if (!pack_conv(p + "conv1.weight", r.cout, r.cin, r.w1, 9 * r.cin)) return false;
    if (!pack_conv(p + "conv2.weight", r.cout, r.cout, r.w2, k2)) return false;
    float *bo, *bs = nullptr;
    if (!fp(p + "conv2.bias", r.cout, &bo)) return false;
    if (r.skip) {  // nin_shortcut (model.py:113-123) appended along K of conv2
      if (!get(p + "nin_shortcut.weight", static_cast<size_t>(r.cout) * r.cin, &t)) return false;
      CUDA_OK(launch_pack_matrix(t->d, r.cout, r.cin, r.w2, k2, 9 * r.cout, 0, 0));
      if (!fp(p + "nin_shortcut.bias", r.cout, &bs)) return false;
    }
    vec_add_kernel<<<(r.cout + 255) / 256, 256>>>(bo, bs, r.b2, r.cout);
    return true;
  }

  bool finalize_attn(VAttn& a) {
    const RawTensor* t;
    const int C = a.C;
    if (!fp(a.prefix + "norm.weight", C, &a.gn_g) || !fp(a.prefix + "norm.bias", C, &a.gn_b) ||
        !fp(a.prefix + "proj_out.bias", C, &a.bo))
      return false;
    a.wqkv = dev_alloc<bf16>(static_cast<size_t>(3) * C * C);
    a.bqkv = dev_alloc<float>(static_cast<size_t>(3) * C);
    a.wo = dev_alloc<bf16>(static_cast<size_t>(C) * C);
    if (!a.wqkv || !a.bqkv || !a.wo) {
      set_error("cudaMalloc failed");
      return false;
    }
    const char* names[3] = {"q", "k", "v"};
    for (int i = 0; i < 3; ++i) {  // 1x1 convs == linears: rows [i*C, (i+1)*C) of the fused [3C][C] matrix
      if (!get(a.prefix + names[i] + ".weight", static_cast<size_t>(C) * C, &t)) return false;
      CUDA_OK(launch_pack_matrix(t->d, C, C, a.wqkv, C, 0, i * C, 0));
      if (!get(a.prefix + names[i] + ".bias", C, &t)) return false;
      CUDA_OK(cudaMemcpy(a.bqkv + static_cast<size_t>(i) * C, t->d, C * sizeof(float), cudaMemcpyDeviceToDevice));
    }
    if (!get(a.prefix + "proj_out.weight", static_cast<size_t>(C) * C, &t)) return false;
    CUDA_OK(launch_pack_matrix(t->d, C, C, a.wo, C, 0, 0, 0));
    return true;
  }

  bool finalize() {
    if (finalized) return true;
    for (const auto& kv : expected_params()) {
      size_t n = 1;
      for (int64_t d : kv.second) n *= static_cast<size_t>(d);
      const RawTensor* t;
      if (!get(kv.first, n, &t)) return false;
    }
    const RawTensor* t;
    const int zc = cfg.z_channels;
    if (cfg.embed_dim != zc) {
      set_error("vae: embed_dim != z_channels is not implemented");
      return false;
    }
    if (!fp("post_quant_conv.weight", static_cast<size_t>(zc) * zc, &w_pq) || !fp("post_quant_conv.bias", zc, &b_pq) ||
        !fp("decoder.conv_in.bias", c_in, &b_in))
      return false;
    // conv_in over im2col rows [tap*zc + c], K padded to 64 (vae_input_pack_kernel)
    w_in = dev_alloc<bf16>(static_cast<size_t>(c_in) * kpad_in, true);
    if (!w_in || !pack_conv("decoder.conv_in.weight", c_in, zc, w_in, kpad_in)) return false;
    for (VRes& r : res)
      if (!finalize_res(r)) return false;
    for (VAttn& a : attn)
      if (!finalize_attn(a)) return false;
    for (VConv& u : up) {
      u.w = dev_alloc<bf16>(static_cast<size_t>(16) * u.cout * u.cin);  // 4 phases x [cout][4*cin]
      if (!u.w || !get(u.prefix + "weight", static_cast<size_t>(u.cout) * u.cin * 9, &t)) return false;
      CUDA_OK(launch_pack_upconv_weight(t->d, u.cout, u.cin, u.w, 0));
      if (!fp(u.prefix + "bias", u.cout, &u.b)) return false;
    }
    if (!fp("decoder.norm_out.weight", c_last, &out_gn_g) || !fp("decoder.norm_out.bias", c_last, &out_gn_b)) return false;
    w_out = dev_alloc<bf16>(static_cast<size_t>(32) * 9 * c_last, true);  // N padded to 32
    b_out = dev_alloc<float>(32, true);
    if (!w_out || !b_out || !pack_conv("decoder.conv_out.weight", cfg.out_ch, c_last, w_out, 9 * c_last)) return false;
    if (!get("decoder.conv_out.bias", cfg.out_ch, &t)) return false;
    CUDA_OK(cudaMemcpy(b_out, t->d, cfg.out_ch * sizeof(float), cudaMemcpyDeviceToDevice));
    if (raw.count("encoder.conv_in.weight") && !finalize_encoder()) return false;
    CUDA_OK(cudaDeviceSynchronize());
    finalized = true;
    return true;
  }

  // Encoder weights (model.py:466-516) + quant_conv (autoencoder.py:84), folded into conv_out: the 1x1 conv
  // is a [2e][2z] matrix applied to conv_out's outputs, so W' = Wq Wc and b' = Wq bc + bq (computed in fp32 here).
  bool finalize_encoder() {
    for (const auto& kv : expected_encoder_params()) {
      size_t n = 1;
      for (int64_t d : kv.second) n *= static_cast<size_t>(d);
      const RawTensor* t;
      if (!get(kv.first, n, &t)) return false;
    }
    const int cin_img = cfg.out_ch, z2 = 2 * cfg.z_channels, e2 = 2 * cfg.embed_dim;
    if (9 * cin_img > kpad_in || e2 > 32) {
      set_error("vae encoder: 9 * image channels must fit 64 and 2 * embed_dim 32");
      return false;
    }
    if (!fp("encoder.conv_in.bias", cfg.ch, &eb_in)) return false;
    ew_in = dev_alloc<bf16>(static_cast<size_t>(cfg.ch) * kpad_in, true);
    if (!ew_in || !pack_conv("encoder.conv_in.weight", cfg.ch, cin_img, ew_in, kpad_in)) return false;
    // the image goes through the latent im2col pack kernel with an identity "post_quant_conv"
    e_ident = dev_alloc<float>(static_cast<size_t>(cin_img) * cin_img, true);
    e_zero = dev_alloc<float>(cin_img, true);
    if (!e_ident || !e_zero) return false;
    {
      std::vector<float> eye(static_cast<size_t>(cin_img) * cin_img, 0.f);
      for (int i = 0; i < cin_img; ++i) eye[static_cast<size_t>(i) * cin_img + i] = 1.f;
      CUDA_OK(cudaMemcpy(e_ident, eye.data(), eye.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    for (VRes& r : eres) {
      if (r.cin % 64 != 0 || r.cout % 32 != 0) {
        set_error("vae encoder: ResnetBlock channels must be multiples of 64 (in) / 32 (out) for the conv kernel");
        return false;
      }
      if (!finalize_res(r)) return false;
    }
    for (VAttn& a : eattn)
      if (!finalize_attn(a)) return false;
    for (VConv& d : down) {
      d.w = dev_alloc<bf16>(static_cast<size_t>(d.cout) * 9 * d.cin);
      if (!d.w || !pack_conv(d.prefix + "weight", d.cout, d.cin, d.w, 9 * d.cin)) return false;
      if (!fp(d.prefix + "bias", d.cout, &d.b)) return false;
    }
    if (!fp("encoder.norm_out.weight", e_c_last, &e_gn_g) || !fp("encoder.norm_out.bias", e_c_last, &e_gn_b)) return false;
    const RawTensor *wc, *bc, *wq, *bq;
    const size_t kc = static_cast<size_t>(e_c_last) * 9;
    if (!get("encoder.conv_out.weight", z2 * kc, &wc) || !get("encoder.conv_out.bias", z2, &bc) ||
        !get("quant_conv.weight", static_cast<size_t>(e2) * z2, &wq) || !get("quant_conv.bias", e2, &bq))
      return false;
    std::vector<float> hwc(z2 * kc), hbc(z2), hwq(static_cast<size_t>(e2) * z2), hbq(e2), fw(e2 * kc, 0.f), fb(32, 0.f);
    CUDA_OK(cudaMemcpy(hwc.data(), wc->d, hwc.size() * sizeof(float), cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(hbc.data(), bc->d, hbc.size() * sizeof(float), cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(hwq.data(), wq->d, hwq.size() * sizeof(float), cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(hbq.data(), bq->d, hbq.size() * sizeof(float), cudaMemcpyDeviceToHost));
    for (int o = 0; o < e2; ++o) {
      double b = hbq[o];
      for (int m = 0; m < z2; ++m) {
        const float q = hwq[static_cast<size_t>(o) * z2 + m];
        b += static_cast<double>(q) * hbc[m];
        for (size_t k = 0; k < kc; ++k) fw[o * kc + k] += q * hwc[m * kc + k];
      }
      fb[o] = static_cast<float>(b);
    }
    float* d_fw = dev_alloc<float>(fw.size());
    ew_out = dev_alloc<bf16>(static_cast<size_t>(32) * kc, true);  // N padded to 32
    eb_out = dev_alloc<float>(32, true);
    if (!d_fw || !ew_out || !eb_out) return false;
    CUDA_OK(cudaMemcpy(d_fw, fw.data(), fw.size() * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(eb_out, fb.data(), 32 * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_OK(launch_pack_conv_weight(d_fw, e2, e_c_last, 3, 3, ew_out, static_cast<int>(kc), 0, 0));
    has_encoder = true;
    return true;
  }

  // ------------------------------------------------------------------ planning
  struct Ctx {
    Arena arena;
    bool dry = true;
    uint8_t* base = nullptr;
    int n_img = 0;
    float* gn_partial = nullptr;
    std::vector<Op>* ops = nullptr;
    template <typename T>
    T* ptr(const Buf& b) const {
      return reinterpret_cast<T*>(base + b.off);
    }
    Buf alloc(size_t M, int C, size_t elem) {
      Buf b;
      b.M = static_cast<int>(M);
      b.C = C;
      b.bytes = M * C * elem;
      b.off = arena.alloc(b.bytes);
      b.valid = true;
      return b;
    }
    void release(Buf& b) {
      if (b.valid) arena.release(b.off, b.bytes);
      b.valid = false;
    }
  };

  void push(Ctx& c, int cls, double flops, double bytes, std::function<cudaError_t(cudaStream_t)> fn) {
    Op op;
    op.cls = cls;
    op.launches = 1;
    op.flops = flops;
    op.bytes = bytes;
    op.run = std::move(fn);
    c.ops->push_back(op);
  }

  bool op_gn(Ctx& c, const Buf& x, int hw, const float* g, const float* b, int silu, const Buf& out, const Buf* raw_out) {
    if (c.dry) return true;
    const float* px = c.ptr<float>(x);
    bf16* po = c.ptr<bf16>(out);
    bf16* pr = raw_out ? c.ptr<bf16>(*raw_out) : nullptr;
    float* partial = c.gn_partial;
    const int n_img = c.n_img, C = x.C;
    push(c, CLS_GN, 0, static_cast<double>(x.M) * C * (4 + 2 + (raw_out ? 2 : 0)), [=](cudaStream_t s) {
      return launch_groupnorm(px, C, nullptr, 0, n_img, hw, g, b, 1e-6f, silu, po, pr, partial, s);  // model.py:46-47
    });
    return true;
  }

  bool conv_op(Ctx& c, const bf16* A, int H, int W, int cin, const bf16* A2, int K2, const bf16* Wt, int cout, float* out,
               const float* bias, const float* residual, bool down = false) {
    GemmPlan p;
    ConvGeom g{c.n_img, H, W, 9, down ? 2 : 1};
    g.asym_pad = down ? 1 : 0;
    if (!make_conv_plan(&p, A, g, cin, A2, K2, Wt, cout, OUT_F32, out, cout, bias, nullptr, 1, 0, residual, cout))
      return false;
    GemmPlan copy = p;
    push(c, CLS_CONV, p.flops, 0, [copy](cudaStream_t s) { return launch_gemm(copy, s); });
    return true;
  }

  // ResnetBlock.forward with temb = None (model.py:125-146)
  bool plan_res(Ctx& c, const VRes& r, const Buf& x, int H, int W, Buf* out) {
    const size_t M = x.M;
    const int hw = H * W;
    Buf a1 = c.alloc(M, r.cin, 2), xb;
    if (r.skip) xb = c.alloc(M, r.cin, 2);
    if (!op_gn(c, x, hw, r.gn1_g, r.gn1_b, 1, a1, r.skip ? &xb : nullptr)) return false;
    Buf h = c.alloc(M, r.cout, 4);
    if (!c.dry && !conv_op(c, c.ptr<bf16>(a1), H, W, r.cin, nullptr, 0, r.w1, r.cout, c.ptr<float>(h), r.b1, nullptr))
      return false;
    c.release(a1);
    Buf a2 = c.alloc(M, r.cout, 2);
    if (!op_gn(c, h, hw, r.gn2_g, r.gn2_b, 1, a2, nullptr)) return false;
    c.release(h);
    *out = c.alloc(M, r.cout, 4);
    if (!c.dry && !conv_op(c, c.ptr<bf16>(a2), H, W, r.cout, r.skip ? c.ptr<bf16>(xb) : nullptr, r.skip ? r.cin : 0, r.w2,
                           r.cout, c.ptr<float>(*out), r.b2, r.skip ? nullptr : c.ptr<float>(x)))
      return false;
    c.release(a2);
    c.release(xb);
    return true;
  }

  // AttnBlock.forward (model.py:176-203): one head over all C channels
  bool plan_attn(Ctx& c, const VAttn& w, const Buf& x, int H, int W, Buf* out) {
    const int C = w.C, L = H * W;
    const size_t M = x.M;
    Buf a = c.alloc(M, C, 2);
    if (!op_gn(c, x, L, w.gn_g, w.gn_b, 0, a, nullptr)) return false;
    Buf qkv = c.alloc(M, 3 * C, 2);
    Buf o = c.alloc(M, C, 2);
    Buf sc = c.alloc(L, L, 4), pr = c.alloc(L, L, 2), vt = c.alloc(C, L, 2);
    *out = c.alloc(M, C, 4);
    if (!c.dry) {
      GemmPlan p;
      if (!make_gemm_plan(&p, c.ptr<bf16>(a), static_cast<int>(M), C, nullptr, 0, w.wqkv, 3 * C, OUT_BF16, c.ptr<bf16>(qkv),
                          3 * C, w.bqkv, nullptr, 1, 0, nullptr, 0))
        return false;
      {
        GemmPlan copy = p;
        push(c, CLS_LINEAR, p.flops, 0, [copy](cudaStream_t s) { return launch_gemm(copy, s); });
      }
      const float scale = 1.0f / sqrtf(static_cast<float>(C));
      for (int n = 0; n < c.n_img; ++n) {
        const bf16* q = c.ptr<bf16>(qkv) + static_cast<size_t>(n) * L * 3 * C;
        float* ps = c.ptr<float>(sc);
        bf16* pp = c.ptr<bf16>(pr);
        bf16* pvt = c.ptr<bf16>(vt);
        // scores[i][j] = q_i . k_j : A = Q (row stride 3C), "weights" = K (row stride 3C)
        if (!make_gemm_plan(&p, q, L, C, nullptr, 0, q + C, L, OUT_F32, ps, L, nullptr, nullptr, 1, 0, nullptr, 0, 3 * C,
                            3 * C))
          return false;
        {
          GemmPlan copy = p;
          push(c, CLS_ATTN, p.flops, 0, [copy](cudaStream_t s) { return launch_gemm(copy, s); });
        }
        push(c, CLS_ATTN, 0, static_cast<double>(L) * L * 6,
             [=](cudaStream_t s) { return launch_softmax_rows(ps, L, L, scale, pp, s); });
        push(c, CLS_ATTN, 0, static_cast<double>(L) * C * 4,
             [=](cudaStream_t s) { return launch_transpose_bf16(q + 2 * C, 3 * C, L, C, pvt, s); });
        // out_i = sum_j p_ij v_j : A = P [L][L], "weights" = V^T [C][L]
        if (!make_gemm_plan(&p, pp, L, L, nullptr, 0, pvt, C, OUT_BF16, c.ptr<bf16>(o) + static_cast<size_t>(n) * L * C, C,
                            nullptr, nullptr, 1, 0, nullptr, 0))
          return false;
        {
          GemmPlan copy = p;
          push(c, CLS_ATTN, p.flops, 0, [copy](cudaStream_t s) { return launch_gemm(copy, s); });
        }
      }
      if (!make_gemm_plan(&p, c.ptr<bf16>(o), static_cast<int>(M), C, nullptr, 0, w.wo, C, OUT_F32, c.ptr<float>(*out), C, w.bo,
                          nullptr, 1, 0, c.ptr<float>(x), C))
        return false;
      GemmPlan copy = p;
      push(c, CLS_LINEAR, p.flops, 0, [copy](cudaStream_t s) { return launch_gemm(copy, s); });
    }
    c.release(a);
    c.release(qkv);
    c.release(o);
    c.release(sc);
    c.release(pr);
    c.release(vt);
    return true;
  }

  // Upsample (model.py:71-76): nearest 2x + conv3x3, folded into four 2x2-tap phase convs
  bool plan_up(Ctx& c, const VConv& w, const Buf& x, int H, int W, Buf* out) {
    const int C = w.cin;
    Buf xb = c.alloc(x.M, C, 2);
    *out = c.alloc(static_cast<size_t>(x.M) * 4, w.cout, 4);
    if (!c.dry) {
      const float* px = c.ptr<float>(x);
      bf16* pb = c.ptr<bf16>(xb);
      const size_t n = static_cast<size_t>(x.M) * C;
      push(c, CLS_OTHER, 0, static_cast<double>(n) * 6, [=](cudaStream_t s) { return launch_cast_bf16(px, n, pb, s); });
      for (int phase = 0; phase < 4; ++phase) {
        GemmPlan p;
        ConvGeom g{c.n_img, H, W, 4, 1};
        g.up_phase = phase;
        if (!make_conv_plan(&p, pb, g, C, nullptr, 0, w.w + static_cast<size_t>(phase) * w.cout * 4 * C, w.cout, OUT_F32,
                            c.ptr<float>(*out), w.cout, w.b, nullptr, 1, 0, nullptr, 0))
          return false;
        p.flops = 2.0 * x.M * static_cast<double>(w.cout) * 9 * C;  // algorithmic: the un-folded conv's share
        GemmPlan copy = p;
        push(c, CLS_CONV, p.flops, 0, [copy](cudaStream_t s) { return launch_gemm(copy, s); });
      }
    }
    c.release(xb);
    return true;
  }

  bool build_plan(Ctx& c, int N, int H, int W) {
    auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
    if (N < 1 || !pow2(H) || !pow2(W)) {
      set_error("vae: latent H and W must be powers of two");
      return false;
    }
    c.n_img = N;
    Buf gnp;
    gnp.bytes = groupnorm_partial_bytes(N);
    gnp.off = c.arena.alloc(gnp.bytes);
    gnp.valid = true;
    if (!c.dry) {
      c.gn_partial = c.ptr<float>(gnp);
      // The GroupNorm kernel's per-image counters must be zero when a forward starts.  Every launch leaves them
      // zeroed, but a plan can be parked while its workspace is reused by the caller (or a launch can have been
      // aborted), so they are cleared on the caller's stream at the head of every forward (a few hundred bytes).
      {
        uint8_t* sync_ptr = c.base + gnp.off + groupnorm_sync_offset(N);
        const size_t sync_bytes = gnp.bytes - groupnorm_sync_offset(N);
        Op op;
        op.cls = CLS_OTHER;
        op.launches = 0;
        op.flops = 0;
        op.bytes = static_cast<double>(sync_bytes);
        op.run = [=](cudaStream_t s) { return cudaMemsetAsync(sync_ptr, 0, sync_bytes, s); };
        c.ops->push_back(op);
      }
    }
    const size_t M0 = static_cast<size_t>(N) * H * W;
    Buf a0 = c.alloc(M0, kpad_in, 2);
    Buf h = c.alloc(M0, c_in, 4);
    if (!c.dry) {
      Vae* self = this;
      bf16* pa0 = c.ptr<bf16>(a0);
      const int zc = cfg.z_channels, kp = kpad_in;
      const float *wpq = w_pq, *bpq = b_pq;
      push(c, CLS_OTHER, 0, static_cast<double>(M0) * (kp * 2 + zc * 4), [=](cudaStream_t s) {
        return launch_vae_input_pack(self->io_z, N, zc, H, W, wpq, bpq, self->io_inv_scale, kp, pa0, s);
      });
      GemmPlan p;
      if (!make_gemm_plan(&p, pa0, static_cast<int>(M0), kp, nullptr, 0, w_in, c_in, OUT_F32, c.ptr<float>(h), c_in, b_in,
                          nullptr, 1, 0, nullptr, 0))
        return false;
      GemmPlan copy = p;
      push(c, CLS_LINEAR, p.flops, 0, [copy](cudaStream_t s) { return launch_gemm(copy, s); });
    }
    c.release(a0);
    Buf cur = h;
    int curH = H, curW = W;
    for (const VLayer& l : layers) {
      Buf nxt;
      if (l.kind == V_RES) {
        if (!plan_res(c, res[l.idx], cur, curH, curW, &nxt)) return false;
      } else if (l.kind == V_ATTN) {
        if (!plan_attn(c, attn[l.idx], cur, curH, curW, &nxt)) return false;
      } else {
        if (!plan_up(c, up[l.idx], cur, curH, curW, &nxt)) return false;
        curH *= 2;
        curW *= 2;
      }
      c.release(cur);
      cur = nxt;
    }
    // norm_out, swish, conv_out (model.py:633-637)
    const size_t Mo = static_cast<size_t>(N) * curH * curW;
    Buf a = c.alloc(Mo, c_last, 2);
    if (!op_gn(c, cur, curH * curW, out_gn_g, out_gn_b, 1, a, nullptr)) return false;
    c.release(cur);
    Buf o32 = c.alloc(Mo, 32, 4);
    if (!c.dry) {
      if (!conv_op(c, c.ptr<bf16>(a), curH, curW, c_last, nullptr, 0, w_out, 32, c.ptr<float>(o32), b_out, nullptr)) return false;
      Vae* self = this;
      const float* po = c.ptr<float>(o32);
      const int cout = cfg.out_ch, oh = curH, ow = curW;
      push(c, CLS_OTHER, 0, static_cast<double>(Mo) * cout * 8, [=](cudaStream_t s) {
        if (self->io_out_u8 != nullptr) return launch_vae_output_u8(po, 32, Mo, self->io_out_u8, s);
        return launch_vae_output(po, 32, N, cout, oh, ow, self->io_out, s);
      });
    }
    c.release(a);
    c.release(o32);
    return true;
  }

  // Downsample (model.py:68-87): F.pad(x, (0,1,0,1)) + conv3x3 stride 2 on the four parity planes of x
  bool plan_down(Ctx& c, const VConv& w, const Buf& x, int H, int W, Buf* out) {
    const int C = w.cin;
    Buf planes = c.alloc(x.M, C, 2);
    *out = c.alloc(static_cast<size_t>(x.M) / 4, w.cout, 4);
    if (!c.dry) {
      const float* px = c.ptr<float>(x);
      bf16* pp = c.ptr<bf16>(planes);
      const int n = c.n_img;
      push(c, CLS_OTHER, 0, static_cast<double>(x.M) * C * 6,
           [=](cudaStream_t s) { return launch_parity_split_bf16(px, n, H, W, C, pp, s); });
      if (!conv_op(c, pp, H / 2, W / 2, C, nullptr, 0, w.w, w.cout, c.ptr<float>(*out), w.b, nullptr, true)) return false;
    }
    c.release(planes);
    return true;
  }

  // Encoder.forward (model.py:518-545) + quant_conv: images [N][3][H][W] -> moments [N][2e][H/f][W/f]
  bool build_enc_plan(Ctx& c, int N, int H, int W) {
    auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
    const int f = 1 << (cfg.n_levels - 1);
    if (N < 1 || !pow2(H) || !pow2(W) || H < f || W < f) {
      set_error("vae encoder: image H and W must be powers of two, at least 2^(levels-1)");
      return false;
    }
    c.n_img = N;
    Buf gnp;
    gnp.bytes = groupnorm_partial_bytes(N);
    gnp.off = c.arena.alloc(gnp.bytes);
    gnp.valid = true;
    if (!c.dry) {
      c.gn_partial = c.ptr<float>(gnp);
      // The GroupNorm kernel's per-image counters must be zero when a forward starts.  Every launch leaves them
      // zeroed, but a plan can be parked while its workspace is reused by the caller (or a launch can have been
      // aborted), so they are cleared on the caller's stream at the head of every forward (a few hundred bytes).
      {
        uint8_t* sync_ptr = c.base + gnp.off + groupnorm_sync_offset(N);
        const size_t sync_bytes = gnp.bytes - groupnorm_sync_offset(N);
        Op op;
        op.cls = CLS_OTHER;
        op.launches = 0;
        op.flops = 0;
        op.bytes = static_cast<double>(sync_bytes);
        op.run = [=](cudaStream_t s) { return cudaMemsetAsync(sync_ptr, 0, sync_bytes, s); };
        c.ops->push_back(op);
      }
    }
    const size_t M0 = static_cast<size_t>(N) * H * W;
    Buf a0 = c.alloc(M0, kpad_in, 2);
    Buf h = c.alloc(M0, cfg.ch, 4);
    if (!c.dry) {
      Vae* self = this;
      bf16* pa0 = c.ptr<bf16>(a0);
      const int ic = cfg.out_ch, kp = kpad_in;
      const float *eye = e_ident, *zero = e_zero;
      push(c, CLS_OTHER, 0, static_cast<double>(M0) * (kp * 2 + ic * 4), [=](cudaStream_t s) {
        return launch_vae_input_pack(self->io_img, N, ic, H, W, eye, zero, 1.0f, kp, pa0, s);
      });
      GemmPlan p;
      if (!make_gemm_plan(&p, pa0, static_cast<int>(M0), kp, nullptr, 0, ew_in, cfg.ch, OUT_F32, c.ptr<float>(h), cfg.ch,
                          eb_in, nullptr, 1, 0, nullptr, 0))
        return false;
      GemmPlan copy = p;
      push(c, CLS_LINEAR, p.flops, 0, [copy](cudaStream_t s) { return launch_gemm(copy, s); });
    }
    c.release(a0);
    Buf cur = h;
    int curH = H, curW = W;
    for (const VLayer& l : elayers) {
      Buf nxt;
      if (l.kind == V_RES) {
        if (!plan_res(c, eres[l.idx], cur, curH, curW, &nxt)) return false;
      } else if (l.kind == V_ATTN) {
        if (!plan_attn(c, eattn[l.idx], cur, curH, curW, &nxt)) return false;
      } else {
        if (!plan_down(c, down[l.idx], cur, curH, curW, &nxt)) return false;
        curH /= 2;
        curW /= 2;
      }
      c.release(cur);
      cur = nxt;
    }
    const size_t Mo = static_cast<size_t>(N) * curH * curW;
    Buf a = c.alloc(Mo, e_c_last, 2);
    if (!op_gn(c, cur, curH * curW, e_gn_g, e_gn_b, 1, a, nullptr)) return false;
    c.release(cur);
    Buf o32 = c.alloc(Mo, 32, 4);
    if (!c.dry) {
      if (!conv_op(c, c.ptr<bf16>(a), curH, curW, e_c_last, nullptr, 0, ew_out, 32, c.ptr<float>(o32), eb_out, nullptr))
        return false;
      Vae* self = this;
      const float* po = c.ptr<float>(o32);
      const int cout = 2 * cfg.embed_dim, oh = curH, ow = curW;
      push(c, CLS_OTHER, 0, static_cast<double>(Mo) * cout * 8,
           [=](cudaStream_t s) { return launch_vae_output(po, 32, N, cout, oh, ow, self->io_moments, s); });
    }
    c.release(a);
    c.release(o32);
    return true;
  }

  bool enc_workspace_bytes(int N, int H, int W, size_t* bytes) {
    if (!has_encoder) {
      set_error("vae: no encoder weights were loaded (encoder.*, quant_conv.*)");
      return false;
    }
    Ctx c;
    c.dry = true;
    std::vector<Op> dummy;
    c.ops = &dummy;
    if (!build_enc_plan(c, N, H, W)) return false;
    *bytes = c.arena.peak + 1024;
    return true;
  }

  bool encode(const float* images, float* moments, int N, int H, int W, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (!finalized) {
      set_error("cap4d_b200_vae_finalize has not been called");
      return false;
    }
    if (!(N == eN && H == eH && W == eW && ws == e_ws && ws_bytes == e_ws_bytes && !eops.empty())) {
      size_t need = 0;
      if (!enc_workspace_bytes(N, H, W, &need)) return false;
      if (ws == nullptr || ws_bytes < need) {
        set_error("workspace too small: need " + std::to_string(need) + " bytes");
        return false;
      }
      eops.clear();
      Ctx c;
      c.dry = false;
      c.base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~static_cast<uintptr_t>(1023));
      c.ops = &eops;
      if (!build_enc_plan(c, N, H, W)) {
        eops.clear();
        return false;
      }
      eN = N, eH = H, eW = W, e_ws = ws, e_ws_bytes = ws_bytes;
    }
    io_img = images;
    io_moments = moments;
    for (size_t i = 0; i < eops.size(); ++i) {
      cudaError_t e = eops[i].run(stream);
      if (e != cudaSuccess) {
        set_error("vae: launch of encoder op " + std::to_string(i) + " failed: " + cudaGetErrorString(e) + " / " +
                  get_error());
        return false;
      }
    }
    return true;
  }

  bool workspace_bytes(int N, int H, int W, size_t* bytes) {
    Ctx c;
    c.dry = true;
    std::vector<Op> dummy;
    c.ops = &dummy;
    if (!build_plan(c, N, H, W)) return false;
    *bytes = c.arena.peak + 1024;
    return true;
  }

  bool decode(const float* z, float* out, unsigned char* out_u8, int N, int H, int W, float scale_factor, void* ws,
              size_t ws_bytes, cudaStream_t stream) {
    if (out_u8 != nullptr && cfg.out_ch != 3) {
      set_error("vae: uint8 BGR output needs out_ch == 3");
      return false;
    }
    if (!finalized) {
      set_error("cap4d_b200_vae_finalize has not been called");
      return false;
    }
    if (!(N == pN && H == pH && W == pW && ws == p_ws && ws_bytes == p_ws_bytes && !ops.empty())) {
      size_t need = 0;
      if (!workspace_bytes(N, H, W, &need)) return false;
      if (ws == nullptr || ws_bytes < need) {
        set_error("workspace too small: need " + std::to_string(need) + " bytes");
        return false;
      }
      ops.clear();
      Ctx c;
      c.dry = false;
      c.base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~static_cast<uintptr_t>(1023));
      c.ops = &ops;
      if (!build_plan(c, N, H, W)) {
        ops.clear();
        return false;
      }
      pN = N, pH = H, pW = W, p_ws = ws, p_ws_bytes = ws_bytes;
    }
    io_z = z;
    io_out = out;
    io_out_u8 = out_u8;
    io_inv_scale = 1.0f / scale_factor;
    for (size_t i = 0; i < ops.size(); ++i) {
      cudaError_t e = ops[i].run(stream);
      if (e != cudaSuccess) {
        set_error("vae: launch of op " + std::to_string(i) + " failed: " + cudaGetErrorString(e) + " / " + get_error());
        return false;
      }
    }
    return true;
  }
};

}  // namespace

}  // namespace cap4d

using namespace cap4d;

extern "C" {

int cap4d_b200_vae_create(const cap4d_b200_vae_config* cfg, void** handle) {
  if (cfg == nullptr || handle == nullptr) {
    set_error("null argument");
    return 1;
  }
  Vae* v = new Vae();
  v->cfg = *cfg;
  if (!v->build_topology()) {
    delete v;
    return 2;
  }
  *handle = v;
  return 0;
}

int cap4d_b200_vae_load_weight(void* handle, const char* name, const float* data, const int64_t* shape, int ndim) {
  Vae* v = static_cast<Vae*>(handle);
  if (v == nullptr || name == nullptr || data == nullptr) {
    set_error("null argument");
    return 1;
  }
  if (v->finalized) {
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
  auto it = v->raw.find(name);
  if (it != v->raw.end() && it->second.d) cudaFree(it->second.d);
  v->raw[name] = t;
  return 0;
}

int cap4d_b200_vae_num_params(void* handle, int* n) {
  Vae* v = static_cast<Vae*>(handle);
  if (v == nullptr || n == nullptr) {
    set_error("null argument");
    return 1;
  }
  *n = static_cast<int>(v->expected_params().size());
  return 0;
}

int cap4d_b200_vae_param_info(void* handle, int index, char* name, int name_cap, int64_t* shape, int* ndim) {
  Vae* v = static_cast<Vae*>(handle);
  if (v == nullptr || name == nullptr || shape == nullptr || ndim == nullptr) {
    set_error("null argument");
    return 1;
  }
  auto params = v->expected_params();
  if (index < 0 || index >= static_cast<int>(params.size())) {
    set_error("parameter index out of range");
    return 1;
  }
  snprintf(name, name_cap, "%s", params[index].first.c_str());
  *ndim = static_cast<int>(params[index].second.size());
  for (int i = 0; i < *ndim; ++i) shape[i] = params[index].second[i];
  return 0;
}

int cap4d_b200_vae_finalize(void* handle) {
  Vae* v = static_cast<Vae*>(handle);
  if (v == nullptr) {
    set_error("null handle");
    return 1;
  }
  return v->finalize() ? 0 : 4;
}

int cap4d_b200_vae_workspace_bytes(void* handle, int N, int H, int W, size_t* bytes) {
  Vae* v = static_cast<Vae*>(handle);
  if (v == nullptr || bytes == nullptr) {
    set_error("null argument");
    return 1;
  }
  return v->workspace_bytes(N, H, W, bytes) ? 0 : 5;
}

int cap4d_b200_vae_decode(void* handle, const float* z, float* images, int N, int H, int W, float scale_factor,
                          void* workspace, size_t workspace_bytes, void* stream) {
  Vae* v = static_cast<Vae*>(handle);
  if (v == nullptr || z == nullptr || images == nullptr || scale_factor == 0.f) {
    set_error("null argument");
    return 1;
  }
  return v->decode(z, images, nullptr, N, H, W, scale_factor, workspace, workspace_bytes,
                   static_cast<cudaStream_t>(stream))
             ? 0
             : 6;
}

int cap4d_b200_vae_decode_u8(void* handle, const float* z, uint8_t* images_bgr, int N, int H, int W, float scale_factor,
                             void* workspace, size_t workspace_bytes, void* stream) {
  Vae* v = static_cast<Vae*>(handle);
  if (v == nullptr || z == nullptr || images_bgr == nullptr || scale_factor == 0.f) {
    set_error("null argument");
    return 1;
  }
  return v->decode(z, nullptr, images_bgr, N, H, W, scale_factor, workspace, workspace_bytes,
                   static_cast<cudaStream_t>(stream))
             ? 0
             : 6;
}

int cap4d_b200_vae_has_encoder(void* handle, int* yes) {
  if (handle == nullptr || yes == nullptr) {
    set_error("null argument");
    return 1;
  }
  *yes = static_cast<Vae*>(handle)->has_encoder ? 1 : 0;
  return 0;
}

int cap4d_b200_vae_encode_workspace_bytes(void* handle, int N, int H, int W, size_t* bytes) {
  if (handle == nullptr || bytes == nullptr) {
    set_error("null argument");
    return 1;
  }
  return static_cast<Vae*>(handle)->enc_workspace_bytes(N, H, W, bytes) ? 0 : 5;
}

int cap4d_b200_vae_encode(void* handle, const float* images, float* moments, int N, int H, int W, void* workspace,
                          size_t workspace_bytes, void* stream) {
  if (handle == nullptr || images == nullptr || moments == nullptr) {
    set_error("null argument");
    return 1;
  }
  return static_cast<Vae*>(handle)->encode(images, moments, N, H, W, workspace, workspace_bytes,
                                           static_cast<cudaStream_t>(stream))
             ? 0
             : 6;
}

int cap4d_b200_vae_num_launches(void* handle, int* n) {
  Vae* v = static_cast<Vae*>(handle);
  if (v == nullptr || n == nullptr) {
    set_error("null argument");
    return 1;
  }
  *n = static_cast<int>(v->ops.size());
  return 0;
}

int cap4d_b200_vae_destroy(void* handle) {
  delete static_cast<Vae*>(handle);
  return 0;
}

}  // extern "C"

// Internal host-side launch API of the cap4d_b200 CUDA kernels (not part of the
// C ABI; see include/cap4d_b200.h for the exported boundary).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace cap4d {

typedef __nv_bfloat16 bf16;

// Thread-local error text returned by cap4d_b200_last_error().
void set_error(const std::string& msg);
const char* get_error();

int sm_count();

// ---------------------------------------------------------------------------
// Tensor-core GEMM / implicit-GEMM convolution (gemm_tc.cu)
//   out[M, N] = epilogue( A[M, K] * W[N, K]^T )
// A is bf16, either a plain row-major matrix or an NHWC activation tensor
// addressed through a tap table (3x3 conv, stride 1 or the stride-2
// parity-plane form); an optional second plain A segment appends extra K
// (fused 1x1 skip convolution).  W is bf16 row-major [N][K_total].
// ---------------------------------------------------------------------------
enum GemmOutMode { OUT_F32 = 0, OUT_BF16 = 1, OUT_GEGLU_BF16 = 2 };

struct GemmParams {
  int M, N;          // logical output size (GEGLU: N counts the interleaved x|gate columns)
  int BN;            // N tile, multiple of 32, <= 256
  int msub;          // 128-row sub-tiles per CTA tile (1 or 2)
  int n_acc;         // TMEM accumulator stages (2 if msub*BN <= 256)
  int epi;           // epilogue instantiation (gemm_tc.cu, epilogue_kind): 0 generic, 1 residual look-ahead, 2 + coalescing
  int bias_smem;     // generic epilogue: each warp keeps the tile's bias slice in shared memory
  int tiles_m, tiles_n;
  int num_kb;        // K / 64 over all segments
  int seg0_kb;       // k-blocks served by A (taps * cin_kb for conv); the rest come from A2
  int a_conv;        // 0: A is 2-D [M][K]; 1: A is 4-D NHWC via the tap table
  int cin_kb;        // conv: k-blocks per tap (Cin / 64)
  int W, H;          // conv: spatial size of the OUTPUT grid (= box geometry)
  int box_h, box_n;  // conv: M tile = box_n images x box_h rows x W columns = 128 pixels
  int n_taps;
  int tap_dx[9], tap_dy[9], tap_dn[9];  // coordinate offsets per tap (dn: image-index offset, parity planes)
  int stages;        // smem pipeline depth
  int up_py, up_px;  // >= 0: output rows are scattered to phase (py, px) of the 2H x 2W grid (folded upsample)
  // epilogue
  int out_mode;
  void* out;          // fp32 or bf16, row-major, leading dimension ldo
  int ldo;
  const float* bias;     // [N] or null (GEGLU: interleaved like the weights)
  const float* rowbias;  // [M / rowbias_div][rowbias_ld] or null: per-image additive vector (timestep embedding)
  int rowbias_div, rowbias_ld;
  const float* residual;  // fp32 [M][ldr] or null
  int ldr;
  int a_f16, b_f16;  // operand storage formats: 0 = bf16, 1 = fp16 (set after make_*_plan; default bf16 x bf16)
};

struct GemmPlan {
  CUtensorMap tmA, tmA2, tmB;
  GemmParams p;
  int two_cta;  // 1: CTA-pair kernel (cta_group::2), tiles_m counts 256-row tiles
  int grid;
  size_t smem_bytes;
  double flops;  // 2*M*N*K, for reporting
  int ldw;       // row stride of the weight matrix in elements (0: dense, = K)
};

struct ConvGeom {
  int n_img, H, W;    // output grid
  int taps;           // 1 (1x1 / plain) or 9
  int stride;         // 1 or 2 (2: A holds the 4 parity planes [4][n_img][H][W][C])
  // up_phase >= 0: one phase (py = up_phase / 2, px = up_phase % 2) of "nearest-2x upsample then conv3x3"
  // folded into a 2x2-tap conv on the LOW-resolution grid H x W; taps = 4, weights [Cout][4*Cin] are the
  // phase-summed 3x3 taps, and output row (n, y, x) is written to pixel (n, 2y+py, 2x+px) of the 2H x 2W map
  int up_phase = -1;
  // stride 2 only: 0 = Conv2d(k 3, stride 2, padding 1) (openaimodel.py Downsample); 1 = the VAE encoder's
  // F.pad(x, (0,1,0,1)) + Conv2d(k 3, stride 2, padding 0) (model.py:80-84): taps reach rows/cols 2y .. 2y+2
  int asym_pad = 0;
};

// Plain GEMM: A [M][K] bf16 (lda == K), optional second segment A2 [M][K2].
// Returns false (and sets the error text) on invalid geometry.
bool make_gemm_plan(GemmPlan* plan, const bf16* A, int M, int K, const bf16* A2, int K2, const bf16* Wt, int N,
                    int out_mode, void* out, int ldo, const float* bias, const float* rowbias, int rowbias_div,
                    int rowbias_ld, const float* residual, int ldr, int lda = 0, int ldw = 0);
// lda / ldw: row strides of A and Wt in elements when they are column slices of wider matrices (0: dense);
// multiples of 8 (TMA needs 16 B strides).

// Implicit-GEMM 3x3 convolution, pad 1: A NHWC bf16 [n_img][H_in][W_in][Cin]
// (stride 2: parity planes, see ConvGeom).  Optional A2 [M][K2] plain segment.
bool make_conv_plan(GemmPlan* plan, const bf16* A, const ConvGeom& g, int Cin, const bf16* A2, int K2,
                    const bf16* Wt, int N, int out_mode, void* out, int ldo, const float* bias, const float* rowbias,
                    int rowbias_div, int rowbias_ld, const float* residual, int ldr);

cudaError_t launch_gemm(const GemmPlan& plan, cudaStream_t stream);

// ---------------------------------------------------------------------------
// Multi-view flash attention, head_dim 64 (attn_tc.cu)
//   qkv: bf16 [M][3C] rows = tokens, columns = [q | k | v], head h at h*64
//   out: bf16 [M][C]
// Tokens [s*L, (s+1)*L) form sequence s (L = h*w for per-view attention,
// V*h*w for the cross-view "3d" attention).
// ---------------------------------------------------------------------------
struct AttnPlan {
  CUtensorMap tmQKV;
  int M, C, L, n_seq, heads;
  bf16* out;
  const bf16* qkv;
  float scale_log2;  // softmax scale * log2(e)
  dim3 grid;
  size_t smem_bytes;
  double flops;  // 4 * n_seq * heads * L * L * 64
};
bool make_attn_plan(AttnPlan* plan, const bf16* qkv, bf16* out, int M, int C, int L, float scale);
// trace != nullptr: instrumented build, clock64 stamps of CTA (0,0,0): [3][16][8] (softmax A, softmax B, MMA thread)
cudaError_t launch_attn(const AttnPlan& plan, cudaStream_t stream, long long* trace = nullptr);

// ---------------------------------------------------------------------------
// Norms (norm.cu)
// ---------------------------------------------------------------------------
// GroupNorm(32 groups) over NHWC fp32, optionally over the channel
// concatenation of two tensors (C1 from x1, C2 from x2); writes bf16
// [M][C1+C2] = (silu?)(gn(x)) and optionally a raw bf16 copy of x.
// partial: scratch of groupnorm_partial_bytes(n_img): fp32 [n_img][GN_MAX_CHUNKS][32][2] + barrier counters.
enum { GN_MAX_CHUNKS = 128 };
// x2_G > 0: x1 / out hold B*x2_G images (the generated views) while x2 still holds B*x2_V images; image n
// reads x2 image (n / x2_G) * x2_V + x2_R + n % x2_G.  n_img_layout: the image count `partial` was sized and
// zeroed for when that is more than n_img (0 = n_img).
cudaError_t launch_groupnorm(const float* x1, int C1, const float* x2, int C2, int n_img, int hw, const float* gamma,
                             const float* beta, float eps, int apply_silu, bf16* out, bf16* raw_out, float* partial,
                             cudaStream_t stream, int x2_G = 0, int x2_V = 0, int x2_R = 0, int n_img_layout = 0,
                             int out_f16 = 0);  // out_f16: 1 = `out` is written as fp16 (raw_out stays bf16); 2 = out AND raw_out are
                                                // fp32 and SiLU is exact (fp32-accuracy mode)
size_t groupnorm_partial_bytes(int n_img);
size_t groupnorm_sync_offset(int n_img);  // the bytes from here to the end must be zero before the first launch

// LayerNorm over the last dim of fp32 [M][C] -> bf16 [M][C]
cudaError_t launch_layernorm(const float* x, int M, int C, const float* gamma, const float* beta, float eps,
                             bf16* out, cudaStream_t stream, int out_f16 = 0);

// ---------------------------------------------------------------------------
// Elementwise / data-movement (elementwise.cu)
// ---------------------------------------------------------------------------
// Input stage (mmdm_unet.py:77-95): ref-mask latent mix, 3x3 im2col of the 4-ch
// latent and the 50-ch pose conditioning packed into one bf16 row of kpad
// columns: [9*cin im2col | ccond pos_enc | 0...].
cudaError_t launch_input_pack(const float* x, const float* z_input, const float* ref_mask, const float* pos_enc,
                              int n_img, int cin, int H, int W, int ccond, int kpad, bf16* out,
                              cudaStream_t stream, int out_f16 = 0);
// Output stage (mmdm_unet.py:118-125): eps = x_input*mask + h*(1-mask), NHWC(ld) -> [n_img][cout][H][W]
cudaError_t launch_output_mix(const float* h, int ldh, const float* x, const float* z_input, const float* ref_mask,
                              int n_img, int cout, int H, int W, int G, int V, int R, float* out,
                              cudaStream_t stream, int* violations = nullptr);
// violations: device counter, incremented once per view that was declared a reference view (v < R with G > 0)
// but has ref_mask == 0 somewhere at its first pixel row; such views come out as NaN.
// dst image (b, g) = src image (b, R + g) for g < V - R: drops the R leading (reference) views of every group
cudaError_t launch_gather_views(const float* src, float* dst, int B, int V, int R, size_t per_img, cudaStream_t stream);
// fp32 NHWC -> bf16 parity planes [4][n_img][H/2][W/2][C] (plane = (y&1)*2 + (x&1))
cudaError_t launch_parity_split_bf16(const float* x, int n_img, int H, int W, int C, bf16* out,
                                     cudaStream_t stream, int out_f32 = 0);  // out_f32: planes stay fp32
// timestep embedding + time_embed MLP + all ResBlock emb_layers (fp32):
//   temb[n] = [cos(t f), sin(t f)] ; e = W2 silu(W1 temb + b1) + b2 ; out[n][:] = Wall silu(e) + ball
cudaError_t launch_time_embed(const long long* t, int n_img, int model_ch, int emb_ch, const float* w1,
                              const float* b1, const float* w2, const float* b2, const float* wall,
                              const float* ball, int n_all, float* scratch, float* out, cudaStream_t stream);
size_t time_embed_scratch_bytes(int n_img, int model_ch, int emb_ch);
// CFG combine + DDIM update (sampler.py:205-231) with scatter:
//   eps = eu + cfg (ec - eu) over the generated views of each group;
//   x[idx[g]] = x[idx[g]] * x_coef + eps * e_coef
// eps: [2*n_groups][V][chw] (uncond halves first, then cond halves), R leading reference views skipped.
cudaError_t launch_cfg_ddim_update(float* latents, const float* eps, const long long* gen_idx, int n_groups, int V,
                                   int R, int chw, float cfg, float x_coef, float e_coef, cudaStream_t stream);
// fp32 -> 16-bit weight repacks (device side).  The target format is bf16 unless set_weight_pack_f16(true) is in
// effect (thread-local; the executor sets it around its finalize): then the same buffers receive fp16 bit patterns
// and the GEMM plans that read them carry b_f16 = 1.
void set_weight_pack_f16(bool f16);
// general form: 0 bf16, 1 fp16, 2 fp32 (the output pointer then addresses floats; fp32-accuracy mode, precise.cu)
void set_weight_pack_format(int fmt);
int weight_pack_format();
cudaError_t launch_pack_conv_weight(const float* w_oihw, int O, int I, int KH, int KW, bf16* out, int ldk,
                                    int k_offset, cudaStream_t stream);  // out[o][k_offset + (kh*KW+kw)*I + i]
// nearest-2x-upsample + conv3x3 folded: out[phase][o][(a*2+b)*I + i] = sum of the 3x3 taps that fall on
// low-res tap (a, b) of phase (py, px); out is [4][O][4*I] bf16
cudaError_t launch_pack_upconv_weight(const float* w_oihw, int O, int I, bf16* out, cudaStream_t stream);
// fp32 -> bf16 copy
cudaError_t launch_cast_bf16(const float* x, size_t n, bf16* out, cudaStream_t stream);

// ---- fp32-accuracy mode (precise.cu): exact bf16 x 3 operand splits and exact pointwise maths ----
// x fp32 [rows][ldx] (columns [0, cols) in groups of group_width) -> out bf16 [rows][ld_out]: every group becomes six
// group_width-wide segments, [l|h|m|m|h|h] (worder 0: activation side) or [h|l|m|h|m|h] (worder 1: weight side)
cudaError_t launch_split6(const float* x, size_t rows, int cols, size_t ldx, int group_width, bf16* out, size_t ld_out,
                          int worder, cudaStream_t stream);
// conv operands: the five small-term segments per group go to out5 [rows][ld5] ([l|h|m|m|h] / weights [h|l|m|h|m]),
// the leading segment (h) to out1 [rows][ld1]
cudaError_t launch_split51(const float* x, size_t rows, int cols, size_t ldx, int group_width, bf16* out5, size_t ld5,
                           bf16* out1, size_t ld1, int worder, cudaStream_t stream);
// out[r][c] = u[r][c] * gelu(u[r][inner + c]) (erf form, exact), u fp32 [rows][2 inner]
cudaError_t launch_geglu_f32(const float* u, size_t rows, int inner, float* out, cudaStream_t stream);
// in-place softmax(scale * s) over the rows of fp32 [rows][L]
cudaError_t launch_softmax_rows_f32(float* s, int rows, int L, float scale, cudaStream_t stream);
cudaError_t launch_transpose_f32(const float* src, size_t ld, int rows, int cols, float* dst, cudaStream_t stream);
// attention on the fp32 fused qkv matrix [M][3C] -> fp32 [M][C] (same sequence convention as make_attn_plan)
cudaError_t launch_attention_f32(const float* qkv, float* out, int M, int C, int L, float scale, cudaStream_t stream);

// ---- VAE decode helpers (elementwise.cu) ----
// z NCHW fp32 -> (z / scale) -> post_quant_conv -> im2col rows [n*H*W][kpad] bf16 for conv_in (3x3, pad 1)
cudaError_t launch_vae_input_pack(const float* z, int n_img, int zc, int H, int W, const float* wpq, const float* bpq,
                                  float inv_scale, int kpad, bf16* out, cudaStream_t stream);
// fp32 NHWC [n*H*W][ldh] -> fp32 NCHW [n][cout][H][W]
cudaError_t launch_vae_output(const float* h, int ldh, int n_img, int cout, int H, int W, float* out,
                              cudaStream_t stream);
// fp32 NHWC [n_pix][ldh] (R, G, B first) -> uint8 [n_pix][3] = BGR of ((x + 1) / 2).clip(0, 1) * 255, truncated
cudaError_t launch_vae_output_u8(const float* h, int ldh, size_t n_pix, unsigned char* out, cudaStream_t stream);
// softmax(scale * s) over rows of fp32 [rows][L] -> bf16
cudaError_t launch_softmax_rows(const float* s, int rows, int L, float scale, bf16* p, cudaStream_t stream);
// dst[cols][rows] = src[rows][cols] (src row stride ld)
cudaError_t launch_transpose_bf16(const bf16* src, int ld, int rows, int cols, bf16* dst, cudaStream_t stream);
cudaError_t launch_pack_matrix(const float* w, int rows, int cols, bf16* out, int ldk, int k_offset, int row_offset,
                               cudaStream_t stream);                     // out[row_offset + r][k_offset + c]

}  // namespace cap4d

/*
 * cap4d_b200 -- C ABI of the B200-native MMDM multi-view denoising hot path.
 *
 * The reference (hitminxuanwang/cap4d) is pure Python and offers no FFI; its seam for this path
 * is three duck-typed Python call conventions.  Each entry point below names the reference
 * interface it replaces.  A Python binding (ctypes, cap4d_b200/_lib.py) is what a maintainer of
 * the reference would add; see INTEGRATION.md.
 *
 * Conventions: every function returns 0 on success and a non-zero status otherwise
 * (cap4d_b200_last_error() returns the text for the calling thread); nothing throws across the
 * boundary.  All tensor pointers are DEVICE pointers unless stated otherwise, fp32, contiguous,
 * in the reference's own layouts.  Calls are stream-ordered on `stream` (a cudaStream_t passed as
 * void*; NULL = the legacy default stream) and never synchronise the device, except
 * load_weight/finalize.  A handle is not re-entrant; use one handle per thread/stream.
 */
#ifndef CAP4D_B200_H_
#define CAP4D_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAP4D_B200_MAX_LEVELS 8

/* Hyper-parameters of MMDMUnetModel (cap4d/mmdm/net/mmdm_unet.py:15-33 +
 * controlnet/ldm/modules/diffusionmodules/openaimodel.py:448-477), i.e. the `unet_config.params`
 * block of configs/mmdm/cap4d_mmdm_final.yaml:95-115.  Fixed by the MMDM and therefore not
 * configurable: use_spatial_transformer=True, transformer_depth=1, conv_resample=True,
 * resblock_updown=False, use_scale_shift_norm=False, temporal_mode="3d", use_context=False. */
typedef struct cap4d_b200_unet_config {
  int in_channels;       /* 4  */
  int out_channels;      /* 4  */
  int model_channels;    /* 320 */
  int condition_channels;/* 50 */
  int num_res_blocks;    /* 2  */
  int n_levels;          /* len(channel_mult) = 4 */
  int channel_mult[CAP4D_B200_MAX_LEVELS];          /* 1,2,4,4 */
  int n_attention_resolutions;                      /* 3 */
  int attention_resolutions[CAP4D_B200_MAX_LEVELS]; /* 4,2,1 */
  int num_head_channels; /* 64 (only 64 is implemented) */
  int time_steps;        /* V: views per group the "3d" attention spans (8) */
} cap4d_b200_unet_config;

/* ---- U-Net handle: replaces MMDMUnetModel (mmdm_unet.py:14-126) ------------------------------ */

/* MMDMUnetModel.__init__ (mmdm_unet.py:15-33): derive the block topology from the config. */
int cap4d_b200_unet_create(const cap4d_b200_unet_config* cfg, void** handle);

/* nn.Module.load_state_dict as used by cap4d/inference/utils.py:44-61: one call per entry of
 * `model.model.diffusion_model.state_dict()`; `name` is the reference key
 * (e.g. "input_blocks.1.1.transformer_blocks.0.attn1.to_q.weight"), `data` fp32 on host or device. */
int cap4d_b200_unet_load_weight(void* handle, const char* name, const float* data, const int64_t* shape, int ndim);

/* Parameter census of the topology (= the keys and shapes of the reference module's state_dict,
 * e.g. 576 tensors / 815 549 764 elements for cap4d_mmdm_final.yaml).  shape must hold 4 entries. */
int cap4d_b200_unet_num_params(void* handle, int* n);
int cap4d_b200_unet_param_info(void* handle, int index, char* name, int name_capacity, int64_t* shape, int* ndim);

/* Arithmetic of this handle; call between create and finalize.
 *   0 (default)  16-bit tensor-core operands, fp32 accumulation: noise prediction within 1e-2 of the fp32 reference
 *                (measured 5.1e-3 .. 6.7e-3 on the production architecture)
 *   1            fp32 accuracy (the reference's own precision, openaimodel.py:522): every operand is the exact
 *                three-way bf16 split of the fp32 tensor and products keep all terms >= 2^-16 (six bf16 MMAs per
 *                product on the same tcgen05 kernels), exact SiLU / GELU / softmax, fp32 attention; within 1e-4 of
 *                the reference, an order of magnitude slower. */
int cap4d_b200_unet_set_precision(void* handle, int fp32_accuracy);

/* Repack all weights for the tensor-core kernels (bf16, K-major, fused QKV / GEGLU / skip-conv
 * layouts).  Fails if any parameter of the topology was not loaded. */
int cap4d_b200_unet_finalize(void* handle);

/* Optional promise about the batches that follow: the first n_ref_views views of every group are
 * reference views, i.e. ref_mask == 1 on them, as StochasticIOSampler builds its groups
 * (cap4d/mmdm/sampler.py:171-195: torch.cat([ref, gen], dim=1)).  Their outputs are x - z_input whatever
 * the network computes (mmdm_unet.py:122-125), so after the last cross-view attention layer the executor
 * keeps only the generated views.  0 (the default) computes every view like the reference.  Affects
 * workspace_bytes and forward calls made after it; outputs are unchanged. */
int cap4d_b200_unet_set_ref_views(void* handle, int n_ref_views);

/* The promise is checked on the device: a view declared a reference view whose ref_mask is 0 was never
 * computed, so its output is written as NaN (never a silently wrong number) and counted.  Returns the number of
 * such views since the last call and resets the count.  Blocking read on the legacy default stream: synchronise
 * non-blocking streams that ran forwards first. */
int cap4d_b200_unet_ref_view_violations(void* handle, int* n);

/* Scratch needed by one forward at this shape (caller owns the buffer; >= 1024 B aligned). */
int cap4d_b200_unet_workspace_bytes(void* handle, int B, int V, int H, int W, size_t* bytes);

/* MMDMUnetModel.forward(x, timesteps, context=None, control) (mmdm_unet.py:67-126), reached from
 * MMLDM.apply_model (cap4d/mmdm/mmdm.py:113-124):
 *   x, z_input, out: [B][V][C][H][W]   ref_mask: [B][V][1][H][W]   pos_enc: [B][V][H][W][Cc]
 *   timesteps: int64 [B][V]. */
int cap4d_b200_unet_forward(void* handle, const float* x, const int64_t* timesteps, const float* z_input,
                            const float* ref_mask, const float* pos_enc, float* out, int B, int V, int H, int W,
                            void* workspace, size_t workspace_bytes, void* stream);

/* Instrumentation (no reference counterpart): kernels launched per forward, algorithmic FLOPs per
 * forward by class, and a forward with CUDA events around every launch.  forward_timed with
 * class_ms != NULL synchronises the stream and returns the per-class ms of that call; with
 * class_ms == NULL it only records the events (no synchronisation, usable inside a timed region)
 * and collect_timings() later returns the per-class sums over all such calls.
 * classes: 0 conv3x3 (tcgen05 implicit GEMM), 1 linear (tcgen05 GEMM), 2 attention core,
 *          3 GroupNorm, 4 LayerNorm, 5 other (pack / mix / resample / embedding). */
#define CAP4D_B200_N_CLASSES 6
/* Build (or fetch from the cache) the launch plan of a batch shape without running it and make it the current plan,
 * which is what num_launches / class_stats / class_exec_flops describe. */
int cap4d_b200_unet_plan(void* handle, int B, int V, int H, int W, void* workspace, size_t workspace_bytes);
int cap4d_b200_unet_num_launches(void* handle, int* n);
int cap4d_b200_unet_class_stats(void* handle, double* flops, double* bytes, int* launches);
/* class_stats reports ALGORITHMIC FLOPs (the reference op's count); this returns what the tensor core executes per
 * class, which is less where an op was restructured (nearest-2x upsample + conv3x3 folded into four 2x2-tap phase
 * convolutions: 4/9 of the multiply-adds). */
int cap4d_b200_unet_class_exec_flops(void* handle, double* flops);
int cap4d_b200_unet_forward_timed(void* handle, const float* x, const int64_t* timesteps, const float* z_input,
                                  const float* ref_mask, const float* pos_enc, float* out, int B, int V, int H,
                                  int W, void* workspace, size_t workspace_bytes, void* stream, float* class_ms);

int cap4d_b200_unet_collect_timings(void* handle, float* class_ms, int* n_runs);

/* Debug taps (no reference counterpart; the reference would use forward hooks): with taps enabled every forward
 * also keeps the fp32 activation after each block of the topology - "input_blocks.<i>", "middle_block",
 * "output_blocks.<i>", the module names of openaimodel.py:544-774 - in the workspace (which therefore grows).
 * tap_info, valid after a forward, returns a DEVICE pointer to the token-major (NHWC) tensor [rows][channels] of tap
 * `index`; n_img is the image count it covers (fewer than B*V once the reference views were dropped, see
 * set_ref_views).  Enabling / disabling drops the cached launch plans. */
int cap4d_b200_unet_enable_taps(void* handle, int on);
int cap4d_b200_unet_num_taps(void* handle, int* n);
int cap4d_b200_unet_tap_info(void* handle, int index, char* name, int name_capacity, const float** data, int64_t* rows,
                             int* channels, int* n_img);

int cap4d_b200_unet_destroy(void* handle);

/* ---- sampler update: replaces cap4d/mmdm/sampler.py:205-208 + 215-231 ----------------------
 * eps: [2*n_groups][V][chw] as returned by the U-Net for the batch
 * [uncond group 0..n-1 | cond group 0..n-1]; for every group g and generated view v >= R:
 *   e = eps_u + cfg * (eps_c - eps_u);  latents[gen_idx[g][v-R]] = latents[..] * x_coef + e * e_coef */
int cap4d_b200_cfg_ddim_update(float* latents, const float* eps, const int64_t* gen_idx, int n_groups, int V, int R,
                               int chw, float cfg_scale, float x_coef, float e_coef, void* stream);

/* ---- device-resident sampler data plane: replaces cap4d/mmdm/sampler.py:141-213 ----------------
 * The reference keeps every tensor on the CPU and, per view group, indexes the conditioning dicts, concatenates
 * reference and generated views (dim 1) and the unconditional and conditional halves (dim 0), and copies the
 * result to the GPU.  Here the stores are uploaded once per sample() call and stay in HBM. */
#define CAP4D_B200_MAX_GROUPS_PER_CALL 16

/* Rows are views: *_z [n][C*H*W] (control["z_input"]), *_mask [n][H*W] (control["ref_mask"]), *_pos [n][H*W*Cc]
 * (control["pos_enc"]); ref_* index the reference views, gen_* the generated views; latents [n_gen][C*H*W] is
 * all_x_T (sampler.py:112).  The *_u members are the unconditional dicts (ref_uncond / gen_uncond): NULL means what
 * CAP4DConditioning produces for unconditional=True (cap4dcond.py:78-88): zeros for z_input and pos_enc, the
 * conditional ref_mask - in which case nothing is stored, uploaded or read for them. */
typedef struct cap4d_b200_sampler_stores {
  const float *ref_z, *ref_mask, *ref_pos;
  const float *gen_z, *gen_mask, *gen_pos;
  const float *ref_z_u, *ref_mask_u, *ref_pos_u;
  const float *gen_z_u, *gen_mask_u, *gen_pos_u;
  float* latents;
} cap4d_b200_sampler_stores;

/* What changes from call to call, read by the kernels from DEVICE memory (a captured CUDA graph is replayed with
 * new contents): the DDIM timestep of the step (sampler.py:124), its x / e_t factors (sampler.py:215-229) and the
 * view groups (rows of the step's index tables) this call processes. */
typedef struct cap4d_b200_sampler_call {
  int64_t timestep;
  float x_coef, e_coef;
  int32_t n_groups, pad_;
  int32_t groups[CAP4D_B200_MAX_GROUPS_PER_CALL];
} cap4d_b200_sampler_call;

/* sampler.py:161-195 for n_groups groups at once: builds x_in / control / t_in of the batch
 * [uncond group 0..n-1 | cond group 0..n-1] x V views (first R views = reference views) in the caller's buffers
 * x_in, z_in [2n][V][C][H][W], mask_in [2n][V][1][H][W], pos_in [2n][V][H][W][Cc], t_in int64 [2n][V].
 * ref_idx int64 [n_its][R] and gen_idx int64 [n_its][G] are the step's ref_batches / gen_batches
 * (sampler.py:131-139) in device memory; `call` is a DEVICE pointer. */
int cap4d_b200_sampler_gather(const cap4d_b200_sampler_stores* stores, const int64_t* ref_idx, const int64_t* gen_idx,
                              const cap4d_b200_sampler_call* call, int n_groups, int V, int R, int C, int H, int W,
                              int Cc, float* x_in, float* z_in, float* mask_in, float* pos_in, int64_t* t_in,
                              void* stream);

/* cap4d_b200_cfg_ddim_update with the groups, gen_idx rows and DDIM factors taken from `call` (device memory). */
int cap4d_b200_sampler_update(float* latents, const float* eps, const int64_t* gen_idx,
                              const cap4d_b200_sampler_call* call, int n_groups, int V, int R, int chw, float cfg_scale,
                              void* stream);

/* Per-step latent exchange between ranks (one process per GPU; rank r owns groups r, r + world, ... like the
 * reference deals groups to its device replicas, sampler.py:151-158): pack the views this rank updated into
 * send [ceil(n_its / world) * G][chw], all-gather (caller, NCCL), then unpack every other rank's block of
 * recv [world][ceil(n_its / world) * G][chw] into the latent store. */
int cap4d_b200_sampler_pack(const float* latents, const int64_t* gen_idx, int n_its, int G, int chw, int rank,
                            int world, float* send, void* stream);
int cap4d_b200_sampler_unpack(float* latents, const float* recv, const int64_t* gen_idx, int n_its, int G, int chw,
                              int rank, int world, void* stream);

/* ---- single-kernel entry points (building blocks; used by tests and micro-benchmarks) --------
 * bf16 tensors are passed as uint16_t*. */

/* torch.nn.functional.linear / 1x1 conv: out[M][N] = A[M][K] W[N][K]^T (+bias[N]) (+residual[M][N]);
 * out_mode 0: fp32, 1: bf16, 2: GEGLU (attention.py:68-75; W/bias rows interleaved [x32|gate32]). */
int cap4d_b200_gemm_bf16(const uint16_t* A, const uint16_t* Wt, int M, int N, int K, const float* bias,
                         const float* residual, void* out, int out_mode, void* stream, float* ms_out, int iters);

/* The same GEMM with the 16-bit storage format of each operand chosen separately (0 = bf16, 1 = fp16): the U-Net
 * executor keeps activations in bf16 and stores the weights as fp16 ("operand formats"). */
int cap4d_b200_gemm_mixed(const uint16_t* A, const uint16_t* Wt, int M, int N, int K, int a_f16, int b_f16,
                          const float* bias, const float* residual, void* out, int out_mode, void* stream,
                          float* ms_out, int iters);

/* nn.Conv2d(k=3, pad=1, stride 1|2) on NHWC bf16 (stride 2: parity planes); weights [Cout][9*Cin]
 * tap-major; optional rowbias[n_img][Cout] (ResBlock emb_out, openaimodel.py:265-274). */
int cap4d_b200_conv3x3_bf16(const uint16_t* A, const uint16_t* Wt, int n_img, int H_out, int W_out, int Cin, int Cout,
                            int stride, const float* bias, const float* rowbias, const float* residual, float* out,
                            void* stream, float* ms_out, int iters);

/* Upsample (openaimodel.py:92-120): F.interpolate(scale 2, nearest) then Conv2d(3, pad 1), evaluated as four
 * 2x2-tap phase convolutions on the low-resolution NHWC bf16 input [n_img][H][W][Cin]; w_oihw is the fp32
 * [Cout][Cin][3][3] device tensor (phase-summed and packed internally); out fp32 [n_img*2H*2W][Cout]. */
int cap4d_b200_upsample_conv3x3_bf16(const uint16_t* A, const float* w_oihw, int n_img, int H, int W, int Cin, int Cout,
                                     const float* bias, float* out, void* stream, float* ms_out, int iters);

/* legacy_attention (attention.py:112-132) for head_dim 64 on the fused qkv matrix [M][3C];
 * sequences are L consecutive rows. */
int cap4d_b200_attention_bf16(const uint16_t* qkv, uint16_t* out, int M, int C, int L, float scale, void* stream,
                              float* ms_out, int iters);

/* Instrumented attention launch: device buffer trace[3][16][8] (int64) receives clock64() stamps of
 * CTA (0,0,0) for the first 16 KV tiles: rows 0/1 = softmax warpgroups, row 2 = MMA issuer. */
int cap4d_b200_attention_trace(const uint16_t* qkv, uint16_t* out, int M, int C, int L, float scale, void* stream,
                               long long* trace);

/* GroupNorm32(32, C1+C2)(cat([x1, x2], channel)) (+SiLU) on NHWC fp32 -> bf16; x2 may be NULL. */
int cap4d_b200_groupnorm_bf16(const float* x1, int C1, const float* x2, int C2, int n_img, int hw,
                              const float* gamma, const float* beta, float eps, int apply_silu, uint16_t* out,
                              uint16_t* raw_out, void* stream, float* ms_out, int iters);

/* LayerNorm32(C) on fp32 [M][C] -> bf16. */
int cap4d_b200_layernorm_bf16(const float* x, int M, int C, const float* gamma, const float* beta, float eps,
                              uint16_t* out, void* stream, float* ms_out, int iters);

/* ---- VAE decoder handle: replaces MMLDM.decode_first_stage -----------------------------
 * controlnet/ldm/models/diffusion/ddpm.py:822-830 (z / scale_factor) -> AutoencoderKL.decode
 * (controlnet/ldm/models/autoencoder.py:87-91) -> Decoder.forward (controlnet/ldm/modules/diffusionmodules/
 * model.py:606-640), called once per generated view by cap4d/inference/utils.py:131-133. */
typedef struct cap4d_b200_vae_config {
  int ch;               /* 128 */
  int n_levels;         /* len(ch_mult) = 4 */
  int ch_mult[CAP4D_B200_MAX_LEVELS]; /* 1,2,4,4 */
  int num_res_blocks;   /* 2 */
  int z_channels;       /* 4 */
  int embed_dim;        /* 4 */
  int out_ch;           /* 3 */
} cap4d_b200_vae_config;

/* Decoder.__init__ (model.py:546-604): derive the layer list (attn_resolutions = [] : mid-block attention only). */
int cap4d_b200_vae_create(const cap4d_b200_vae_config* cfg, void** handle);
/* fp32 parameters under the reference's state_dict names ("decoder.*", "post_quant_conv.*" are required;
 * "encoder.*" + "quant_conv.*" are optional and enable cap4d_b200_vae_encode); host or device pointers. */
int cap4d_b200_vae_load_weight(void* handle, const char* name, const float* data, const int64_t* shape, int ndim);
int cap4d_b200_vae_num_params(void* handle, int* n);
int cap4d_b200_vae_param_info(void* handle, int index, char* name, int name_cap, int64_t* shape, int* ndim);
int cap4d_b200_vae_finalize(void* handle);
int cap4d_b200_vae_workspace_bytes(void* handle, int N, int H, int W, size_t* bytes);
/* z: fp32 [N][z_channels][H][W] (sampler latents, still multiplied by scale_factor);
 * images: fp32 [N][out_ch][8H][8W] in about [-1, 1] (what decode_first_stage returns). */
int cap4d_b200_vae_decode(void* handle, const float* z, float* images, int N, int H, int W, float scale_factor,
                          void* workspace, size_t workspace_bytes, void* stream);
/* Same decode, but the result is the uint8 array the reference writes to disk (cap4d/inference/utils.py:134-137):
 * images_bgr: uint8 [N][8H][8W][3] = ((x + 1) / 2).clip(0, 1) * 255 truncated, channels reversed for cv2.imwrite. */
int cap4d_b200_vae_decode_u8(void* handle, const float* z, uint8_t* images_bgr, int N, int H, int W, float scale_factor,
                             void* workspace, size_t workspace_bytes, void* stream);
/* Encoder half (the rest of SURVEY 8f rank 1): AutoencoderKL.encode (controlnet/ldm/models/autoencoder.py:82-85) =
 * Encoder.forward (controlnet/ldm/modules/diffusionmodules/model.py:518-545) + quant_conv, reached from
 * MMLDM.get_input (cap4d/mmdm/mmdm.py:60-63) once per reference image.  Available when the "encoder.*" and
 * "quant_conv.*" entries of the state_dict were loaded before finalize (they are optional as a whole).
 * images: fp32 [N][out_ch][H][W] in [-1, 1]; moments: fp32 [N][2*embed_dim][H/f][W/f] with f = 2^(n_levels-1), i.e.
 * DiagonalGaussianDistribution.parameters (mean | logvar); sampling and the scale factor stay with the caller. */
int cap4d_b200_vae_has_encoder(void* handle, int* yes);
int cap4d_b200_vae_encode_workspace_bytes(void* handle, int N, int H, int W, size_t* bytes);
int cap4d_b200_vae_encode(void* handle, const float* images, float* moments, int N, int H, int W, void* workspace,
                          size_t workspace_bytes, void* stream);
int cap4d_b200_vae_num_launches(void* handle, int* n);
int cap4d_b200_vae_destroy(void* handle);

/* ---- conditioning maps: replaces CAP4DConditioning.forward(unconditional=False) ----------
 * cap4d/mmdm/conditioning/cap4dcond.py:91-133 with PropRenderer.render (cap4d/mmdm/conditioning/mesh2img.py:334-379:
 * pytorch3d 0.7.8 rasterize_meshes with cameras=None, blur_radius 0, one face per pixel, clipped barycentrics, no
 * culling; interpolate_face_attributes), fused per view into one kernel: rasterise at (S*sr)^2 -> interpolate the
 * template positions `props` and the expression offsets / std -> [sin(2^k p), cos(2^k p)] -> render mask -> sr x sr
 * area average -> concat.  Channel order of pos_enc[n][S][S][C] (cap4dcond.py:104-133):
 *   positional_channels | 3 offsets (if offsets_3d) | 3 ray map (if ray_map) | ref_mask | crop mask (if crop_mask).
 * verts_2d / offsets_3d: [n][n_verts][3] (x, y in pytorch3d NDC, z = depth; batch["verts_2d"], batch["offsets_3d"]);
 * faces int32 [n_faces][3], props [n_verts][3], face_mask uint8 [n_faces] = PropRenderer's buffers
 * (mesh2img.py:348-366); ray_map [n][3][S][S], ref_mask / crop_mask [n][S][S] as the dataset provides them
 * (cap4d/inference/data/inference_data.py:108-114).  pix_to_face (optional, int32 [n][S*sr][S*sr]) receives
 * Fragments.pix_to_face with per-mesh face indices (-1 = background).  workspace: caller-owned device scratch of
 * cap4d_b200_cond_workspace_bytes(n_views, n_faces) bytes (the per-face tile ranges of the pre-pass).  The
 * unconditional branch (cap4dcond.py:78-88) is all zeros and has no entry point. */
int cap4d_b200_cond_workspace_bytes(int n_views, int n_faces, size_t* bytes);
int cap4d_b200_cond_pos_enc(const float* verts_2d, const float* offsets_3d, const int32_t* faces, const float* props,
                            const uint8_t* face_mask, const float* ray_map, const float* ref_mask,
                            const float* crop_mask, float* pos_enc, int32_t* pix_to_face, int n_views, int n_verts,
                            int n_faces, int image_size, int super_resolution, int positional_channels,
                            float positional_multiplier, float std_expr_deformation, void* workspace,
                            size_t workspace_bytes, void* stream);

/* load_camera_rays (cap4d/datasets/utils.py:161-186) followed by the rotation into the reference camera's frame
 * (cap4d/inference/data/inference_data.py:89-100), in fp64 like the numpy original, result fp32 [n][3][S][S].
 * cam: fp64 [n][22] on the device = new_fx, new_fy, new_cx, new_cy (utils.py:169-173), inv(extr[:3,:3]) row-major,
 * ref_extr[:3,:3] row-major. */
int cap4d_b200_cond_ray_map(const double* cam, float* ray_map, int n_views, int image_size, void* stream);

const char* cap4d_b200_last_error(void);
const char* cap4d_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CAP4D_B200_H_ */

/* deepv_b200 — C ABI of the B200-native DeepVerse denoise + decode hot path.
 *
 * The reference (lorenzocean/deepv) has no FFI: its seam is Python duck typing at three call
 * sites fed by InferencePipeline._create_models (pipeline.py:203-223).  This header is the
 * boundary a binding for those call sites talks to (INTEGRATION.md shows the ctypes stub):
 *
 *   dv_mmdit_*        replaces MMDiT.forward                      model/mmdit.py:1467-1530
 *   dv_cfg_euler_step replaces the CFG combine + scheduler.step   pipeline.py:502-520,
 *                                                                model/scheduler.py:230-294
 *   dv_stage_renoise  replaces the pyramid stage transition       pipeline.py:452-465
 *   dv_block_noise    replaces sample_block_noise                 pipeline.py:431-437
 *   dv_vae_*          replaces CausalVideoVAE.decode (tiled)      model/vae.py:885-920,989-1014
 *
 * Conventions: plain pointers and sizes only; every pointer named *_dev is DEVICE memory owned
 * by the caller and must stay alive until the stream has consumed it; calls enqueue work on
 * `stream` (a cudaStream_t passed as void*) and never synchronise or allocate, except the
 * *_create / *_plan_create calls, which allocate and may synchronise.  Every call returns 0 on
 * success or a negative code; dv_last_error() returns a thread-local message.  Handles are not
 * thread-safe: one per GPU / process.  There is NO CPU fallback: without a CUDA device every
 * compute entry point fails with DV_ERR_CUDA.
 */
#ifndef DEEPV_B200_H_
#define DEEPV_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DV_OK 0
#define DV_ERR_INVALID (-1)
#define DV_ERR_CUDA (-2)

#define DV_DTYPE_F32 0
#define DV_DTYPE_BF16 1

const char* dv_last_error(void);
int dv_version(void);
/* number of kernels launched by this library in this process since the last reset */
long long dv_launch_count(void);
void dv_launch_count_reset(void);
/* Per-launch CUDA-event profiler (off by default).  kind: 0 dense tcgen05 GEMM, 1 implicit-GEMM
 * conv3d, 2 attention, 3 other.  dv_profile_summary sums launches, device milliseconds and the
 * algorithmic FLOPs / bytes of every launch of that kind since the last reset; synchronise the
 * stream first.                                                                               */
void dv_profile_enable(int on);
void dv_profile_reset(void);
/* one CSV line per recorded launch (kind,tag,flops,bytes,ms); synchronise first */
int dv_profile_dump(const char* path);
int dv_profile_summary(int kind, long long* count, double* ms, double* flops, double* bytes);

/* ------------------------------------------------------------------------------------------
 * Sampler step  (bit-exact with the reference's ATen arithmetic)
 * ------------------------------------------------------------------------------------------ */
/* noise_pred_dev: [n_branch][numel] (uncond, text[, text+history]); sample/out: [numel].
 * n_branch 1: no guidance; 2: u + w_text (t - u); 3: ... + w_hist (h - t).
 * out = dtype( fp32(sample) + (sigma_next - sigma) * guided )                                */
int dv_cfg_euler_step(const void* noise_pred_dev, int n_branch, const void* sample_dev,
                      void* out_dev, long long numel, float w_text, float w_hist, double sigma,
                      double sigma_next, int dtype, void* stream);
/* lat_lo_dev: [planes][h][w]; noise_dev/out_dev: [planes][2h][2w];
 * out = alpha * nearest_up2(lat_lo) + beta * noise                                            */
int dv_stage_renoise(const void* lat_lo_dev, const void* noise_dev, void* out_dev, int planes,
                     int h, int w, double alpha, double beta, int dtype, void* stream);
/* z_dev: iid N(0,1) fp32 [planes][h/2][w/2][4]; out_dev: [planes][h][w] with
 * cov(2x2 block) = (1+gamma) I - gamma 11^T                                                   */
/* Bilinear down-sampling by 2 of [planes][H][W] (get_pyramid_latent pipeline.py:226-240; the initial
 * noise pyramid pipeline.py:554-557 with scale = 2): bit-equal to F.interpolate(mode='bilinear').   */
int dv_resize_half(const void* in_dev, void* out_dev, long long planes, int H, int W, float scale,
                   int dtype, void* stream);
int dv_block_noise(const float* z_dev, void* out_dev, int planes, int h, int w, float gamma,
                   int dtype, void* stream);

/* ------------------------------------------------------------------------------------------
 * Building blocks (exported for the parity tests and for other bindings)
 * ------------------------------------------------------------------------------------------ */
/* C[b][m][n] = epi( sum_k A[b][m][k] W[n][k] + bias[n] );  A, W, C bf16; K % 64 == 0, N % 32 == 0.
 * epi: 0 identity, 1 gelu-tanh.                                                               */
int dv_gemm_bf16(const void* A_dev, const void* W_dev, const float* bias_dev, void* C_dev,
                 int batch, int M, int N, int K, int epi, void* stream);
/* joint attention over fused qkv [B][L][3*H*64] bf16 -> out [B][L][H*64] bf16.
 * kv_end_dev: int32 [L] (keys visible to each query = prefix length, non-decreasing);
 * key_bias_dev: fp32 [B][Lpad], 0 for live keys, -inf for dead keys and k >= L; Lpad % 128 == 0 */
int dv_attention(const void* qkv_dev, void* out_dev, const int* kv_end_dev,
                 const float* key_bias_dev, int B, int L, int Lpad, int H, void* stream);
/* channels-last causal conv3d (zero padded, stride 1): x [B][T][H][W][Cin] bf16,
 * w [Cout_rows][kt*kh*kw*Cin] bf16 (tap-major, channel-minor), bias fp32 [>= Cout, padded to 32].
 * store: 0 plain, 1 pixel-shuffle (rows ordered (p1,p2,c)), 2 frame-interleave (rows (p,c)).   */
int dv_conv3d_cl(const void* x_dev, const void* w_dev, const float* bias_dev,
                 const void* residual_dev, void* out_dev, int B, int T, int H, int W, int Cin,
                 int Cout, int w_rows, int ksize, int store, int drop_first, void* stream);
/* The strided causal convs of the VAE encoder (vae.py:322,346; CausalConv3d :229-231,251): 3x3x3,
 * stride (1,2,2) or (2,1,1), zero padding 1 pixel around / 2 frames in front, then a VALID conv:
 * out [B][(T-1)/sT+1][H/sH][W/sW][Cout] bf16.  The strides are realised by TMA tensor maps over the
 * input parities (no im2col, no copies).                                                          */
int dv_conv3d_strided_cl(const void* x_dev, const void* w_dev, const float* bias_dev, void* out_dev,
                         int B, int T, int H, int W, int Cin, int Cout, int w_rows, int sT, int sH,
                         int sW, void* stream);

/* ------------------------------------------------------------------------------------------
 * MMDiT denoiser
 * ------------------------------------------------------------------------------------------ */
typedef struct dv_mmdit dv_mmdit;
typedef struct dv_mmdit_plan dv_mmdit_plan;

typedef struct {
  int num_layers;        /* 24 */
  int num_heads;         /* 24 */
  int head_dim;          /* 64 */
  int in_channels;       /* 38 */
  int patch_size;        /* 2  */
  int joint_dim;         /* 4096 (T5 width) */
  int pooled_dim;        /* 2048 */
  int pos_embed_max;     /* 192 */
  int pos_base_size;     /* sample_size / patch_size = 64 */
  int patch_k_pad;       /* in_channels * 4 rounded up to 64 = 192 */
} dv_mmdit_config;

/* All weights are DEVICE pointers owned by the caller; matrices bf16 row-major [out][in],
 * vectors fp32.  Per-layer arrays have num_layers entries; *_c entries of the last layer that
 * the reference does not have (context_pre_only block, mmdit.py:378-383) are NULL.            */
typedef struct {
  const void* const* w_qkv_x;  const float* const* b_qkv_x;   /* [3*D][D]: to_q | to_k | to_v      */
  const void* const* w_qkv_c;  const float* const* b_qkv_c;   /* add_q_proj | add_k_proj | add_v_proj */
  const float* const* qk_norm_x;                               /* [2][64]: norm_q.weight, norm_k.weight */
  const float* const* qk_norm_c;                               /* norm_add_q, norm_add_k            */
  const void* const* w_out_x;  const float* const* b_out_x;   /* to_out[0]                          */
  const void* const* w_out_c;  const float* const* b_out_c;   /* to_add_out (NULL in last layer)    */
  const void* const* w_ff1_x;  const float* const* b_ff1_x;   /* ff.net[0].proj      [4D][D]        */
  const void* const* w_ff2_x;  const float* const* b_ff2_x;   /* ff.net[2]           [D][4D]        */
  const void* const* w_ff1_c;  const float* const* b_ff1_c;   /* ff_context (NULL in last layer)    */
  const void* const* w_ff2_c;  const float* const* b_ff2_c;
  /* all adaLN linears concatenated along the output dim, in forward order:
   * per layer: norm1.linear (6D) then norm1_context.linear (6D; last layer 2D), then
   * norm_out.linear (2D).  rows = (L-1)*12D + 6D + 2D + 2D                                     */
  const void* w_mod;  const float* b_mod;  int mod_rows;
  const void* w_t1;   const float* b_t1;   /* time_text_embed.timestep_embedder.linear_1 [D][256] */
  const void* w_t2;   const float* b_t2;   /* ...linear_2 [D][D]                                  */
  const void* w_p1;   const float* b_p1;   /* time_text_embed.text_embedder.linear_1 [D][pooled]  */
  const void* w_p2;   const float* b_p2;   /* ...linear_2 [D][D]                                  */
  const void* w_ctx;  const float* b_ctx;  /* context_embedder [D][joint_dim]                     */
  const void* w_patch;      const float* b_patch;       /* pos_embed.proj as [D][patch_k_pad]      */
  const void* w_patch_hist; const float* b_patch_hist;  /* pos_embed.proj_history                  */
  const void* w_proj_out;   const float* b_proj_out;    /* proj_out [4*C][D], bias padded to 32k   */
} dv_mmdit_weights;

int dv_mmdit_create(const dv_mmdit_config* cfg, const dv_mmdit_weights* w, dv_mmdit** out);
void dv_mmdit_destroy(dv_mmdit* m);
/* debug: which = 0 -> the persistent block kernel's barrier word + (DV_PBK_TRACE=1) per-phase time stamps of every CTA */
int dv_mmdit_debug_buffer(dv_mmdit* m, int which, void** dev_ptr, long long* bytes);

/* A plan fixes one token layout: batch, the clips of one sample (oldest
 * first, the LAST one is the noisy clip and the only one returned), context length and the
 * optional history frame.  clip_thw: n_clips x (t, h, w) in LATENT pixels.  Allocates all
 * workspace the forward needs.                                                                 */
int dv_mmdit_plan_create(dv_mmdit* m, int batch, int n_clips, const int* clip_thw, int text_len,
                         int has_history, int hist_h, int hist_w, int hist_downsample,
                         dv_mmdit_plan** out);
void dv_mmdit_plan_destroy(dv_mmdit_plan* p);
long long dv_mmdit_plan_workspace_bytes(const dv_mmdit_plan* p);
/* algorithmic FLOPs of one forward with this plan (SURVEY.md §8d formula, dense attention pairs
 * restricted to frame-causal ones), for roofline accounting                                   */
double dv_mmdit_plan_flops(const dv_mmdit_plan* p);

/* clips_dev[i]: [B][C][t_i][h_i][w_i] contiguous, dtype io_dtype.
 * enc_dev: [B][text_len][joint_dim] (enc_dtype); ctx_mask_dev: fp32 [B][hist_tokens + text_len]
 * (history mask first, then the text mask; non-zero = live); pooled_dev: fp32 [B][pooled_dim];
 * timestep_dev: fp32 [B]; history_dev: [B][C][1][hist_h][hist_w] (io_dtype) or NULL.
 * out_dev: [B][C][1][h][w] of the noisy clip, dtype out_dtype.                                 */
int dv_mmdit_forward(dv_mmdit_plan* p, const void* const* clips_dev, int io_dtype,
                     const void* enc_dev, int enc_dtype, const float* ctx_mask_dev,
                     const float* pooled_dev, const float* timestep_dev, const void* history_dev,
                     void* out_dev, int out_dtype, void* stream);

/* ------------------------------------------------------------------------------------------
 * Ulysses sequence parallelism over the ranks of one rollout group (reference: none — the
 * reference is single-GPU; BASELINE.json north_star / SURVEY.md §8e.2).  The video tokens of a
 * forward are sharded over sp_world ranks (each runs LN / projections / MLP on its rows; the
 * <= 269 context tokens are replicated); around every attention the ranks exchange "my tokens,
 * your heads" for "all tokens, my heads" with an equal-block all-to-all, and the fp32 video
 * stream is all-gathered once before the output head, so that every rank returns the full
 * prediction.  The exchange is a callback: block j of send_dev goes to rank j, block i of
 * recv_dev comes from rank i, enqueued on `stream`.  dv_comm_* is the NCCL implementation of it
 * (libnccl.so.2 is resolved with dlopen; the communicator is built from a unique id that the
 * host broadcasts over its own process group).
 * ------------------------------------------------------------------------------------------ */
typedef int (*dv_exchange_fn)(void* user, const void* send_dev, void* recv_dev,
                              long long bytes_per_peer, void* stream);
/* Lv %% sp_world == 0 and num_heads %% sp_world == 0 are required; sp_world = 1 switches it off. */
int dv_mmdit_plan_set_sp(dv_mmdit_plan* p, int sp_rank, int sp_world, dv_exchange_fn fn, void* user);

/* Peer-memory variant of the exchange (the transfer fused into the producing kernels): when every
 * rank's qkv / attention buffers are mapped here (CUDA IPC between the processes of one box, or
 * plain pointers inside one process), the QKV-projection epilogue stores each head straight into
 * the buffer of the rank that owns it and the attention epilogue stores each output row into its
 * owner's buffer — NVLink stores issued tile by tile while the tensor pipe keeps working — and the
 * exchange callback is only called as a barrier (bytes_per_peer == 0, null buffers).
 * qkv_ptrs / attn_ptrs: sp_world device pointers each, entry sp_rank = this plan's own buffers
 * (dv_mmdit_plan_buffers); NULL switches back to the staged all-to-all.
 * x_ptrs / flag_ptrs (optional, both or neither): every rank's fp32 video stream and int[8] arrival
 * words.  With them the barrier is a one-warp kernel (release store of an epoch into every peer's
 * flag word, acquire-polling of the own words) and the final all-gather of the stream is a peer-store
 * kernel: the forward then contains no NCCL call and no host callback, and is replayed as a CUDA
 * graph.  Without them the exchange callback serves as the barrier (bytes_per_peer == 0).
 * x_dev / flags_dev of dv_mmdit_plan_buffers may be NULL; flags exist after dv_mmdit_plan_set_sp.   */
int dv_mmdit_plan_buffers(dv_mmdit_plan* p, void** qkv_dev, void** attn_dev, void** x_dev, void** flags_dev);
int dv_mmdit_plan_set_sp_peers(dv_mmdit_plan* p, void* const* qkv_ptrs, void* const* attn_ptrs,
                               void* const* x_ptrs, void* const* flag_ptrs);
int dv_ipc_get_handle(const void* dev_ptr, void* handle64);      /* cudaIpcGetMemHandle  */
int dv_ipc_open_handle(const void* handle64, void** dev_ptr);    /* cudaIpcOpenMemHandle */
int dv_ipc_close_handle(void* dev_ptr);                          /* cudaIpcCloseMemHandle (after the peers' plans are idle) */

typedef struct dv_comm dv_comm;
int dv_comm_unique_id(const char* nccl_path, void* id128);   /* rank 0: 128-byte id to broadcast */
int dv_comm_create(const char* nccl_path, const void* id128, int rank, int world, dv_comm** out);
void dv_comm_destroy(dv_comm* c);
/* a dv_exchange_fn; user = dv_comm* */
int dv_comm_exchange(void* user, const void* send_dev, void* recv_dev, long long bytes_per_peer,
                     void* stream);

/* ------------------------------------------------------------------------------------------
 * Causal video VAE decoder
 * ------------------------------------------------------------------------------------------ */
typedef struct dv_vae dv_vae;
typedef struct dv_vae_plan dv_vae_plan;

typedef struct {
  int latent_channels;        /* 16 */
  int out_channels;           /* 3  */
  int block_channels[4];      /* decoder_block_out_channels, e.g. 128,256,512,512 */
  int layers_per_block[4];    /* e.g. 3,3,3,3 */
  int spatial_up[4];          /* per up block (in decoder order): 1,1,1,0 */
  int temporal_up[4];         /* 1,1,1,0 */
  int norm_groups;            /* 32 */
  /* encoder (all zero when only the decoder's weights are loaded) */
  int enc_in_channels;        /* 3  */
  int enc_block_channels[4];  /* encoder_block_out_channels, e.g. 128,256,512,512 */
  int enc_layers_per_block[4];/* e.g. 2,2,2,2 */
  int enc_spatial_down[4];    /* 1,1,1,0 */
  int enc_temporal_down[4];   /* 1,1,1,0 */
} dv_vae_config;

/* Weights are passed as a flat name-ordered table built by the host side (see
 * deepv_b200/vae.py:pack_decoder_weights); conv weights bf16 [Cout_rows][taps*Cin_pad],
 * biases / norm affine fp32.                                                                   */
typedef struct {
  const char* name;
  const void* ptr;
  long long numel;
} dv_tensor_ref;

int dv_vae_create(const dv_vae_config* cfg, const dv_tensor_ref* tensors, int n_tensors,
                  dv_vae** out);
void dv_vae_destroy(dv_vae* v);
/* plan for decoding latents [1][16][T][h][w] with the reference's tiling (tile 32 latent px,
 * stride 24, 64 px blends; vae.py:989-1014)                                                    */
int dv_vae_plan_create(dv_vae* v, int T, int h, int w, int tile_latent, dv_vae_plan** out);
void dv_vae_plan_destroy(dv_vae_plan* p);
double dv_vae_plan_flops(const dv_vae_plan* p);
/* Trimmed decode: from now on dv_vae_decode / dv_vae_decode_tiles / dv_vae_blend of this plan produce output frames
 * [first_frame, Tout) only (0 = all); frames in front are neither computed nor written.  The kept frames are bit-identical
 * to a full decode: every causal conv only computes the frames the kept ones depend on (two frames of look-back per 3x3x3
 * conv at its own temporal resolution).  For the continuation iterations of a rollout, which discard the 25 re-decoded
 * input frames (pipeline.py:327-328).                                                                                   */
int dv_vae_plan_set_first_frame(dv_vae_plan* p, int first_frame);
/* z_dev: [1][16][T][h][w] (dtype), already un-normalised (pipeline.py:705-709);
 * out_dev: [1][3][8(T-1)+1][8h][8w] (dtype)                                                    */
int dv_vae_decode(dv_vae_plan* p, const void* z_dev, int z_dtype, void* out_dev, int out_dtype,
                  void* stream);
/* Tile-granular decode for multi-GPU sharding (the 6 tiles x {rgb, disparity} of one iteration
 * are independent, vae.py:994-1000; only the blend needs all of them):
 *   dv_vae_plan_geometry   tile grid (rows x cols) and output frame count
 *   dv_vae_plan_tile_info  decoded size of tile i (row-major), buffer = [t_out][H][W][3] bf16
 *   dv_vae_plan_bind_tile  make tile i live in a caller-owned device buffer (so the host can
 *                          broadcast / all_gather it between ranks)
 *   dv_vae_decode_tiles    decode the tiles whose bit is set in tile_mask
 *   dv_vae_blend           the sequential-order blend + crop + concat of all tiles -> out_dev   */
int dv_vae_plan_geometry(const dv_vae_plan* p, int* rows, int* cols, int* t_out);
int dv_vae_plan_tile_info(const dv_vae_plan* p, int tile, int* H, int* W);
int dv_vae_plan_bind_tile(dv_vae_plan* p, int tile, void* buf_dev);
int dv_vae_decode_tiles(dv_vae_plan* p, const void* z_dev, int z_dtype,
                        unsigned long long tile_mask, void* stream);
int dv_vae_blend(dv_vae_plan* p, void* out_dev, int out_dtype, void* stream);

/* ------------------------------------------------------------------------------------------
 * Causal video VAE encoder (row f1): `vae.encode(x)` as the rollout calls it (pipeline.py:250-251,
 * 569,574) -> CausalVideoVAE.encode with tiling on and temporal_chunk=False (vae.py:844-883,
 * 954-987): 256-px tiles every 192 px, encoder + quant_conv per tile (vae.py:630-689), 8-latent
 * linear blends, 24-latent crops, DiagonalGaussianDistribution (vae.py:599-615).
 * A plan fixes the clip [1][3][T][H][W]; tiles must be multiples of 64 x 128 pixels.
 *   x_dev        [1][3][T][H][W] (x_dtype)
 *   moments_dev  fp32 [1][2z][T'][H/8][W/8] (mean | logvar), T' = three causal halvings of T; or NULL
 *   noise_dev    fp32 [1][z][T'][H/8][W/8] standard-normal draw, or NULL (then no sample)
 *   sample_dev   [1][z][T'][H/8][W/8] (sample_dtype) = mean + exp(0.5 clamp(logvar,-30,20)) noise; or NULL
 * ------------------------------------------------------------------------------------------ */
typedef struct dv_vae_enc_plan dv_vae_enc_plan;
int dv_vae_enc_plan_create(dv_vae* v, int T, int H, int W, int tile_px, dv_vae_enc_plan** out);
void dv_vae_enc_plan_destroy(dv_vae_enc_plan* p);
double dv_vae_enc_plan_flops(const dv_vae_enc_plan* p);
int dv_vae_enc_plan_latent_dims(const dv_vae_enc_plan* p, int* t, int* h, int* w);
/* DiagonalGaussianDistribution.sample (vae.py:602-615) on moments [mean (n) | logvar (n)] */
int dv_gaussian_sample(const float* moments_dev, const float* noise_dev, void* sample_dev, long long n,
                       int sample_dtype, void* stream);
int dv_vae_encode(dv_vae_enc_plan* p, const void* x_dev, int x_dtype, float* moments_dev,
                  const float* noise_dev, void* sample_dev, int sample_dtype, void* stream);

/* ------------------------------------------------------------------------------------------
 * Rollout feedback on the device (row f3): what `InferencePipeline.generate` does on the host
 * between two `generate_i2v` calls (pipeline.py:311-414), with no host synchronisation.
 * Videos are [1][3][T][H][W]; poses are fp32 [n][4][4] row-major on the device.
 * ------------------------------------------------------------------------------------------ */
/* frames [t0, t0+n) -> uint8 (truncating, pipeline.py:341) -> ToTensor + Normalize(0.5, 0.5)
 * (pipeline.py:564-567): out [1][3][n][H][W] (out_dtype); u8_dev (optional) [n][H][W][3], the PIL frames */
int dv_frames_requantise(const void* frames_dev, int dtype, int T, int H, int W, int t0, int n,
                         void* out_dev, int out_dtype, unsigned char* u8_dev, void* stream);
/* pipeline.py:311-313: clamp(mean_c(raw) * 0.5 + 0.5, 0, 1)^2 / scale / 0.95 on 3 equal channels;
 * scale_dev: device scalar of the previous iteration, NULL = 1 (first iteration); out fp32 */
int dv_disparity_post(const void* raw_dev, int dtype, int T, int H, int W, const float* scale_dev,
                      float* out_dev, void* stream);
/* pipeline.py:346-350 (and :399-401 with clamp): if compute_scale, *scale_dev = 1 / max(disp[:, :, t0]);
 * out [1][3][n][H][W] (out_dtype) = sqrt(disp * scale * 0.95) * 2 - 1 for frames [t0, t0+n) */
int dv_disparity_renorm(const float* disp_dev, int T, int H, int W, int t0, int n, float* scale_dev,
                        int compute_scale, int clamp, void* out_dev, int out_dtype, void* stream);
/* pipeline.py:688-692 + raymap_to_trans_matrix :77-163 (append_first_reference, relative -> absolute):
 * latents [1][C][T][h][w] (dtype) whose channels [c0, c0+6) hold the normalised ray map; frames 1..T-1
 * are decoded.  trans3d_dev / trans2d_dev: fp32 [T][4][4] camera-to-world (frame 0 = identity) / intrinsics */
int dv_raymap_to_pose(const void* latents_dev, int dtype, int C, int c0, int T, int h, int w, int ds,
                      float* trans3d_dev, float* trans2d_dev, void* stream);
/* get_raymap_from_camera_parameters(_batchversion) pipeline.py:29-75 for n cameras at H x W pixels:
 * out [1][6][n][H/ds][W/ds] (out_dtype), (x - mean) / std applied when normalise (pipeline.py:300-301,258-259) */
int dv_camera_raymap(const float* trans2d_dev, const float* trans3d_dev, int n, int H, int W, int ds,
                     int normalise, void* out_dev, int out_dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEEPV_B200_H_ */

// deepv_b200 — launchers of the non-GEMM kernels (all enqueue on `stream`, no sync,
// no allocation).  Every function returns 0 or a negative error code.
#pragma once
#include "common.cuh"

namespace dv {

// ---- attention (attention.cu) ------------------------------------------------------
// tile_dead: [B][Lpad/128] "tile holds a dead key" flags from launch_key_bias, or nullptr
int launch_attention(const void* qkv, void* out, const int* kv_end, const float* key_bias,
                     const int* tile_dead, int B, int L, int Lpad, int H, cudaStream_t stream,
                     double flops = 0.0, int head0 = 0, int n_heads = 0, void* const* out_peers = nullptr,
                     int n_peers = 0, int peer_Lc = 0, int peer_Lw = 0);

// ---- MMDiT elementwise (elementwise.cu) -----------------------------------------------
// out_bf16[b][l][:] = LN(x[b][l][:]) * (1 + scale[b][:]) + shift[b][:]   (eps inside sqrt)
struct LnRows {
  const float* x;        // [B][L][D] fp32, batch stride x_bs
  long long x_bs;
  __nv_bfloat16* out;    // [B][L][D] bf16, batch stride out_bs
  long long out_bs;
  const float* shift;    // + b * mod_bs
  const float* scale;
  int L;
};
// two row sets (video + context stream) in one launch; r1 may be null
int launch_ln_modulate2(const LnRows& r0, const LnRows* r1, int mod_bs, int B, int D, float eps,
                        cudaStream_t stream);
int launch_ln_modulate(const float* x, long long x_batch_stride, __nv_bfloat16* out,
                       long long out_batch_stride, const float* shift, const float* scale,
                       int mod_batch_stride, int B, int L, int D, float eps, cudaStream_t stream);

// out[b][n] (+)= act_out( W[n][:] . act_in(in[b][:]) + bias[n] ),  W bf16 [N][K], B <= 4
// need (optional, device int[B]): the launch does nothing when every entry is 0 (conditioning cache hit)
int launch_gemv(const __nv_bfloat16* W, const float* bias, const float* in, int in_stride,
                float* out, int out_stride, int B, int N, int K, int silu_in, int accumulate,
                cudaStream_t stream, const int* need = nullptr);
// conditioning cache: the adaLN modulation table is a pure function of (timestep, pooled embedding) per batch row
// (mmdit.py:747-753,548,495) and a rollout asks for the same 15 timesteps x few prompts over and over.
// keys [slots][1 + pooled_dim], valid [slots], tables [slots][row_floats]; slot_of / need / store: device int[B]
int launch_cond_lookup(const float* t, const float* pooled, int pooled_dim, int B, float* keys, int* valid, int slots,
                       int* slot_of, int* need, int* store, cudaStream_t stream);
int launch_cond_finish(float* mod, long long row_floats, float* cache, const int* slot_of, const int* need,
                       const int* store, int B, cudaStream_t stream);

// sinusoidal timestep features, cos first (mmdit.py:645-683 with flip_sin_to_cos, shift 0)
int launch_timestep_features(const float* t, float* out, int B, cudaStream_t stream);

// latent [B][C][T][H][W] (fp32 or bf16) -> patch rows [B][rows_total][Kpad] bf16 at row_offset,
// k = c*4 + p1*2 + p2, zero padded to Kpad.  pool2 = 1 first averages 2x2 pixel blocks
// (bilinear /2 of mmdit.py:990) before patchifying.
int launch_patchify(const void* latent, int is_bf16, __nv_bfloat16* out, int rows_total,
                    int row_offset, int Kpad, int B, int C, int T, int H, int W, int pool2,
                    cudaStream_t stream);

// fp32/bf16 -> bf16 copy (context embeddings staging)
int launch_to_bf16(const void* in, int is_bf16, __nv_bfloat16* out, long long n,
                   cudaStream_t stream);

// base sincos table [S*S][D] (mmdit.py:590-642, grid coords / (S / base))
int launch_pos_base(float* out, int S, int D, int base_size, cudaStream_t stream);
// crop [top:top+oh, left:left+ow] of the base table, bilinear resize to (h, w) (align_corners
// False, mmdit.py:864-871), write T copies to out rows [row_offset, row_offset + T*h*w)
int launch_pos_clip(const float* base, int S, int D, float* out, int row_offset, int T, int h,
                    int w, int oh, int ow, cudaStream_t stream);

// key_bias[b][k] = 0 for live keys, -inf for dead context keys and k >= L
// (+ tile_dead[b][tile] = 1 when the 128-key tile holds any masked key; may be nullptr)
int launch_key_bias(const float* ctx_mask, int Lc, float* key_bias, int* tile_dead, int B, int L,
                    int Lpad, cudaStream_t stream);

// ---- sampler step (sampler.cu) -------------------------------------------------------
// CFG combine (pipeline.py:502-513) + Euler step (scheduler.py:278-286), rounding sequence of
// the ATen bf16 path reproduced exactly when is_bf16 = 1; plain fp32 otherwise.
int launch_cfg_euler(const void* noise_pred, int n_branch, const void* sample, void* out,
                     long long numel, float w_text, float w_hist, double sigma, double sigma_next,
                     int is_bf16, cudaStream_t stream);
// stage transition (pipeline.py:453-465): nearest x2 upsample, alpha * x + beta * noise
int launch_stage_renoise(const void* lat_lo, const void* noise, void* out, int planes, int h, int w,
                         double alpha, double beta, int is_bf16, cudaStream_t stream);
// correlated 2x2 block noise (pipeline.py:431-437): noise block = Lchol(4x4) @ z
int launch_block_noise(const float* z, void* out, int planes, int h, int w, float gamma,
                       int is_bf16, cudaStream_t stream);

// bilinear 1/2 down-sampling of [planes][H][W], ATen's rounding sequence (pipeline.py:226-240,554-557)
int launch_resize_half(const void* in, void* out, long long planes, int H, int W, float scale, int is_bf16,
                       cudaStream_t stream);

// ---- VAE elementwise (vae_kernels.cu) ---------------------------------------------------
// y = silu?( (x - mean) * rstd * gamma + beta ) over channels-last bf16 [frames][HW][C]
// acc: `replicas` copies (stride in doubles) of [frames][G][2] fp64 (sum, sum of squares)
int launch_gn_apply(const __nv_bfloat16* x, const double* acc, int replicas, long long replica_stride,
                    const float* gamma, const float* beta, __nv_bfloat16* y, int frames, int HW, int C,
                    int G, float eps, int silu_on, cudaStream_t stream);

// ---- Ulysses sequence parallelism: staging copies around the all-to-all (elementwise.cu) ----
int launch_sp_qkv_pack(const void* qkv, void* stage, int B, int L, int D, int row0, int Lw, int P,
                       int Hc, cudaStream_t stream);
int launch_sp_qkv_unpack(const void* stage, void* qkv, int B, int L, int D, int Lc, int Lw, int P,
                         int Hc, int r, cudaStream_t stream);
int launch_sp_attn_pack(const void* attn, void* stage, int B, int L, int D, int Lc, int Lw, int P,
                        int Hc, int r, cudaStream_t stream);
int launch_sp_attn_unpack(const void* stage, void* attn, int B, int L, int D, int Lc, int Lw, int P,
                          int Hc, int r, cudaStream_t stream);
int launch_sp_x_pack(const float* x, void* stage, int B, int Lv, int D, int Lw, int P, int r,
                     cudaStream_t stream);
int launch_sp_x_unpack(const void* stage, float* x, int B, int Lv, int D, int Lw, int P, int r,
                       cudaStream_t stream);
// peer-memory variant without NCCL: flag_peers[i] = rank i's int[8] arrival words, epoch = this rank's device-resident
// barrier counter; x_peers[i] = rank i's fp32 stream [B][Lv][D]
int launch_sp_barrier(void* const* flag_peers, int* epoch, int rank, int P, cudaStream_t stream);
int launch_sp_x_share(void* const* x_peers, int B, int Lv, int D, int row0, int rows, int rank, int P,
                      cudaStream_t stream);

}  // namespace dv

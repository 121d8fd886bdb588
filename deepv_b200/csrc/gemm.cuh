// deepv_b200 — the one tensor-core contraction kernel of the library.
//
// C[b][m][n] = sum_k A[b][m][k] * W[n][k]   (bf16 x bf16 -> fp32 in TMEM)
//
// * A comes either from a dense (K, rows, batch) tensor or, for the VAE's causal
//   conv3d, from a channels-last (C, W, H, T, B) activation through a 5-D TMA box
//   per filter tap (implicit GEMM; TMA out-of-bounds zero fill IS the causal /
//   spatial zero padding of reference model/vae.py:192-199,229-236).
// * W is always a K-major [N][K] bf16 matrix (nn.Linear layout; conv weights are
//   repacked to [Cout][tap][Cin] at load time).
// * The epilogue is selected at run time (EpiMode) and fuses what the reference
//   runs as separate ATen kernels.
#pragma once
#include "common.cuh"
#include "kernels.cuh"

namespace dv {

enum EpiMode : int {
  EPI_BF16 = 0,        // out_bf16 = acc + bias (+ residual_bf16, same layout as out)
  EPI_GELU = 1,        // out_bf16 = gelu_tanh(acc + bias)            (mmdit.py:97)
  EPI_RESID_GATE = 2,  // x_f32 += gate[b][n] * (acc + bias)          (mmdit.py:409-410,416-418,424-431)
  EPI_F32_ADD = 3,     // out_f32 = acc + bias + addend[row_map[m]][n] (patch-embed + pos, mmdit.py:894-935)
  EPI_QKV = 4,         // per-head RMSNorm + temporal RoPE on q,k; v plain (mmdit.py:282-307,131-136)
  EPI_UNPATCH = 5,     // proj_out + unpatchify store               (mmdit.py:1527,1452-1457)
  EPI_CONV = 6,        // conv3d: bias (+residual), NDHWC store with optional pixel-shuffle /
                       // frame-interleave address map              (vae.py:251,309,382,407-409)
  EPI_BF16_ROWBIAS = 7, // out_bf16 = acc + bias[m]  (transposed projections, V^T for VAE attention)
  EPI_TAPS = 8         // out_f32[n/4][b*M + m][4] = acc: per-tap partial products of a conv with <= 4
                       // output channels (plane-major so that the tap gather is coalesced); ldo = batch*M
};

enum ConvStore : int { CONV_PLAIN = 0, CONV_SHUFFLE_HW = 1, CONV_INTERLEAVE_T = 2 };

struct GemmDesc {
  // ---- problem -------------------------------------------------------------
  int batch;           // dense: batch count; conv: B
  int M;               // dense: valid rows per batch; conv: ignored (T*H*W)
  int N;               // valid output columns
  int K;               // dense: reduction length (multiple of 64 after padding)
  // ---- A operand -------------------------------------------------------------
  const void* A;       // bf16
  long long a_batch_stride;  // elements (dense)
  int lda;             // elements (dense)
  // conv geometry (a_mode == 1): activation [B][T][H][W][C] bf16
  int a_mode;          // 0 dense, 1 conv
  int cT, cH, cW, cC;  // input dims (C multiple of 64)
  int kt, kh, kw;      // 3,3,3 or 1,1,1
  int sT, sH, sW;      // conv strides: (1,1,1) [0 = 1], (1,2,2) or (2,1,1) (VAE encoder, vae.py:322,346);
                       // output dims T' = (T-1)/sT + 1 (causal), H' = H/sH, W' = W/sW
  // ---- W operand -------------------------------------------------------------
  const void* W;       // bf16 [N_rows][K]
  int w_rows;          // rows present in memory (>= N)
  long long w_batch_stride;  // elements; 0 = one W shared by every batch entry
  int ldw;             // elements between W rows; 0 = K (dense)
  // ---- epilogue ---------------------------------------------------------------
  int mode;
  void* out;
  long long out_batch_stride;  // elements
  int ldo;                     // elements
  int out_row_offset;          // rows added to m (joint-sequence placement)
  const float* bias;           // [N] (or [M] for ROWBIAS); may be null
  const float* gate;           // RESID_GATE: gate[b * gate_batch_stride + n]
  int gate_batch_stride;
  const float* addend;         // F32_ADD: [rows][N] fp32, may be null
  const int* row_map;          // F32_ADD: per-m row into addend (null -> m)
  // QKV
  const float* qk_norm_w;      // [2][64]: q weight then k weight
  const float* rope_cs;        // [frames][32][2] (cos, sin)
  const int* frame_id;         // [M]
  int heads_dim;               // = 1536 (q|k|v block width)
  // Ulysses over peer memory: the 64-column head starting at column n of the q, k or v block is
  // stored into out_peer[(n % heads_dim) / peer_cols] (same layout on every rank) instead of `out`
  void* out_peer[8];           // null-filled = off
  int peer_cols;               // attention columns (heads * 64) per rank
  // UNPATCH: out bf16 [B][C][1][2*gh][2*gw]
  int up_gh, up_gw, up_C;
  int out_f32;                 // UNPATCH: store fp32 instead of bf16
  // CONV
  int conv_store;              // ConvStore
  int conv_drop_first;         // INTERLEAVE_T: drop output frame 0 (vae.py:408-409)
  int conv_t0;                 // compute conv output frames [conv_t0, T) only (stride-1 convs; frame indices stay absolute)
  const void* residual;        // bf16, same layout as out (PLAIN store only), may be null
  int out_C;                   // channels of the stored tensor (N, N/4 or N/2)
  // CONV: optional fused GroupNorm statistics of the STORED tensor (vae.py:161-167): the epilogue
  // adds sum / sum of squares of the bf16-rounded outputs into gn_acc[(frame * G + group) * 2 + {0,1}]
  double* gn_acc;              // fp64 accumulators (zeroed by the caller), or null
  int gn_cpg;                  // channels per group of the stored tensor (4, 8 or 16)
  int gn_replicas;             // accumulator copies (CTA i adds into copy i % replicas; the reader
  long long gn_replica_stride; //   sums them) — spreads the same-address atomics; stride in doubles
};

// Enqueue on `stream`.  Returns 0 or a negative error (see dv_last_error()).
int launch_gemm(const GemmDesc& d, cudaStream_t stream);
// Two dense problems of the same epilogue mode in ONE launch (video + context stream of a joint
// block); d1 may be null.  Falls back to two launches when wave quantisation makes that cheaper.
int launch_gemm_pair(const GemmDesc& d0, const GemmDesc* d1, cudaStream_t stream);
// conv_halo.cu: 0 = launched, 1 = not a problem that kernel takes (use the generic path), < 0 = error
int launch_conv_halo(const GemmDesc& d, cudaStream_t stream);

// Persistent block kernel (gemm.cu, pbk_kernel): a sequence of LN-modulate and dense GEMM phases of the joint transformer
// blocks in ONE launch, grid barriers between the phases.  For token layouts whose GEMMs have fewer tiles than SMs.
struct PbkPhaseIn {
  int kind;          // 0 = LN-modulate (ln0, optional ln1), 1 = GEMM (g0, optional g1, same epilogue mode / N / K)
  LnRows ln0, ln1;
  int has_ln1;
  int mod_bs, B;
  float eps;
  GemmDesc g0, g1;
  int has_g1;
};
// workspace: fp32 scratch for split-K partials (pbk_workspace_floats() floats); bar: one zero-initialised device word
long long pbk_workspace_floats();
int launch_pbk(const PbkPhaseIn* phases, int n, float* workspace, unsigned* bar, cudaStream_t stream);

}  // namespace dv

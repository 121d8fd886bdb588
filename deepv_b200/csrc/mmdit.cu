// deepv_b200 — MMDiT denoiser forward: host-side orchestration of the sm_100a kernels.
//
// Mirrors reference model/mmdit.py:1467-1530 (MMDiT.forward) for the runnable configuration
// (sincos spatial pos-emb, temporal RoPE, temporal-causal joint attention, num_stages == 1):
//   temb      time_text_embed                      mmdit.py:739-753   -> GEMV chain
//   adaLN     all 2L+1 modulation linears at once  mmdit.py:548,495   -> one GEMV (temb is
//             block independent, so they are hoisted out of the block loop)
//   context   context_embedder (+ history tokens)  mmdit.py:1480-1485 -> tcgen05 GEMM
//   patches   PatchEmbed3D + cropped/interp pos    mmdit.py:882-975   -> patchify + GEMM(+pos)
//   blocks    JointTransformerBlock x L            mmdit.py:385-433   -> LN-mod, QKV GEMM with
//             RMSNorm+RoPE epilogue, attention, out-proj/FF GEMMs with gate-residual epilogue
//   head      norm_out + proj_out + unpatchify (last clip only, mmdit.py:1450,1526-1528)
// The residual streams are kept in fp32; GEMM operands are bf16.
#include <vector>

#include "../../include/deepv_b200.h"
#include <cstdlib>
#include <cstring>

#include "gemm.cuh"
#include "kernels.cuh"

using namespace dv;

struct dv_mmdit {
  dv_mmdit_config cfg;
  dv_mmdit_weights w;
  std::vector<const void*> p_w_qkv_x, p_w_qkv_c, p_w_out_x, p_w_out_c, p_w_ff1_x, p_w_ff2_x,
      p_w_ff1_c, p_w_ff2_c;
  std::vector<const float*> p_b_qkv_x, p_b_qkv_c, p_qk_norm_x, p_qk_norm_c, p_b_out_x, p_b_out_c,
      p_b_ff1_x, p_b_ff2_x, p_b_ff1_c, p_b_ff2_c;
  float* pos_base = nullptr;  // [S*S][D]
  float* rope_cs = nullptr;   // [kMaxFrames][32][2]
  int D = 0;
  // conditioning cache: modulation tables memoised per (timestep, pooled embedding) row (DV_MOD_CACHE_SLOTS, 0 = off)
  int cc_slots = 0;
  float *cc_keys = nullptr, *cc_tables = nullptr;
  int* cc_valid = nullptr;
  // persistent block kernel: split-K partial workspace (one 128 x 256 fp32 tile per SM) and the grid-barrier word
  float* pbk_work = nullptr;
  unsigned* pbk_bar = nullptr;
};

static constexpr int kMaxFrames = 256;

struct ClipInfo {
  int t, h, w;       // latent
  int gh, gw;        // token grid
  int row0, rows;    // rows inside the video stream
};

struct dv_mmdit_plan {
  dv_mmdit* m = nullptr;
  int B = 0, text_len = 0, hist_tokens = 0, hist_h = 0, hist_w = 0, hist_ds = 1;
  int Lc = 0, Lv = 0, L = 0, Lpad = 0, n_last = 0;
  std::vector<ClipInfo> clips;
  double flops = 0, attn_flops_layer = 0;
  long long bytes = 0;
  std::vector<void*> allocs;
  // device buffers
  int *frame_x = nullptr, *frame_c = nullptr, *kv_end = nullptr, *tile_dead = nullptr;
  float *pos_x = nullptr, *pos_h = nullptr;
  float *x = nullptr, *c = nullptr, *key_bias = nullptr;
  float *tfeat = nullptr, *g1 = nullptr, *g3 = nullptr, *temb = nullptr, *mod = nullptr;
  int* cc = nullptr;  // conditioning cache verdict of this forward: slot_of[4] | need[4] | store[4]
  __nv_bfloat16 *xn = nullptr, *cn = nullptr, *qkv = nullptr, *attn = nullptr, *ffh = nullptr, *ffh_c = nullptr;
  __nv_bfloat16 *patch_a = nullptr, *hist_a = nullptr, *enc_bf = nullptr, *xo = nullptr;
  // Ulysses sequence parallelism (dv_mmdit_plan_set_sp): this rank owns the video rows
  // [sp_rank * Lv / sp_world, ...) and the heads [sp_rank * H / sp_world, ...)
  int sp_rank = 0, sp_world = 1;
  dv_exchange_fn sp_exchange = nullptr;
  void* sp_user = nullptr;
  void *sp_send = nullptr, *sp_recv = nullptr;  // staging buffers of the all-to-all
  // peer-memory variant: every rank's qkv / attn buffer (own entry = own buffer); when set, the
  // QKV and attention epilogues store straight into the owner's buffer over NVLink and the
  // exchange callback is only used as a barrier (bytes_per_peer == 0)
  void* qkv_peer[8] = {nullptr};
  void* attn_peer[8] = {nullptr};
  bool peers_set = false;
  // device-side barrier + all-gather of the stream over peer memory (no NCCL, no host callback in the forward, so
  // a sequence-parallel forward can be replayed as a CUDA graph): every rank's fp32 stream and flag words mapped
  void* x_peer[8] = {nullptr};
  void* flag_peer[8] = {nullptr};
  bool dev_barrier = false;
  int* sp_flags = nullptr;   // [8] arrival epochs written by the peers (slot = source rank)
  int* sp_epoch = nullptr;   // [1] this rank's barrier count (device resident: graph replays keep counting)
  // CUDA-graph replay of the forward (DV_MMDIT_GRAPH=1, experimental): the inputs are copied into
  // plan-owned staging buffers so that one captured graph serves every call of this token layout
  cudaGraphExec_t graph_exec = nullptr;
  cudaStream_t cap_stream = nullptr;  // capture happens on a plan-owned stream (the caller's may be the legacy one)
  long long graph_launches = 0;       // kernels inside the captured forward (added to the launch counter per replay)
  int graph_io_dtype = -1, graph_enc_dtype = -1, graph_out_dtype = -1;
  std::vector<void*> stage_clips;
  void *stage_enc = nullptr, *stage_hist = nullptr, *stage_out = nullptr;
  float *stage_mask = nullptr, *stage_pooled = nullptr, *stage_t = nullptr;
};

namespace {

template <typename T>
int dev_alloc(dv_mmdit_plan* p, T** out, long long n) {
  void* ptr = nullptr;
  const size_t bytes = static_cast<size_t>(n) * sizeof(T);
  DV_CHECK_CUDA(cudaMalloc(&ptr, bytes ? bytes : 16));
  DV_CHECK_CUDA(cudaMemset(ptr, 0, bytes ? bytes : 16));
  p->allocs.push_back(ptr);
  p->bytes += bytes;
  *out = reinterpret_cast<T*>(ptr);
  return 0;
}

template <typename T, typename S>
void copy_ptrs(std::vector<T>& dst, S const* src, int n) {
  dst.assign(src, src + n);
}

GemmDesc dense_desc(const void* A, long long a_bs, int lda, const void* W, int w_rows,
                    const float* bias, int B, int M, int N, int K) {
  GemmDesc d = {};
  d.batch = B;
  d.M = M;
  d.N = N;
  d.K = K;
  d.A = A;
  d.a_batch_stride = a_bs;
  d.lda = lda;
  d.a_mode = 0;
  d.W = W;
  d.w_rows = w_rows;
  d.bias = bias;
  return d;
}

}  // namespace

extern "C" int dv_mmdit_create(const dv_mmdit_config* cfg, const dv_mmdit_weights* w,
                               dv_mmdit** out) {
  DV_REQUIRE(cfg && w && out, "dv_mmdit_create: null argument");
  DV_REQUIRE(cfg->head_dim == 64, "dv_mmdit_create: head_dim %d != 64", cfg->head_dim);
  DV_REQUIRE(cfg->patch_size == 2, "dv_mmdit_create: patch_size %d != 2", cfg->patch_size);
  const int D = cfg->num_heads * cfg->head_dim;
  DV_REQUIRE(D == 1536, "dv_mmdit_create: inner dim %d != 1536 (LN kernel is specialised)", D);
  DV_REQUIRE(cfg->patch_k_pad % 64 == 0 && cfg->patch_k_pad >= cfg->in_channels * 4,
             "dv_mmdit_create: patch_k_pad=%d", cfg->patch_k_pad);
  DV_REQUIRE(cfg->joint_dim % 64 == 0 && cfg->pooled_dim % 8 == 0, "dv_mmdit_create: dims");
  const int NL = cfg->num_layers;
  const int expect_rows = (NL - 1) * 12 * D + 6 * D + 2 * D + 2 * D;
  DV_REQUIRE(w->mod_rows == expect_rows, "dv_mmdit_create: mod_rows=%d, expected %d", w->mod_rows,
             expect_rows);
  dv_mmdit* m = new dv_mmdit();
  m->cfg = *cfg;
  m->w = *w;
  m->D = D;
  copy_ptrs(m->p_w_qkv_x, w->w_qkv_x, NL);
  copy_ptrs(m->p_b_qkv_x, w->b_qkv_x, NL);
  copy_ptrs(m->p_w_qkv_c, w->w_qkv_c, NL);
  copy_ptrs(m->p_b_qkv_c, w->b_qkv_c, NL);
  copy_ptrs(m->p_qk_norm_x, w->qk_norm_x, NL);
  copy_ptrs(m->p_qk_norm_c, w->qk_norm_c, NL);
  copy_ptrs(m->p_w_out_x, w->w_out_x, NL);
  copy_ptrs(m->p_b_out_x, w->b_out_x, NL);
  copy_ptrs(m->p_w_out_c, w->w_out_c, NL);
  copy_ptrs(m->p_b_out_c, w->b_out_c, NL);
  copy_ptrs(m->p_w_ff1_x, w->w_ff1_x, NL);
  copy_ptrs(m->p_b_ff1_x, w->b_ff1_x, NL);
  copy_ptrs(m->p_w_ff2_x, w->w_ff2_x, NL);
  copy_ptrs(m->p_b_ff2_x, w->b_ff2_x, NL);
  copy_ptrs(m->p_w_ff1_c, w->w_ff1_c, NL);
  copy_ptrs(m->p_b_ff1_c, w->b_ff1_c, NL);
  copy_ptrs(m->p_w_ff2_c, w->w_ff2_c, NL);
  copy_ptrs(m->p_b_ff2_c, w->b_ff2_c, NL);

  // base sincos table (PatchEmbed3D buffer, mmdit.py:820-824) and the temporal RoPE table
  // (EmbedNDRoPE / rope(), mmdit.py:999-1028: fp64 angles, cast to fp32).
  const int S = cfg->pos_embed_max;
  cudaError_t e = cudaMalloc(&m->pos_base, static_cast<size_t>(S) * S * D * sizeof(float));
  if (e != cudaSuccess) {
    set_error("dv_mmdit_create: cudaMalloc pos table: %s", cudaGetErrorString(e));
    delete m;
    return DV_ERR_CUDA;
  }
  int rc = launch_pos_base(m->pos_base, S, D, cfg->pos_base_size, 0);
  if (rc) {
    delete m;
    return rc;
  }
  std::vector<float> cs(static_cast<size_t>(kMaxFrames) * 64);
  for (int f = 0; f < kMaxFrames; ++f)
    for (int i = 0; i < 32; ++i) {
      const double omega = 1.0 / pow(10000.0, static_cast<double>(2 * i) / 64.0);
      const double a = static_cast<double>(f) * omega;
      cs[(f * 32 + i) * 2 + 0] = static_cast<float>(cos(a));
      cs[(f * 32 + i) * 2 + 1] = static_cast<float>(sin(a));
    }
  e = cudaMalloc(&m->rope_cs, cs.size() * sizeof(float));
  if (e == cudaSuccess)
    e = cudaMemcpy(m->rope_cs, cs.data(), cs.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    set_error("dv_mmdit_create: rope table: %s", cudaGetErrorString(e));
    delete m;
    return DV_ERR_CUDA;
  }
  {
    // 15 timesteps x (negative + a handful of prompts) of a rollout: 128 slots of 1.76 MB, four probes per lookup
    const char* env = getenv("DV_MOD_CACHE_SLOTS");
    m->cc_slots = env ? atoi(env) : 128;
    if (m->cc_slots > 0) {
      const size_t klen = static_cast<size_t>(cfg->pooled_dim) + 1;
      e = cudaMalloc(&m->cc_keys, m->cc_slots * klen * sizeof(float));
      if (e == cudaSuccess) e = cudaMalloc(&m->cc_tables, static_cast<size_t>(m->cc_slots) * w->mod_rows * sizeof(float));
      if (e == cudaSuccess) e = cudaMalloc(&m->cc_valid, m->cc_slots * sizeof(int));
      if (e == cudaSuccess) e = cudaMemset(m->cc_valid, 0, m->cc_slots * sizeof(int));
      if (e != cudaSuccess) {
        set_error("dv_mmdit_create: conditioning cache: %s", cudaGetErrorString(e));
        dv_mmdit_destroy(m);
        return DV_ERR_CUDA;
      }
    }
  }
  e = cudaMalloc(&m->pbk_work, static_cast<size_t>(pbk_workspace_floats()) * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&m->pbk_bar, 64 + 16 * 160 * 8);   // barrier word (+ room for DV_PBK_TRACE stamps)
  if (e != cudaSuccess) {
    set_error("dv_mmdit_create: persistent-kernel workspace: %s", cudaGetErrorString(e));
    dv_mmdit_destroy(m);
    return DV_ERR_CUDA;
  }
  *out = m;
  return DV_OK;
}

extern "C" int dv_mmdit_debug_buffer(dv_mmdit* m, int which, void** dev_ptr, long long* bytes) {
  DV_REQUIRE(m && dev_ptr && bytes, "dv_mmdit_debug_buffer: null argument");
  DV_REQUIRE(which == 0, "dv_mmdit_debug_buffer: unknown buffer %d", which);
  *dev_ptr = m->pbk_bar;   // [0] grid-barrier word; from byte 64: DV_PBK_TRACE stamps [phase + 1][SM] u64 (globaltimer ns)
  *bytes = 64 + 16 * 160 * 8;
  return DV_OK;
}

extern "C" void dv_mmdit_destroy(dv_mmdit* m) {
  if (!m) return;
  cudaFree(m->pbk_work);
  cudaFree(m->pbk_bar);
  cudaFree(m->pos_base);
  cudaFree(m->rope_cs);
  cudaFree(m->cc_keys);
  cudaFree(m->cc_tables);
  cudaFree(m->cc_valid);
  delete m;
}

extern "C" int dv_mmdit_plan_create(dv_mmdit* m, int batch, int n_clips, const int* clip_thw,
                                    int text_len, int has_history, int hist_h, int hist_w,
                                    int hist_downsample, dv_mmdit_plan** out) {
  DV_REQUIRE(m && clip_thw && out, "dv_mmdit_plan_create: null argument");
  DV_REQUIRE(batch >= 1 && batch <= 4, "dv_mmdit_plan_create: batch %d not in [1,4]", batch);
  DV_REQUIRE(n_clips >= 1, "dv_mmdit_plan_create: n_clips=%d", n_clips);
  const int D = m->D;
  const int S = m->cfg.pos_embed_max;
  dv_mmdit_plan* p = new dv_mmdit_plan();
  p->m = m;
  p->B = batch;
  p->text_len = text_len;
  auto fail = [&](int rc) {
    dv_mmdit_plan_destroy(p);
    return rc;
  };

  int row = 0;
  for (int i = 0; i < n_clips; ++i) {
    ClipInfo ci;
    ci.t = clip_thw[3 * i];
    ci.h = clip_thw[3 * i + 1];
    ci.w = clip_thw[3 * i + 2];
    if (ci.t < 1 || ci.h % 2 || ci.w % 2 || ci.h < 2 || ci.w < 2) {
      set_error("dv_mmdit_plan_create: clip %d has invalid dims (%d,%d,%d)", i, ci.t, ci.h, ci.w);
      return fail(DV_ERR_INVALID);
    }
    ci.gh = ci.h / 2;
    ci.gw = ci.w / 2;
    ci.row0 = row;
    ci.rows = ci.t * ci.gh * ci.gw;
    row += ci.rows;
    p->clips.push_back(ci);
  }
  p->Lv = row;
  const ClipInfo& last = p->clips.back();
  p->n_last = last.rows;
  if (has_history) {
    p->hist_h = hist_h;
    p->hist_w = hist_w;
    p->hist_ds = hist_downsample;
    if (hist_downsample != 2 || hist_h % 4 || hist_w % 4) {
      set_error("dv_mmdit_plan_create: history %dx%d / %d unsupported (need ratio 2)", hist_h,
                hist_w, hist_downsample);
      return fail(DV_ERR_INVALID);
    }
    p->hist_tokens = (hist_h / 4) * (hist_w / 4);
  }
  p->Lc = p->hist_tokens + text_len;
  p->L = p->Lc + p->Lv;
  p->Lpad = (p->L + 127) / 128 * 128;
  // every clip's positional crop is taken at the noisy clip's size (mmdit.py:960-962)
  for (const ClipInfo& ci : p->clips) {
    if (ci.gh > last.gh || ci.gw > last.gw || last.gh > S || last.gw > S) {
      set_error("dv_mmdit_plan_create: clip grid %dx%d exceeds the noisy clip's %dx%d", ci.gh,
                ci.gw, last.gh, last.gw);  // mmdit.py:851-852
      return fail(DV_ERR_INVALID);
    }
  }

  const int B = batch, Lv = p->Lv, Lc = p->Lc, L = p->L;
  const int Kp = m->cfg.patch_k_pad;
  int rc = 0;
#define DV_A(ptr, n)                          \
  if ((rc = dev_alloc(p, &(ptr), (n))) != 0) return fail(rc)
  DV_A(p->frame_x, Lv);
  DV_A(p->frame_c, Lc);
  DV_A(p->kv_end, L);
  DV_A(p->pos_x, static_cast<long long>(Lv) * D);
  DV_A(p->pos_h, static_cast<long long>(p->hist_tokens ? p->hist_tokens : 1) * D);
  DV_A(p->x, static_cast<long long>(B) * Lv * D);
  DV_A(p->c, static_cast<long long>(B) * Lc * D);
  DV_A(p->key_bias, static_cast<long long>(B) * p->Lpad);
  DV_A(p->tile_dead, static_cast<long long>(B) * (p->Lpad / 128));
  DV_A(p->tfeat, B * 256);
  DV_A(p->g1, B * D);
  DV_A(p->g3, B * D);
  DV_A(p->temb, B * D);
  DV_A(p->mod, static_cast<long long>(B) * m->w.mod_rows);
  DV_A(p->cc, 12);
  DV_A(p->xn, static_cast<long long>(B) * Lv * D);
  DV_A(p->cn, static_cast<long long>(B) * Lc * D);
  DV_A(p->qkv, static_cast<long long>(B) * L * 3 * D);
  DV_A(p->attn, static_cast<long long>(B) * L * D);
  DV_A(p->ffh, static_cast<long long>(B) * Lv * 4 * D);
  DV_A(p->ffh_c, static_cast<long long>(B) * Lc * 4 * D);
  DV_A(p->patch_a, static_cast<long long>(B) * Lv * Kp);
  DV_A(p->hist_a, static_cast<long long>(B) * (p->hist_tokens ? p->hist_tokens : 1) * Kp);
  DV_A(p->enc_bf, static_cast<long long>(B) * text_len * m->cfg.joint_dim);
  DV_A(p->xo, static_cast<long long>(B) * p->n_last * D);
#undef DV_A

  // token -> frame ids (mmdit.py:1336-1356: cumulative over clips; context = 0) and the
  // visible-key prefix of every query (frame(q) >= frame(k), mmdit.py:1430-1433)
  std::vector<int> fx(Lv), kv(L);
  std::vector<int> frame_end;  // tokens (joint index) with frame <= f
  int fbase = 0;
  for (const ClipInfo& ci : p->clips) {
    for (int t = 0; t < ci.t; ++t) {
      for (int r = 0; r < ci.gh * ci.gw; ++r) fx[ci.row0 + t * ci.gh * ci.gw + r] = fbase + t;
      frame_end.push_back(Lc + ci.row0 + (t + 1) * ci.gh * ci.gw);
    }
    fbase += ci.t;
  }
  if (fbase > kMaxFrames) {
    set_error("dv_mmdit_plan_create: %d frames > %d", fbase, kMaxFrames);
    return fail(DV_ERR_INVALID);
  }
  for (int i = 0; i < Lc; ++i) kv[i] = frame_end[0];
  for (int i = 0; i < Lv; ++i) kv[Lc + i] = frame_end[fx[i]];
  cudaError_t e = cudaMemcpy(p->frame_x, fx.data(), Lv * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(p->kv_end, kv.data(), L * sizeof(int), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    set_error("dv_mmdit_plan_create: upload: %s", cudaGetErrorString(e));
    return fail(DV_ERR_CUDA);
  }
  // positional tables (mmdit.py:841-880): crop at the noisy clip's size, resize per clip
  for (const ClipInfo& ci : p->clips) {
    rc = launch_pos_clip(m->pos_base, S, D, p->pos_x, ci.row0, ci.t, ci.gh, ci.gw, last.gh, last.gw, 0);
    if (rc) return fail(rc);
  }
  if (p->hist_tokens) {
    // forward_history_v2 (mmdit.py:985-993): crop at the down-sampled history size, no resize
    rc = launch_pos_clip(m->pos_base, S, D, p->pos_h, 0, 1, hist_h / 4, hist_w / 4, hist_h / 4,
                         hist_w / 4, 0);
    if (rc) return fail(rc);
  }
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    set_error("dv_mmdit_plan_create: %s", cudaGetErrorString(e));
    return fail(DV_ERR_CUDA);
  }

  // algorithmic FLOPs
  {
    const double d = D, NL = m->cfg.num_layers, Bd = B;
    double lin = 2.0 * Bd * (Lv * 12.0 * d * d * NL + Lc * (12.0 * d * d * (NL - 1) + 3.0 * d * d));
    lin += 2.0 * Bd * Lv * (m->cfg.in_channels * 4.0) * d + 2.0 * Bd * p->n_last * (m->cfg.in_channels * 4.0) * d;
    lin += 2.0 * Bd * text_len * m->cfg.joint_dim * d;
    lin += 2.0 * Bd * p->hist_tokens * (m->cfg.in_channels * 4.0) * d;
    lin += 2.0 * Bd * d * m->w.mod_rows + 2.0 * Bd * (256.0 * d + d * d + m->cfg.pooled_dim * d + d * d);
    double pairs = 0;
    for (int i = 0; i < L; ++i) pairs += kv[i];
    const double attn = 4.0 * d * NL * Bd * pairs;
    p->flops = lin + attn;
    p->attn_flops_layer = 4.0 * d * Bd * pairs;
  }
  *out = p;
  return DV_OK;
}

extern "C" void dv_mmdit_plan_destroy(dv_mmdit_plan* p) {
  if (!p) return;
  if (p->graph_exec) cudaGraphExecDestroy(p->graph_exec);
  if (p->cap_stream) cudaStreamDestroy(p->cap_stream);
  for (void* a : p->allocs) cudaFree(a);
  delete p;
}

extern "C" int dv_mmdit_plan_set_sp(dv_mmdit_plan* p, int sp_rank, int sp_world, dv_exchange_fn fn,
                                    void* user) {
  DV_REQUIRE(p, "dv_mmdit_plan_set_sp: null plan");
  DV_REQUIRE(sp_world >= 1 && sp_rank >= 0 && sp_rank < sp_world, "dv_mmdit_plan_set_sp: rank %d of %d",
             sp_rank, sp_world);
  const int H = p->m->cfg.num_heads;
  DV_REQUIRE(H % sp_world == 0, "dv_mmdit_plan_set_sp: %d heads do not split over %d ranks", H, sp_world);
  DV_REQUIRE(p->Lv % sp_world == 0, "dv_mmdit_plan_set_sp: %d video tokens do not split over %d ranks",
             p->Lv, sp_world);
  DV_REQUIRE(sp_world == 1 || fn != nullptr, "dv_mmdit_plan_set_sp: exchange function missing");
  if (sp_world > 1 && p->sp_send == nullptr) {
    const int D = p->m->D;
    // sized for every sp_world >= 2 (a cached plan may be re-sharded): q|k|v of my rows (bf16), the attention rows
    // of all context + my video rows (bf16), the fp32 stream
    const long long half = (p->Lv + 1) / 2;
    long long bytes = static_cast<long long>(p->B) * half * 3 * D * 2;
    const long long b2 = static_cast<long long>(p->B) * (p->Lc + half) * D * 2;
    const long long b3 = static_cast<long long>(p->B) * p->Lv * D * 4;
    bytes = bytes > b2 ? bytes : b2;
    bytes = bytes > b3 ? bytes : b3;
    char *s0 = nullptr, *s1 = nullptr;
    int rc = dev_alloc(p, &s0, bytes);
    if (rc == 0) rc = dev_alloc(p, &s1, bytes);
    if (rc == 0) rc = dev_alloc(p, &p->sp_flags, 8);
    if (rc == 0) rc = dev_alloc(p, &p->sp_epoch, 1);
    if (rc) return rc;
    p->sp_send = s0;
    p->sp_recv = s1;
  }
  if (sp_world != p->sp_world || sp_rank != p->sp_rank) {   // a different sharding: peers and the captured graph are stale
    p->peers_set = false;
    p->dev_barrier = false;
    if (p->graph_exec) {
      cudaGraphExecDestroy(p->graph_exec);
      p->graph_exec = nullptr;
    }
  }
  p->sp_rank = sp_rank;
  p->sp_world = sp_world;
  p->sp_exchange = fn;
  p->sp_user = user;
  return DV_OK;
}

extern "C" int dv_mmdit_plan_buffers(dv_mmdit_plan* p, void** qkv_dev, void** attn_dev, void** x_dev,
                                     void** flags_dev) {
  DV_REQUIRE(p && qkv_dev && attn_dev, "dv_mmdit_plan_buffers: null argument");
  *qkv_dev = p->qkv;
  *attn_dev = p->attn;
  if (x_dev) *x_dev = p->x;
  if (flags_dev) *flags_dev = p->sp_flags;   // null before dv_mmdit_plan_set_sp(sp_world > 1)
  return DV_OK;
}

extern "C" int dv_mmdit_plan_set_sp_peers(dv_mmdit_plan* p, void* const* qkv_ptrs, void* const* attn_ptrs,
                                          void* const* x_ptrs, void* const* flag_ptrs) {
  DV_REQUIRE(p, "dv_mmdit_plan_set_sp_peers: null plan");
  if (p->graph_exec) {   // the captured forward holds the old pointers
    cudaGraphExecDestroy(p->graph_exec);
    p->graph_exec = nullptr;
  }
  p->dev_barrier = false;
  if (qkv_ptrs == nullptr || attn_ptrs == nullptr) {
    p->peers_set = false;
    return DV_OK;
  }
  DV_REQUIRE(p->sp_world > 1 && p->sp_world <= 8, "dv_mmdit_plan_set_sp_peers: call dv_mmdit_plan_set_sp first (sp_world=%d)",
             p->sp_world);
  for (int i = 0; i < p->sp_world; ++i) {
    DV_REQUIRE(qkv_ptrs[i] && attn_ptrs[i], "dv_mmdit_plan_set_sp_peers: null pointer for rank %d", i);
    p->qkv_peer[i] = qkv_ptrs[i];
    p->attn_peer[i] = attn_ptrs[i];
  }
  DV_REQUIRE(p->qkv_peer[p->sp_rank] == p->qkv && p->attn_peer[p->sp_rank] == p->attn,
             "dv_mmdit_plan_set_sp_peers: entry %d must be this plan's own buffers", p->sp_rank);
  p->peers_set = true;
  if (x_ptrs != nullptr && flag_ptrs != nullptr) {
    for (int i = 0; i < p->sp_world; ++i) {
      DV_REQUIRE(x_ptrs[i] && flag_ptrs[i], "dv_mmdit_plan_set_sp_peers: null stream / flag pointer for rank %d", i);
      p->x_peer[i] = x_ptrs[i];
      p->flag_peer[i] = flag_ptrs[i];
    }
    DV_REQUIRE(p->x_peer[p->sp_rank] == p->x && p->flag_peer[p->sp_rank] == p->sp_flags,
               "dv_mmdit_plan_set_sp_peers: entry %d must be this plan's own stream / flags", p->sp_rank);
    p->dev_barrier = true;
  }
  return DV_OK;
}

extern "C" int dv_ipc_get_handle(const void* dev_ptr, void* handle64) {
  DV_REQUIRE(dev_ptr && handle64, "dv_ipc_get_handle: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  DV_CHECK_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)));
  memcpy(handle64, &h, sizeof(h));
  return DV_OK;
}

extern "C" int dv_ipc_open_handle(const void* handle64, void** dev_ptr) {
  DV_REQUIRE(handle64 && dev_ptr, "dv_ipc_open_handle: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  DV_CHECK_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return DV_OK;
}

extern "C" int dv_ipc_close_handle(void* dev_ptr) {
  DV_REQUIRE(dev_ptr, "dv_ipc_close_handle: null pointer");
  DV_CHECK_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return DV_OK;
}

extern "C" long long dv_mmdit_plan_workspace_bytes(const dv_mmdit_plan* p) { return p ? p->bytes : 0; }
extern "C" double dv_mmdit_plan_flops(const dv_mmdit_plan* p) { return p ? p->flops : 0.0; }

static int forward_body(dv_mmdit_plan* p, const void* const* clips_dev, int io_dtype,
                        const void* enc_dev, int enc_dtype, const float* ctx_mask_dev,
                        const float* pooled_dev, const float* timestep_dev,
                        const void* history_dev, void* out_dev, int out_dtype, cudaStream_t st) {
  dv_mmdit* m = p->m;
  const dv_mmdit_weights& w = m->w;
  const int B = p->B, D = m->D, Lv = p->Lv, Lc = p->Lc, L = p->L;
  const int C = m->cfg.in_channels, Kp = m->cfg.patch_k_pad, NL = m->cfg.num_layers;
  const int MR = w.mod_rows;
  const int is_bf16 = io_dtype == DV_DTYPE_BF16;
  int rc;
#define DV_RUN(call) \
  if ((rc = (call)) != 0) return rc

  // ---- conditioning: temb and every adaLN modulation vector -------------------------
  // all of it is a pure function of (timestep, pooled embedding) per batch row: looked up in the model's cache
  // first; the GEMVs return at once when every row was found, and found rows are copied from their slots
  const int* need = nullptr;
  if (m->cc_slots > 0 && MR % 4 == 0) {
    DV_RUN(launch_cond_lookup(timestep_dev, pooled_dev, m->cfg.pooled_dim, B, m->cc_keys, m->cc_valid, m->cc_slots,
                              p->cc, p->cc + 4, p->cc + 8, st));
    need = p->cc + 4;
  }
  DV_RUN(launch_timestep_features(timestep_dev, p->tfeat, B, st));
  DV_RUN(launch_gemv(reinterpret_cast<const __nv_bfloat16*>(w.w_t1), w.b_t1, p->tfeat, 256, p->g1, D,
                     B, D, 256, 0, 0, st, need));
  DV_RUN(launch_gemv(reinterpret_cast<const __nv_bfloat16*>(w.w_t2), w.b_t2, p->g1, D, p->temb, D, B,
                     D, D, 1, 0, st, need));
  DV_RUN(launch_gemv(reinterpret_cast<const __nv_bfloat16*>(w.w_p1), w.b_p1, pooled_dev,
                     m->cfg.pooled_dim, p->g3, D, B, D, m->cfg.pooled_dim, 0, 0, st, need));
  DV_RUN(launch_gemv(reinterpret_cast<const __nv_bfloat16*>(w.w_p2), w.b_p2, p->g3, D, p->temb, D, B,
                     D, D, 1, 1, st, need));
  DV_RUN(launch_gemv(reinterpret_cast<const __nv_bfloat16*>(w.w_mod), w.b_mod, p->temb, D, p->mod, MR,
                     B, MR, D, 1, 0, st, need));
  if (need != nullptr) DV_RUN(launch_cond_finish(p->mod, MR, m->cc_tables, p->cc, p->cc + 4, p->cc + 8, B, st));

  // ---- context stream: [history tokens | text tokens] ---------------------------------
  DV_RUN(launch_to_bf16(enc_dev, enc_dtype == DV_DTYPE_BF16, p->enc_bf,
                        static_cast<long long>(B) * p->text_len * m->cfg.joint_dim, st));
  {
    GemmDesc d = dense_desc(p->enc_bf, static_cast<long long>(p->text_len) * m->cfg.joint_dim,
                            m->cfg.joint_dim, w.w_ctx, D, w.b_ctx, B, p->text_len, D,
                            m->cfg.joint_dim);
    d.mode = EPI_F32_ADD;
    d.out = p->c;
    d.out_batch_stride = static_cast<long long>(Lc) * D;
    d.ldo = D;
    d.out_row_offset = p->hist_tokens;
    DV_RUN(launch_gemm(d, st));
  }
  if (p->hist_tokens) {
    DV_RUN(launch_patchify(history_dev, is_bf16, p->hist_a, p->hist_tokens, 0, Kp, B, C, 1,
                           p->hist_h, p->hist_w, 1, st));
    GemmDesc d = dense_desc(p->hist_a, static_cast<long long>(p->hist_tokens) * Kp, Kp,
                            w.w_patch_hist, D, w.b_patch_hist, B, p->hist_tokens, D, Kp);
    d.mode = EPI_F32_ADD;
    d.out = p->c;
    d.out_batch_stride = static_cast<long long>(Lc) * D;
    d.ldo = D;
    d.addend = p->pos_h;
    DV_RUN(launch_gemm(d, st));
  }
  // ---- video stream: patch embed + positional table -------------------------------------
  for (size_t i = 0; i < p->clips.size(); ++i) {
    const ClipInfo& ci = p->clips[i];
    DV_RUN(launch_patchify(clips_dev[i], is_bf16, p->patch_a, Lv, ci.row0, Kp, B, C, ci.t, ci.h,
                           ci.w, 0, st));
  }
  {
    GemmDesc d = dense_desc(p->patch_a, static_cast<long long>(Lv) * Kp, Kp, w.w_patch, D,
                            w.b_patch, B, Lv, D, Kp);
    d.mode = EPI_F32_ADD;
    d.out = p->x;
    d.out_batch_stride = static_cast<long long>(Lv) * D;
    d.ldo = D;
    d.addend = p->pos_x;
    DV_RUN(launch_gemm(d, st));
  }
  DV_RUN(launch_key_bias(ctx_mask_dev, Lc, p->key_bias, p->tile_dead, B, L, p->Lpad, st));

  const long long xs = static_cast<long long>(Lv) * D, cs = static_cast<long long>(Lc) * D;
  const long long js = static_cast<long long>(L) * D;  // joint (attention output) batch stride
  // Ulysses: this rank's window of the video rows (the whole stream when sp_world == 1); every
  // per-token step below runs on the window, the buffers keep their full-sequence layout
  const int P = p->sp_world, R = p->sp_rank;
  const int Lw = Lv / P, v0 = R * Lw;
  const int Hc = D / P;  // attention columns (heads * 64) per rank
  float* xw = p->x + static_cast<long long>(v0) * D;
  __nv_bfloat16* xnw = p->xn + static_cast<long long>(v0) * D;
  __nv_bfloat16* ffhw = p->ffh + static_cast<long long>(v0) * 4 * D;
  auto exchange = [&](long long bytes_per_peer) -> int {
    int erc = p->sp_exchange(p->sp_user, p->sp_send, p->sp_recv, bytes_per_peer, st);
    if (erc != 0 && last_error()[0] == 0) set_error("dv_mmdit_forward: exchange callback failed (%d)", erc);
    return erc;
  };
  const bool peer = P > 1 && p->peers_set;
  const bool devbar = peer && p->dev_barrier;
  // barrier of the peer-memory variant: flag stores + polling on the device, or NCCL's one-word all-reduce
  auto sp_sync = [&]() -> int {
    if (devbar) return launch_sp_barrier(p->flag_peer, p->sp_epoch, R, P, st);
    return exchange(0);
  };
  // ---- transformer blocks ------------------------------------------------------------------
  // Every step of a joint block covers the video and the context stream at once (the reference runs them as separate
  // modules, mmdit.py:385-433).  The six LN / GEMM steps between two attentions are described as phases and either
  // launched one by one or, for token layouts whose GEMMs have fewer tiles than SMs, handed to the persistent block
  // kernel in one launch (gemm.cu, pbk_kernel).
  auto ln_phase = [&](int i, bool second) {
    const bool last = (i == NL - 1);
    const float* mx = p->mod + static_cast<long long>(i) * 12 * D;  // video modulation, 6 chunks
    const float* mc = mx + 6 * D;                                   // context modulation
    PbkPhaseIn ph = {};
    ph.kind = 0;
    ph.mod_bs = MR;
    ph.B = B;
    ph.eps = 1e-6f;
    if (!second) {
      // norm1 / norm1_context: chunk order shift, scale, gate (msa), shift, scale, gate (mlp);
      // last block: AdaLayerNormContinuous on the context, chunk order scale, shift (mmdit.py:513)
      ph.ln0 = LnRows{xw, xs, xnw, xs, mx + 0 * D, mx + 1 * D, Lw};
      ph.ln1 = LnRows{p->c, cs, p->cn, cs, last ? mc + 1 * D : mc + 0 * D, last ? mc + 0 * D : mc + 1 * D, Lc};
      ph.has_ln1 = 1;
    } else {
      ph.ln0 = LnRows{xw, xs, xnw, xs, mx + 3 * D, mx + 4 * D, Lw};
      ph.ln1 = LnRows{p->c, cs, p->cn, cs, mc + 3 * D, mc + 4 * D, Lc};
      ph.has_ln1 = last ? 0 : 1;
    }
    return ph;
  };
  // fused q|k|v projection + per-head RMSNorm + temporal RoPE, written into the joint layout
  auto qkv_phase = [&](int i) {
    PbkPhaseIn ph = {};
    ph.kind = 1;
    ph.has_g1 = 1;
    for (int s = 0; s < 2; ++s) {
      const bool vid = (s == 0);
      GemmDesc d = dense_desc(vid ? xnw : p->cn, vid ? xs : cs, D,
                              vid ? m->p_w_qkv_x[i] : m->p_w_qkv_c[i], 3 * D,
                              vid ? m->p_b_qkv_x[i] : m->p_b_qkv_c[i], B, vid ? Lw : Lc, 3 * D, D);
      d.mode = EPI_QKV;
      d.out = p->qkv;
      d.out_batch_stride = static_cast<long long>(L) * 3 * D;
      d.ldo = 3 * D;
      d.out_row_offset = vid ? Lc + v0 : 0;
      d.qk_norm_w = vid ? m->p_qk_norm_x[i] : m->p_qk_norm_c[i];
      d.rope_cs = m->rope_cs;
      d.frame_id = vid ? p->frame_x + v0 : p->frame_c;
      d.heads_dim = D;
      if (vid && P > 1 && p->peers_set) {  // each head straight into the buffer of the rank that owns it
        for (int j = 0; j < P; ++j) d.out_peer[j] = p->qkv_peer[j];
        d.peer_cols = Hc;
      }
      (vid ? ph.g0 : ph.g1) = d;
    }
    return ph;
  };
  // x += gate_msa * to_out(attn);  c += c_gate_msa * to_add_out(attn_c)  (not in the last block)
  auto out_phase = [&](int i) {
    const bool last = (i == NL - 1);
    const float* mx = p->mod + static_cast<long long>(i) * 12 * D;
    const float* mc = mx + 6 * D;
    PbkPhaseIn ph = {};
    ph.kind = 1;
    GemmDesc dx = dense_desc(p->attn + static_cast<long long>(Lc + v0) * D, js, D, m->p_w_out_x[i], D,
                             m->p_b_out_x[i], B, Lw, D, D);
    dx.mode = EPI_RESID_GATE;
    dx.out = xw;
    dx.out_batch_stride = xs;
    dx.ldo = D;
    dx.gate = mx + 2 * D;
    dx.gate_batch_stride = MR;
    ph.g0 = dx;
    if (!last) {
      GemmDesc dc = dense_desc(p->attn, js, D, m->p_w_out_c[i], D, m->p_b_out_c[i], B, Lc, D, D);
      dc.mode = EPI_RESID_GATE;
      dc.out = p->c;
      dc.out_batch_stride = cs;
      dc.ldo = D;
      dc.gate = mc + 2 * D;
      dc.gate_batch_stride = MR;
      ph.g1 = dc;
      ph.has_g1 = 1;
    }
    return ph;
  };
  // feed-forward, both streams: W1 + GELU (which = 0), W2 + gated residual (which = 1)
  auto ff_phase = [&](int i, int which) {
    const bool last = (i == NL - 1);
    const float* mx = p->mod + static_cast<long long>(i) * 12 * D;
    const float* mc = mx + 6 * D;
    PbkPhaseIn ph = {};
    ph.kind = 1;
    ph.has_g1 = last ? 0 : 1;
    for (int s = 0; s < 2; ++s) {
      const bool vid = (s == 0);
      const float* mm = vid ? mx : mc;
      float* res = vid ? xw : p->c;
      __nv_bfloat16* nrm = vid ? xnw : p->cn;
      __nv_bfloat16* hid = vid ? ffhw : p->ffh_c;
      const long long rs = vid ? xs : cs;
      const int Ls = vid ? Lw : Lc;
      const long long hs = static_cast<long long>(vid ? Lv : Lc) * 4 * D;  // hidden batch stride
      GemmDesc d;
      if (which == 0) {
        d = dense_desc(nrm, rs, D, vid ? m->p_w_ff1_x[i] : m->p_w_ff1_c[i], 4 * D,
                       vid ? m->p_b_ff1_x[i] : m->p_b_ff1_c[i], B, Ls, 4 * D, D);
        d.mode = EPI_GELU;
        d.out = hid;
        d.out_batch_stride = hs;
        d.ldo = 4 * D;
      } else {
        d = dense_desc(hid, hs, 4 * D, vid ? m->p_w_ff2_x[i] : m->p_w_ff2_c[i], D,
                       vid ? m->p_b_ff2_x[i] : m->p_b_ff2_c[i], B, Ls, D, 4 * D);
        d.mode = EPI_RESID_GATE;
        d.out = res;
        d.out_batch_stride = rs;
        d.ldo = D;
        d.gate = mm + 5 * D;
        d.gate_batch_stride = MR;
      }
      (vid ? ph.g0 : ph.g1) = d;
    }
    return ph;
  };
  auto run_phase = [&](const PbkPhaseIn& ph) -> int {
    if (ph.kind == 0) return launch_ln_modulate2(ph.ln0, ph.has_ln1 ? &ph.ln1 : nullptr, ph.mod_bs, ph.B, D, ph.eps, st);
    return launch_gemm_pair(ph.g0, ph.has_g1 ? &ph.g1 : nullptr, st);
  };
  // persistent block kernel (EXPERIMENT, off by default): DV_MMDIT_PBK=1 uses it for layouts with at most
  // DV_PBK_MAX_ROWS (default 512) rows per rank.  Correct (the parity suites pass with it) but not faster: 122 us for the
  // six steps of the smallest block against ~120 us as six launches — a grid barrier + pipeline refill per phase costs what
  // a dependent launch costs (profiles/r02k_summary.txt, r02l_summary.txt).  Read per call so that tests can switch it.
  const char* pbk_s = getenv("DV_MMDIT_PBK");
  const char* pbk_r = getenv("DV_PBK_MAX_ROWS");
  const int pbk_max_rows = pbk_r ? atoi(pbk_r) : 512;
  const bool use_pbk = pbk_s != nullptr && atoi(pbk_s) != 0 && m->pbk_work != nullptr && B * (Lw + Lc) <= pbk_max_rows;
  auto run_phases = [&](const PbkPhaseIn* ph, int n) -> int {
    if (use_pbk) return launch_pbk(ph, n, m->pbk_work, m->pbk_bar, st);
    for (int j = 0; j < n; ++j) {
      const int prc = run_phase(ph[j]);
      if (prc) return prc;
    }
    return 0;
  };
  {
    PbkPhaseIn head[2] = {ln_phase(0, false), qkv_phase(0)};
    DV_RUN(run_phases(head, 2));
  }
  for (int i = 0; i < NL; ++i) {
    if (peer) {
      DV_RUN(sp_sync());  // barrier: every rank's q|k|v stores have landed
    } else if (P > 1) {
      // my rows, every rank's heads  ->  every rank's rows, my heads
      DV_RUN(launch_sp_qkv_pack(p->qkv, p->sp_send, B, L, D, Lc + v0, Lw, P, Hc, st));
      DV_RUN(exchange(static_cast<long long>(B) * Lw * 3 * Hc * 2));
      DV_RUN(launch_sp_qkv_unpack(p->sp_recv, p->qkv, B, L, D, Lc, Lw, P, Hc, R, st));
    }
    DV_RUN(launch_attention(p->qkv, p->attn, p->kv_end, p->key_bias, p->tile_dead, B, L, p->Lpad,
                            m->cfg.num_heads, st, p->attn_flops_layer / P, R * (m->cfg.num_heads / P),
                            m->cfg.num_heads / P, peer ? p->attn_peer : nullptr, peer ? P : 0, Lc, Lw));
    if (peer) {
      DV_RUN(sp_sync());  // barrier: every rank's attention rows have landed
    } else if (P > 1) {
      // all context rows + rank j's video rows of my heads -> rank j; back come the other heads
      DV_RUN(launch_sp_attn_pack(p->attn, p->sp_send, B, L, D, Lc, Lw, P, Hc, R, st));
      DV_RUN(exchange(static_cast<long long>(B) * (Lc + Lw) * Hc * 2));
      DV_RUN(launch_sp_attn_unpack(p->sp_recv, p->attn, B, L, D, Lc, Lw, P, Hc, R, st));
    }
    // the rest of block i and, up to its q|k|v projection, block i + 1
    PbkPhaseIn tail[6] = {out_phase(i), ln_phase(i, true), ff_phase(i, 0), ff_phase(i, 1)};
    int nt = 4;
    if (i + 1 < NL) {
      tail[nt++] = ln_phase(i + 1, false);
      tail[nt++] = qkv_phase(i + 1);
    }
    DV_RUN(run_phases(tail, nt));
  }

  if (P > 1) {
    // every rank runs the (cheap) output head itself: all-gather the rows of the noisy clip
    if (devbar) {
      const ClipInfo& lcl = p->clips.back();
      const int lo = v0 > lcl.row0 ? v0 : lcl.row0, hi = v0 + Lw;
      DV_RUN(launch_sp_x_share(p->x_peer, B, Lv, D, lo, hi > lo ? hi - lo : 0, R, P, st));
      DV_RUN(sp_sync());
    } else {
      DV_RUN(launch_sp_x_pack(p->x, p->sp_send, B, Lv, D, Lw, P, R, st));
      DV_RUN(exchange(static_cast<long long>(B) * Lw * D * 4));
      DV_RUN(launch_sp_x_unpack(p->sp_recv, p->x, B, Lv, D, Lw, P, R, st));
    }
  }
  // ---- head: norm_out (scale, shift) + proj_out + unpatchify, noisy clip only ---------------
  {
    const ClipInfo& lc = p->clips.back();
    const float* mo = p->mod + static_cast<long long>(NL - 1) * 12 * D + 6 * D + 2 * D;
    DV_RUN(launch_ln_modulate(p->x + static_cast<long long>(lc.row0) * D, xs, p->xo,
                              static_cast<long long>(p->n_last) * D, mo + 1 * D, mo + 0 * D, MR, B,
                              p->n_last, D, 1e-6f, st));
    GemmDesc d = dense_desc(p->xo, static_cast<long long>(p->n_last) * D, D, w.w_proj_out, 4 * C,
                            w.b_proj_out, B, p->n_last, 4 * C, D);
    d.mode = EPI_UNPATCH;
    d.out = out_dev;
    d.out_batch_stride = static_cast<long long>(C) * lc.t * lc.h * lc.w;
    d.up_gh = lc.gh;
    d.up_gw = lc.gw;
    d.up_C = C;
    d.out_f32 = (out_dtype == DV_DTYPE_F32);
    DV_REQUIRE(lc.t == 1, "dv_mmdit_forward: noisy clip must have one latent frame (t=%d)", lc.t);
    DV_RUN(launch_gemm(d, st));
  }
#undef DV_RUN
  return DV_OK;
}

// One captured graph per plan (DV_MMDIT_GRAPH=1, experimental, off by default; not yet run on hardware):
// every input is first copied into plan-owned staging buffers (a few MB of device-to-device copies), so the
// ~170 launches of a forward replay as one graph launch whatever tensors the caller passes.  Not used with
// sequence parallelism (host exchange callback) or while the per-launch profiler is on.
static int forward_graph(dv_mmdit_plan* p, const void* const* clips_dev, int io_dtype, const void* enc_dev,
                         int enc_dtype, const float* ctx_mask_dev, const float* pooled_dev,
                         const float* timestep_dev, const void* history_dev, void* out_dev, int out_dtype,
                         cudaStream_t st) {
  dv_mmdit* m = p->m;
  const int B = p->B, C = m->cfg.in_channels;
  const size_t io_sz = io_dtype == DV_DTYPE_BF16 ? 2 : 4;
  const size_t enc_sz = enc_dtype == DV_DTYPE_BF16 ? 2 : 4;
  const size_t out_sz = out_dtype == DV_DTYPE_BF16 ? 2 : 4;
  const ClipInfo& lc = p->clips.back();
  const size_t enc_bytes = static_cast<size_t>(B) * p->text_len * m->cfg.joint_dim * enc_sz;
  const size_t hist_bytes = static_cast<size_t>(B) * C * p->hist_h * p->hist_w * io_sz;
  const size_t out_bytes = static_cast<size_t>(B) * C * lc.t * lc.h * lc.w * out_sz;
  if (p->stage_t == nullptr || p->graph_io_dtype != io_dtype || p->graph_enc_dtype != enc_dtype ||
      p->graph_out_dtype != out_dtype) {
    // (re)build the staging buffers for these dtypes; fp32-sized so that a dtype change fits as well
    if (p->stage_t == nullptr) {
      int rc;
      p->stage_clips.resize(p->clips.size());
      for (size_t i = 0; i < p->clips.size(); ++i) {
        const ClipInfo& ci = p->clips[i];
        float* buf = nullptr;
        if ((rc = dev_alloc(p, &buf, static_cast<long long>(B) * C * ci.t * ci.h * ci.w)) != 0) return rc;
        p->stage_clips[i] = buf;
      }
      float* f = nullptr;
      if ((rc = dev_alloc(p, &f, static_cast<long long>(B) * p->text_len * m->cfg.joint_dim)) != 0) return rc;
      p->stage_enc = f;
      if (p->hist_tokens) {
        if ((rc = dev_alloc(p, &f, static_cast<long long>(B) * C * p->hist_h * p->hist_w)) != 0) return rc;
        p->stage_hist = f;
      }
      if ((rc = dev_alloc(p, &f, static_cast<long long>(B) * C * lc.t * lc.h * lc.w)) != 0) return rc;
      p->stage_out = f;
      if ((rc = dev_alloc(p, &p->stage_mask, static_cast<long long>(B) * p->Lc)) != 0) return rc;
      if ((rc = dev_alloc(p, &p->stage_pooled, static_cast<long long>(B) * m->cfg.pooled_dim)) != 0) return rc;
      if ((rc = dev_alloc(p, &p->stage_t, B)) != 0) return rc;
    }
    if (p->graph_exec) {
      cudaGraphExecDestroy(p->graph_exec);
      p->graph_exec = nullptr;
    }
    p->graph_io_dtype = io_dtype;
    p->graph_enc_dtype = enc_dtype;
    p->graph_out_dtype = out_dtype;
  }
  for (size_t i = 0; i < p->clips.size(); ++i) {
    const ClipInfo& ci = p->clips[i];
    DV_CHECK_CUDA(cudaMemcpyAsync(p->stage_clips[i], clips_dev[i], static_cast<size_t>(B) * C * ci.t * ci.h * ci.w * io_sz,
                                  cudaMemcpyDeviceToDevice, st));
  }
  DV_CHECK_CUDA(cudaMemcpyAsync(p->stage_enc, enc_dev, enc_bytes, cudaMemcpyDeviceToDevice, st));
  if (p->hist_tokens) DV_CHECK_CUDA(cudaMemcpyAsync(p->stage_hist, history_dev, hist_bytes, cudaMemcpyDeviceToDevice, st));
  DV_CHECK_CUDA(cudaMemcpyAsync(p->stage_mask, ctx_mask_dev, static_cast<size_t>(B) * p->Lc * 4, cudaMemcpyDeviceToDevice, st));
  DV_CHECK_CUDA(cudaMemcpyAsync(p->stage_pooled, pooled_dev, static_cast<size_t>(B) * m->cfg.pooled_dim * 4,
                                cudaMemcpyDeviceToDevice, st));
  DV_CHECK_CUDA(cudaMemcpyAsync(p->stage_t, timestep_dev, static_cast<size_t>(B) * 4, cudaMemcpyDeviceToDevice, st));
  if (p->graph_exec == nullptr) {
    // capture on the plan's own stream: the caller's stream may be the legacy default stream, which cannot
    // be captured; nothing executes during capture, so no ordering with `st` is needed here
    if (p->cap_stream == nullptr) DV_CHECK_CUDA(cudaStreamCreateWithFlags(&p->cap_stream, cudaStreamNonBlocking));
    std::vector<const void*> cl(p->stage_clips.begin(), p->stage_clips.end());
    const long long before = launch_count();
    DV_CHECK_CUDA(cudaStreamBeginCapture(p->cap_stream, cudaStreamCaptureModeRelaxed));
    const int rc = forward_body(p, cl.data(), io_dtype, p->stage_enc, enc_dtype, p->stage_mask, p->stage_pooled,
                                p->stage_t, p->hist_tokens ? p->stage_hist : nullptr, p->stage_out, out_dtype,
                                p->cap_stream);
    cudaGraph_t g = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(p->cap_stream, &g);
    p->graph_launches = launch_count() - before;
    note_launch(static_cast<int>(-p->graph_launches));  // nothing ran yet: the replay below counts them
    if (rc != 0) {
      if (g) cudaGraphDestroy(g);
      return rc;
    }
    DV_CHECK_CUDA(ce);
    DV_CHECK_CUDA(cudaGraphInstantiate(&p->graph_exec, g, 0));
    cudaGraphDestroy(g);
  }
  DV_CHECK_CUDA(cudaGraphLaunch(p->graph_exec, st));
  note_launch(static_cast<int>(p->graph_launches));
  DV_CHECK_CUDA(cudaMemcpyAsync(out_dev, p->stage_out, out_bytes, cudaMemcpyDeviceToDevice, st));
  return DV_OK;
}

extern "C" int dv_mmdit_forward(dv_mmdit_plan* p, const void* const* clips_dev, int io_dtype,
                                const void* enc_dev, int enc_dtype, const float* ctx_mask_dev,
                                const float* pooled_dev, const float* timestep_dev,
                                const void* history_dev, void* out_dev, int out_dtype,
                                void* stream_v) {
  DV_REQUIRE(p && clips_dev && enc_dev && ctx_mask_dev && pooled_dev && timestep_dev && out_dev,
             "dv_mmdit_forward: null argument");
  DV_REQUIRE((p->hist_tokens > 0) == (history_dev != nullptr),
             "dv_mmdit_forward: history pointer does not match the plan (plan has %d history tokens)",
             p->hist_tokens);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_v);
  // CUDA-graph replay: DV_MMDIT_GRAPH=1 always (when possible), =0 never; default: for sequence-parallel forwards
  // over peer memory (row windows make every launch latency-bound and the host issue rate the limit), not for
  // single-rank forwards (GPU-bound: measured 172.2 vs 171.6 ms per C2 step, profiles/r02a_summary.txt)
  static const char* genv = getenv("DV_MMDIT_GRAPH");
  static const int gmode = genv ? (atoi(genv) != 0 ? 1 : 0) : -1;
  const bool capturable = p->sp_world == 1 || (p->peers_set && p->dev_barrier);   // no host callback inside
  const bool use_graph = gmode == 1 || (gmode == -1 && p->sp_world > 1);
  if (use_graph && capturable && !prof_on())
    return forward_graph(p, clips_dev, io_dtype, enc_dev, enc_dtype, ctx_mask_dev, pooled_dev, timestep_dev,
                         history_dev, out_dev, out_dtype, st);
  return forward_body(p, clips_dev, io_dtype, enc_dev, enc_dtype, ctx_mask_dev, pooled_dev, timestep_dev,
                      history_dev, out_dev, out_dtype, st);
}

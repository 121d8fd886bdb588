// deepv_b200 — causal video-VAE decode (latent -> RGB / disparity frames) on the sm_100a kernels.
//
// Mirrors reference model/vae.py CausalVideoVAE.decode as the pipeline calls it
// (pipeline.py:713: temporal_chunk=True, window_size=1, tile_sample_min_size=256):
//   tiled_decode  (vae.py:989-1014)  6 overlapping 32x32-latent tiles, sequential in-place blends
//   chunk_decode  (vae.py:903-920)   temporal windows with a 2-frame conv cache
//   decoder       (vae.py:731-751)   conv_in, mid (res, attn, res), 4 up blocks, GN+SiLU, conv_out
// B200-first re-design: every tile is decoded in ONE pass over all latent frames (the causal
// zero padding + cache of the windowed reference is mathematically the same contraction —
// SURVEY.md App. E.2 — and gives 7x fewer, larger launches); activations are channels-last bf16 so
// that each 3x3x3 conv is an implicit GEMM whose A tiles are 5-D TMA boxes with hardware zero
// fill (gemm.cu); pixel-shuffle / frame-interleave are store-address maps in the conv epilogue;
// the 6-tile geometry and the blend ORDER are kept exactly (per-tile GroupNorm statistics and
// mid-block attention make tiling observable: 35.7 dB vs untiled).
#include <cmath>
#include <map>
#include <string>
#include <vector>

#include "../../include/deepv_b200.h"
#include "gemm.cuh"
#include "kernels.cuh"

using namespace dv;

namespace dv {
int launch_gn_stats(const __nv_bfloat16* x, double* acc, int frames, int HW, int C, int G,
                    cudaStream_t stream);
int launch_softmax_rows(const float* s, __nv_bfloat16* p, long long rows, int cols, float scale,
                        cudaStream_t stream);
int launch_transpose(const __nv_bfloat16* in, int ld_in, __nv_bfloat16* out, int frames, int rows,
                     int cols, cudaStream_t stream);
int launch_latent_tile(const void* z, int is_bf16, __nv_bfloat16* out, int C, int T, int h, int w,
                       int y0, int x0, int th, int tw, int Cpad, cudaStream_t stream);
// conv with <= 4 output channels from per-tap partial products (EPI_TAPS planes), vae.py:225-252
int launch_tap_gather(const float* planes, const float* bias, __nv_bfloat16* out, int T, int H, int W,
                      int Cout, int t_first, cudaStream_t stream);
}  // namespace dv

struct dv_vae {
  dv_vae_config cfg;
  std::map<std::string, const void*> t;
  const void* get(const std::string& n) const {
    auto it = t.find(n);
    return it == t.end() ? nullptr : it->second;
  }
};

struct TileOut {
  __nv_bfloat16* px;  // [Tout][H][W][C] (decoder: C = 3 frames; encoder: C = 64, 2z moments used)
  int H, W;
  int C = 3;
};

struct dv_vae_plan {
  dv_vae* v = nullptr;
  int T = 0, h = 0, w = 0, tile = 0, Tout = 0;
  int first_frame = 0;          // decode output frames [first_frame, Tout) only (dv_vae_plan_set_first_frame)
  int rows = 0, cols = 0;
  std::vector<int> ys, xs;      // latent tile origins
  std::vector<TileOut> tiles;   // rows * cols
  __nv_bfloat16* buf[4] = {nullptr, nullptr, nullptr, nullptr};
  long long buf_elems = 0;
  __nv_bfloat16 *qk = nullptr, *vt = nullptr, *pr = nullptr;
  float* sc = nullptr;
  // GroupNorm (sum, sumsq) accumulators: one slot of [frames][groups][2] fp64 per normalised tensor
  // of a tile run, zeroed by one memset per tile; filled by the producing conv's epilogue
  double* gn_acc = nullptr;
  int gn_slots = 0, gn_slot_elems = 0, gn_replica_elems = 0;
  static constexpr int kGnReplicas = 8;
  float* taps = nullptr;        // conv_out per-tap planes [32][T*H*W][4] fp32 of the largest tile
  long long taps_elems = 0;
  TileOut* tiles_dev = nullptr;
  double flops = 0;
  std::vector<void*> allocs;
};

namespace {

struct BlendArgs {
  const TileOut* tiles;
  int rows, cols, Tout, Hout, Wout, limit, extent;
  int nch = 3;  // channels written (decoder: 3; encoder: 2z moments)
  int t_first = 0;  // frames [t_first, Tout) are blended and written (trimmed decode)
};

template <int DEPTH>
__device__ float blended(const BlendArgs& a, int i, int j, int t, int y, int x, int c) {
  const TileOut& tl = a.tiles[i * a.cols + j];
  float v = __bfloat162float(tl.px[((static_cast<long long>(t) * tl.H + y) * tl.W + x) * tl.C + c]);
  if constexpr (DEPTH > 0) {
    if (i > 0) {  // blend_v with the (already blended) tile above, vae.py:942-946
      const TileOut& up = a.tiles[(i - 1) * a.cols + j];
      const int e = min(min(up.H, tl.H), a.extent);
      if (y < e) {
        const float av = blended<DEPTH - 1>(a, i - 1, j, t, up.H - e + y, x, c);
        const float wy = static_cast<float>(static_cast<double>(y) / e);
        const float wn = static_cast<float>(1.0 - static_cast<double>(y) / e);
        v = av * wn + v * wy;
      }
    }
    if (j > 0) {  // blend_h with the (already blended) tile on the left, vae.py:948-952
      const TileOut& lf = a.tiles[i * a.cols + j - 1];
      const int e = min(min(lf.W, tl.W), a.extent);
      if (x < e) {
        const float av = blended<DEPTH - 1>(a, i, j - 1, t, y, lf.W - e + x, c);
        const float wx = static_cast<float>(static_cast<double>(x) / e);
        const float wn = static_cast<float>(1.0 - static_cast<double>(x) / e);
        v = av * wn + v * wx;
      }
    }
  }
  return v;
}

template <typename T>
__global__ void blend_kernel(BlendArgs a, T* __restrict__ out) {
  // out: [nch][Tout][Hout][Wout]; x fastest -> coalesced stores
  const int nT = a.Tout - a.t_first;
  const long long total = static_cast<long long>(a.nch) * nT * a.Hout * a.Wout;
  long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int X = idx % a.Wout;
  long long r = idx / a.Wout;
  const int Y = r % a.Hout;
  r /= a.Hout;
  const int t = a.t_first + static_cast<int>(r % nT);
  const int c = r / nT;
  idx = ((static_cast<long long>(c) * a.Tout + t) * a.Hout + Y) * a.Wout + X;
  const int i = min(Y / a.limit, a.rows - 1), j = min(X / a.limit, a.cols - 1);
  // depth 3 covers every dependency chain of a row-major sweep (corner -> up -> up-left)
  const float v = blended<3>(a, i, j, t, Y - i * a.limit, X - j * a.limit, c);
  if constexpr (sizeof(T) == 2)
    out[idx] = __float2bfloat16(v);
  else
    out[idx] = v;
}

template <typename T>
int plan_alloc(dv_vae_plan* p, T** out, long long n) {
  void* ptr = nullptr;
  DV_CHECK_CUDA(cudaMalloc(&ptr, static_cast<size_t>(n) * sizeof(T)));
  p->allocs.push_back(ptr);
  *out = reinterpret_cast<T*>(ptr);
  return 0;
}

struct Act {  // channels-last activation
  __nv_bfloat16* p;
  int T, H, W, C;
  double* gn = nullptr;  // GroupNorm statistics accumulated by the producer, or null
  int t0 = 0;            // frames [t0, T) hold data (a trimmed decode computes only what the kept frames depend on)
  long long elems() const { return static_cast<long long>(T) * H * W * C; }
};

struct Runner {
  dv_vae_plan* pl;
  cudaStream_t st;
  int rc = 0;
  double flops = 0;
  bool dry = false;  // count flops / buffer sizes only
  long long max_elems = 0;
  int gn_slot = 0;   // next free accumulator slot of this tile run
  double* new_gn_slot() {
    const int i = gn_slot++;
    if (dry) {
      if (gn_slot > pl->gn_slots) pl->gn_slots = gn_slot;
      return nullptr;
    }
    return pl->gn_acc + static_cast<long long>(i) * pl->gn_slot_elems;
  }

  const void* W(const std::string& n) {
    const void* p = pl->v->get(n);
    if (!p && rc == 0) {
      set_error("vae: missing tensor '%s'", n.c_str());
      rc = DV_ERR_INVALID;
    }
    return p;
  }
  void note(const Act& a) { max_elems = a.elems() > max_elems ? a.elems() : max_elems; }

  // causal conv (vae.py:225-252) as implicit GEMM
  // `stats`: also accumulate the GroupNorm statistics of the output (its consumer is a norm)
  // `t0`: first frame of the CONV output to compute (before an interleave store doubles the frame index)
  Act conv(const std::string& name, const Act& x, int cout, int ks, int store, int drop_first,
           const __nv_bfloat16* residual, __nv_bfloat16* outbuf, int w_rows = -1, bool stats = false,
           int sT = 1, int sH = 1, int sW = 1, int t0 = 0) {
    Act y;
    y.p = outbuf;
    if ((std::max(0, t0 - (ks - 1)) < x.t0 || (t0 > 0 && sT * sH * sW != 1)) && rc == 0) {
      set_error("vae: conv %s from frame %d needs input frames from %d, have %d", name.c_str(), t0, t0 - (ks - 1), x.t0);
      rc = DV_ERR_INVALID;
    }
    y.T = (x.T - 1) / sT + 1;  // causal: (kt-1) zero frames in front, then a VALID conv (vae.py:229-231)
    y.H = x.H / sH;
    y.W = x.W / sW;
    y.C = cout;
    if (store == CONV_SHUFFLE_HW) {
      y.H *= 2;
      y.W *= 2;
      y.C = cout / 4;
    } else if (store == CONV_INTERLEAVE_T) {
      y.T = 2 * x.T - (drop_first ? 1 : 0);
      y.C = cout / 2;
    }
    y.t0 = store == CONV_INTERLEAVE_T ? std::max(0, 2 * t0 - (drop_first ? 1 : 0)) : t0;
    note(y);
    flops += 2.0 * ((x.T - 1) / sT + 1 - t0) * (x.H / sH) * (x.W / sW) * cout * static_cast<double>(ks * ks * ks) * x.C;
    const int G = pl->v->cfg.norm_groups;
    if (stats && y.C % G == 0 && (y.C / G == 4 || y.C / G == 8 || y.C / G == 16)) y.gn = new_gn_slot();
    if (dry || rc) return y;
    GemmDesc d = {};
    d.batch = 1;
    d.N = cout;
    d.gn_acc = y.gn;
    d.gn_cpg = y.gn ? y.C / G : 0;
    d.gn_replicas = dv_vae_plan::kGnReplicas;
    d.gn_replica_stride = pl->gn_replica_elems;
    d.A = x.p;
    d.a_mode = 1;
    d.cT = x.T;
    d.cH = x.H;
    d.cW = x.W;
    d.cC = x.C;
    d.kt = d.kh = d.kw = ks;
    d.sT = sT;
    d.sH = sH;
    d.sW = sW;
    d.W = W(name + ".weight");
    d.w_rows = w_rows > 0 ? w_rows : cout;
    d.bias = reinterpret_cast<const float*>(W(name + ".bias"));
    d.mode = EPI_CONV;
    d.out = outbuf;
    d.conv_store = store;
    d.conv_drop_first = drop_first;
    d.conv_t0 = t0;
    d.residual = residual;
    d.out_C = y.C;
    if (rc == 0) rc = launch_gemm(d, st);
    return y;
  }

  // per-frame GroupNorm (+ SiLU), vae.py:161-167,298-304
  Act gn(const std::string& name, const Act& x, bool act, __nv_bfloat16* outbuf) {
    Act y = x;
    y.p = outbuf;
    y.gn = nullptr;
    if (dry && x.gn == nullptr) new_gn_slot();  // (sizes the accumulator ring)
    if (dry || rc) return y;
    const int G = pl->v->cfg.norm_groups;
    const double* acc = x.gn;
    // frames [x.t0, T): the statistics are per frame, indexed by the absolute frame
    const long long foff = static_cast<long long>(x.t0) * x.H * x.W * x.C;
    const long long aoff = static_cast<long long>(x.t0) * G * 2;
    if (acc == nullptr) {  // producer without fused statistics: one read pass
      double* slot = new_gn_slot();
      rc = launch_gn_stats(x.p + foff, slot + aoff, x.T - x.t0, x.H * x.W, x.C, G, st);
      acc = slot;
    }
    if (rc == 0)
      rc = launch_gn_apply(x.p + foff, acc + aoff, dv_vae_plan::kGnReplicas, pl->gn_replica_elems,
                           reinterpret_cast<const float*>(W(name + ".weight")),
                           reinterpret_cast<const float*>(W(name + ".bias")), outbuf + foff, x.T - x.t0, x.H * x.W,
                           x.C, G, 1e-6f, act ? 1 : 0, st);
    return y;
  }

  // CausalResnetBlock3D (vae.py:293-310); x lives in buf[ix], result goes to buf[iout]
  // `t_out`: first output frame anything downstream depends on; conv1 then starts two frames earlier and the input
  // must hold frames from t_out - 4 (need_in())
  Act resnet(const std::string& name, const Act& x, int cout, int ia, int ib, int iout, int t_out = 0) {
    __nv_bfloat16** B = pl->buf;
    const int t1 = std::max(0, t_out - 2);
    Act a = gn(name + ".norm1", x, true, B[ia]);
    Act h = conv(name + ".conv1.conv", a, cout, 3, CONV_PLAIN, 0, nullptr, B[ib], -1, true, 1, 1, 1, t1);
    Act a2 = gn(name + ".norm2", h, true, B[ia]);
    const __nv_bfloat16* res = x.p;
    if (x.C != cout) {
      Act sc = conv(name + ".conv_shortcut.conv", x, cout, 1, CONV_PLAIN, 0, nullptr, B[ib], -1, false, 1, 1, 1, t_out);
      res = sc.p;  // conv1's output (in B[ib]) has been consumed by norm2 already
    }
    return conv(name + ".conv2.conv", a2, cout, 3, CONV_PLAIN, 0, res, B[iout], -1, true, 1, 1, 1, t_out);
  }

  // diffusers Attention of the mid block (vae.py:439-445,463-467): per frame, one head of width C
  Act attention(const std::string& name, const Act& x, int ia, int iout) {
    __nv_bfloat16** B = pl->buf;
    const int C = x.C, pix = x.H * x.W, F = x.T;
    Act xn = gn(name + ".group_norm", x, false, B[ia]);
    flops += 2.0 * F * pix * C * C * 4 + 4.0 * F * pix * static_cast<double>(pix) * C;
    Act y = x;
    y.p = B[iout];
    y.gn = nullptr;  // the attention output has no fused statistics
    note(y);
    if (dry || rc) return y;
    auto dense = [&](const void* A, long long abs_, int lda, const void* Wp, int wrows, long long wbs,
                     int ldw, const float* bias, int M, int N, int K, void* out, long long obs,
                     int ldo, const void* residual, int mode) {
      GemmDesc d = {};
      d.batch = F;
      d.M = M;
      d.N = N;
      d.K = K;
      d.A = A;
      d.a_batch_stride = abs_;
      d.lda = lda;
      d.W = Wp;
      d.w_rows = wrows;
      d.w_batch_stride = wbs;
      d.ldw = ldw;
      d.bias = bias;
      d.mode = mode;
      d.out = out;
      d.out_batch_stride = obs;
      d.ldo = ldo;
      d.residual = residual;
      if (rc == 0) rc = launch_gemm(d, st);
    };
    const long long pc = static_cast<long long>(pix) * C;
    // q | k | v in one GEMM: [pix][3C]
    dense(xn.p, pc, C, W(name + ".to_qkv.weight"), 3 * C, 0, 0,
          reinterpret_cast<const float*>(W(name + ".to_qkv.bias")), pix, 3 * C, C, pl->qk, 3 * pc,
          3 * C, nullptr, EPI_BF16);
    // V^T [C][pix] per frame: the K-major B operand of P V
    if (rc == 0) rc = launch_transpose(pl->qk + 2 * C, 3 * C, pl->vt, F, pix, C, st);
    // S = Q K^T in fp32 (A = q columns, B = k columns of the same frame; rows are 3C apart)
    dense(pl->qk, 3 * pc, 3 * C, pl->qk + C, pix, 3 * pc, 3 * C, nullptr, pix, pix, C, pl->sc,
          static_cast<long long>(pix) * pix, pix, nullptr, EPI_F32_ADD);
    if (rc == 0)
      rc = launch_softmax_rows(pl->sc, pl->pr, static_cast<long long>(F) * pix, pix,
                               1.0f / sqrtf(static_cast<float>(C)), st);
    // O = P V
    dense(pl->pr, static_cast<long long>(pix) * pix, pix, pl->vt, C, pc, 0, nullptr, pix, C, pix,
          B[ia], pc, C, nullptr, EPI_BF16);
    // out projection + residual
    dense(B[ia], pc, C, W(name + ".to_out.0.weight"), C, 0, 0,
          reinterpret_cast<const float*>(W(name + ".to_out.0.bias")), pix, C, C, y.p, pc, C, x.p,
          EPI_BF16);
    return y;
  }
};

}  // namespace

extern "C" int dv_vae_create(const dv_vae_config* cfg, const dv_tensor_ref* tensors, int n_tensors,
                             dv_vae** out) {
  DV_REQUIRE(cfg && tensors && out, "dv_vae_create: null argument");
  DV_REQUIRE(cfg->norm_groups == 32, "dv_vae_create: norm_groups=%d (32 supported)", cfg->norm_groups);
  dv_vae* v = new dv_vae();
  v->cfg = *cfg;
  for (int i = 0; i < n_tensors; ++i) v->t[tensors[i].name] = tensors[i].ptr;
  *out = v;
  return DV_OK;
}

extern "C" void dv_vae_destroy(dv_vae* v) { delete v; }

static int run_tile(dv_vae_plan* p, const void* z, int z_bf16, int ti, int tj, cudaStream_t st,
                    bool dry, double* flops, long long* max_elems) {
  const dv_vae_config& c = p->v->cfg;
  Runner r;
  r.pl = p;
  r.st = st;
  r.dry = dry;
  const int y0 = p->ys[ti], x0 = p->xs[tj];
  const int th = std::min(p->tile, p->h - y0), tw = std::min(p->tile, p->w - x0);
  if (th % 8 || tw % 16) {
    set_error("vae: latent tile %dx%d must be a multiple of 8x16", th, tw);
    return DV_ERR_INVALID;
  }
  __nv_bfloat16** B = p->buf;
  Act x{B[0], p->T, th, tw, 64};
  r.note(x);
  if (!dry) {
    cudaError_t e = cudaMemsetAsync(p->gn_acc, 0,
                                    static_cast<size_t>(p->gn_slots + 1) * p->gn_slot_elems * sizeof(double), st);
    if (e != cudaSuccess) {
      set_error("vae: memset of the GroupNorm accumulators: %s", cudaGetErrorString(e));
      return DV_ERR_CUDA;
    }
    r.rc = launch_latent_tile(z, z_bf16, B[0], c.latent_channels, p->T, p->h, p->w, y0, x0, th, tw, 64, st);
  }
  const int top = c.block_channels[3];
  // post_quant_conv (1x1x1, output padded to 64 channels so it can feed conv_in's TMA boxes)
  x = r.conv("post_quant_conv.conv", x, 64, 1, CONV_PLAIN, 0, nullptr, B[1], 64);
  x = r.conv("decoder.conv_in.conv", x, top, 3, CONV_PLAIN, 0, nullptr, B[0], -1, true);
  // buffers: x in B[0]; scratch B[1], B[2]; result B[3] -> rotate
  int cur = 0;
  auto other = [&](int k) { return (cur + k) & 3; };
  // Trimmed decode (dv_vae_plan_set_first_frame): only output frames >= keep are wanted.  Every 3x3x3 causal conv looks
  // two frames back at its own temporal resolution, GroupNorm and the mid-block attention are per frame, so walking the
  // decoder backwards gives the first frame each conv has to compute; the results for the kept frames are bit-identical
  // to a full decode.  (A continuation iteration of the rollout drops the 25 re-decoded input frames, pipeline.py:327.)
  const int keep = p->first_frame;
  int res_t[4][8], sp_t0[4] = {0, 0, 0, 0}, tp_t0[4] = {0, 0, 0, 0};
  {
    int n = std::max(0, keep - 2);   // conv_out reads two frames back
    for (int i = 3; i >= 0; --i) {
      if (c.temporal_up[i]) {        // output frame o = 2 t - 1 + {0, 1}
        tp_t0[i] = (n + 1) / 2;
        n = std::max(0, tp_t0[i] - 2);
      }
      if (c.spatial_up[i]) {
        sp_t0[i] = n;
        n = std::max(0, n - 2);
      }
      for (int j = c.layers_per_block[i] - 1; j >= 0; --j) {
        res_t[i][j] = n;
        n = std::max(0, n - 4);      // conv2 and conv1
      }
    }
    // (everything in front of the up blocks runs on all frames)
  }
  {
    x = r.resnet("decoder.mid_block.resnets.0", x, top, other(1), other(2), other(3));
    cur = other(3);
    x = r.attention("decoder.mid_block.attentions.0", x, other(1), other(3));
    cur = other(3);
    x = r.resnet("decoder.mid_block.resnets.1", x, top, other(1), other(2), other(3));
    cur = other(3);
  }
  for (int i = 0; i < 4; ++i) {
    const int co = c.block_channels[3 - i];
    for (int j = 0; j < c.layers_per_block[i]; ++j) {
      x = r.resnet("decoder.up_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), x, co,
                   other(1), other(2), other(3), res_t[i][j]);
      cur = other(3);
    }
    if (c.spatial_up[i]) {
      x = r.conv("decoder.up_blocks." + std::to_string(i) + ".upsamplers.0.conv.conv", x, co * 4, 3,
                 CONV_SHUFFLE_HW, 0, nullptr, B[other(1)], -1, !c.temporal_up[i], 1, 1, 1, sp_t0[i]);
      cur = other(1);
    }
    if (c.temporal_up[i]) {
      x = r.conv("decoder.up_blocks." + std::to_string(i) + ".temporal_upsamplers.0.conv.conv", x,
                 co * 2, 3, CONV_INTERLEAVE_T, 1, nullptr, B[other(1)], -1, true, 1, 1, 1, tp_t0[i]);
      cur = other(1);
    }
  }
  x = r.gn("decoder.conv_norm_out", x, true, B[other(1)]);
  TileOut& to = p->tiles[ti * p->cols + tj];
  // conv_out (C -> 3): a full implicit GEMM would spend a 128-row MMA tile on 3 output channels.
  // Instead ONE K = C GEMM produces the 27 x 3 per-tap partial products of every pixel and a
  // gather kernel sums the 27 shifted planes (same terms, fp32 throughout).
  Act y = x;
  y.p = to.px;
  y.C = c.out_channels;
  {
    const long long npix = static_cast<long long>(x.T) * x.H * x.W;
    const int tA = std::max(0, keep - 2);                       // first frame whose tap planes are needed
    const long long pixA = static_cast<long long>(tA) * x.H * x.W;
    r.flops += 2.0 * (npix - pixA) * c.out_channels * 27.0 * x.C;
    if (dry) {
      if (32 * npix * 4 > p->taps_elems) p->taps_elems = 32 * npix * 4;
    } else if (r.rc == 0) {
      GemmDesc d = {};
      d.batch = 1;
      d.M = static_cast<int>(npix - pixA);
      d.N = 128;  // 27 taps x 4 columns = 108, padded to the 32-column epilogue chunk (weight rows
                  // beyond 108 are zero-filled by TMA; their planes are written but never read)
      d.K = x.C;
      d.A = x.p + pixA * x.C;
      d.a_batch_stride = npix * x.C;
      d.lda = x.C;
      d.W = r.W("decoder.conv_out.conv.weight_taps");
      d.w_rows = 27 * 4;
      d.mode = EPI_TAPS;
      d.out = p->taps + pixA * 4;   // planes are [tap][pixel][4] fp32 with the full pixel count as the plane stride
      d.ldo = static_cast<int>(npix);
      if (npix >= (1ll << 31) || x.C % 64 != 0) {
        set_error("conv_out: %lld pixels, C=%d", npix, x.C);
        r.rc = DV_ERR_INVALID;
      }
      if (r.rc == 0) r.rc = launch_gemm(d, st);
      if (r.rc == 0)
        r.rc = launch_tap_gather(p->taps, reinterpret_cast<const float*>(r.W("decoder.conv_out.conv.bias")),
                                 to.px, x.T, x.H, x.W, c.out_channels, std::min(keep, x.T - 1), st);
    }
  }
  if (dry) {
    to.H = y.H;
    to.W = y.W;
    p->Tout = y.T;
  }
  if (flops) *flops += r.flops;
  if (max_elems && r.max_elems > *max_elems) *max_elems = r.max_elems;
  return r.rc;
}

extern "C" int dv_vae_plan_create(dv_vae* v, int T, int h, int w, int tile_latent,
                                  dv_vae_plan** out) {
  DV_REQUIRE(v && out, "dv_vae_plan_create: null argument");
  DV_REQUIRE(T >= 1 && h >= 8 && w >= 16 && tile_latent >= 8, "dv_vae_plan_create: T=%d h=%d w=%d", T,
             h, w);
  dv_vae_plan* p = new dv_vae_plan();
  p->v = v;
  p->T = T;
  p->h = h;
  p->w = w;
  p->tile = tile_latent;
  auto fail = [&](int rc) {
    dv_vae_plan_destroy(p);
    return rc;
  };
  // tile origins (vae.py:990-997): stride = 3/4 tile when the latent exceeds one tile
  const bool tiled = (h > tile_latent) || (w > tile_latent);
  const int stride = tiled ? (tile_latent * 3) / 4 : (h > w ? h : w);
  for (int y = 0; y < h; y += stride) p->ys.push_back(y);
  for (int x = 0; x < w; x += stride) p->xs.push_back(x);
  if (!tiled) {
    p->ys.assign(1, 0);
    p->xs.assign(1, 0);
    p->tile = h > w ? h : w;
  }
  p->rows = static_cast<int>(p->ys.size());
  p->cols = static_cast<int>(p->xs.size());
  p->tiles.resize(static_cast<size_t>(p->rows) * p->cols);
  long long max_elems = 0;
  for (int i = 0; i < p->rows; ++i)
    for (int j = 0; j < p->cols; ++j) {
      int rc = run_tile(p, nullptr, 0, i, j, 0, true, &p->flops, &max_elems);
      if (rc) return fail(rc);
    }
  p->buf_elems = max_elems;
  int rc = 0;
  for (int k = 0; k < 4; ++k)
    if ((rc = plan_alloc(p, &p->buf[k], max_elems)) != 0) return fail(rc);
  const int top = v->cfg.block_channels[3];
  const long long pix = static_cast<long long>(p->tile) * p->tile;
  if ((rc = plan_alloc(p, &p->qk, static_cast<long long>(T) * pix * 3 * top)) != 0) return fail(rc);
  if ((rc = plan_alloc(p, &p->vt, static_cast<long long>(T) * pix * top)) != 0) return fail(rc);
  if ((rc = plan_alloc(p, &p->sc, static_cast<long long>(T) * pix * pix)) != 0) return fail(rc);
  if ((rc = plan_alloc(p, &p->pr, static_cast<long long>(T) * pix * pix)) != 0) return fail(rc);
  const int max_frames = 8 * T + 8;
  p->gn_replica_elems = max_frames * 64 * 2;
  p->gn_slot_elems = dv_vae_plan::kGnReplicas * p->gn_replica_elems;
  if ((rc = plan_alloc(p, &p->gn_acc, static_cast<long long>(p->gn_slots + 1) * p->gn_slot_elems)) != 0)
    return fail(rc);
  cudaError_t e = cudaSuccess;
  if ((rc = plan_alloc(p, &p->taps, p->taps_elems)) != 0) return fail(rc);
  for (auto& to : p->tiles) {
    if ((rc = plan_alloc(p, &to.px, static_cast<long long>(p->Tout) * to.H * to.W * 3)) != 0)
      return fail(rc);
  }
  if ((rc = plan_alloc(p, &p->tiles_dev, static_cast<long long>(p->tiles.size()))) != 0) return fail(rc);
  if (e == cudaSuccess)
    e = cudaMemcpy(p->tiles_dev, p->tiles.data(), p->tiles.size() * sizeof(TileOut),
                   cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    set_error("dv_vae_plan_create: %s", cudaGetErrorString(e));
    return fail(DV_ERR_CUDA);
  }
  *out = p;
  return DV_OK;
}

extern "C" void dv_vae_plan_destroy(dv_vae_plan* p) {
  if (!p) return;
  for (void* a : p->allocs) cudaFree(a);
  delete p;
}

extern "C" double dv_vae_plan_flops(const dv_vae_plan* p) { return p ? p->flops : 0.0; }

// Decode only output frames [first_frame, Tout) from now on (0 = everything): the frames in front are neither computed
// nor written.  The kept frames are bit-identical to a full decode (see run_tile).
extern "C" int dv_vae_plan_set_first_frame(dv_vae_plan* p, int first_frame) {
  DV_REQUIRE(p && first_frame >= 0 && first_frame < p->Tout, "dv_vae_plan_set_first_frame: frame %d of %d", first_frame,
             p ? p->Tout : 0);
  p->first_frame = first_frame;
  return DV_OK;
}

extern "C" int dv_vae_plan_geometry(const dv_vae_plan* p, int* rows, int* cols, int* t_out) {
  DV_REQUIRE(p && rows && cols && t_out, "dv_vae_plan_geometry: null argument");
  *rows = p->rows;
  *cols = p->cols;
  *t_out = p->Tout;
  return DV_OK;
}

extern "C" int dv_vae_plan_tile_info(const dv_vae_plan* p, int tile, int* H, int* W) {
  DV_REQUIRE(p && H && W && tile >= 0 && tile < p->rows * p->cols, "dv_vae_plan_tile_info: bad tile %d",
             tile);
  *H = p->tiles[tile].H;
  *W = p->tiles[tile].W;
  return DV_OK;
}

// Use a caller-owned buffer ([Tout][H][W][3] bf16) for one decoded tile, so that the host can
// exchange tiles between ranks (NCCL broadcast / all_gather) before the blend.
extern "C" int dv_vae_plan_bind_tile(dv_vae_plan* p, int tile, void* buf_dev) {
  DV_REQUIRE(p && buf_dev && tile >= 0 && tile < p->rows * p->cols, "dv_vae_plan_bind_tile: bad tile %d",
             tile);
  p->tiles[tile].px = reinterpret_cast<__nv_bfloat16*>(buf_dev);
  DV_CHECK_CUDA(cudaMemcpy(p->tiles_dev, p->tiles.data(), p->tiles.size() * sizeof(TileOut),
                           cudaMemcpyHostToDevice));
  return DV_OK;
}

extern "C" int dv_vae_decode_tiles(dv_vae_plan* p, const void* z_dev, int z_dtype,
                                   unsigned long long tile_mask, void* stream) {
  DV_REQUIRE(p && z_dev, "dv_vae_decode_tiles: null argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (int i = 0; i < p->rows; ++i)
    for (int j = 0; j < p->cols; ++j) {
      if (!((tile_mask >> (i * p->cols + j)) & 1ull)) continue;
      int rc = run_tile(p, z_dev, z_dtype == DV_DTYPE_BF16, i, j, st, false, nullptr, nullptr);
      if (rc) return rc;
    }
  return DV_OK;
}

extern "C" int dv_vae_decode(dv_vae_plan* p, const void* z_dev, int z_dtype, void* out_dev,
                             int out_dtype, void* stream) {
  DV_REQUIRE(p && z_dev && out_dev, "dv_vae_decode: null argument");
  int rc = dv_vae_decode_tiles(p, z_dev, z_dtype, ~0ull, stream);
  if (rc) return rc;
  return dv_vae_blend(p, out_dev, out_dtype, stream);
}

extern "C" int dv_vae_blend(dv_vae_plan* p, void* out_dev, int out_dtype, void* stream) {
  DV_REQUIRE(p && out_dev, "dv_vae_blend: null argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  BlendArgs a;
  a.tiles = p->tiles_dev;
  a.rows = p->rows;
  a.cols = p->cols;
  a.Tout = p->Tout;
  const int scale = 8;
  a.Hout = p->h * scale;
  a.Wout = p->w * scale;
  a.extent = (p->tile * scale) / 4;          // blend_extent = tile_sample_min_size * 0.25
  a.limit = p->tile * scale - a.extent;      // row_limit
  if (p->rows == 1 && p->cols == 1) a.limit = (a.Hout > a.Wout ? a.Hout : a.Wout) + 1;
  a.t_first = p->first_frame;
  const long long total = 3LL * (a.Tout - a.t_first) * a.Hout * a.Wout;
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  if (out_dtype == DV_DTYPE_BF16)
    blend_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(a, reinterpret_cast<__nv_bfloat16*>(out_dev));
  else
    blend_kernel<float><<<blocks, 256, 0, st>>>(a, reinterpret_cast<float*>(out_dev));
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return DV_OK;
}

// =================================================================================================
// Encoder.  Reuses the decoder's plan machinery (rotating NDHWC buffers,
// GroupNorm accumulator ring, mid-block attention scratch, tile table + blend kernel).
// =================================================================================================
struct dv_vae_enc_plan {
  dv_vae_plan base;            // base.T/h/w: input frames, PIXEL height/width; base.tile: tile size in pixels
  int Tl = 0, lh = 0, lw = 0;  // latent dims of the whole clip
  float* moments = nullptr;    // fp32 [2z][Tl][lh][lw] scratch when the caller does not want the moments
};

namespace {

__global__ void gaussian_sample_kernel(const float* __restrict__ moments, const float* __restrict__ noise,
                                       void* __restrict__ out, long long n, int out_bf16) {
  // moments: [mean (n) | logvar (n)]; DiagonalGaussianDistribution.sample (vae.py:602-615)
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float lv = fminf(fmaxf(moments[n + i], -30.0f), 20.0f);
  const float v = moments[i] + expf(0.5f * lv) * noise[i];
  if (out_bf16)
    reinterpret_cast<__nv_bfloat16*>(out)[i] = __float2bfloat16(v);
  else
    reinterpret_cast<float*>(out)[i] = v;
}

}  // namespace

// encoder + quant_conv of one pixel tile (vae.py:671-689, :859-860)
static int run_enc_tile(dv_vae_enc_plan* ep, const void* x, int x_bf16, int ti, int tj, cudaStream_t st,
                        bool dry, double* flops, long long* max_elems) {
  dv_vae_plan* p = &ep->base;
  const dv_vae_config& c = p->v->cfg;
  Runner r;
  r.pl = p;
  r.st = st;
  r.dry = dry;
  const int y0 = p->ys[ti], x0 = p->xs[tj];
  const int th = std::min(p->tile, p->h - y0), tw = std::min(p->tile, p->w - x0);
  if (th % 64 || tw % 128) {
    set_error("vae encode: pixel tile %dx%d must be a multiple of 64x128", th, tw);
    return DV_ERR_INVALID;
  }
  __nv_bfloat16** B = p->buf;
  Act a{B[0], p->T, th, tw, 64};  // 3 image channels padded to one 64-channel TMA box
  r.note(a);
  if (!dry) {
    cudaError_t e = cudaMemsetAsync(p->gn_acc, 0,
                                    static_cast<size_t>(p->gn_slots + 1) * p->gn_slot_elems * sizeof(double), st);
    if (e != cudaSuccess) {
      set_error("vae encode: memset of the GroupNorm accumulators: %s", cudaGetErrorString(e));
      return DV_ERR_CUDA;
    }
    r.rc = launch_latent_tile(x, x_bf16, B[0], c.enc_in_channels, p->T, p->h, p->w, y0, x0, th, tw, 64, st);
  }
  int cur = 0;
  auto other = [&](int k) { return (cur + k) & 3; };
  a = r.conv("encoder.conv_in.conv", a, c.enc_block_channels[0], 3, CONV_PLAIN, 0, nullptr, B[other(1)], -1, true);
  cur = other(1);
  for (int i = 0; i < 4; ++i) {
    const int co = c.enc_block_channels[i];
    for (int j = 0; j < c.enc_layers_per_block[i]; ++j) {
      a = r.resnet("encoder.down_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), a, co, other(1),
                   other(2), other(3));
      cur = other(3);
    }
    if (c.enc_spatial_down[i]) {
      a = r.conv("encoder.down_blocks." + std::to_string(i) + ".downsamplers.0.conv.conv", a, co, 3, CONV_PLAIN, 0,
                 nullptr, B[other(1)], -1, !c.enc_temporal_down[i], 1, 2, 2);
      cur = other(1);
    }
    if (c.enc_temporal_down[i]) {
      a = r.conv("encoder.down_blocks." + std::to_string(i) + ".temporal_downsamplers.0.conv.conv", a, co, 3,
                 CONV_PLAIN, 0, nullptr, B[other(1)], -1, true, 2, 1, 1);
      cur = other(1);
    }
  }
  const int top = c.enc_block_channels[3];
  a = r.resnet("encoder.mid_block.resnets.0", a, top, other(1), other(2), other(3));
  cur = other(3);
  a = r.attention("encoder.mid_block.attentions.0", a, other(1), other(3));
  cur = other(3);
  a = r.resnet("encoder.mid_block.resnets.1", a, top, other(1), other(2), other(3));
  cur = other(3);
  a = r.gn("encoder.conv_norm_out", a, true, B[other(1)]);
  cur = other(1);
  // conv_out (top -> 2z) and quant_conv (1x1x1), both padded to 64 channels so that they chain
  a = r.conv("encoder.conv_out.conv", a, 64, 3, CONV_PLAIN, 0, nullptr, B[other(1)], 64);
  cur = other(1);
  TileOut& to = p->tiles[ti * p->cols + tj];
  Act m = r.conv("quant_conv.conv", a, 64, 1, CONV_PLAIN, 0, nullptr, to.px, 64);
  if (dry) {
    to.H = m.H;
    to.W = m.W;
    to.C = 64;
    p->Tout = m.T;
  }
  if (flops) *flops += r.flops;
  if (max_elems && r.max_elems > *max_elems) *max_elems = r.max_elems;
  return r.rc;
}

extern "C" void dv_vae_enc_plan_destroy(dv_vae_enc_plan* ep) {
  if (!ep) return;
  for (void* a : ep->base.allocs) cudaFree(a);
  delete ep;
}

extern "C" int dv_vae_enc_plan_create(dv_vae* v, int T, int H, int W, int tile_px, dv_vae_enc_plan** out) {
  DV_REQUIRE(v && out, "dv_vae_enc_plan_create: null argument");
  DV_REQUIRE(v->cfg.enc_block_channels[0] > 0, "dv_vae_enc_plan_create: the encoder's weights were not loaded");
  DV_REQUIRE(T >= 1 && H >= 64 && W >= 128 && tile_px >= 128, "dv_vae_enc_plan_create: T=%d H=%d W=%d tile=%d", T, H,
             W, tile_px);
  int downs = 0;
  for (int i = 0; i < 4; ++i) downs += v->cfg.enc_spatial_down[i] ? 1 : 0;
  DV_REQUIRE(downs == 3, "dv_vae_enc_plan_create: %d spatial down-samplings (3 supported: /8)", downs);
  dv_vae_enc_plan* ep = new dv_vae_enc_plan();
  dv_vae_plan* p = &ep->base;
  p->v = v;
  p->T = T;
  p->h = H;
  p->w = W;
  p->tile = tile_px;
  auto fail = [&](int rc) {
    dv_vae_enc_plan_destroy(ep);
    return rc;
  };
  // tile origins (vae.py:955,961-963): every 3/4 tile when the frame exceeds one tile
  const bool tiled = (H > tile_px) || (W > tile_px);
  if (tiled) {
    const int stride = (tile_px * 3) / 4;
    for (int y = 0; y < H; y += stride) p->ys.push_back(y);
    for (int x = 0; x < W; x += stride) p->xs.push_back(x);
  } else {
    p->ys.assign(1, 0);
    p->xs.assign(1, 0);
    p->tile = H > W ? H : W;
  }
  p->rows = static_cast<int>(p->ys.size());
  p->cols = static_cast<int>(p->xs.size());
  p->tiles.resize(static_cast<size_t>(p->rows) * p->cols);
  long long max_elems = 0;
  for (int i = 0; i < p->rows; ++i)
    for (int j = 0; j < p->cols; ++j) {
      int rc = run_enc_tile(ep, nullptr, 0, i, j, 0, true, &p->flops, &max_elems);
      if (rc) return fail(rc);
    }
  ep->Tl = p->Tout;
  ep->lh = H / 8;
  ep->lw = W / 8;
  p->buf_elems = max_elems;
  int rc = 0;
  for (int k = 0; k < 4; ++k)
    if ((rc = plan_alloc(p, &p->buf[k], max_elems)) != 0) return fail(rc);
  const int top = v->cfg.enc_block_channels[3];
  const long long pix = static_cast<long long>(p->tile / 8) * (p->tile / 8);
  const int Tm = ep->Tl;  // the mid block runs at latent resolution
  if ((rc = plan_alloc(p, &p->qk, static_cast<long long>(Tm) * pix * 3 * top)) != 0) return fail(rc);
  if ((rc = plan_alloc(p, &p->vt, static_cast<long long>(Tm) * pix * top)) != 0) return fail(rc);
  if ((rc = plan_alloc(p, &p->sc, static_cast<long long>(Tm) * pix * pix)) != 0) return fail(rc);
  if ((rc = plan_alloc(p, &p->pr, static_cast<long long>(Tm) * pix * pix)) != 0) return fail(rc);
  const int max_frames = T + 8;
  p->gn_replica_elems = max_frames * 64 * 2;
  p->gn_slot_elems = dv_vae_plan::kGnReplicas * p->gn_replica_elems;
  if ((rc = plan_alloc(p, &p->gn_acc, static_cast<long long>(p->gn_slots + 1) * p->gn_slot_elems)) != 0)
    return fail(rc);
  for (auto& to : p->tiles)
    if ((rc = plan_alloc(p, &to.px, static_cast<long long>(p->Tout) * to.H * to.W * to.C)) != 0) return fail(rc);
  if ((rc = plan_alloc(p, &p->tiles_dev, static_cast<long long>(p->tiles.size()))) != 0) return fail(rc);
  const long long mom = 2LL * v->cfg.latent_channels * ep->Tl * ep->lh * ep->lw;
  if ((rc = plan_alloc(p, &ep->moments, mom)) != 0) return fail(rc);
  cudaError_t e = cudaMemcpy(p->tiles_dev, p->tiles.data(), p->tiles.size() * sizeof(TileOut), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    set_error("dv_vae_enc_plan_create: %s", cudaGetErrorString(e));
    return fail(DV_ERR_CUDA);
  }
  *out = ep;
  return DV_OK;
}

extern "C" double dv_vae_enc_plan_flops(const dv_vae_enc_plan* p) { return p ? p->base.flops : 0.0; }

extern "C" int dv_vae_enc_plan_latent_dims(const dv_vae_enc_plan* p, int* t, int* h, int* w) {
  DV_REQUIRE(p && t && h && w, "dv_vae_enc_plan_latent_dims: null argument");
  *t = p->Tl;
  *h = p->lh;
  *w = p->lw;
  return DV_OK;
}

extern "C" int dv_gaussian_sample(const float* moments_dev, const float* noise_dev, void* sample_dev,
                                  long long n, int sample_dtype, void* stream) {
  DV_REQUIRE(moments_dev && noise_dev && sample_dev && n > 0, "dv_gaussian_sample: bad argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  gaussian_sample_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(
      moments_dev, noise_dev, sample_dev, n, sample_dtype == DV_DTYPE_BF16);
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return DV_OK;
}

extern "C" int dv_vae_encode(dv_vae_enc_plan* ep, const void* x_dev, int x_dtype, float* moments_dev,
                             const float* noise_dev, void* sample_dev, int sample_dtype, void* stream) {
  DV_REQUIRE(ep && x_dev, "dv_vae_encode: null argument");
  DV_REQUIRE((noise_dev == nullptr) == (sample_dev == nullptr), "dv_vae_encode: noise and sample go together");
  dv_vae_plan* p = &ep->base;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (int i = 0; i < p->rows; ++i)
    for (int j = 0; j < p->cols; ++j) {
      int rc = run_enc_tile(ep, x_dev, x_dtype == DV_DTYPE_BF16, i, j, st, false, nullptr, nullptr);
      if (rc) return rc;
    }
  float* mom = moments_dev ? moments_dev : ep->moments;
  // blend_v / blend_h + crop + concat of the moment tiles (vae.py:970-983)
  BlendArgs a;
  a.tiles = p->tiles_dev;
  a.rows = p->rows;
  a.cols = p->cols;
  a.Tout = ep->Tl;
  a.Hout = ep->lh;
  a.Wout = ep->lw;
  const int tl = p->tile / 8;
  a.extent = tl / 4;          // blend_extent = tile_latent_min_size * 0.25
  a.limit = tl - a.extent;    // row_limit
  if (p->rows == 1 && p->cols == 1) a.limit = (a.Hout > a.Wout ? a.Hout : a.Wout) + 1;
  a.nch = 2 * p->v->cfg.latent_channels;
  const long long total = static_cast<long long>(a.nch) * a.Tout * a.Hout * a.Wout;
  blend_kernel<float><<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(a, mom);
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  if (sample_dev != nullptr) {
    const long long n = total / 2;
    gaussian_sample_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(
        mom, noise_dev, sample_dev, n, sample_dtype == DV_DTYPE_BF16);
    DV_CHECK_CUDA(cudaGetLastError());
    note_launch();
  }
  return DV_OK;
}

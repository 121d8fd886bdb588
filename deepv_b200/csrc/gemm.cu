// deepv_b200 — persistent, warp-specialised tcgen05 GEMM / implicit-GEMM conv3d.
//
// One CTA per SM, 320 threads:
//   warp 0      TMA producer: A/B k-blocks -> 128B-swizzled smem ring
//   warp 1      MMA issuer: tcgen05.mma M=128 (CTA pairs: 256), N=256, K=16, fp32 accum in TMEM; also owns TMEM alloc/dealloc
//               (both run their loops as whole converged warps; the issuing lane is elected inside the instruction —
//               common.cuh w_* forms — so the operands stay in uniform registers instead of an ELECT / R2UR loop per issue)
//   warps 2..9  epilogue (two warps per TMEM lane quarter, one accumulator row per thread): tcgen05.ld -> fused epilogue
//               (compile-time EpiMode) -> shared-memory staging + TMA store / reduce-add, or direct vector stores
// ONE tile shape, 128 x 256 x 64: measured on B200, an M128 tcgen05.mma with both operands in
// shared memory takes >= 128 cycles whatever N is (the 128-row A read), so N = 256 is the only
// width that runs the tensor pipe at rate; a CTA therefore needs ~K*8 cycles per tile and the
// lever for problems with fewer tiles than SMs is split-K, not narrower tiles:
//   * splits == 1: persistent grid, two TMEM accumulator stages (epilogue of tile i overlaps the
//     mainloop of tile i+1);
//   * splits  > 1: the S CTAs that share an output tile form a thread-block cluster; each
//     reduces K/S, parks its fp32 partial tile in its own shared memory, and after a cluster
//     barrier every CTA sums one row-slice over all peers through distributed shared memory in
//     rank order (deterministic, no atomics, no workspace) and runs the epilogue for that slice.
// Convolutions with Cout <= 128 swap the operand roles (weights = M side, a 16x16-pixel box =
// N side).  Dense problems with contiguous batch rows are flattened into one row space (Problem::flat_M); a conv can
// be asked for output frames [conv_t0, T) only (the trimmed decode).  See gemm.cuh for the operand / epilogue contract.
#include "gemm.cuh"

#include <cstdio>
#include <cstdlib>

namespace dv {

namespace {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int kEpiWarps = 8;  // two warps per TMEM lane quarter: each takes half of a tile's columns
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kStages = 4;
constexpr int kABytes = BM * BK * 2;
constexpr int kBBytes = BN * BK * 2;
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kBarBytes = 256;
constexpr int kGnBytes = 4 * 8 * 8 * 2 * 4;  // fused GroupNorm partials: [warp][chunk][group][2] floats
// epilogue staging for TMA stores: 2 column halves x 2 buffers of 128 rows x 64 B (64B-swizzled); the GroupNorm
// partials of the conv epilogue (which stores directly) live at the start of the same region
constexpr int kOutBufBytes = 128 * 64;
constexpr int kOutBytes = 4 * kOutBufBytes;
static_assert(kGnBytes <= kOutBytes, "GroupNorm partials share the staging region");
constexpr int kSmemBytes = kStages * kStageBytes + kOutBytes + kBarBytes + 1024;  // + align slack
constexpr uint32_t kTmemCols = 2 * BN;
constexpr int kMaxSplits = 8;  // portable cluster size

struct Problem {
  alignas(64) CUtensorMap tmA;
  alignas(64) CUtensorMap tmB;
  alignas(64) CUtensorMap tmBh;  // CTA-pair mode: same matrix, 128-row box (each CTA loads half of N)
  alignas(64) CUtensorMap tmAp[3];  // strided conv: the other input parities (tmA = parity 0), see setup_problem
  alignas(64) CUtensorMap tmOut;    // tma_out: the output as (N, M, batch), box 64 B x 128 rows, 64B swizzle
  int tma_out;                      // epilogue stages tiles in shared memory and stores / reduce-adds them with TMA
  GemmDesc d;
  int m_tiles, n_tiles, k_blocks, tiles;
  int kb_per_split;      // split-K over a cluster of `splits` CTAs
  int m_pairs, pair_tiles;  // CTA-pair mode: 256-row tiles (two m-tiles) x n_tiles x batch
  int tiles_w, tiles_h;  // conv: M-tile grid inside one frame
  int c_blocks;          // conv: Cin / 64
  int swap;              // conv, Cout <= 128: weights are the M operand, 16x16 pixels the N operand
  int oT, oH, oW;        // conv: output dims (input dims over the strides)
  int oT0;               // conv: first output frame computed (GemmDesc::conv_t0); m-tiles cover frames [oT0, oT)
  int sT, sH, sW;        // conv: strides (1 or 2)
  // dense problems whose batch rows are contiguous in A and in the output run as ONE problem of flat_B * flat_M rows
  // (d.M = flat_B * flat_M, d.batch = 1): a 128-row (CTA pair: 256-row) tile then wastes at most one partial tile per
  // launch instead of one per batch row — B = 3 x 269 context rows are 4 pair tiles instead of 6, B = 2 x 384 video
  // rows 3 instead of 4.  flat_M = 0: not flattened.  The logical batch of a row (per-batch gate vectors) is row / flat_M.
  int flat_M, flat_B;
};

// One launch runs up to two problems of the same epilogue mode (the video and the context stream
// of a joint transformer block): tiles [0, p[0].tiles) belong to p[0], the rest to p[1].
struct KArgs {
  Problem p[2];
  int total_tiles;
  int splits;
  int total_pair_tiles;
  int pdl_early;
#ifdef DV_GEMM_TRACE
  long long* trace;   // probe build only (scripts/probe/gemm_trace.py): per-CTA globaltimer stamps
#endif
};

// probe build: slot k of this CTA's record <- globaltimer (slot 0 holds the SM id)
#ifdef DV_GEMM_TRACE
#define DV_GT(on, slot)                                                                                       \
  do {                                                                                                        \
    if ((on) && blockIdx.x < 4096) k.trace[blockIdx.x * 12 + (slot)] = static_cast<long long>(global_ns());    \
  } while (0)
#else
#define DV_GT(on, slot) \
  do {                  \
  } while (0)
#endif

struct TileCoord {
  int b, m_tile, n_tile, kb0, kb1;
};

__device__ __forceinline__ const Problem& problem_of(const KArgs& k, int& tile) {
  if (tile < k.p[0].tiles) return k.p[0];
  tile -= k.p[0].tiles;
  return k.p[1];
}

__device__ __forceinline__ TileCoord decode_tile(const Problem& a, int tile, int split) {
  TileCoord tc;
  tc.n_tile = tile % a.n_tiles;
  const int rest = tile / a.n_tiles;
  tc.m_tile = rest % a.m_tiles;
  tc.b = rest / a.m_tiles;
  tc.kb0 = split * a.kb_per_split;
  tc.kb1 = min(tc.kb0 + a.kb_per_split, a.k_blocks);
  return tc;
}

// ---- cluster helpers ------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the second barrier of a split-K launch orders nothing in memory: it only keeps a CTA (its shared memory) alive until
// every peer has consumed the values it loaded from it, which the in-order issue of the consuming stores already implies
__device__ __forceinline__ void cluster_sync_relaxed() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// (not volatile, no memory clobber: the partial tiles are written before the cluster barrier, which is the compiler fence;
// the loads may then be scheduled together with the epilogue's global loads instead of in front of them)
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t addr) {
  float4 v;
  asm("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// ------------------------------------------------------------------------------
// epilogue: one thread = one output row, W (32 or 64) accumulator columns at a time
// ------------------------------------------------------------------------------
struct RowCtx {
  int b, m;        // batch, dense row index inside the batch (conv: unused)
  int gb;          // logical batch of the row (differs from b for a flattened problem)
  bool ok;         // row exists
  int ct, ch, cw;  // conv: output pixel
};

__device__ __forceinline__ RowCtx make_row(const Problem& a, const TileCoord& tc, int row_in_tile) {
  const GemmDesc& d = a.d;
  RowCtx r;
  r.b = tc.b;
  r.m = tc.m_tile * BM + row_in_tile;
  r.ct = r.ch = r.cw = 0;
  if (d.a_mode == 1) {
    const int per_frame = a.tiles_w * a.tiles_h;
    r.ct = a.oT0 + tc.m_tile / per_frame;
    const int q = tc.m_tile % per_frame;
    r.ch = (q / a.tiles_w) * 8 + (row_in_tile >> 4);
    r.cw = (q % a.tiles_w) * 16 + (row_in_tile & 15);
    r.ok = (r.ct < a.oT) && (r.ch < a.oH) && (r.cw < a.oW);
  } else {
    r.ok = r.m < d.M;
  }
  r.gb = a.flat_M ? min(r.m / a.flat_M, a.flat_B - 1) : tc.b;
  return r;
}

__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const float* v) {
  uint4* p = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 q;
    q.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
    q.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
    q.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
    q.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
    p[i] = q;
  }
}

template <int W>
__device__ __forceinline__ void add_bias(const float* bias, int n, float (&v)[W]) {
  if (bias != nullptr) {
    const float4* bp = reinterpret_cast<const float4*>(bias + n);
#pragma unroll
    for (int i = 0; i < W / 4; ++i) {
      const float4 t = __ldg(bp + i);
      v[4 * i + 0] += t.x;
      v[4 * i + 1] += t.y;
      v[4 * i + 2] += t.z;
      v[4 * i + 3] += t.w;
    }
  }
}

template <int W>
__device__ __forceinline__ void add_bf16x32(const __nv_bfloat16* src, float (&v)[W]) {
  const uint4* r4 = reinterpret_cast<const uint4*>(src);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 t = __ldg(r4 + i);
    v[8 * i + 0] += bf16_lo(t.x);
    v[8 * i + 1] += bf16_hi(t.x);
    v[8 * i + 2] += bf16_lo(t.y);
    v[8 * i + 3] += bf16_hi(t.y);
    v[8 * i + 4] += bf16_lo(t.z);
    v[8 * i + 5] += bf16_hi(t.z);
    v[8 * i + 6] += bf16_lo(t.w);
    v[8 * i + 7] += bf16_hi(t.w);
  }
}

// sum / sum of squares of the bf16-rounded values of 32 / CPG GroupNorm groups over the rows a warp
// holds.  `uniform` (all 32 lanes on the same chunk): butterfly over the lanes, then lane 0 either
// parks the warp's partial in shared memory (`part`: [8 groups][2] floats of this warp and chunk;
// the tile flush merges the four warps in a fixed order) or adds it to the fp64 accumulators.
template <int CPG>
__device__ __forceinline__ void gn_accumulate(const float (&v)[32], bool row_ok, double* acc,
                                              bool uniform, float* part) {
#pragma unroll
  for (int g = 0; g < 32 / CPG; ++g) {
    float sm = 0.f, sq = 0.f;
#pragma unroll
    for (int i = 0; i < CPG; ++i) {
      const float x = row_ok ? __bfloat162float(__float2bfloat16(v[g * CPG + i])) : 0.f;
      sm += x;
      sq += x * x;
    }
    if (uniform) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sm += __shfl_xor_sync(0xffffffffu, sm, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
      }
      if ((threadIdx.x & 31) == 0) {
        if (part != nullptr) {
          part[2 * g] = sm;
          part[2 * g + 1] = sq;
        } else {
          atomicAdd(acc + 2 * g, static_cast<double>(sm));
          atomicAdd(acc + 2 * g + 1, static_cast<double>(sq));
        }
      }
    } else if (row_ok) {
      atomicAdd(acc + 2 * g, static_cast<double>(sm));
      atomicAdd(acc + 2 * g + 1, static_cast<double>(sq));
    }
  }
}

// output frame / first stored channel of accumulator column n of a conv tile whose rows sit in
// input frame ct (the store-mode address maps of vae.py:382,407-409)
__device__ __forceinline__ void conv_out_coord(const GemmDesc& d, int ct, int n, int* ot, int* oc) {
  *ot = ct;
  *oc = n;
  if (d.conv_store == CONV_SHUFFLE_HW) {
    *oc = n % d.out_C;
  } else if (d.conv_store == CONV_INTERLEAVE_T) {
    *oc = n % d.out_C;
    *ot = 2 * ct + n / d.out_C - (d.conv_drop_first ? 1 : 0);
  }
}

__device__ __forceinline__ double* gn_replica(const GemmDesc& d) {
  return d.gn_acc + static_cast<long long>(blockIdx.x % d.gn_replicas) * d.gn_replica_stride;
}

template <int MODE>
struct EpiW {
  static constexpr int value = (MODE == EPI_QKV) ? 64 : 32;
};

// one head (64 columns) of the fused q|k|v projection of row m: bias, per-head RMSNorm (q, k), temporal RoPE on
// adjacent pairs (mmdit.py:282-307,131-136); v stays as it is
__device__ __forceinline__ void qkv_head(const GemmDesc& d, int m, int n, float (&v)[64]) {
  add_bias<64>(d.bias, n, v);
  const int region = n / d.heads_dim;  // 0 q, 1 k, 2 v
  if (region < 2) {
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) ss += v[i] * v[i];
    const float rs = rsqrtf(ss * (1.0f / 64.0f) + 1e-5f);  // RMSNorm eps, mmdit.py:195,453-454
    const float2* w2 = reinterpret_cast<const float2*>(d.qk_norm_w + region * 64);
    const int fid = __ldg(d.frame_id + m);
    const float2* cs = reinterpret_cast<const float2*>(d.rope_cs) + fid * 32;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float2 w = __ldg(w2 + i);
      const float x0 = v[2 * i] * rs * w.x;
      const float x1 = v[2 * i + 1] * rs * w.y;
      const float2 t = __ldg(cs + i);  // (cos, sin) of frame * 10000^(-2i/64)
      v[2 * i] = t.x * x0 - t.y * x1;
      v[2 * i + 1] = t.y * x0 + t.x * x1;
    }
  }
}

// v = W accumulator columns [n, n + W) of output row `r`; n < d.N and r.ok hold.
// `uniform`: every lane of the warp handles the same column chunk n of rows of ONE tile (true on the
// TMEM path; false in the split-K reduction, where lanes may sit on different chunks).
template <int MODE, int W>
__device__ __forceinline__ void epi_row(const Problem& a, const RowCtx& r, int n, float (&v)[W],
                                        bool uniform = true, float* gn_part = nullptr) {
  const GemmDesc& d = a.d;
  if constexpr (MODE == EPI_BF16 || MODE == EPI_GELU) {
    add_bias<W>(d.bias, n, v);
    if constexpr (MODE == EPI_GELU) {
#pragma unroll
      for (int i = 0; i < W; ++i) v[i] = gelu_tanh(v[i]);
    }
    const long long off = static_cast<long long>(r.b) * d.out_batch_stride +
                          static_cast<long long>(r.m + d.out_row_offset) * d.ldo + n;
    if (d.residual != nullptr)
      add_bf16x32<W>(reinterpret_cast<const __nv_bfloat16*>(d.residual) + off, v);
    store_bf16x32(reinterpret_cast<__nv_bfloat16*>(d.out) + off, v);
  } else if constexpr (MODE == EPI_BF16_ROWBIAS) {
    const float bm = d.bias ? __ldg(d.bias + r.m) : 0.f;
#pragma unroll
    for (int i = 0; i < W; ++i) v[i] += bm;
    store_bf16x32(reinterpret_cast<__nv_bfloat16*>(d.out) +
                      static_cast<long long>(r.b) * d.out_batch_stride +
                      static_cast<long long>(r.m + d.out_row_offset) * d.ldo + n,
                  v);
  } else if constexpr (MODE == EPI_RESID_GATE) {
    add_bias<W>(d.bias, n, v);
    float4* x4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(d.out) +
                                           static_cast<long long>(r.b) * d.out_batch_stride +
                                           static_cast<long long>(r.m + d.out_row_offset) * d.ldo + n);
    const float4* g4 = reinterpret_cast<const float4*>(d.gate + r.gb * d.gate_batch_stride + n);
#pragma unroll
    for (int i = 0; i < W / 4; ++i) {
      const float4 g = __ldg(g4 + i);
      float4 t = __ldcg(x4 + i);   // L2: inside the persistent block kernel x is updated by other SMs between phases
      t.x += g.x * v[4 * i + 0];
      t.y += g.y * v[4 * i + 1];
      t.z += g.z * v[4 * i + 2];
      t.w += g.w * v[4 * i + 3];
      x4[i] = t;
    }
  } else if constexpr (MODE == EPI_F32_ADD) {
    add_bias<W>(d.bias, n, v);
    if (d.addend != nullptr) {
      const int ar = d.row_map ? __ldg(d.row_map + r.m) : r.m;
      const float4* a4 =
          reinterpret_cast<const float4*>(d.addend + static_cast<long long>(ar) * d.N + n);
#pragma unroll
      for (int i = 0; i < W / 4; ++i) {
        const float4 t = __ldg(a4 + i);
        v[4 * i + 0] += t.x;
        v[4 * i + 1] += t.y;
        v[4 * i + 2] += t.z;
        v[4 * i + 3] += t.w;
      }
    }
    float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(d.out) +
                                           static_cast<long long>(r.b) * d.out_batch_stride +
                                           static_cast<long long>(r.m + d.out_row_offset) * d.ldo + n);
#pragma unroll
    for (int i = 0; i < W / 4; ++i)
      o4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else if constexpr (MODE == EPI_QKV) {
    qkv_head(d, r.m, n, v);
    void* base = d.out;
    if (d.peer_cols > 0) base = d.out_peer[(n % d.heads_dim) / d.peer_cols];  // the rank that owns this head
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(base) +
                         static_cast<long long>(r.b) * d.out_batch_stride +
                         static_cast<long long>(r.m + d.out_row_offset) * d.ldo + n;
    store_bf16x32(out, v);
    store_bf16x32(out + 32, v + 32);
  } else if constexpr (MODE == EPI_UNPATCH) {
    // row m = (gy, gx) of the noisy clip's token grid; column n = (p1, p2, c)
    // out[b][c][0][2*gy+p1][2*gx+p2]   (mmdit.py:1453-1457, patch 2)
    const int gy = r.m / d.up_gw, gx = r.m % d.up_gw;
    const int H2 = 2 * d.up_gh, W2 = 2 * d.up_gw;
    const long long ob = static_cast<long long>(r.b) * d.out_batch_stride;
#pragma unroll
    for (int i = 0; i < W; ++i) {
      const int nn = n + i;
      if (nn < d.N) {
        const float val = v[i] + (d.bias ? __ldg(d.bias + nn) : 0.f);
        const int cch = nn % d.up_C;
        const int pq = nn / d.up_C;
        const int p1 = pq >> 1, p2 = pq & 1;
        const long long o =
            ob + (static_cast<long long>(cch) * H2 + (2 * gy + p1)) * W2 + (2 * gx + p2);
        if (d.out_f32)
          reinterpret_cast<float*>(d.out)[o] = val;
        else
          reinterpret_cast<__nv_bfloat16*>(d.out)[o] = __float2bfloat16(val);
      }
    }
  } else if constexpr (MODE == EPI_TAPS) {
    float4* o4 = reinterpret_cast<float4*>(d.out);
    const long long row = static_cast<long long>(r.b) * d.M + r.m;
#pragma unroll
    for (int i = 0; i < W / 4; ++i)
      o4[static_cast<long long>(n / 4 + i) * d.ldo + row] =
          make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else if constexpr (MODE == EPI_CONV) {
    add_bias<W>(d.bias, n, v);
    int ot = r.ct, oh = r.ch, ow = r.cw, oc = n, oH = a.oH, oW = a.oW;
    if (d.conv_store == CONV_SHUFFLE_HW) {
      // packed weight rows are ordered (p1, p2, c): vae.py:382
      const int q = n / d.out_C;
      oc = n % d.out_C;
      oh = 2 * r.ch + (q >> 1);
      ow = 2 * r.cw + (q & 1);
      oH = 2 * a.oH;
      oW = 2 * a.oW;
    } else if (d.conv_store == CONV_INTERLEAVE_T) {
      // packed weight rows are ordered (p, c): vae.py:407-409
      const int p = n / d.out_C;
      oc = n % d.out_C;
      ot = 2 * r.ct + p - (d.conv_drop_first ? 1 : 0);
      if (ot < 0) return;
    }
    const int oT =
        (d.conv_store == CONV_INTERLEAVE_T) ? (2 * a.oT - (d.conv_drop_first ? 1 : 0)) : a.oT;
    const long long off =
        (((static_cast<long long>(r.b) * oT + ot) * oH + oh) * oW + ow) * d.out_C + oc;
    if (r.ok) {
      if (d.residual != nullptr)
        add_bf16x32<W>(reinterpret_cast<const __nv_bfloat16*>(d.residual) + off, v);
      store_bf16x32(reinterpret_cast<__nv_bfloat16*>(d.out) + off, v);
    }
    if (d.gn_acc != nullptr) {
      // fused GroupNorm statistics of what was just stored (bf16-rounded), per (frame, group)
      double* acc = gn_replica(d) + (static_cast<long long>(ot) * (d.out_C / d.gn_cpg) + oc / d.gn_cpg) * 2;
      if (d.gn_cpg == 4)
        gn_accumulate<4>(v, r.ok, acc, uniform, gn_part);
      else if (d.gn_cpg == 8)
        gn_accumulate<8>(v, r.ok, acc, uniform, gn_part);
      else
        gn_accumulate<16>(v, r.ok, acc, uniform, gn_part);
    }
  }
}

// Accumulator tile in TMEM -> epilogue, 32 columns in flight while 32 are processed.
// gn_smem: [4 warps][8 chunks][8 groups][2] floats for the fused GroupNorm statistics (conv only)
// `half`: which 128 of the tile's 256 columns this warp handles (two warps share a lane quarter)
template <int MODE>
__device__ __forceinline__ void epilogue_from_tmem(const Problem& a, const TileCoord& tc,
                                                   uint32_t tmem_acc, int row_in_tile, int quarter,
                                                   int half, float* gn_smem = nullptr) {
  constexpr int W = EpiW<MODE>::value;
  const GemmDesc& d = a.d;
  const RowCtx r = make_row(a, tc, row_in_tile);
  const int n0 = tc.n_tile * BN;
  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16);
  if constexpr (W == 64) {
#pragma unroll 1
    for (int c = half * (BN / 128); c < (half + 1) * (BN / 128); ++c) {
      const int n = n0 + c * 64;
      if (n >= d.N) break;  // warp-uniform
      uint32_t r0[32], r1[32];
      tmem_ld_32x32(taddr + c * 64, r0);
      tmem_ld_32x32(taddr + c * 64 + 32, r1);
      tmem_ld_wait();
      float v[64];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        v[i] = __uint_as_float(r0[i]);
        v[32 + i] = __uint_as_float(r1[i]);
      }
      if (r.ok) epi_row<MODE, 64>(a, r, n, v);
    }
  } else {
    uint32_t buf[2][32];
    const int c0 = half * (BN / 64);  // 4 chunks of 32 columns per warp
    tmem_ld_32x32(taddr + c0 * 32, buf[0]);
#pragma unroll
    for (int cc = 0; cc < BN / 64; ++cc) {
      const int c = c0 + cc;
      const int n = n0 + c * 32;
      tmem_ld_wait();
      if (cc + 1 < BN / 64) tmem_ld_32x32(taddr + (c + 1) * 32, buf[(cc + 1) & 1]);
      if (n < d.N && (r.ok || MODE == EPI_CONV)) {  // conv rows always enter: warp-wide GN reduction
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(buf[cc & 1][i]);
        epi_row<MODE, 32>(a, r, n, v, true, gn_smem ? gn_smem + ((quarter * 8 + c) * 8) * 2 : nullptr);
      }
    }
    tmem_ld_wait();
    if (MODE == EPI_CONV && gn_smem != nullptr && d.gn_acc != nullptr) {
      // merge the four warps' partials in a fixed order: one fp64 atomic pair per (chunk, group)
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
      const int e = row_in_tile;  // 0..127 -> (chunk, group, sum | sumsq); the first four warps flush
      const int c = e >> 4, g = (e >> 1) & 7, which = e & 1;
      const int n = n0 + c * 32;
      if (half == 0 && n < d.N && g < 32 / d.gn_cpg) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) tot += gn_smem[((w * 8 + c) * 8 + g) * 2 + which];
        int ot, oc;
        conv_out_coord(d, r.ct, n, &ot, &oc);
        if (ot >= 0)
          atomicAdd(gn_replica(d) + (static_cast<long long>(ot) * (d.out_C / d.gn_cpg) + oc / d.gn_cpg + g) * 2 + which,
                    static_cast<double>(tot));
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");  // partials may be overwritten by the next tile
    }
  }
}

template <int MODE>
struct EpiTma {
  static constexpr bool value = MODE == EPI_BF16 || MODE == EPI_GELU || MODE == EPI_QKV || MODE == EPI_RESID_GATE;
};

// Epilogue through shared memory and TMA (dense problems, no split-K): the TMEM layout gives every thread one output
// ROW, so direct stores touch 32 different rows per instruction (half-used sectors, and a read-modify-write of the fp32
// stream for the gated residual).  Here the 128 threads of a column half write their rows into a 64B-swizzled
// 128 x 64 B staging buffer (conflict-free 16-byte stores) and one thread hands it to TMA: full-line writes for bf16
// outputs, `cp.reduce.async.bulk ... add` for the gated residual (x += gate * (acc + bias) happens in L2, x is never
// read by the SM).  Two buffers per half = the two 64-byte column groups of one step; rows / columns beyond the
// problem are clipped by the tensor map.  `issuer`: the one thread per column half that owns the bulk groups.
template <int MODE>
__device__ __forceinline__ void epilogue_tma(const Problem& a, const TileCoord& tc, uint32_t tmem_acc,
                                             int row_in_tile, int quarter, int half, uint8_t* out_stage, bool issuer) {
  const GemmDesc& d = a.d;
  const int nbase = tc.n_tile * BN + half * 128;
  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16) + half * 128;
  const uint32_t bar = 2 + half;
  uint8_t* buf0 = out_stage + half * (2 * kOutBufBytes);
  uint8_t* buf1 = buf0 + kOutBufBytes;
  const uint32_t sw = (row_in_tile >> 1) & 3;  // 64B swizzle: 16-byte chunk index ^= bits [7,9) of the byte address
  uint4* row0 = reinterpret_cast<uint4*>(buf0 + row_in_tile * 64);
  uint4* row1 = reinterpret_cast<uint4*>(buf1 + row_in_tile * 64);
  const int m0 = tc.m_tile * BM;
  const bool live = m0 < d.M;  // CTA pairs: the odd last m-tile does not exist
  if constexpr (MODE == EPI_RESID_GATE) {
    uint32_t buf[2][32];
    tmem_ld_32x32(taddr, buf[0]);
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int n = nbase + it * 32;
      tmem_ld_wait();
      if (it + 1 < 4) tmem_ld_32x32(taddr + (it + 1) * 32, buf[(it + 1) & 1]);
      if (n < d.N) {  // uniform over the column half
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(buf[it & 1][i]);
        add_bias<32>(d.bias, n, v);
        const int gb = a.flat_M ? min((m0 + row_in_tile) / a.flat_M, a.flat_B - 1) : tc.b;
        const float4* g4 = reinterpret_cast<const float4*>(d.gate + gb * d.gate_batch_stride + n);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 g = __ldg(g4 + i);
          v[4 * i + 0] *= g.x;
          v[4 * i + 1] *= g.y;
          v[4 * i + 2] *= g.z;
          v[4 * i + 3] *= g.w;
        }
        if (issuer) bulk_wait_read<0>();  // both buffers have been read by their previous stores
        named_bar_sync(bar, 128);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          row0[c ^ sw] = make_uint4(__float_as_uint(v[4 * c]), __float_as_uint(v[4 * c + 1]),
                                    __float_as_uint(v[4 * c + 2]), __float_as_uint(v[4 * c + 3]));
          row1[c ^ sw] = make_uint4(__float_as_uint(v[16 + 4 * c]), __float_as_uint(v[16 + 4 * c + 1]),
                                    __float_as_uint(v[16 + 4 * c + 2]), __float_as_uint(v[16 + 4 * c + 3]));
        }
        fence_proxy_async_smem();
        named_bar_sync(bar, 128);
        if (issuer && live) {
          tma_reduce_add_3d(&a.tmOut, buf0, n, m0, tc.b);
          tma_reduce_add_3d(&a.tmOut, buf1, n + 16, m0, tc.b);
          bulk_commit();
        }
      }
    }
    tmem_ld_wait();
  } else {
#pragma unroll 1
    for (int it = 0; it < 2; ++it) {
      const int n = nbase + it * 64;
      if (n >= d.N) break;  // uniform over the column half
      uint32_t r0[32], r1[32];
      tmem_ld_32x32(taddr + it * 64, r0);
      tmem_ld_32x32(taddr + it * 64 + 32, r1);
      tmem_ld_wait();
      float v[64];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        v[i] = __uint_as_float(r0[i]);
        v[32 + i] = __uint_as_float(r1[i]);
      }
      const bool second = n + 32 < d.N;
      if constexpr (MODE == EPI_QKV) {
        const int m = m0 + row_in_tile;
        qkv_head(d, m < d.M ? m : 0, n, v);  // rows beyond M are clipped by the store; keep their table reads in range
      } else {
        if (d.bias != nullptr) {
          const float4* bp = reinterpret_cast<const float4*>(d.bias + n);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (i < 8 || second) {
              const float4 t = __ldg(bp + i);
              v[4 * i + 0] += t.x;
              v[4 * i + 1] += t.y;
              v[4 * i + 2] += t.z;
              v[4 * i + 3] += t.w;
            }
          }
        }
        if constexpr (MODE == EPI_GELU) {
#pragma unroll
          for (int i = 0; i < 64; ++i) v[i] = gelu_tanh(v[i]);
        }
      }
      if (issuer) bulk_wait_read<0>();
      named_bar_sync(bar, 128);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        row0[c ^ sw] = make_uint4(pack_bf16x2(v[8 * c], v[8 * c + 1]), pack_bf16x2(v[8 * c + 2], v[8 * c + 3]),
                                  pack_bf16x2(v[8 * c + 4], v[8 * c + 5]), pack_bf16x2(v[8 * c + 6], v[8 * c + 7]));
        row1[c ^ sw] = make_uint4(pack_bf16x2(v[32 + 8 * c], v[32 + 8 * c + 1]), pack_bf16x2(v[32 + 8 * c + 2], v[32 + 8 * c + 3]),
                                  pack_bf16x2(v[32 + 8 * c + 4], v[32 + 8 * c + 5]), pack_bf16x2(v[32 + 8 * c + 6], v[32 + 8 * c + 7]));
      }
      fence_proxy_async_smem();
      named_bar_sync(bar, 128);
      if (issuer && live) {
        tma_store_3d(&a.tmOut, buf0, n, m0, tc.b);
        if (second) tma_store_3d(&a.tmOut, buf1, n + 32, m0, tc.b);
        bulk_commit();
      }
    }
  }
}

// Swapped-operand conv tile (Cout <= 128): TMEM lane = output channel, column = pixel of the
// 16x16 tile (w fastest).  One thread owns one channel; for every pixel the warp writes a
// contiguous run of 32 channels (64 B) into the NDHWC output.  Plain store mode only.
__device__ __forceinline__ void epilogue_tile_swapped(const Problem& a, const TileCoord& tc,
                                                      uint32_t tmem_acc, int quarter, int lane,
                                                      int half) {
  const GemmDesc& d = a.d;
  const int per_frame = a.tiles_w * a.tiles_h;
  const int ct = a.oT0 + tc.m_tile / per_frame;
  const int r = tc.m_tile % per_frame;
  const int h0 = (r / a.tiles_w) * 16, w0 = (r % a.tiles_w) * 16;
  const int c = quarter * 32 + lane;  // output channel of this thread
  const bool c_ok = c < d.N;
  const float bias = (c_ok && d.bias != nullptr) ? __ldg(d.bias + c) : 0.f;
  const uint32_t taddr_row = tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16);
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(d.out);
  const __nv_bfloat16* res = reinterpret_cast<const __nv_bfloat16*>(d.residual);
  const long long frame_off = (static_cast<long long>(tc.b) * a.oT + ct) * a.oH;
  if (quarter * 32 >= d.N) return;  // warp-uniform: no live channel in this lane quarter
  float gsum = 0.f, gsq = 0.f;      // fused GroupNorm statistics of this thread's channel
  uint32_t buf[2][32];
  tmem_ld_32x32(taddr_row + half * 128, buf[0]);
#pragma unroll
  for (int ci = 0; ci < 4; ++ci) {  // this warp's 4 of the 8 chunks of 32 pixels (2 tile rows each)
    const int cc = half * 4 + ci;
    tmem_ld_wait();
    if (ci + 1 < 4) tmem_ld_32x32(taddr_row + (cc + 1) * 32, buf[(ci + 1) & 1]);
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      const int oh = h0 + cc * 2 + hr;
      if (oh >= a.oH || !c_ok) continue;
      const long long row_off = ((frame_off + oh) * a.oW + w0) * d.out_C + c;
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(buf[ci & 1][hr * 16 + i]) + bias;
      if (res != nullptr) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          v[i] += __bfloat162float(res[row_off + static_cast<long long>(i) * d.out_C]);
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const __nv_bfloat16 q = __float2bfloat16(v[i]);
        out[row_off + static_cast<long long>(i) * d.out_C] = q;
        const float x = __bfloat162float(q);
        gsum += x;
        gsq += x * x;
      }
    }
  }
  tmem_ld_wait();
  if (d.gn_acc != nullptr) {
    // channels of a group sit on adjacent lanes: fold them, one fp64 atomic pair per group
    for (int o = 1; o < d.gn_cpg; o <<= 1) {
      gsum += __shfl_xor_sync(0xffffffffu, gsum, o);
      gsq += __shfl_xor_sync(0xffffffffu, gsq, o);
    }
    if (c_ok && (c % d.gn_cpg) == 0) {
      double* acc = gn_replica(d) + (static_cast<long long>(ct) * (d.out_C / d.gn_cpg) + c / d.gn_cpg) * 2;
      atomicAdd(acc, static_cast<double>(gsum));
      atomicAdd(acc + 1, static_cast<double>(gsq));
    }
  }
}

// ------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(const __grid_constant__ KArgs k) {
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle needs 1024-byte aligned tiles.
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* out_stage = smem + kStages * kStageBytes;  // TMA-store staging (1024-byte aligned) / GroupNorm partials
  uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + kOutBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
#ifdef DV_GEMM_TRACE
  if (threadIdx.x == 0 && blockIdx.x < 4096) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    k.trace[blockIdx.x * 12] = smid;
  }
#endif
  DV_GT(threadIdx.x == 0, 1);   // entry

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&k.p[0].tmA);
    tma_prefetch_desc(&k.p[0].tmB);
    if (k.p[0].tma_out) tma_prefetch_desc(&k.p[0].tmOut);
    if (k.p[1].tiles > 0) {
      tma_prefetch_desc(&k.p[1].tmA);
      tma_prefetch_desc(&k.p[1].tmB);
      if (k.p[1].tma_out) tma_prefetch_desc(&k.p[1].tmOut);
    }
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 32 * kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc<kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  DV_GT(threadIdx.x == 0, 2);   // barriers + tensor memory set up
  pdl_wait();  // everything above overlapped the previous kernel's tail
  if (k.pdl_early) pdl_trigger();
  DV_GT(threadIdx.x == 0, 3);   // predecessor's data visible

  // split-K: the cluster = the splits of ONE tile (grid = tiles * splits, one pass);
  // otherwise a persistent loop over tiles
  const bool split_mode = k.splits > 1;
  const int split = split_mode ? static_cast<int>(cluster_ctarank()) : 0;
  const int tile0 =
      split_mode ? static_cast<int>(blockIdx.x) / k.splits : static_cast<int>(blockIdx.x);
  const int tile_step = split_mode ? k.total_tiles : static_cast<int>(gridDim.x);

  if (warp == 0) {
    // ============================ TMA producer ============================
    // (the whole converged warp runs the loop; the w_ forms elect the issuing lane and keep operands in uniform registers)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int gtile = tile0; gtile < k.total_tiles; gtile += tile_step) {
        int tile = gtile;
        const Problem& a = problem_of(k, tile);
        const GemmDesc& d = a.d;
        const TileCoord tc = decode_tile(a, tile, split);
        int ct = 0, h0 = 0, w0 = 0;
        if (d.a_mode == 1) {
          const int per_frame = a.tiles_w * a.tiles_h;
          ct = a.oT0 + tc.m_tile / per_frame;
          const int r = tc.m_tile % per_frame;
          h0 = (r / a.tiles_w) * (a.swap ? 16 : 8);
          w0 = (r % a.tiles_w) * 16;
        }
        for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + kABytes;
          w_mbar_expect_tx(&full_bar[stage], kStageBytes);
          if (d.a_mode == 1) {
            const int tap = kb / a.c_blocks;
            const int cb = kb - tap * a.c_blocks;
            const int dt = tap / (d.kh * d.kw);
            const int dh = (tap / d.kw) % d.kh;
            const int dw = tap % d.kw;
            // causal in time (kt-1 frames of zero history), centred in space; TMA's
            // out-of-bounds zero fill is the padding
            int x = w0 + dw - d.kw / 2, y = h0 + dh - d.kh / 2, t = ct + dt - (d.kt - 1);
            const CUtensorMap* tma = &a.tmA;
            if (a.sW == 2) {
              // stride (1,2,2): input pixel 2*o + e, e in {-1,0,1}: parity (e & 1) selects one of four
              // maps over the half-resolution grid, (e - parity) / 2 is the offset inside it
              const int ew = dw - d.kw / 2, eh = dh - d.kh / 2;
              const int pw = ew & 1, ph = eh & 1;
              x = w0 + (ew - pw) / 2;
              y = h0 + (eh - ph) / 2;
              const int q = ph * 2 + pw;
              if (q) tma = &a.tmAp[q - 1];
            } else if (a.sT == 2) {
              // stride (2,1,1), causal: input frame 2*o + e, e in {-2,-1,0}
              const int et = dt - (d.kt - 1);
              const int pt = et & 1;
              t = ct + (et - pt) / 2;
              if (pt) tma = &a.tmAp[0];
            }
            if (a.swap) {
              // weights (<= 128 rows) feed the M side, a 16x16-pixel box the N side
              w_tma_load_2d(&a.tmB, &full_bar[stage], sa, kb * BK, 0);
              w_tma_load_5d(tma, &full_bar[stage], sb, cb * 64, x, y, t, tc.b);
            } else {
              w_tma_load_5d(tma, &full_bar[stage], sa, cb * 64, x, y, t, tc.b);
              w_tma_load_2d(&a.tmB, &full_bar[stage], sb, kb * BK, tc.n_tile * BN);
            }
          } else {
            w_tma_load_3d(&a.tmA, &full_bar[stage], sa, kb * BK, tc.m_tile * BM, tc.b);
            if (d.w_batch_stride != 0)
              w_tma_load_3d(&a.tmB, &full_bar[stage], sb, kb * BK, tc.n_tile * BN, tc.b);
            else
              w_tma_load_2d(&a.tmB, &full_bar[stage], sb, kb * BK, tc.n_tile * BN);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ==============================
    {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int gtile = tile0; gtile < k.total_tiles; gtile += tile_step, ++it) {
        const int as = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + as * BN;
        int tile = gtile;
        const Problem& a = problem_of(k, tile);
        const TileCoord tc = decode_tile(a, tile, split);
        for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (it == 0 && kb == tc.kb0) DV_GT(lane == 0, 4);   // first operands landed
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint32_t sb = sa + kABytes;
          const uint64_t da = umma_desc_sw128(sa, 16, 1024);
          const uint64_t db = umma_desc_sw128(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // +32 B per K=16 slice inside the 128 B swizzle row (address field is >>4)
            w_umma_bf16_ss(tmem_acc, da + 2 * k, db + 2 * k, idesc, ((kb - tc.kb0) | k) != 0);
          }
          w_umma_commit(&empty_bar[stage]);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        DV_GT(lane == 0, 5);   // (last tile's) MMAs issued
        if (tc.kb0 < tc.kb1)
          w_umma_commit(&tmem_full[as]);
        else if (lane == 0)
          mbar_arrive(&tmem_full[as]);  // empty K range (tail split): nothing was issued
      }
    }
  } else {
    // ============================ epilogue ================================
    const int quarter = warp & 3;       // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;   // which 128 columns of the tile
    const int row_in_tile = quarter * 32 + lane;
    if (!split_mode) {
      int it = 0;
      for (int gtile = tile0; gtile < k.total_tiles; gtile += tile_step, ++it) {
        const int as = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        int tile = gtile;
        const Problem& a = problem_of(k, tile);
        const TileCoord tc = decode_tile(a, tile, 0);
        mbar_wait(&tmem_full[as], aph);
        tc_fence_after();
        DV_GT(threadIdx.x == 64, 6);   // (last tile's) accumulator complete
        if (MODE == EPI_CONV && a.swap) {
          epilogue_tile_swapped(a, tc, tmem_base + as * BN, quarter, lane, half);
        } else if (EpiTma<MODE>::value && a.tma_out) {
          if constexpr (EpiTma<MODE>::value)
            epilogue_tma<MODE>(a, tc, tmem_base + as * BN, row_in_tile, quarter, half, out_stage,
                               quarter == 0 && lane == 0);
        } else {
          epilogue_from_tmem<MODE>(a, tc, tmem_base + as * BN, row_in_tile, quarter, half,
                                   reinterpret_cast<float*>(out_stage));
        }
        tc_fence_before();
        mbar_arrive(&tmem_empty[as]);
      }
      if (EpiTma<MODE>::value && quarter == 0 && lane == 0) bulk_wait_all();  // my TMA stores have landed
    } else {
      // park the fp32 partial tile in this CTA's smem as [col4][row] float4 (the operand ring is
      // idle: every MMA that read it has retired once tmem_full fires)
      mbar_wait(&tmem_full[0], 0);
      tc_fence_after();
      DV_GT(threadIdx.x == 64, 6);   // accumulator complete
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
      float4* part = reinterpret_cast<float4*>(smem);
      int tile = tile0;
      const Problem& a = problem_of(k, tile);
      const bool empty = split * a.kb_per_split >= a.k_blocks;  // tail split without k-blocks
#pragma unroll 1
      for (int c = half * (BN / 64); c < (half + 1) * (BN / 64); ++c) {
        uint32_t raw[32];
        tmem_ld_32x32(taddr + c * 32, raw);
        tmem_ld_wait();
        if (empty) {
#pragma unroll
          for (int i = 0; i < 32; ++i) raw[i] = 0u;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
          part[(c * 8 + i) * BM + row_in_tile] =
              make_float4(__uint_as_float(raw[4 * i]), __uint_as_float(raw[4 * i + 1]),
                          __uint_as_float(raw[4 * i + 2]), __uint_as_float(raw[4 * i + 3]));
      }
      tc_fence_before();
    }
  }

  if (split_mode) {
    // every CTA of the cluster has parked its partial tile
    __syncwarp();
    cluster_sync_all();
    DV_GT(threadIdx.x == 64, 7);   // every split parked its partial tile
    if (warp >= 2) {
      constexpr int W = EpiW<MODE>::value;
      const int t = threadIdx.x - 64;      // 0..255
      const int rows_per = BM / k.splits;  // rows of the tile this CTA finishes
      const int row_in_tile = split * rows_per + (t % rows_per);
      const int group = t / rows_per;      // 0..2*splits-1: which column chunks
      int tile = tile0;
      const Problem& a = problem_of(k, tile);
      const GemmDesc& d = a.d;
      const TileCoord tc = decode_tile(a, tile, split);
      const RowCtx r = make_row(a, tc, row_in_tile);
      const uint32_t part0 = smem_u32(smem);
      for (int c = group; c < BN / W; c += 2 * k.splits) {
        const int n = tc.n_tile * BN + c * W;
        // (conv rows always enter when the warp is uniform: warp-wide GroupNorm reduction)
        if (n >= d.N || (!r.ok && !(MODE == EPI_CONV && rows_per >= 32))) continue;
        float v[W];
#pragma unroll
        for (int i = 0; i < W; ++i) v[i] = 0.f;
#pragma unroll 2
        for (int p = 0; p < k.splits; ++p) {  // fixed order: bit-reproducible sums (splits is 2, 4 or 8)
          const uint32_t peer = map_to_cta(part0, static_cast<uint32_t>(p));
#pragma unroll
          for (int i = 0; i < W / 4; ++i) {
            const float4 q = ld_dsmem_f4(
                peer + static_cast<uint32_t>(((c * (W / 4) + i) * BM + row_in_tile) * 16));
            v[4 * i + 0] += q.x;
            v[4 * i + 1] += q.y;
            v[4 * i + 2] += q.z;
            v[4 * i + 3] += q.w;
          }
        }
        epi_row<MODE, W>(a, r, n, v, rows_per >= 32);
      }
    }
    // nobody leaves while a peer may still read its shared memory
    __syncwarp();
    cluster_sync_relaxed();
  }

  DV_GT(threadIdx.x == 64, 8);   // epilogue / reduction done
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
  DV_GT(threadIdx.x == 0, 9);   // exit
}

// ------------------------------------------------------------------------------
// CTA-pair kernel (tcgen05 cta_group::2): a cluster of two CTAs on one TPC computes a 256 x 256
// tile.  Each CTA loads its own 128 A rows and HALF of the B tile (128 of the 256 N rows), so the
// L2 -> SM and shared-memory traffic per SM drops from 48 KB to 32 KB per k-block; the leader CTA
// issues one M=256 MMA that reads both CTAs' shared memory and writes both CTAs' TMEM.  Used when
// there are at least as many 256-row tiles as SM pairs (no split-K).
// ------------------------------------------------------------------------------
constexpr int kPairStages = 6;
constexpr int kPairStageBytes = kABytes + kBBytes / 2;  // 32 KB
constexpr int kPairSmemBytes = kPairStages * kPairStageBytes + kOutBytes + kBarBytes + 1024;

__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* m, uint32_t leader_bar, void* dst,
                                                 int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(const CUtensorMap* m, uint32_t leader_bar, void* dst,
                                                 int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1),
        "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_pair(const CUtensorMap* m, uint32_t leader_bar, void* dst,
                                                 int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :
      : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                                  uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the same-offset mbarrier of BOTH CTAs of the pair when the issued MMAs retire
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
// whole-warp forms (see common.cuh): all 32 converged lanes execute them, one elected lane issues
__device__ __forceinline__ void w_tma_load_2d_pair(const CUtensorMap* m, uint32_t leader_bar, void* dst, int c0, int c1) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];\n\t"
      "}\n"
      :
      : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void w_tma_load_3d_pair(const CUtensorMap* m, uint32_t leader_bar, void* dst, int c0, int c1,
                                                   int c2) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];\n\t"
      "}\n"
      :
      : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void w_tma_load_5d_pair(const CUtensorMap* m, uint32_t leader_bar, void* dst, int c0, int c1,
                                                   int c2, int c3, int c4) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n\t"
      "}\n"
      :
      : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
        "r"(c4)
      : "memory");
}
__device__ __forceinline__ void w_umma_bf16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                    uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void w_umma_commit_pair(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr)
               : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* slot_in_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(slot_in_smem)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols)
               : "memory");
}

__device__ __forceinline__ const Problem& pair_problem_of(const KArgs& k, int& tile) {
  if (tile < k.p[0].pair_tiles) return k.p[0];
  tile -= k.p[0].pair_tiles;
  return k.p[1];
}
__device__ __forceinline__ TileCoord decode_pair_tile(const Problem& a, int tile, int rank) {
  TileCoord tc;
  tc.n_tile = tile % a.n_tiles;
  const int rest = tile / a.n_tiles;
  tc.m_tile = 2 * (rest % a.m_pairs) + rank;  // may be == m_tiles (odd count): all rows out of range
  tc.b = rest / a.m_pairs;
  tc.kb0 = 0;
  tc.kb1 = a.k_blocks;
  return tc;
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) gemm_pair_kernel(const __grid_constant__ KArgs k) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* out_stage = smem + kPairStages * kPairStageBytes;  // TMA-store staging / GroupNorm partials
  uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + kOutBytes);
  uint64_t* full_bar = bars;                      // leader's are used (both CTAs' TMA signal them)
  uint64_t* empty_bar = bars + kPairStages;       // local, arrived by the leader's multicast commit
  uint64_t* tmem_full = bars + 2 * kPairStages;   // local, multicast commit
  uint64_t* tmem_empty = tmem_full + 2;           // leader's: 4 epilogue warps x 2 CTAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
#ifdef DV_GEMM_TRACE
  if (threadIdx.x == 0 && blockIdx.x < 4096) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    k.trace[blockIdx.x * 12] = smid;
  }
#endif
  DV_GT(threadIdx.x == 0, 1);   // entry
  const int rank = static_cast<int>(cluster_ctarank());
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&k.p[0].tmA);
    tma_prefetch_desc(&k.p[0].tmBh);
    if (k.p[0].tma_out) tma_prefetch_desc(&k.p[0].tmOut);
    if (k.p[1].tiles > 0) {
      tma_prefetch_desc(&k.p[1].tmA);
      tma_prefetch_desc(&k.p[1].tmBh);
      if (k.p[1].tma_out) tma_prefetch_desc(&k.p[1].tmOut);
    }
    for (int i = 0; i < kPairStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_pair<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncwarp();
  cluster_sync_all();  // peer barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  DV_GT(threadIdx.x == 0, 2);   // barriers + tensor memory set up
  pdl_wait();  // everything above overlapped the previous kernel's tail
  if (k.pdl_early) pdl_trigger();
  DV_GT(threadIdx.x == 0, 3);   // predecessor's data visible

  const int cluster_id = static_cast<int>(blockIdx.x) >> 1;
  const int n_clusters = static_cast<int>(gridDim.x) >> 1;

  if (warp == 0) {
    // ============================ TMA producer (both CTAs) ============================
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int gtile = cluster_id; gtile < k.total_pair_tiles; gtile += n_clusters) {
        int tile = gtile;
        const Problem& a = pair_problem_of(k, tile);
        const GemmDesc& d = a.d;
        const TileCoord tc = decode_pair_tile(a, tile, rank);
        int ct = 0, h0 = 0, w0 = 0;
        if (d.a_mode == 1) {
          const int per_frame = a.tiles_w * a.tiles_h;
          ct = a.oT0 + tc.m_tile / per_frame;
          const int r = tc.m_tile % per_frame;
          h0 = (r / a.tiles_w) * 8;
          w0 = (r % a.tiles_w) * 16;
        }
        const int n_row = tc.n_tile * BN + rank * (BN / 2);  // this CTA's half of the B tile
        for (int kb = 0; kb < a.k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kPairStageBytes;
          uint8_t* sb = sa + kABytes;
          const uint32_t lbar = map_to_cta(smem_u32(&full_bar[stage]), 0);
          if (leader) w_mbar_expect_tx(&full_bar[stage], 2 * kPairStageBytes);
          if (d.a_mode == 1) {
            const int tap = kb / a.c_blocks;
            const int cb = kb - tap * a.c_blocks;
            const int dt = tap / (d.kh * d.kw);
            const int dh = (tap / d.kw) % d.kh;
            const int dw = tap % d.kw;
            w_tma_load_5d_pair(&a.tmA, lbar, sa, cb * 64, w0 + dw - d.kw / 2, h0 + dh - d.kh / 2,
                             ct + dt - (d.kt - 1), tc.b);
            w_tma_load_2d_pair(&a.tmBh, lbar, sb, kb * BK, n_row);
          } else {
            w_tma_load_3d_pair(&a.tmA, lbar, sa, kb * BK, tc.m_tile * BM, tc.b);
            if (d.w_batch_stride != 0)
              w_tma_load_3d_pair(&a.tmBh, lbar, sb, kb * BK, n_row, tc.b);
            else
              w_tma_load_2d_pair(&a.tmBh, lbar, sb, kb * BK, n_row);
          }
          if (++stage == kPairStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer (leader CTA only) ========================
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int gtile = cluster_id; gtile < k.total_pair_tiles; gtile += n_clusters, ++it) {
        const int as = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + as * BN;
        int tile = gtile;
        const Problem& a = pair_problem_of(k, tile);
        for (int kb = 0; kb < a.k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (it == 0 && kb == 0) DV_GT(lane == 0, 4);   // first operands landed
          const uint32_t sa = smem_u32(smem + stage * kPairStageBytes);
          const uint32_t sb = sa + kABytes;
          const uint64_t da = umma_desc_sw128(sa, 16, 1024);
          const uint64_t db = umma_desc_sw128(sb, 16, 1024);
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk)
            w_umma_bf16_ss_pair(tmem_acc, da + 2 * kk, db + 2 * kk, idesc, (kb | kk) != 0);
          w_umma_commit_pair(&empty_bar[stage]);
          if (++stage == kPairStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        DV_GT(lane == 0, 5);   // (last tile's) MMAs issued
        w_umma_commit_pair(&tmem_full[as]);
      }
    }
  } else {
    // ============================ epilogue (both CTAs, own 128 rows) ==================
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row_in_tile = quarter * 32 + lane;
    int it = 0;
    for (int gtile = cluster_id; gtile < k.total_pair_tiles; gtile += n_clusters, ++it) {
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      int tile = gtile;
      const Problem& a = pair_problem_of(k, tile);
      const TileCoord tc = decode_pair_tile(a, tile, rank);
      mbar_wait(&tmem_full[as], aph);
      tc_fence_after();
      DV_GT(threadIdx.x == 64, 6);   // (last tile's) accumulator complete
      if (EpiTma<MODE>::value && a.tma_out) {
        if constexpr (EpiTma<MODE>::value)
          epilogue_tma<MODE>(a, tc, tmem_base + as * BN, row_in_tile, quarter, half, out_stage,
                             quarter == 0 && lane == 0);
      } else {
        epilogue_from_tmem<MODE>(a, tc, tmem_base + as * BN, row_in_tile, quarter, half,
                                 reinterpret_cast<float*>(out_stage));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(map_to_cta(smem_u32(&tmem_empty[as]), 0));
    }
    if (EpiTma<MODE>::value && quarter == 0 && lane == 0) bulk_wait_all();  // my TMA stores have landed
  }

  DV_GT(threadIdx.x == 64, 8);   // epilogue done
  pdl_trigger();
  tc_fence_before();
  __syncwarp();
  cluster_sync_all();  // the peer's MMAs / commits no longer touch this CTA
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair<kTmemCols>(tmem_base);
  }
  DV_GT(threadIdx.x == 0, 9);   // exit
}

// ------------------------------------------------------------------------------
// Persistent block kernel: the LN / GEMM steps between two attentions of the joint transformer blocks as PHASES of one
// launch, separated by grid barriers (one CTA per SM, all co-resident).  For token layouts whose GEMMs have fewer
// tiles than SMs a launch costs 17-30 us whatever its size (launch gap + set-up + pipeline fill + epilogue,
// DESIGN.md 3.1 finding 7); here those costs are paid once per block instead of six times.
//   kind LN      LayerNorm + adaLN modulate of the fp32 streams (one warp per row, rows over all warps of the grid)
//   kind GEMM    units = (tile, k-split) dealt round-robin over the CTAs; splits == 1: the usual epilogue (TMA store /
//                reduce-add or direct); splits > 1: the fp32 partial tile goes to a global workspace ([unit][col4][row])
//   kind REDUCE  finishes the preceding split GEMM: items = (tile, 32-row slice); the partials are summed in split order
//                (bit-reproducible) and the direct epilogue runs
// Data written in one phase and read in a later one crosses SMs inside ONE kernel: generic-proxy readers use ld.global.cg
// (L1 is not coherent), TMA stores are drained (wait_group 0) and both proxies fenced before every grid barrier.
// ------------------------------------------------------------------------------
constexpr int kPbkMaxPhases = 12;
constexpr int kPbkMaxProblems = 8;
enum PbkKind : int { PBK_LN = 0, PBK_GEMM = 1, PBK_REDUCE = 2 };

struct PbkLnSeg {
  const float* x;
  long long x_bs;
  __nv_bfloat16* out;
  long long out_bs;
  const float* shift;
  const float* scale;
  int L;
};
struct PbkPhase {
  int kind;
  int p0, np;      // problems of a GEMM / REDUCE phase: [p0, p0 + np)
  int mode;        // EpiMode
  int splits;
  int tiles;       // over the phase's problems
  PbkLnSeg ln[2];  // kind LN (ln[1].L == 0: one segment)
  int mod_bs, B;
  float eps;
};
struct PbkArgs {
  Problem p[kPbkMaxProblems];
  PbkPhase ph[kPbkMaxPhases];
  int n_phases;
  unsigned* bar;   // grid barrier counter, zero at launch
  float4* work;    // split-K partials
  int pdl_early;
  unsigned long long* trace;  // debug (DV_PBK_TRACE): [phase + 1][CTA] globaltimer stamps, or null
};

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// all threads of all CTAs; `epoch` counts the barriers of this launch
__device__ __forceinline__ void pbk_grid_sync(unsigned* bar, unsigned& epoch) {
  fence_proxy_async_all();   // my TMA writes (already complete) / generic writes, ordered before the release below
  __threadfence();
  __syncthreads();
  ++epoch;
  if (threadIdx.x == 0) {
    const unsigned target = epoch * gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    const uint64_t t0 = global_ns();
    while (true) {
      unsigned v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if (v >= target) break;
      if (global_ns() - t0 > 4000000000ull) asm volatile("trap;");
    }
    __threadfence();
  }
  __syncthreads();
  fence_proxy_async_all();   // later TMA loads see what the other CTAs wrote with generic stores
}

__device__ __forceinline__ const Problem& pbk_problem(const PbkArgs& k, const PbkPhase& ph, int& tile) {
  if (ph.np == 1 || tile < k.p[ph.p0].tiles) return k.p[ph.p0];
  tile -= k.p[ph.p0].tiles;
  return k.p[ph.p0 + 1];
}

__device__ __forceinline__ void pbk_ln_rows(const PbkPhase& ph, int gwarp, int nwarps, int lane) {
  constexpr int D = 1536, V = D / 128;
  const int rows0 = ph.B * ph.ln[0].L;
  const int rows = rows0 + ph.B * ph.ln[1].L;
  for (int w = gwarp; w < rows; w += nwarps) {
    const bool second = w >= rows0;
    const int wr = second ? w - rows0 : w;
    const PbkLnSeg& sg = second ? ph.ln[1] : ph.ln[0];
    const int b = wr / sg.L, l = wr - b * sg.L;
    const float4* xr = reinterpret_cast<const float4*>(sg.x + b * sg.x_bs + static_cast<long long>(l) * D);
    float4 v[V];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      v[i] = __ldcg(xr + lane + 32 * i);   // the stream was updated by other SMs in the previous phase
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
      q += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.0f / D) + ph.eps);
    const float4* sh = reinterpret_cast<const float4*>(sg.shift + static_cast<long long>(b) * ph.mod_bs);
    const float4* sc = reinterpret_cast<const float4*>(sg.scale + static_cast<long long>(b) * ph.mod_bs);
    uint2* o2 = reinterpret_cast<uint2*>(sg.out + b * sg.out_bs + static_cast<long long>(l) * D);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float4 h = __ldg(sh + lane + 32 * i);
      const float4 c = __ldg(sc + lane + 32 * i);
      uint2 pk;
      pk.x = pack_bf16x2((v[i].x - mean) * rstd * (1.0f + c.x) + h.x, (v[i].y - mean) * rstd * (1.0f + c.y) + h.y);
      pk.y = pack_bf16x2((v[i].z - mean) * rstd * (1.0f + c.z) + h.z, (v[i].w - mean) * rstd * (1.0f + c.w) + h.w);
      o2[lane + 32 * i] = pk;
    }
  }
}

template <int MODE>
__device__ __forceinline__ void pbk_epilogue(const Problem& a, const TileCoord& tc, uint32_t tmem_acc, int row_in_tile,
                                             int quarter, int half, uint8_t* out_stage, bool issuer) {
  if (a.tma_out)
    epilogue_tma<MODE>(a, tc, tmem_acc, row_in_tile, quarter, half, out_stage, issuer);
  else
    epilogue_from_tmem<MODE>(a, tc, tmem_acc, row_in_tile, quarter, half, nullptr);
}

// REDUCE item: rows [slice * 32, +32) of one tile; t = 0..255.  The partials of two splits are fetched together (their
// loads are independent, the L2 round trips overlap); the sum still runs in split order.
template <int MODE>
__device__ __forceinline__ void pbk_reduce_item(const Problem& a, const TileCoord& tc, const float4* work,
                                                int unit0, int splits, int slice, int t) {
  constexpr int W = EpiW<MODE>::value;
  const int row_in_tile = slice * 32 + (t & 31);
  const int group = t >> 5;  // 0..7
  const RowCtx r = make_row(a, tc, row_in_tile);
  if (!r.ok) return;
  for (int c = group; c < BN / W; c += 8) {
    const int n = tc.n_tile * BN + c * W;
    if (n >= a.d.N) continue;
    float v[W];
#pragma unroll
    for (int i = 0; i < W; ++i) v[i] = 0.f;
    const float4* part = work + static_cast<long long>(unit0) * (BM * BN / 4) + (c * (W / 4)) * BM + row_in_tile;
    constexpr int STEP = (W == 32) ? 2 : 1;   // W = 64 (q|k|v heads): one split at a time keeps the registers
#pragma unroll 1
    for (int sp = 0; sp < splits; sp += STEP) {  // fixed order: bit-reproducible sums
      float4 q[STEP][W / 4];
#pragma unroll
      for (int j = 0; j < STEP; ++j) {
        const bool live = sp + j < splits;
#pragma unroll
        for (int i = 0; i < W / 4; ++i)
          q[j][i] = live ? __ldcg(part + static_cast<long long>(sp + j) * (BM * BN / 4) + i * BM)
                         : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < STEP; ++j) {
#pragma unroll
        for (int i = 0; i < W / 4; ++i) {
          v[4 * i + 0] += q[j][i].x;
          v[4 * i + 1] += q[j][i].y;
          v[4 * i + 2] += q[j][i].z;
          v[4 * i + 3] += q[j][i].w;
        }
      }
    }
    epi_row<MODE, W>(a, r, n, v, false);
  }
}

__global__ void __launch_bounds__(kThreads, 1) pbk_kernel(const __grid_constant__ PbkArgs k) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* out_stage = smem + kStages * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + kOutBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 32 * kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  if (k.pdl_early) pdl_trigger();

  // per-role pipeline state, carried across the phases
  int stage = 0;
  uint32_t phase = 0;
  int acc_it = 0;
  unsigned epoch = 0;
  const int quarter = warp & 3;
  const int half = (warp - 2) >> 2;
  const int row_in_tile = quarter * 32 + lane;
  const bool issuer = warp >= 2 && quarter == 0 && lane == 0;
  constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0, 0);

  if (k.trace != nullptr && threadIdx.x == 0) k.trace[blockIdx.x] = global_ns();
  for (int pi = 0; pi < k.n_phases; ++pi) {
    const PbkPhase& ph = k.ph[pi];
    if (ph.kind == PBK_LN) {
      pbk_ln_rows(ph, static_cast<int>(blockIdx.x) * (kThreads / 32) + warp, static_cast<int>(gridDim.x) * (kThreads / 32), lane);
    } else if (ph.kind == PBK_GEMM) {
      const int units = ph.tiles * ph.splits;
      if (warp == 0) {
        if (lane == 0) {
          for (int u = blockIdx.x; u < units; u += gridDim.x) {
            int tile = u / ph.splits;
            const Problem& a = pbk_problem(k, ph, tile);
            const TileCoord tc = decode_tile(a, tile, u % ph.splits);
            for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* sa = smem + stage * kStageBytes;
              uint8_t* sb = sa + kABytes;
              mbar_expect_tx(&full_bar[stage], kStageBytes);
              tma_load_3d(&a.tmA, &full_bar[stage], sa, kb * BK, tc.m_tile * BM, tc.b);
              tma_load_2d(&a.tmB, &full_bar[stage], sb, kb * BK, tc.n_tile * BN);
              if (++stage == kStages) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
      } else if (warp == 1) {
        if (lane == 0) {
          for (int u = blockIdx.x; u < units; u += gridDim.x, ++acc_it) {
            const int as = acc_it & 1;
            const uint32_t aph = (acc_it >> 1) & 1;
            mbar_wait(&tmem_empty[as], aph ^ 1);
            tc_fence_after();
            const uint32_t tmem_acc = tmem_base + as * BN;
            int tile = u / ph.splits;
            const Problem& a = pbk_problem(k, ph, tile);
            const TileCoord tc = decode_tile(a, tile, u % ph.splits);
            for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
              mbar_wait(&full_bar[stage], phase);
              tc_fence_after();
              const uint32_t sa = smem_u32(smem + stage * kStageBytes);
              const uint32_t sb = sa + kABytes;
              const uint64_t da = umma_desc_sw128(sa, 16, 1024);
              const uint64_t db = umma_desc_sw128(sb, 16, 1024);
#pragma unroll
              for (int kk = 0; kk < BK / 16; ++kk)
                umma_bf16_ss(tmem_acc, da + 2 * kk, db + 2 * kk, idesc, ((kb - tc.kb0) | kk) != 0);
              umma_commit(&empty_bar[stage]);
              if (++stage == kStages) {
                stage = 0;
                phase ^= 1;
              }
            }
            if (tc.kb0 < tc.kb1)
              umma_commit(&tmem_full[as]);
            else
              mbar_arrive(&tmem_full[as]);
          }
        }
      } else {
        for (int u = blockIdx.x; u < units; u += gridDim.x, ++acc_it) {
          const int as = acc_it & 1;
          const uint32_t aph = (acc_it >> 1) & 1;
          int tile = u / ph.splits;
          const Problem& a = pbk_problem(k, ph, tile);
          const TileCoord tc = decode_tile(a, tile, u % ph.splits);
          mbar_wait(&tmem_full[as], aph);
          tc_fence_after();
          const uint32_t tmem_acc = tmem_base + as * BN;
          if (ph.splits == 1) {
            if (ph.mode == EPI_GELU)
              pbk_epilogue<EPI_GELU>(a, tc, tmem_acc, row_in_tile, quarter, half, out_stage, issuer);
            else if (ph.mode == EPI_QKV)
              pbk_epilogue<EPI_QKV>(a, tc, tmem_acc, row_in_tile, quarter, half, out_stage, issuer);
            else
              pbk_epilogue<EPI_RESID_GATE>(a, tc, tmem_acc, row_in_tile, quarter, half, out_stage, issuer);
          } else {
            // park the fp32 partial tile of this unit in the workspace as [col4][row] float4
            const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16);
            float4* part = k.work + static_cast<long long>(u) * (BM * BN / 4);
            const bool empty = tc.kb0 >= tc.kb1;
#pragma unroll 1
            for (int c = half * (BN / 64); c < (half + 1) * (BN / 64); ++c) {
              uint32_t raw[32];
              tmem_ld_32x32(taddr + c * 32, raw);
              tmem_ld_wait();
              if (empty) {
#pragma unroll
                for (int i = 0; i < 32; ++i) raw[i] = 0u;
              }
#pragma unroll
              for (int i = 0; i < 8; ++i)
                __stcg(part + (c * 8 + i) * BM + row_in_tile,
                       make_float4(__uint_as_float(raw[4 * i]), __uint_as_float(raw[4 * i + 1]),
                                   __uint_as_float(raw[4 * i + 2]), __uint_as_float(raw[4 * i + 3])));
            }
          }
          tc_fence_before();
          mbar_arrive(&tmem_empty[as]);
        }
        if (issuer) bulk_wait_all();  // my TMA stores / reduce-adds of this phase are performed
      }
    } else {  // PBK_REDUCE
      if (warp >= 2) {
        const int t = static_cast<int>(threadIdx.x) - 64;
        const int items = ph.tiles * (BM / 32);
        for (int it = blockIdx.x; it < items; it += gridDim.x) {
          int tile = it / (BM / 32);
          const int slice = it % (BM / 32);
          const int unit0 = tile * ph.splits;
          const Problem& a = pbk_problem(k, ph, tile);
          const TileCoord tc = decode_tile(a, tile, 0);
          if (ph.mode == EPI_GELU)
            pbk_reduce_item<EPI_GELU>(a, tc, k.work, unit0, ph.splits, slice, t);
          else if (ph.mode == EPI_QKV)
            pbk_reduce_item<EPI_QKV>(a, tc, k.work, unit0, ph.splits, slice, t);
          else
            pbk_reduce_item<EPI_RESID_GATE>(a, tc, k.work, unit0, ph.splits, slice, t);
        }
      }
    }
    if (k.trace != nullptr) {   // end of this CTA's work in the phase (before the barrier)
      __syncthreads();
      if (threadIdx.x == 0) k.trace[static_cast<long long>(pi + 1) * gridDim.x + blockIdx.x] = global_ns();
    }
    if (pi + 1 < k.n_phases) pbk_grid_sync(k.bar, epoch);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

template <int MODE>
int launch_mode(const KArgs& ka, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    DV_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<MODE>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  const dim3 grid(ka.splits > 1 ? ka.total_tiles * ka.splits
                                : (ka.total_tiles < sm_count() ? ka.total_tiles : sm_count()));
  DV_CHECK_CUDA(launch_pdl(gemm_tc_kernel<MODE>, grid, dim3(kThreads), kSmemBytes, stream, ka.splits, ka));
  note_launch();
  return 0;
}

template <int MODE>
int launch_pair_mode(const KArgs& ka, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    DV_CHECK_CUDA(cudaFuncSetAttribute(gemm_pair_kernel<MODE>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemBytes));
    attr_set = true;
  }
  const int max_clusters = sm_count() / 2;
  const int clusters = ka.total_pair_tiles < max_clusters ? ka.total_pair_tiles : max_clusters;
  DV_CHECK_CUDA(launch_pdl(gemm_pair_kernel<MODE>, dim3(2 * clusters), dim3(kThreads), kPairSmemBytes,
                           stream, 2, ka));
  note_launch();
  return 0;
}

// CTAs that can be resident at once when launched as clusters of `s` (GPC packing leaves a few SMs
// idle); the kernel's footprint does not depend on the epilogue mode
int cluster_capacity(int s) {
  static int cap[kMaxSplits + 1] = {0};
  if (cap[s] != 0) return cap[s];
  cudaFuncSetAttribute(gemm_tc_kernel<EPI_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kThreads);
  cfg.gridDim = dim3(s * 64);
  cfg.dynamicSmemBytes = kSmemBytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = s;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, gemm_tc_kernel<EPI_BF16>, &cfg) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    cap[s] = -1;
  } else {
    cap[s] = n * s;
  }
  return cap[s];
}

}  // namespace

// geometry + tensor maps of one problem
static int setup_problem(const GemmDesc& d_in, Problem& pr, bool allow_flat = true) {
  GemmDesc d = d_in;
  pr.flat_M = 0;
  pr.flat_B = 1;
  static const bool no_flat = getenv("DV_GEMM_NO_FLAT") != nullptr;
  if (allow_flat && !no_flat && d.a_mode == 0 && d.batch > 1 && d.w_batch_stride == 0 && d.peer_cols == 0 &&
      (d.mode == EPI_BF16 || d.mode == EPI_GELU || d.mode == EPI_RESID_GATE) && d.out_row_offset == 0 &&
      d.a_batch_stride == static_cast<long long>(d.M) * d.lda &&
      d.out_batch_stride == static_cast<long long>(d.M) * d.ldo) {
    pr.flat_M = d.M;
    pr.flat_B = d.batch;
    d.M *= d.batch;
    d.batch = 1;
  }
  pr.d = d;
  pr.swap = 0;
  DV_REQUIRE(d.batch > 0 && d.N > 0, "gemm: empty problem (batch=%d N=%d)", d.batch, d.N);
  int K;
  if (d.a_mode == 1) {
    DV_REQUIRE(d.mode == EPI_CONV, "conv operand needs EPI_CONV (mode %d)", d.mode);
    DV_REQUIRE(d.cC % 64 == 0, "conv: Cin=%d must be a multiple of 64 (pad at pack time)", d.cC);
    DV_REQUIRE(d.cW % 16 == 0 && d.cH % 8 == 0, "conv: H=%d W=%d must be multiples of 8/16", d.cH,
               d.cW);
    pr.c_blocks = d.cC / 64;
    K = d.kt * d.kh * d.kw * d.cC;
    pr.swap = (d.N <= 128 && d.conv_store == CONV_PLAIN && d.w_batch_stride == 0) ? 1 : 0;
    DV_REQUIRE(pr.swap || (d.N % 32 == 0 && d.out_C % 32 == 0),
               "conv: Cout=%d (stored channels %d) must be <= 128 with a plain store or multiples of 32",
               d.N, d.out_C);
    pr.sT = d.sT > 1 ? d.sT : 1;
    pr.sH = d.sH > 1 ? d.sH : 1;
    pr.sW = d.sW > 1 ? d.sW : 1;
    DV_REQUIRE((pr.sT == 1 && pr.sH == pr.sW && pr.sW <= 2) || (pr.sT == 2 && pr.sH == 1 && pr.sW == 1),
               "conv: stride (%d,%d,%d) unsupported (only (1,1,1), (1,2,2), (2,1,1))", pr.sT, pr.sH, pr.sW);
    DV_REQUIRE(pr.sT * pr.sH * pr.sW == 1 || (d.kt == 3 && d.conv_store == CONV_PLAIN),
               "conv: strided convs are 3x3x3 with a plain store");
    DV_REQUIRE(d.cH % pr.sH == 0 && d.cW % pr.sW == 0, "conv: H=%d W=%d not divisible by the stride", d.cH, d.cW);
    pr.oT = (d.cT - 1) / pr.sT + 1;
    pr.oH = d.cH / pr.sH;
    pr.oW = d.cW / pr.sW;
    DV_REQUIRE(pr.oW % 16 == 0 && pr.oH % 8 == 0, "conv: output H=%d W=%d must be multiples of 8/16", pr.oH,
               pr.oW);
    pr.tiles_w = pr.oW / 16;
    pr.tiles_h = pr.swap ? (pr.oH + 15) / 16 : pr.oH / 8;
    pr.oT0 = d.conv_t0;
    DV_REQUIRE(pr.oT0 >= 0 && pr.oT0 < pr.oT && (pr.oT0 == 0 || pr.sT * pr.sH * pr.sW == 1),
               "conv: first frame %d of %d (stride-1 convs only)", pr.oT0, pr.oT);
    pr.m_tiles = (pr.oT - pr.oT0) * pr.tiles_w * pr.tiles_h;
  } else {
    DV_REQUIRE(d.mode != EPI_CONV, "EPI_CONV needs the conv operand");
    DV_REQUIRE(d.K % 64 == 0, "gemm: K=%d must be a multiple of 64", d.K);
    DV_REQUIRE(d.M > 0, "gemm: M=%d", d.M);
    K = d.K;
    pr.c_blocks = 1;
    pr.tiles_w = pr.tiles_h = 1;
    pr.m_tiles = (d.M + BM - 1) / BM;
    pr.oT = pr.oH = pr.oW = 0;
    pr.oT0 = 0;
    pr.sT = pr.sH = pr.sW = 1;
  }
  pr.k_blocks = K / BK;
  DV_REQUIRE(d.mode == EPI_UNPATCH || d.mode == EPI_CONV || d.N % 32 == 0,
             "gemm: N=%d must be a multiple of 32 for epilogue mode %d", d.N, d.mode);
  DV_REQUIRE(d.mode != EPI_QKV || d.N % 64 == 0, "gemm: QKV epilogue needs N %% 64 == 0 (N=%d)",
             d.N);
  pr.n_tiles = pr.swap ? 1 : (d.N + BN - 1) / BN;
  const long long tiles_ll = static_cast<long long>(pr.m_tiles) * d.batch * pr.n_tiles;
  DV_REQUIRE(tiles_ll < (1ll << 28), "gemm: too many tiles");
  pr.tiles = static_cast<int>(tiles_ll);
  pr.kb_per_split = pr.k_blocks;

  if (d.a_mode == 1) {
    uint64_t dims[5] = {(uint64_t)d.cC, (uint64_t)d.cW, (uint64_t)d.cH, (uint64_t)d.cT,
                        (uint64_t)d.batch};
    uint64_t strides[4] = {(uint64_t)d.cC * 2, (uint64_t)d.cC * d.cW * 2,
                           (uint64_t)d.cC * d.cW * d.cH * 2,
                           (uint64_t)d.cC * d.cW * d.cH * d.cT * 2};
    uint32_t box[5] = {64, 16, pr.swap ? 16u : 8u, 1, 1};
    const char* base = reinterpret_cast<const char*>(d.A);
    const uint64_t px = (uint64_t)d.cC * 2, row = px * d.cW, frame = row * d.cH;
    if (pr.sW == 2) {
      // four maps over the half-resolution grid, one per (row parity, column parity)
      dims[1] = d.cW / 2;
      dims[2] = d.cH / 2;
      strides[0] = 2 * px;
      strides[1] = 2 * row;
      for (int q = 0; q < 4; ++q) {
        int rc = make_tensor_map_bf16(q ? &pr.tmAp[q - 1] : &pr.tmA, base + (q >> 1) * row + (q & 1) * px, 5,
                                      dims, strides, box, 1);
        if (rc) return rc;
      }
    } else if (pr.sT == 2) {
      // two maps, one per frame parity (even frames: ceil(T/2), odd frames: floor(T/2))
      strides[2] = 2 * frame;
      for (int q = 0; q < 2; ++q) {
        const int frames = q == 0 ? (d.cT + 1) / 2 : d.cT / 2;
        dims[3] = frames > 0 ? frames : 1;  // (T = 1: the odd map is only ever addressed out of bounds)
        int rc = make_tensor_map_bf16(q ? &pr.tmAp[0] : &pr.tmA, base + q * frame, 5, dims, strides, box, 1);
        if (rc) return rc;
      }
    } else {
      int rc = make_tensor_map_bf16(&pr.tmA, d.A, 5, dims, strides, box, 1);
      if (rc) return rc;
    }
  } else {
    uint64_t dims[3] = {(uint64_t)d.K, (uint64_t)d.M, (uint64_t)d.batch};
    uint64_t strides[2] = {(uint64_t)d.lda * 2, (uint64_t)d.a_batch_stride * 2};
    if (d.batch == 1) strides[1] = (uint64_t)d.lda * 2 * (uint64_t)d.M;
    uint32_t box[3] = {64, 128, 1};
    int rc = make_tensor_map_bf16(&pr.tmA, d.A, 3, dims, strides, box, 1);
    if (rc) return rc;
  }
  const uint64_t ldw_bytes = (uint64_t)(d.ldw > 0 ? d.ldw : K) * 2;
  if (d.w_batch_stride != 0) {
    uint64_t dims[3] = {(uint64_t)K, (uint64_t)d.w_rows, (uint64_t)d.batch};
    uint64_t strides[2] = {ldw_bytes, (uint64_t)d.w_batch_stride * 2};
    uint32_t box[3] = {64, (uint32_t)BN, 1};
    int rc = make_tensor_map_bf16(&pr.tmB, d.W, 3, dims, strides, box, 1);
    if (rc) return rc;
    box[1] = BN / 2;
    rc = make_tensor_map_bf16(&pr.tmBh, d.W, 3, dims, strides, box, 1);
    if (rc) return rc;
  } else {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)d.w_rows};
    uint64_t strides[1] = {ldw_bytes};
    uint32_t box[2] = {64, pr.swap ? 128u : (uint32_t)BN};
    int rc = make_tensor_map_bf16(&pr.tmB, d.W, 2, dims, strides, box, 1);
    if (rc) return rc;
    box[1] = BN / 2;
    rc = make_tensor_map_bf16(&pr.tmBh, d.W, 2, dims, strides, box, 1);
    if (rc) return rc;
  }
  pr.m_pairs = (pr.m_tiles + 1) / 2;
  pr.pair_tiles = pr.m_pairs * d.batch * pr.n_tiles;

  // TMA-store epilogue (dense bf16 outputs and the gated fp32 residual; not with peer stores or a bf16 residual)
  pr.tma_out = 0;
  static const bool no_tma_store = getenv("DV_GEMM_NO_TMA_STORE") != nullptr;
  const bool mode_ok = d.mode == EPI_BF16 || d.mode == EPI_GELU || d.mode == EPI_QKV || d.mode == EPI_RESID_GATE;
  if (!no_tma_store && d.a_mode == 0 && mode_ok && d.residual == nullptr && d.peer_cols == 0) {
    const uint64_t esz = d.mode == EPI_RESID_GATE ? 4 : 2;
    const uint64_t row_bytes = static_cast<uint64_t>(d.ldo) * esz;
    const uint64_t batch_bytes = d.batch == 1 ? row_bytes * static_cast<uint64_t>(d.M)
                                              : static_cast<uint64_t>(d.out_batch_stride) * esz;
    const char* base = reinterpret_cast<const char*>(d.out) + static_cast<uint64_t>(d.out_row_offset) * row_bytes;
    if (row_bytes % 16 == 0 && batch_bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0) {
      uint64_t dims[3] = {(uint64_t)d.N, (uint64_t)d.M, (uint64_t)d.batch};
      uint64_t strides[2] = {row_bytes, batch_bytes};
      uint32_t box[3] = {static_cast<uint32_t>(64 / esz), 128, 1};
      int rc = make_tensor_map(&pr.tmOut, base, esz == 4 ? 1 : 0, 3, dims, strides, box, 2);
      if (rc) return rc;
      pr.tma_out = 1;
    }
  }
  return 0;
}

// ---- split-K choice: a CTA needs ~570 cycles per k-block (4 MMAs at the 128-cycle floor) plus
// ~3000 cycles of fixed latency (first TMA round trip, epilogue); with fewer tiles than SMs the
// K loop is cut over a cluster.  Returns the modelled cycles, the split factor in *splits_out.
static double plan_splits(long long tiles, int k_blocks, bool allow_split, int* splits_out) {
  const int nsm = sm_count();
  double best = 1e30;
  int splits = 1;
  static const bool split_large = getenv("DV_GEMM_SPLIT_LARGE") != nullptr;  // round-1 behaviour (A/B switch)
  for (int s = 1; s <= kMaxSplits; s *= 2) {
    if (s > 1 && (!allow_split || k_blocks / s < 2)) break;
    // split-K CTAs run one pass each (no overlap of their reduction + epilogue with the next mainloop), which the
    // model below does not see: with at least one tile per SM the persistent kernel / CTA pairs are used instead
    if (s > 1 && tiles >= nsm && !split_large) break;
    const int usable = s == 1 ? nsm : cluster_capacity(s);
    if (usable <= 0) break;
    const long long ctas = tiles * s;
    const double waves = static_cast<double>((ctas + usable - 1) / usable);
    const double kbs = static_cast<double>((k_blocks + s - 1) / s);
    const double t = waves * (kbs * 570.0 + 3000.0) + (s > 1 ? 2500.0 : 0.0);
    if (t < best * 0.93) {  // a larger cluster has to pay for itself
      best = t;
      splits = s;
    }
  }
  *splits_out = splits;
  return best;
}

static double problem_flops(const GemmDesc& d, int K, const Problem& pr) {
  const double rows = static_cast<double>(d.a_mode == 1 ? (double)(pr.oT - pr.oT0) * pr.oH * pr.oW : d.M) * d.batch;
  return 2.0 * rows * d.N * K;
}

static int launch_args(KArgs& ka, int mode, bool pair, cudaStream_t stream) {
  if (pair) {
    switch (mode) {
      case EPI_BF16: return launch_pair_mode<EPI_BF16>(ka, stream);
      case EPI_GELU: return launch_pair_mode<EPI_GELU>(ka, stream);
      case EPI_RESID_GATE: return launch_pair_mode<EPI_RESID_GATE>(ka, stream);
      case EPI_F32_ADD: return launch_pair_mode<EPI_F32_ADD>(ka, stream);
      case EPI_QKV: return launch_pair_mode<EPI_QKV>(ka, stream);
      case EPI_CONV: return launch_pair_mode<EPI_CONV>(ka, stream);
      case EPI_BF16_ROWBIAS: return launch_pair_mode<EPI_BF16_ROWBIAS>(ka, stream);
      case EPI_TAPS: return launch_pair_mode<EPI_TAPS>(ka, stream);
      default: break;  // no pair variant: fall through to the single-CTA kernel
    }
  }
  switch (mode) {
    case EPI_BF16: return launch_mode<EPI_BF16>(ka, stream);
    case EPI_GELU: return launch_mode<EPI_GELU>(ka, stream);
    case EPI_RESID_GATE: return launch_mode<EPI_RESID_GATE>(ka, stream);
    case EPI_F32_ADD: return launch_mode<EPI_F32_ADD>(ka, stream);
    case EPI_QKV: return launch_mode<EPI_QKV>(ka, stream);
    case EPI_UNPATCH: return launch_mode<EPI_UNPATCH>(ka, stream);
    case EPI_CONV: return launch_mode<EPI_CONV>(ka, stream);
    case EPI_BF16_ROWBIAS: return launch_mode<EPI_BF16_ROWBIAS>(ka, stream);
    case EPI_TAPS: return launch_mode<EPI_TAPS>(ka, stream);
    default:
      set_error("gemm: unknown epilogue mode %d", mode);
      return -1;
  }
}

long long pbk_workspace_floats() { return static_cast<long long>(sm_count()) * BM * BN; }

int launch_pbk(const PbkPhaseIn* phases, int n, float* workspace, unsigned* bar, cudaStream_t stream) {
  DV_REQUIRE(phases && n >= 1 && workspace && bar, "pbk: bad argument");
  static bool attr_set = false;
  if (!attr_set) {
    DV_CHECK_CUDA(cudaFuncSetAttribute(pbk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    int occ = 0;
    DV_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pbk_kernel, kThreads, kSmemBytes));
    DV_REQUIRE(occ >= 1, "pbk: the kernel does not fit an SM");
    attr_set = true;
  }
  const int grid = sm_count();   // one CTA per SM, all co-resident: the grid barrier relies on it
  PbkArgs ka;                    // ~13 KB, passed by value (large kernel parameters)
  ka.n_phases = 0;
  ka.bar = bar;
  ka.work = reinterpret_cast<float4*>(workspace);
  ka.pdl_early = pdl_early() ? 1 : 0;
  // DV_PBK_TRACE=1: per-phase time stamps of every CTA behind the barrier word ([kPbkMaxPhases + 1][grid] u64; the
  // caller's `bar` buffer must then be that large — scripts/probe/pbk_trace.py)
  ka.trace = getenv("DV_PBK_TRACE") ? reinterpret_cast<unsigned long long*>(bar) + 8 : nullptr;
  int np = 0;
  double flops = 0.0, bytes = 0.0;
  for (int i = 0; i < n; ++i) {
    const PbkPhaseIn& in = phases[i];
    DV_REQUIRE(ka.n_phases + 2 <= kPbkMaxPhases, "pbk: too many phases");
    PbkPhase& ph = ka.ph[ka.n_phases++];
    ph = PbkPhase{};
    if (in.kind == 0) {
      ph.kind = PBK_LN;
      auto fill = [](PbkLnSeg& sg, const LnRows& r) {
        sg.x = r.x;
        sg.x_bs = r.x_bs;
        sg.out = r.out;
        sg.out_bs = r.out_bs;
        sg.shift = r.shift;
        sg.scale = r.scale;
        sg.L = r.L;
      };
      fill(ph.ln[0], in.ln0);
      if (in.has_ln1) {
        fill(ph.ln[1], in.ln1);
      } else {
        ph.ln[1] = ph.ln[0];
        ph.ln[1].L = 0;
      }
      ph.mod_bs = in.mod_bs;
      ph.B = in.B;
      ph.eps = in.eps;
      bytes += static_cast<double>(in.B) * (in.ln0.L + (in.has_ln1 ? in.ln1.L : 0)) * 1536.0 * 6.0;
      // probe (DV_PBK_TRACE=2): run every LN phase twice (it is idempotent) to see what the second, instruction-cache-warm
      // pass costs
      if (getenv("DV_PBK_TRACE") && atoi(getenv("DV_PBK_TRACE")) == 2 && ka.n_phases + 3 <= kPbkMaxPhases) {
        ka.ph[ka.n_phases] = ph;
        ++ka.n_phases;
      }
      continue;
    }
    DV_REQUIRE(in.kind == 1, "pbk: phase kind %d", in.kind);
    DV_REQUIRE(np + 1 + in.has_g1 <= kPbkMaxProblems, "pbk: too many GEMM problems");
    DV_REQUIRE(in.g0.a_mode == 0 && (in.g0.mode == EPI_GELU || in.g0.mode == EPI_QKV || in.g0.mode == EPI_RESID_GATE),
               "pbk: dense GELU / QKV / gated-residual problems only (mode %d)", in.g0.mode);
    ph.kind = PBK_GEMM;
    ph.p0 = np;
    ph.np = 1 + (in.has_g1 ? 1 : 0);
    ph.mode = in.g0.mode;
    int rc = setup_problem(in.g0, ka.p[np], false);
    if (rc) return rc;
    if (in.has_g1) {
      DV_REQUIRE(in.g1.mode == in.g0.mode && in.g1.a_mode == 0 && in.g1.K == in.g0.K && in.g1.N == in.g0.N,
                 "pbk: the two problems of a phase must share mode, N and K");
      rc = setup_problem(in.g1, ka.p[np + 1], false);
      if (rc) return rc;
    }
    ph.tiles = ka.p[np].tiles + (in.has_g1 ? ka.p[np + 1].tiles : 0);
    const int kb = ka.p[np].k_blocks;
    int splits = 1;
    if (ph.tiles < grid) {
      splits = grid / ph.tiles;
      if (splits > 6) splits = 6;
      while (splits > 1 && kb / splits < 4) --splits;
    }
    ph.splits = splits;
    for (int j = 0; j < ph.np; ++j) ka.p[np + j].kb_per_split = (kb + splits - 1) / splits;
    for (int j = 0; j < ph.np; ++j) {
      const GemmDesc& d = j ? in.g1 : in.g0;
      const double rows = static_cast<double>(d.M) * d.batch;
      flops += 2.0 * rows * d.N * d.K;
      bytes += 2.0 * (rows * d.K + static_cast<double>(d.N) * d.K + rows * d.N);
    }
    np += ph.np;
    if (splits > 1) {
      DV_REQUIRE(static_cast<long long>(ph.tiles) * splits <= grid, "pbk: %d units of partials exceed the workspace", ph.tiles * splits);
      PbkPhase& red = ka.ph[ka.n_phases++];
      red = ph;
      red.kind = PBK_REDUCE;
    }
  }
  DV_CHECK_CUDA(cudaMemsetAsync(bar, 0, sizeof(unsigned), stream));
  char tag[56] = "pbk";
  if (prof_on()) {
    int rows = 0;
    for (int i = 0; i < n && rows == 0; ++i)
      if (phases[i].kind == 1) rows = phases[i].g0.batch * (phases[i].g0.M + (phases[i].has_g1 ? phases[i].g1.M : 0));
    snprintf(tag, sizeof(tag), "pbk rows%d phases%d", rows, ka.n_phases);
  }
  ProfScope ps(PROF_GEMM, flops, bytes, stream, tag);
  DV_CHECK_CUDA(launch_pdl(pbk_kernel, dim3(grid), dim3(kThreads), kSmemBytes, stream, 1, ka));
  note_launch();
  return 0;
}

int launch_gemm(const GemmDesc& d, cudaStream_t stream) { return launch_gemm_pair(d, nullptr, stream); }

#ifdef DV_GEMM_TRACE
long long* g_gemm_trace = nullptr;
extern "C" long long* dv_gemm_trace_buffer() {
  if (!g_gemm_trace) {
    cudaMalloc(&g_gemm_trace, 4096 * 12 * sizeof(long long));
    cudaMemset(g_gemm_trace, 0, 4096 * 12 * sizeof(long long));
  }
  return g_gemm_trace;
}
#endif

int launch_gemm_pair(const GemmDesc& d0, const GemmDesc* d1, cudaStream_t stream) {
  static const bool no_split = getenv("DV_GEMM_NOSPLIT") != nullptr;
  static const bool no_group = getenv("DV_GEMM_NOGROUP") != nullptr;
  if (d1 == nullptr && d0.a_mode == 1) {
    // 3x3x3 stride-1 convs with <= 128 output channels: pixel tile + halo shared by nine taps (conv_halo.cu)
    const int hrc = launch_conv_halo(d0, stream);
    if (hrc <= 0) return hrc;
  }
  KArgs ka;
  int rc = setup_problem(d0, ka.p[0]);
  if (rc) return rc;
  ka.p[1].tiles = 0;
  const bool split_ok0 = !ka.p[0].swap && !no_split;
  if (d1 != nullptr) {
    DV_REQUIRE(d1->mode == d0.mode, "gemm pair: epilogue modes differ (%d vs %d)", d0.mode, d1->mode);
    DV_REQUIRE(d0.a_mode == 0 && d1->a_mode == 0, "gemm pair: dense problems only");
    rc = setup_problem(*d1, ka.p[1]);
    if (rc) return rc;
    // one launch for both (the small problem fills the tail of the big one) unless wave
    // quantisation makes two launches cheaper
    int s_joint, s0, s1;
    const int kb_max = ka.p[0].k_blocks > ka.p[1].k_blocks ? ka.p[0].k_blocks : ka.p[1].k_blocks;
    const int kb_min = ka.p[0].k_blocks < ka.p[1].k_blocks ? ka.p[0].k_blocks : ka.p[1].k_blocks;
    const double t_joint = plan_splits(static_cast<long long>(ka.p[0].tiles) + ka.p[1].tiles, kb_max,
                                       !no_split && kb_min >= 2, &s_joint);
    const double t_sep = plan_splits(ka.p[0].tiles, ka.p[0].k_blocks, !no_split, &s0) +
                         plan_splits(ka.p[1].tiles, ka.p[1].k_blocks, !no_split, &s1) + 6000.0;
    if (no_group || t_sep < t_joint) {
      rc = launch_gemm_pair(d0, nullptr, stream);
      if (rc) return rc;
      return launch_gemm_pair(*d1, nullptr, stream);
    }
    while (s_joint > 1 && kb_min / s_joint < 1) s_joint /= 2;
    ka.splits = s_joint;
  } else {
    int s;
    plan_splits(ka.p[0].tiles, ka.p[0].k_blocks, split_ok0, &s);
    ka.splits = s;
  }
  for (int i = 0; i < 2; ++i)
    if (ka.p[i].tiles > 0) ka.p[i].kb_per_split = (ka.p[i].k_blocks + ka.splits - 1) / ka.splits;
  ka.total_tiles = ka.p[0].tiles + ka.p[1].tiles;
  ka.pdl_early = pdl_early() ? 1 : 0;
#ifdef DV_GEMM_TRACE
  ka.trace = dv_gemm_trace_buffer();
#endif
  if (ka.p[1].tiles == 0) ka.p[1].pair_tiles = 0;
  ka.total_pair_tiles = ka.p[0].pair_tiles + ka.p[1].pair_tiles;
  // CTA pairs (cta_group::2) once every SM pair has a 256-row tile: ~570 cycles per k-block instead
  // of the ~850 a lone CTA gets when all SMs pull 48 KB per k-block through L2
  bool pair = false;
  // Stride-1 implicit-GEMM convs with >= 256 output channels (not operand-swapped) pair as well: half of the
  // weight tile per CTA takes the k-block from ~1000 to ~880 cycles (1170 -> 1300 TFLOP/s, A/B on one box:
  // profiles/r01_exp_conv_pair.txt).  DV_CONV_NOPAIR=1 switches it off.
  static const bool conv_nopair = getenv("DV_CONV_NOPAIR") != nullptr;
  const bool conv_ok = !conv_nopair && d0.a_mode == 1 && d1 == nullptr && !ka.p[0].swap && d0.sT <= 1 &&
                       d0.sH <= 1 && d0.sW <= 1;
  if (ka.splits == 1 && (d0.a_mode == 0 || conv_ok) && d0.mode != EPI_UNPATCH) {
    const int nsm = sm_count();
    const double w1 = static_cast<double>((ka.total_tiles + nsm - 1) / nsm);
    const double w2 = static_cast<double>((ka.total_pair_tiles + nsm / 2 - 1) / (nsm / 2));
    pair = ka.total_tiles >= nsm && w2 * 0.75 < w1;   // (pairing a 144-tile problem measured 0 / -8 %)
    static const char* ep = getenv("DV_GEMM_PAIR");
    if (ep) pair = atoi(ep) != 0 && ka.total_pair_tiles > 0;
  }

  const int K0 = ka.p[0].k_blocks * BK;
  double flops = problem_flops(d0, K0, ka.p[0]);
  double rows = static_cast<double>(d0.a_mode == 1 ? (double)(ka.p[0].oT - ka.p[0].oT0) * ka.p[0].oH * ka.p[0].oW : d0.M) * d0.batch;
  double bytes = 2.0 * (rows * K0 / (d0.a_mode == 1 ? d0.kt * d0.kh * d0.kw : 1) + (double)d0.N * K0 + rows * d0.N);
  if (d1 != nullptr) {
    const int K1 = ka.p[1].k_blocks * BK;
    flops += problem_flops(*d1, K1, ka.p[1]);
    bytes += 2.0 * ((double)d1->M * d1->batch * (K1 + d1->N) + (double)d1->N * K1);
  }
  char tag[56] = "";
  if (prof_on()) {
    if (d0.a_mode == 1)
      snprintf(tag, sizeof(tag), "conv T%d H%d W%d Ci%d N%d k%d%s%s s%d e%d", d0.cT - d0.conv_t0, d0.cH, d0.cW,
               d0.cC, d0.N, d0.kt, ka.p[0].swap ? " sw" : "", pair ? " 2cta" : "", ka.splits,
               d0.conv_store);
    else if (d1 != nullptr)
      snprintf(tag, sizeof(tag), "gemm B%d M%d+%d N%d K%d s%d%s e%d%s", d0.batch, d0.M, d1->M, d0.N, K0,
               ka.splits, pair ? " 2cta" : "", d0.mode, ka.p[0].flat_M ? " flat" : "");
    else
      snprintf(tag, sizeof(tag), "gemm B%d M%d N%d K%d s%d%s e%d%s", d0.batch, d0.M, d0.N, K0, ka.splits,
               pair ? " 2cta" : "", d0.mode, ka.p[0].flat_M ? " flat" : "");
  }
  const int pid = prof_begin(d0.a_mode == 1 ? PROF_CONV : PROF_GEMM, flops, bytes, stream, tag);
  rc = launch_args(ka, d0.mode, pair, stream);
  prof_end(pid, stream);
  return rc;
}

}  // namespace dv

// deepv_b200 — HBM-bound helper kernels of the MMDiT forward (vectorised, coalesced).
#include "kernels.cuh"

namespace dv {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------
// LayerNorm (no affine) + adaLN modulate, one warp per token row, D = 1536.
// Reference: AdaLayerNormZero / AdaLayerNormContinuous / norm2 (+ modulate)
// model/mmdit.py:558,505,412-413,427-428; LN eps 1e-6 inside the sqrt, biased variance.
// ---------------------------------------------------------------------------------
struct LnSeg {
  const float* x;
  long long x_bs;
  __nv_bfloat16* out;
  long long out_bs;
  const float* shift;
  const float* scale;
  int L;
};
struct LnArgs {
  LnSeg seg[2];  // seg[1].L == 0: single segment (rows of seg[0] come first)
  int mod_bs, B;
  float eps;
  int pdl_early;
};

template <int D>
__global__ void __launch_bounds__(256) ln_modulate_kernel(const __grid_constant__ LnArgs a) {
  constexpr int V = D / 128;  // float4 per lane
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int rows0 = a.B * a.seg[0].L;
  pdl_wait();
  if (a.pdl_early) pdl_trigger();
  if (warp >= rows0 + a.B * a.seg[1].L) return;
  const bool second = warp >= rows0;
  if (second) warp -= rows0;
  const LnSeg& sg = second ? a.seg[1] : a.seg[0];
  const int b = warp / sg.L, l = warp - b * sg.L;
  const float4* xr = reinterpret_cast<const float4*>(sg.x + b * sg.x_bs + static_cast<long long>(l) * D);
  float4 v[V];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    v[i] = xr[lane + 32 * i];
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
    q += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + a.eps);
  const float4* sh = reinterpret_cast<const float4*>(sg.shift + static_cast<long long>(b) * a.mod_bs);
  const float4* sc = reinterpret_cast<const float4*>(sg.scale + static_cast<long long>(b) * a.mod_bs);
  uint2* o = reinterpret_cast<uint2*>(sg.out + b * sg.out_bs + static_cast<long long>(l) * D);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float4 h = __ldg(sh + lane + 32 * i);
    const float4 c = __ldg(sc + lane + 32 * i);
    float y0 = (v[i].x - mean) * rstd * (1.0f + c.x) + h.x;
    float y1 = (v[i].y - mean) * rstd * (1.0f + c.y) + h.y;
    float y2 = (v[i].z - mean) * rstd * (1.0f + c.z) + h.z;
    float y3 = (v[i].w - mean) * rstd * (1.0f + c.w) + h.w;
    uint2 pk;
    pk.x = pack_bf16x2(y0, y1);
    pk.y = pack_bf16x2(y2, y3);
    o[lane + 32 * i] = pk;
  }
}

// ---------------------------------------------------------------------------------
// Skinny GEMV for the conditioning path (M = CFG batch <= 4): pure weight streaming.
// One warp per output row, 16-byte loads, activations staged once per CTA in smem.
// Reference: TimestepEmbedding / TextProjection / adaLN linears, mmdit.py:718-736,548,495.
// ---------------------------------------------------------------------------------
constexpr int kGemvRowsPerWarp = 8;
constexpr int kGemvWarps = 8;

template <int NB>
__global__ void __launch_bounds__(kGemvWarps * 32) gemv_kernel(
    const __nv_bfloat16* __restrict__ W, const float* __restrict__ bias,
    const float* __restrict__ in, int in_stride, float* __restrict__ out, int out_stride, int N,
    int K, int silu_in, int accumulate, int row_blocks, const int* __restrict__ need) {
  extern __shared__ float s_in[];  // [NB][K]
  if (need != nullptr) {  // conditioning cache: every batch row was found, nothing to compute (uniform over the grid)
    int any = 0;
#pragma unroll
    for (int b = 0; b < NB; ++b) any |= need[b];
    if (!any) return;
  }
  for (int i = threadIdx.x; i < NB * K; i += blockDim.x) {
    const int b = i / K, k = i - b * K;
    float v = in[b * in_stride + k];
    s_in[i] = silu_in ? silu(v) : v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // four weight rows in flight per warp (4 x 16 B loads per lane and k-step): the kernel is pure
  // weight streaming, so memory-level parallelism is what sets its bandwidth.  CTAs are persistent
  // over row blocks, so the activation staging above (and its bubble) is paid once per CTA.
  constexpr int R = 4;
  for (int blk = blockIdx.x; blk < row_blocks; blk += gridDim.x) {
  const int row0 = (blk * kGemvWarps + warp) * kGemvRowsPerWarp;
  for (int rr = 0; rr < kGemvRowsPerWarp; rr += R) {
    const int n0 = row0 + rr;
    if (n0 >= N) break;
    const uint4* wr[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int n = (n0 + r < N) ? n0 + r : N - 1;  // clamp: tail rows re-read the last row
      wr[r] = reinterpret_cast<const uint4*>(W + static_cast<long long>(n) * K);
    }
    float acc[R][NB];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int b = 0; b < NB; ++b) acc[r][b] = 0.f;
#pragma unroll 2
    for (int k8 = lane; k8 < K / 8; k8 += 32) {
      uint4 w[R];
#pragma unroll
      for (int r = 0; r < R; ++r) w[r] = __ldg(wr[r] + k8);
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const float4 a0 = *reinterpret_cast<const float4*>(s_in + b * K + k8 * 8);
        const float4 a1 = *reinterpret_cast<const float4*>(s_in + b * K + k8 * 8 + 4);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          acc[r][b] += bf16_lo(w[r].x) * a0.x + bf16_hi(w[r].x) * a0.y + bf16_lo(w[r].y) * a0.z +
                       bf16_hi(w[r].y) * a0.w + bf16_lo(w[r].z) * a1.x + bf16_hi(w[r].z) * a1.y +
                       bf16_lo(w[r].w) * a1.z + bf16_hi(w[r].w) * a1.w;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int n = n0 + r;
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        float v = warp_sum(acc[r][b]);
        if (lane == 0 && n < N) {
          v += bias ? bias[n] : 0.f;
          float* o = out + b * out_stride + n;
          *o = accumulate ? (*o + v) : v;
        }
      }
    }
  }
  }
}

__global__ void timestep_features_kernel(const float* __restrict__ t, float* __restrict__ out,
                                         int B) {
  // emb[j] = t * exp(-ln(10000) * j / 128); out = [cos(emb) (128) | sin(emb) (128)]
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 128) return;
  const int b = i / 128, j = i - b * 128;
  const float expo = -9.210340371976184f * static_cast<float>(j) / 128.0f;
  const float arg = t[b] * expf(expo);
  out[b * 256 + j] = cosf(arg);
  out[b * 256 + 128 + j] = sinf(arg);
}

template <typename T>
__device__ __forceinline__ float ldf(const T* p, long long i);
template <>
__device__ __forceinline__ float ldf<float>(const float* p, long long i) {
  return p[i];
}
template <>
__device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p, long long i) {
  return __bfloat162float(p[i]);
}

// one thread per (b, t, gy, gx, k) output element; k fastest -> coalesced bf16 stores
template <typename T>
__global__ void patchify_kernel(const T* __restrict__ lat, __nv_bfloat16* __restrict__ out,
                                int rows_total, int row_offset, int Kpad, int B, int C, int Tn,
                                int H, int W, int pool2) {
  const int gh = (pool2 ? H / 2 : H) / 2, gw = (pool2 ? W / 2 : W) / 2;
  const long long total = static_cast<long long>(B) * Tn * gh * gw * Kpad;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int k = idx % Kpad;
  long long r = idx / Kpad;
  const int gx = r % gw;
  r /= gw;
  const int gy = r % gh;
  r /= gh;
  const int tt = r % Tn;
  const int b = r / Tn;
  float val = 0.f;
  if (k < C * 4) {
    const int c = k >> 2, p1 = (k >> 1) & 1, p2 = k & 1;
    const int y = 2 * gy + p1, x = 2 * gx + p2;
    const long long plane = ((static_cast<long long>(b) * C + c) * Tn + tt) * H * W;
    if (pool2) {
      // bilinear resize by exactly 1/2 (align_corners=False) == 2x2 mean
      const long long o = plane + static_cast<long long>(2 * y) * W + 2 * x;
      val = 0.25f * (ldf(lat, o) + ldf(lat, o + 1) + ldf(lat, o + W) + ldf(lat, o + W + 1));
    } else {
      val = ldf(lat, plane + static_cast<long long>(y) * W + x);
    }
  }
  const long long orow = static_cast<long long>(b) * rows_total + row_offset +
                         (static_cast<long long>(tt) * gh + gy) * gw + gx;
  out[orow * Kpad + k] = __float2bfloat16(val);
}

template <typename T>
__global__ void to_bf16_kernel(const T* __restrict__ in, __nv_bfloat16* __restrict__ out,
                               long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2bfloat16(ldf(in, i));
}

// base 2-D sincos table: channel layout [sin(x w) | cos(x w) | sin(y w) | cos(y w)], each D/4,
// w_d = 10000^(-d / (D/4)), coordinates = index / (S / base_size); evaluated in fp64 like numpy.
__global__ void pos_base_kernel(float* __restrict__ out, int S, int D, int base_size) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(S) * S * D;
  if (idx >= total) return;
  const int ch = idx % D;
  const long long p = idx / D;
  const int xi = p % S, yi = p / S;
  const int quarter = D / 4;
  const int sel = ch / quarter, dd = ch % quarter;
  // numpy: arange(S, float32) / (S / base) -> float32 coordinate
  const float coord_f = static_cast<float>(sel < 2 ? xi : yi) /
                        static_cast<float>(static_cast<double>(S) / base_size);
  const double omega = 1.0 / pow(10000.0, static_cast<double>(dd) / static_cast<double>(quarter));
  const double a = static_cast<double>(coord_f) * omega;
  out[idx] = static_cast<float>((sel & 1) ? cos(a) : sin(a));
}

// torch F.interpolate(mode='bilinear', align_corners=False) source index
__device__ __forceinline__ void bilinear_src(int dst, int in_size, int out_size, int& i0, int& i1,
                                             float& lam) {
  const float scale = static_cast<float>(in_size) / static_cast<float>(out_size);
  float src = (static_cast<float>(dst) + 0.5f) * scale - 0.5f;
  src = src < 0.f ? 0.f : src;
  i0 = static_cast<int>(src);
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  lam = src - static_cast<float>(i0);
}

__global__ void pos_clip_kernel(const float* __restrict__ base, int S, int D,
                                float* __restrict__ out, int row_offset, int Tn, int h, int w,
                                int oh, int ow) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(h) * w * D;
  if (idx >= total) return;
  const int ch = idx % D;
  const int p = idx / D;
  const int x = p % w, y = p / w;
  const int top = (S - oh) / 2, left = (S - ow) / 2;
  float val;
  if (oh == h && ow == w) {
    val = base[(static_cast<long long>(top + y) * S + left + x) * D + ch];
  } else {
    int y0, y1, x0, x1;
    float ly, lx;
    bilinear_src(y, oh, h, y0, y1, ly);
    bilinear_src(x, ow, w, x0, x1, lx);
    auto at = [&](int yy, int xx) {
      return base[(static_cast<long long>(top + yy) * S + left + xx) * D + ch];
    };
    // ATen upsample_bilinear2d: w0 = 1 - lambda
    val = (1.f - ly) * ((1.f - lx) * at(y0, x0) + lx * at(y0, x1)) +
          ly * ((1.f - lx) * at(y1, x0) + lx * at(y1, x1));
  }
  for (int t = 0; t < Tn; ++t)
    out[(static_cast<long long>(row_offset) + static_cast<long long>(t) * h * w + p) * D + ch] = val;
}

// one CTA of 128 threads per (sample, 128-key tile): key_bias plus a per-tile "holds a dead key"
// flag that lets the attention kernel skip all masking work on fully-live tiles
__global__ void __launch_bounds__(128) key_bias_kernel(const float* __restrict__ ctx_mask, int Lc,
                                                       float* __restrict__ key_bias,
                                                       int* __restrict__ tile_dead, int L,
                                                       int Lpad) {
  const int b = blockIdx.y, tile = blockIdx.x;
  const int k = tile * 128 + threadIdx.x;
  float v = 0.f;
  if (k >= L) {
    v = -INFINITY;
  } else if (k < Lc) {
    v = (ctx_mask[b * Lc + k] != 0.f) ? 0.f : -INFINITY;
  }
  key_bias[static_cast<long long>(b) * Lpad + k] = v;
  const int any = __syncthreads_or(v != 0.f);
  if (tile_dead != nullptr && threadIdx.x == 0) tile_dead[b * (Lpad / 128) + tile] = any;
}

// ---------------------------------------------------------------------------------
// Ulysses sequence parallelism: staging copies around the all-to-all.
// qkv [B][L][3D] / attn [B][L][D] are the joint buffers; rank r owns the video rows
// [Lc + r*Lw, Lc + (r+1)*Lw) and the heads [r*Hc/64, ...) (Hc = columns per rank).
// 16-byte vectors; one thread per vector.
// ---------------------------------------------------------------------------------
// stage[j][b][l][part][Hc] <- qkv[b][row0 + l][part*D + j*Hc + :]        (my rows, peer j's heads)
__global__ void sp_qkv_pack_kernel(const uint4* __restrict__ qkv, uint4* __restrict__ stage, int B,
                                   int L, int D, int row0, int Lw, int P, int Hc) {
  const int hv = Hc / 8, dv = D / 8;
  const long long total = static_cast<long long>(P) * B * Lw * 3 * hv;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = idx % hv;
  long long t = idx / hv;
  const int part = t % 3;
  t /= 3;
  const int l = t % Lw;
  t /= Lw;
  const int b = t % B;
  const int j = t / B;
  stage[idx] = qkv[(static_cast<long long>(b) * L + row0 + l) * 3 * dv + part * dv + j * hv + c];
}
// qkv[b][Lc + i*Lw + l][part*D + r*Hc + :] <- stage[i][b][l][part][Hc]    (peer i's rows, my heads)
__global__ void sp_qkv_unpack_kernel(const uint4* __restrict__ stage, uint4* __restrict__ qkv, int B,
                                     int L, int D, int Lc, int Lw, int P, int Hc, int r) {
  const int hv = Hc / 8, dv = D / 8;
  const long long total = static_cast<long long>(P) * B * Lw * 3 * hv;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = idx % hv;
  long long t = idx / hv;
  const int part = t % 3;
  t /= 3;
  const int l = t % Lw;
  t /= Lw;
  const int b = t % B;
  const int i = t / B;
  if (i == r) return;  // my own rows are already in place
  qkv[(static_cast<long long>(b) * L + Lc + i * Lw + l) * 3 * dv + part * dv + r * hv + c] = stage[idx];
}
// stage[j][b][row][Hc] <- attn[b][row < Lc ? row : Lc + j*Lw + (row - Lc)][r*Hc + :]
//   (all context rows and peer j's video rows, my heads)
__global__ void sp_attn_pack_kernel(const uint4* __restrict__ attn, uint4* __restrict__ stage, int B,
                                    int L, int D, int Lc, int Lw, int P, int Hc, int r) {
  const int hv = Hc / 8, dv = D / 8, rows = Lc + Lw;
  const long long total = static_cast<long long>(P) * B * rows * hv;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = idx % hv;
  long long t = idx / hv;
  const int row = t % rows;
  t /= rows;
  const int b = t % B;
  const int j = t / B;
  const int src_row = row < Lc ? row : Lc + j * Lw + (row - Lc);
  stage[idx] = attn[(static_cast<long long>(b) * L + src_row) * dv + r * hv + c];
}
// attn[b][row < Lc ? row : Lc + r*Lw + (row - Lc)][i*Hc + :] <- stage[i][b][row][Hc]
__global__ void sp_attn_unpack_kernel(const uint4* __restrict__ stage, uint4* __restrict__ attn, int B,
                                      int L, int D, int Lc, int Lw, int P, int Hc, int r) {
  const int hv = Hc / 8, dv = D / 8, rows = Lc + Lw;
  const long long total = static_cast<long long>(P) * B * rows * hv;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = idx % hv;
  long long t = idx / hv;
  const int row = t % rows;
  t /= rows;
  const int b = t % B;
  const int i = t / B;
  if (i == r) return;  // my own heads are already in place
  const int dst_row = row < Lc ? row : Lc + r * Lw + (row - Lc);
  attn[(static_cast<long long>(b) * L + dst_row) * dv + i * hv + c] = stage[idx];
}
// all-gather of the fp32 video stream: stage[j][b][l][D] <- x[b][r*Lw + l][:] (same block for every
// peer), then x[b][i*Lw + l][:] <- stage[i][b][l][D]
__global__ void sp_x_pack_kernel(const float4* __restrict__ x, float4* __restrict__ stage, int B, int Lv,
                                 int D, int Lw, int P, int r) {
  const int dv = D / 4;
  const long long total = static_cast<long long>(P) * B * Lw * dv;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = idx % dv;
  long long t = idx / dv;
  const int l = t % Lw;
  t /= Lw;
  const int b = t % B;
  stage[idx] = x[(static_cast<long long>(b) * Lv + r * Lw + l) * dv + c];
}
__global__ void sp_x_unpack_kernel(const float4* __restrict__ stage, float4* __restrict__ x, int B, int Lv,
                                   int D, int Lw, int P, int r) {
  const int dv = D / 4;
  const long long total = static_cast<long long>(P) * B * Lw * dv;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = idx % dv;
  long long t = idx / dv;
  const int l = t % Lw;
  t /= Lw;
  const int b = t % B;
  const int i = t / B;
  if (i == r) return;
  x[(static_cast<long long>(b) * Lv + i * Lw + l) * dv + c] = stage[idx];
}

// ---- conditioning cache (adaLN modulation table memoised per (timestep, pooled embedding)) -----------------
// One CTA walks the B rows in order.  Row b's key = (t[b], pooled[b][:]) bit for bit; its slot is looked for in four
// probes from hash % slots.  hit: the slot holds exactly this key and was NOT installed during this call (its table would not be
// there yet: rows 1 and 2 of a 3-branch CFG batch carry the same key).  miss: the key is installed; `store[b]`
// says whether row b is the last row of this call that maps to its slot (it then owns the slot's table).
constexpr int kCondProbes = 4;
__global__ void __launch_bounds__(256) cond_lookup_kernel(const float* __restrict__ t, const float* __restrict__ pooled,
                                                          int pooled_dim, int B, float* __restrict__ keys,
                                                          int* __restrict__ valid, int slots, int* __restrict__ slot_of,
                                                          int* __restrict__ need, int* __restrict__ store) {
  __shared__ unsigned long long red[8];
  __shared__ unsigned long long s_hash;
  __shared__ int s_slot[4], s_need[4], s_flag;
  const int klen = pooled_dim + 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int b = 0; b < B; ++b) {
    unsigned long long h = 0;
    for (int i = threadIdx.x; i < klen; i += blockDim.x) {
      const unsigned v = __float_as_uint(i == 0 ? t[b] : pooled[static_cast<long long>(b) * pooled_dim + i - 1]);
      h += (static_cast<unsigned long long>(v) + 0x9E3779B97F4A7C15ull) * (2ull * i + 0x100000001B3ull);
    }
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    if (lane == 0) red[warp] = h;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long tot = 0;
      for (int w = 0; w < 8; ++w) tot += red[w];
      tot ^= tot >> 29;
      s_hash = tot;
    }
    __syncthreads();
    // four probes (linear from hash % slots): the first slot that holds this key, else the first free one, else a
    // victim picked by other bits of the hash
    const int base = static_cast<int>(s_hash % static_cast<unsigned long long>(slots));
    int found = -1, free_slot = -1;
    for (int pr = 0; pr < kCondProbes && found < 0; ++pr) {
      const int sl = (base + pr) % slots;
      if (threadIdx.x == 0) s_flag = 1;
      __syncthreads();
      int same = valid[sl] != 0;
      if (!same && free_slot < 0) free_slot = sl;
      if (same) {
        const float* key = keys + static_cast<long long>(sl) * klen;
        for (int i = threadIdx.x; i < klen; i += blockDim.x) {
          const unsigned v = __float_as_uint(i == 0 ? t[b] : pooled[static_cast<long long>(b) * pooled_dim + i - 1]);
          if (v != __float_as_uint(key[i])) same = 0;
        }
      }
      if (!same) s_flag = 0;   // benign race: every writer writes 0
      __syncthreads();
      if (s_flag != 0) found = sl;
      __syncthreads();
    }
    const int sl = found >= 0 ? found
                              : (free_slot >= 0 ? free_slot
                                                : (base + static_cast<int>((s_hash >> 40) % kCondProbes)) % slots);
    bool installed_now = false;
    for (int j = 0; j < b; ++j) installed_now |= (s_slot[j] == sl && s_need[j]);
    const bool hit = found >= 0 && !installed_now;
    if (found < 0) {   // install this key (the table follows after the GEMVs)
      float* key = keys + static_cast<long long>(sl) * klen;
      for (int i = threadIdx.x; i < klen; i += blockDim.x)
        key[i] = i == 0 ? t[b] : pooled[static_cast<long long>(b) * pooled_dim + i - 1];
    }
    __syncthreads();
    if (threadIdx.x == 0) s_slot[b] = sl;
    if (threadIdx.x == 0) {
      s_need[b] = hit ? 0 : 1;
      if (!hit) valid[sl] = 1;
      __threadfence();
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    int st[4];
    for (int b = 0; b < B; ++b) {
      int last = 1;
      for (int j = b + 1; j < B; ++j) last &= !(s_slot[j] == s_slot[b] && s_need[j]);
      st[b] = s_need[b] && last;
    }
    // a found row whose slot is re-keyed by another row of this call must not read the slot while it is rewritten
    for (int b = 0; b < B; ++b) {
      if (s_need[b]) continue;
      for (int j = 0; j < B; ++j)
        if (j != b && s_slot[j] == s_slot[b] && st[j]) s_need[b] = 1;
    }
    for (int b = 0; b < B; ++b) {
      slot_of[b] = s_slot[b];
      need[b] = s_need[b];
      store[b] = st[b];
    }
  }
}
// hit rows: cache -> plan table; owning miss rows: plan table -> cache
__global__ void cond_finish_kernel(float4* __restrict__ mod, long long row4, float4* __restrict__ cache,
                                   const int* __restrict__ slot_of, const int* __restrict__ need,
                                   const int* __restrict__ store) {
  const int b = blockIdx.y;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= row4) return;
  float4* m = mod + b * row4 + i;
  float4* c = cache + slot_of[b] * row4 + i;
  if (!need[b])
    *m = *c;
  else if (store[b])
    *c = *m;
}

// ---- sequence parallelism over peer memory: barrier and all-gather without NCCL ---------------------------
struct SpPeers {
  void* p[8];
};
// One warp.  Lane j announces this rank's arrival (epoch e) in rank j's flag array and then waits until rank j's
// arrival shows up in this rank's own array.  Epochs only grow, so a peer that is already one barrier ahead is
// still seen as arrived.  The previous kernels of this stream have completed (the barrier is launched without the
// programmatic-serialization attribute), i.e. their stores into the peers' buffers are performed before the
// release store of the flag.  A peer that never arrives traps after 30 s instead of hanging the box.
__global__ void sp_barrier_kernel(SpPeers flags, int* __restrict__ epoch, int rank, int P) {
  __shared__ int e_sh;
  if (threadIdx.x == 0) {
    e_sh = *epoch + 1;
    *epoch = e_sh;
  }
  __syncthreads();
  const int e = e_sh;
  const int j = threadIdx.x;
  if (j < P) {
    __threadfence_system();
    int* dst = reinterpret_cast<int*>(flags.p[j]) + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(e) : "memory");
    const int* src = reinterpret_cast<const int*>(flags.p[rank]) + j;
    const uint64_t t0 = global_ns();
    while (true) {
      int v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
      if (v - e >= 0) break;
      if (global_ns() - t0 > 30000000000ull) asm volatile("trap;");
    }
  }
}
// rows [row0, row0 + rows) of my fp32 stream -> the same rows of every peer's stream (plain stores over NVLink)
__global__ void sp_x_share_kernel(SpPeers xs, int B, int Lv, int D, int row0, int rows, int rank, int P) {
  const int dv = D / 4;
  const long long per_peer = static_cast<long long>(B) * rows * dv;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= per_peer * (P - 1)) return;
  int j = static_cast<int>(idx / per_peer);
  if (j >= rank) ++j;
  long long t = idx % per_peer;
  const int c = t % dv;
  t /= dv;
  const int l = t % rows;
  const int b = static_cast<int>(t / rows);
  const long long off = (static_cast<long long>(b) * Lv + row0 + l) * dv + c;
  reinterpret_cast<float4*>(xs.p[j])[off] = reinterpret_cast<const float4*>(xs.p[rank])[off];
}

}  // namespace

int launch_cond_lookup(const float* t, const float* pooled, int pooled_dim, int B, float* keys, int* valid, int slots,
                       int* slot_of, int* need, int* store, cudaStream_t stream) {
  DV_REQUIRE(B >= 1 && B <= 4 && slots > 0, "cond_lookup: B=%d slots=%d", B, slots);
  cond_lookup_kernel<<<1, 256, 0, stream>>>(t, pooled, pooled_dim, B, keys, valid, slots, slot_of, need, store);
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

int launch_cond_finish(float* mod, long long row_floats, float* cache, const int* slot_of, const int* need,
                       const int* store, int B, cudaStream_t stream) {
  DV_REQUIRE(row_floats % 4 == 0, "cond_finish: row length %lld", row_floats);
  const long long row4 = row_floats / 4;
  ProfScope ps(PROF_OTHER, 0.0, static_cast<double>(B) * row_floats * 8.0, stream, "cond_cache_copy");
  cond_finish_kernel<<<dim3(static_cast<unsigned>((row4 + 255) / 256), B), 256, 0, stream>>>(
      reinterpret_cast<float4*>(mod), row4, reinterpret_cast<float4*>(cache), slot_of, need, store);
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

int launch_sp_barrier(void* const* flag_peers, int* epoch, int rank, int P, cudaStream_t stream) {
  DV_REQUIRE(flag_peers && epoch && P >= 2 && P <= 8 && rank >= 0 && rank < P, "sp_barrier: bad argument");
  SpPeers f;
  for (int i = 0; i < 8; ++i) f.p[i] = i < P ? flag_peers[i] : nullptr;
  sp_barrier_kernel<<<1, 32, 0, stream>>>(f, epoch, rank, P);
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

int launch_sp_x_share(void* const* x_peers, int B, int Lv, int D, int row0, int rows, int rank, int P,
                      cudaStream_t stream) {
  DV_REQUIRE(x_peers && P >= 2 && P <= 8 && rank >= 0 && rank < P && D % 4 == 0, "sp_x_share: bad argument");
  if (rows <= 0) return 0;
  SpPeers x;
  for (int i = 0; i < 8; ++i) x.p[i] = i < P ? x_peers[i] : nullptr;
  const long long n = static_cast<long long>(P - 1) * B * rows * (D / 4);
  sp_x_share_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(x, B, Lv, D, row0, rows, rank, P);
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

int launch_ln_modulate2(const LnRows& r0, const LnRows* r1, int mod_bs, int B, int D, float eps,
                        cudaStream_t stream) {
  DV_REQUIRE(D == 1536 || D == 512, "ln_modulate: D=%d unsupported", D);
  LnArgs a;
  auto fill = [](LnSeg& s, const LnRows& r) {
    s.x = r.x;
    s.x_bs = r.x_bs;
    s.out = r.out;
    s.out_bs = r.out_bs;
    s.shift = r.shift;
    s.scale = r.scale;
    s.L = r.L;
  };
  fill(a.seg[0], r0);
  if (r1 != nullptr) {
    fill(a.seg[1], *r1);
  } else {
    a.seg[1] = a.seg[0];
    a.seg[1].L = 0;
  }
  a.mod_bs = mod_bs;
  a.B = B;
  a.eps = eps;
  a.pdl_early = pdl_early() ? 1 : 0;
  const long long rows = static_cast<long long>(B) * (a.seg[0].L + a.seg[1].L);
  if (rows == 0) return 0;
  const int blocks = static_cast<int>((rows * 32 + 255) / 256);
  ProfScope ps(PROF_OTHER, 0.0, static_cast<double>(rows) * D * 6.0, stream, "ln_modulate");
  if (D == 1536)
    DV_CHECK_CUDA(launch_pdl(ln_modulate_kernel<1536>, dim3(blocks), dim3(256), 0, stream, 1, a));
  else
    DV_CHECK_CUDA(launch_pdl(ln_modulate_kernel<512>, dim3(blocks), dim3(256), 0, stream, 1, a));
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

int launch_ln_modulate(const float* x, long long x_bs, __nv_bfloat16* out, long long out_bs,
                       const float* shift, const float* scale, int mod_bs, int B, int L, int D,
                       float eps, cudaStream_t stream) {
  LnRows r = {x, x_bs, out, out_bs, shift, scale, L};
  return launch_ln_modulate2(r, nullptr, mod_bs, B, D, eps, stream);
}

int launch_gemv(const __nv_bfloat16* W, const float* bias, const float* in, int in_stride,
                float* out, int out_stride, int B, int N, int K, int silu_in, int accumulate,
                cudaStream_t stream, const int* need) {
  DV_REQUIRE(B >= 1 && B <= 4, "gemv: batch %d not in [1,4]", B);
  DV_REQUIRE(K % 8 == 0, "gemv: K=%d must be a multiple of 8", K);
  const int rows_per_cta = kGemvWarps * kGemvRowsPerWarp;
  const int row_blocks = (N + rows_per_cta - 1) / rows_per_cta;
  const size_t smem = static_cast<size_t>(B) * K * sizeof(float);
#define DV_GEMV(NB)                                                                              \
  do {                                                                                           \
    static int occ = 0;                                                                          \
    if (!occ) {                                                                                  \
      DV_CHECK_CUDA(cudaFuncSetAttribute(gemv_kernel<NB>,                                        \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); \
      DV_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gemv_kernel<NB>,         \
                                                                  kGemvWarps * 32, smem));       \
      if (occ < 1) occ = 1;                                                                      \
    }                                                                                            \
    const int resident = sm_count() * occ; /* persistent: one wave of resident CTAs */           \
    const int blocks = row_blocks < resident ? row_blocks : resident;                            \
    gemv_kernel<NB><<<blocks, kGemvWarps * 32, smem, stream>>>(W, bias, in, in_stride, out,      \
                                                               out_stride, N, K, silu_in,        \
                                                               accumulate, row_blocks, need);    \
  } while (0)
  DV_REQUIRE(smem <= 96 * 1024, "gemv: B*K too large for smem staging");
  ProfScope ps(PROF_OTHER, 0.0, static_cast<double>(N) * K * 2.0, stream, N > 100000 ? "gemv_adaln" : "gemv");
  switch (B) {
    case 1: DV_GEMV(1); break;
    case 2: DV_GEMV(2); break;
    case 3: DV_GEMV(3); break;
    default: DV_GEMV(4); break;
  }
#undef DV_GEMV
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

int launch_timestep_features(const float* t, float* out, int B, cudaStream_t stream) {
  timestep_features_kernel<<<(B * 128 + 127) / 128, 128, 0, stream>>>(t, out, B);
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

int launch_patchify(const void* latent, int is_bf16, __nv_bfloat16* out, int rows_total,
                    int row_offset, int Kpad, int B, int C, int T, int H, int W, int pool2,
                    cudaStream_t stream) {
  DV_REQUIRE(Kpad >= C * 4, "patchify: Kpad=%d < C*4=%d", Kpad, C * 4);
  DV_REQUIRE(H % (pool2 ? 4 : 2) == 0 && W % (pool2 ? 4 : 2) == 0, "patchify: H=%d W=%d", H, W);
  const int gh = (pool2 ? H / 2 : H) / 2, gw = (pool2 ? W / 2 : W) / 2;
  const long long total = static_cast<long long>(B) * T * gh * gw * Kpad;
  const int blocks = static_cast<int>((total + 255) / 256);
  if (is_bf16)
    patchify_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(latent), out, rows_total, row_offset, Kpad, B, C, T,
        H, W, pool2);
  else
    patchify_kernel<float><<<blocks, 256, 0, stream>>>(reinterpret_cast<const float*>(latent), out,
                                                       rows_total, row_offset, Kpad, B, C, T, H, W,
                                                       pool2);
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

int launch_to_bf16(const void* in, int is_bf16, __nv_bfloat16* out, long long n,
                   cudaStream_t stream) {
  const int blocks = static_cast<int>((n + 255) / 256);
  if (is_bf16)
    to_bf16_kernel<__nv_bfloat16>
        <<<blocks, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(in), out, n);
  else
    to_bf16_kernel<float><<<blocks, 256, 0, stream>>>(reinterpret_cast<const float*>(in), out, n);
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

int launch_pos_base(float* out, int S, int D, int base_size, cudaStream_t stream) {
  const long long total = static_cast<long long>(S) * S * D;
  pos_base_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, stream>>>(out, S, D, base_size);
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

int launch_pos_clip(const float* base, int S, int D, float* out, int row_offset, int T, int h,
                    int w, int oh, int ow, cudaStream_t stream) {
  DV_REQUIRE(oh >= h && ow >= w && oh <= S && ow <= S, "pos_clip: crop %dx%d vs clip %dx%d", oh, ow,
             h, w);  // mmdit.py:851-852
  const long long total = static_cast<long long>(h) * w * D;
  pos_clip_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, stream>>>(base, S, D, out,
                                                                           row_offset, T, h, w, oh,
                                                                           ow);
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

int launch_key_bias(const float* ctx_mask, int Lc, float* key_bias, int* tile_dead, int B, int L,
                    int Lpad, cudaStream_t stream) {
  DV_REQUIRE(Lpad % 128 == 0, "key_bias: Lpad=%d must be a multiple of 128", Lpad);
  key_bias_kernel<<<dim3(Lpad / 128, B), 128, 0, stream>>>(ctx_mask, Lc, key_bias, tile_dead, L, Lpad);
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

// ---- Ulysses staging copies ------------------------------------------------------------
#define DV_SP_LAUNCH(kernel, total, ...)                                                  \
  do {                                                                                    \
    const long long _n = (total);                                                         \
    if (_n > 0) {                                                                         \
      kernel<<<static_cast<unsigned>((_n + 255) / 256), 256, 0, stream>>>(__VA_ARGS__);   \
      DV_CHECK_CUDA(cudaGetLastError());                                                  \
      note_launch();                                                                      \
    }                                                                                     \
  } while (0)

int launch_sp_qkv_pack(const void* qkv, void* stage, int B, int L, int D, int row0, int Lw, int P,
                       int Hc, cudaStream_t stream) {
  DV_SP_LAUNCH(sp_qkv_pack_kernel, static_cast<long long>(P) * B * Lw * 3 * (Hc / 8),
               reinterpret_cast<const uint4*>(qkv), reinterpret_cast<uint4*>(stage), B, L, D, row0, Lw, P, Hc);
  return 0;
}
int launch_sp_qkv_unpack(const void* stage, void* qkv, int B, int L, int D, int Lc, int Lw, int P,
                         int Hc, int r, cudaStream_t stream) {
  DV_SP_LAUNCH(sp_qkv_unpack_kernel, static_cast<long long>(P) * B * Lw * 3 * (Hc / 8),
               reinterpret_cast<const uint4*>(stage), reinterpret_cast<uint4*>(qkv), B, L, D, Lc, Lw, P, Hc, r);
  return 0;
}
int launch_sp_attn_pack(const void* attn, void* stage, int B, int L, int D, int Lc, int Lw, int P,
                        int Hc, int r, cudaStream_t stream) {
  DV_SP_LAUNCH(sp_attn_pack_kernel, static_cast<long long>(P) * B * (Lc + Lw) * (Hc / 8),
               reinterpret_cast<const uint4*>(attn), reinterpret_cast<uint4*>(stage), B, L, D, Lc, Lw, P, Hc, r);
  return 0;
}
int launch_sp_attn_unpack(const void* stage, void* attn, int B, int L, int D, int Lc, int Lw, int P,
                          int Hc, int r, cudaStream_t stream) {
  DV_SP_LAUNCH(sp_attn_unpack_kernel, static_cast<long long>(P) * B * (Lc + Lw) * (Hc / 8),
               reinterpret_cast<const uint4*>(stage), reinterpret_cast<uint4*>(attn), B, L, D, Lc, Lw, P, Hc, r);
  return 0;
}
int launch_sp_x_pack(const float* x, void* stage, int B, int Lv, int D, int Lw, int P, int r,
                     cudaStream_t stream) {
  DV_SP_LAUNCH(sp_x_pack_kernel, static_cast<long long>(P) * B * Lw * (D / 4),
               reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(stage), B, Lv, D, Lw, P, r);
  return 0;
}
int launch_sp_x_unpack(const void* stage, float* x, int B, int Lv, int D, int Lw, int P, int r,
                       cudaStream_t stream) {
  DV_SP_LAUNCH(sp_x_unpack_kernel, static_cast<long long>(P) * B * Lw * (D / 4),
               reinterpret_cast<const float4*>(stage), reinterpret_cast<float4*>(x), B, Lv, D, Lw, P, r);
  return 0;
}
#undef DV_SP_LAUNCH

}  // namespace dv

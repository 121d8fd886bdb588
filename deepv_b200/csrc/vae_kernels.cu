// deepv_b200 — HBM-bound kernels of the VAE decoder: per-frame GroupNorm (+SiLU), row
// softmax for the mid-block attention, layout conversion and the tile blend.
// Activations are channels-last bf16: [frames][H*W][C].
#include "kernels.cuh"

namespace dv {
namespace {

// ---------------------------------------------------------------------------------
// GroupNorm statistics (reference CausalGroupNorm, vae.py:161-167: frames folded into the
// batch, 32 groups, eps 1e-6).  Pass 1: per-CTA partial (sum, sumsq) per group in fp32,
// merged with fp64 atomics; pass 2 (gn_finalize) turns them into (mean, rstd).
// ---------------------------------------------------------------------------------
constexpr int kGnThreads = 256;

// Thread t of a CTA owns the 8-channel vector (t % (C/8)) of every pixel it visits, so its group
// (two groups when a group has 4 channels) never changes; pixels are strided by 256 / (C/8) and
// four loads are kept in flight per thread.
__global__ void __launch_bounds__(kGnThreads) gn_partial_kernel(const __nv_bfloat16* __restrict__ x,
                                                                double* __restrict__ acc, int HW,
                                                                int C, int G, int pix_per_cta) {
  // grid: (ceil(HW / pix_per_cta), frames)
  const int f = blockIdx.y;
  const int vec_per_pix = C / 8;
  const int cpg = C / G;  // channels per group: 4, 8 or 16
  const int p0 = blockIdx.x * pix_per_cta;
  const int p1 = min(p0 + pix_per_cta, HW);
  const uint4* xf = reinterpret_cast<const uint4*>(x + static_cast<long long>(f) * HW * C);
  __shared__ float4 s_part[kGnThreads];  // (sum, sumsq) of channels 0-3 and 4-7 of every thread
  const int v = threadIdx.x % vec_per_pix;           // requires blockDim % vec_per_pix == 0
  const int pstep = kGnThreads / vec_per_pix;
  float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;      // cpg == 4 -> two groups per 8-vector
  auto add = [&](const uint4& t) {
    const float e[8] = {bf16_lo(t.x), bf16_hi(t.x), bf16_lo(t.y), bf16_hi(t.y),
                        bf16_lo(t.z), bf16_hi(t.z), bf16_lo(t.w), bf16_hi(t.w)};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      s0 += e[i];
      q0 += e[i] * e[i];
    }
#pragma unroll
    for (int i = 4; i < 8; ++i) {
      s1 += e[i];
      q1 += e[i] * e[i];
    }
  };
  int p = p0 + threadIdx.x / vec_per_pix;
  for (; p + 3 * pstep < p1; p += 4 * pstep) {
    uint4 t[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) t[u] = __ldg(xf + static_cast<long long>(p + u * pstep) * vec_per_pix + v);
#pragma unroll
    for (int u = 0; u < 4; ++u) add(t[u]);
  }
  for (; p < p1; p += pstep) add(__ldg(xf + static_cast<long long>(p) * vec_per_pix + v));
  // fixed-order reduction inside the CTA (bit-reproducible), fp64 atomics between CTAs
  s_part[threadIdx.x] = make_float4(s0, q0, s1, q1);
  __syncthreads();
  if (threadIdx.x < G) {
    const int g = threadIdx.x;
    float sum = 0.f, sq = 0.f;
    if (cpg == 4) {
      const int vv = g >> 1;
      for (int k = 0; k < pstep; ++k) {
        const float4 t = s_part[vv + k * vec_per_pix];
        sum += (g & 1) ? t.z : t.x;
        sq += (g & 1) ? t.w : t.y;
      }
    } else {
      const int v0 = g * (cpg / 8), v1 = v0 + cpg / 8;
      for (int vv = v0; vv < v1; ++vv)
        for (int k = 0; k < pstep; ++k) {
          const float4 t = s_part[vv + k * vec_per_pix];
          sum += t.x + t.z;
          sq += t.y + t.w;
        }
    }
    atomicAdd(&acc[(static_cast<long long>(f) * G + g) * 2 + 0], static_cast<double>(sum));
    atomicAdd(&acc[(static_cast<long long>(f) * G + g) * 2 + 1], static_cast<double>(sq));
  }
}

__device__ __forceinline__ float silu_fast(float x) {
  return __fdividef(x, 1.0f + __expf(-x));
}
// x * sigmoid(x) = 0.5 x (1 + tanh(0.5 x)): ONE trip through the XU pipe (MUFU.TANH) instead of two (EX2 + RCP)
__device__ __forceinline__ float silu_tanh(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

// y = silu?( (x - mean) * rstd * gamma + beta ): same thread -> channel-vector mapping as the
// statistics pass, so the per-channel affine a*x + b is hoisted out of the pixel loop.
__global__ void __launch_bounds__(kGnThreads) gn_apply_kernel(const __nv_bfloat16* __restrict__ x,
                                                              const double* __restrict__ acc,
                                                              int replicas, long long replica_stride,
                                                              double inv_count, float eps,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta,
                                                              __nv_bfloat16* __restrict__ y, int HW,
                                                              int C, int G, int do_silu,
                                                              int pix_per_cta) {
  // do_silu: 0 none, 1 = x / (1 + e^-x) (EX2 + RCP), 2 = 0.5 x (1 + tanh(0.5 x)) (one MUFU)
  const int f = blockIdx.y;
  const int vec_per_pix = C / 8;
  const int cpg = C / G;
  const int p0 = blockIdx.x * pix_per_cta;
  const int p1 = min(p0 + pix_per_cta, HW);
  const int v = threadIdx.x % vec_per_pix;
  const int pstep = kGnThreads / vec_per_pix;
  const int c0 = v * 8;
  // mean / rstd of this frame's groups once per CTA: (sum, sum of squares) over the accumulator
  // replicas -> mean, 1/sqrt(var + eps) in fp64 (biased variance)
  __shared__ float2 s_ms[64];
  pdl_wait();
  if (threadIdx.x < G) {
    const int g = threadIdx.x;
    double sx = 0.0, sy = 0.0;
    for (int rp = 0; rp < replicas; ++rp) {
      const double2 sq = __ldg(reinterpret_cast<const double2*>(acc + rp * replica_stride) +
                               static_cast<long long>(f) * G + g);
      sx += sq.x;
      sy += sq.y;
    }
    const double mean = sx * inv_count;
    double var = sy * inv_count - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    s_ms[g] = make_float2(static_cast<float>(mean), static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps))));
  }
  __syncthreads();
  float a[8], b[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float2 ms = s_ms[(c0 + k) / cpg];
    const float gm = __ldg(gamma + c0 + k);
    a[k] = ms.y * gm;
    b[k] = __ldg(beta + c0 + k) - ms.x * a[k];
  }
  const long long base = static_cast<long long>(f) * HW * vec_per_pix;
  const uint4* xf = reinterpret_cast<const uint4*>(x) + base;
  uint4* yf = reinterpret_cast<uint4*>(y) + base;
  auto apply = [&](const uint4& t) {
    float e[8] = {bf16_lo(t.x), bf16_hi(t.x), bf16_lo(t.y), bf16_hi(t.y),
                  bf16_lo(t.z), bf16_hi(t.z), bf16_lo(t.w), bf16_hi(t.w)};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float r = fmaf(e[k], a[k], b[k]);
      e[k] = do_silu == 2 ? silu_tanh(r) : (do_silu ? silu_fast(r) : r);
    }
    uint4 o;
    o.x = pack_bf16x2(e[0], e[1]);
    o.y = pack_bf16x2(e[2], e[3]);
    o.z = pack_bf16x2(e[4], e[5]);
    o.w = pack_bf16x2(e[6], e[7]);
    return o;
  };
  int p = p0 + threadIdx.x / vec_per_pix;
  for (; p + 3 * pstep < p1; p += 4 * pstep) {
    uint4 t[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) t[u] = __ldg(xf + static_cast<long long>(p + u * pstep) * vec_per_pix + v);
#pragma unroll
    for (int u = 0; u < 4; ++u) yf[static_cast<long long>(p + u * pstep) * vec_per_pix + v] = apply(t[u]);
  }
  for (; p < p1; p += pstep)
    yf[static_cast<long long>(p) * vec_per_pix + v] = apply(__ldg(xf + static_cast<long long>(p) * vec_per_pix + v));
}

// softmax over rows of fp32 scores (single 512-wide head of the mid block: diffusers
// Attention at vae.py:439-445, scale = 1/sqrt(C)); one warp per row, in place allowed.
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s,
                                                           __nv_bfloat16* __restrict__ p,
                                                           long long rows, int cols, float scale) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* sr = s + row * cols;
  __nv_bfloat16* pr = p + row * cols;
  float m = -INFINITY;
  for (int c = lane; c < cols; c += 32) m = fmaxf(m, sr[c] * scale);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float sum = 0.f;
  for (int c = lane; c < cols; c += 32) sum += __expf(sr[c] * scale - m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.0f / sum;
  for (int c = lane; c < cols; c += 32)
    pr[c] = __float2bfloat16(__expf(sr[c] * scale - m) * inv);
}

// [frames][rows][cols] -> [frames][cols][rows]  (V^T for the P V product)
__global__ void transpose_kernel(const __nv_bfloat16* __restrict__ in, int ld_in,
                                 __nv_bfloat16* __restrict__ out, int rows, int cols) {
  __shared__ __nv_bfloat16 tile[32][33];
  const int f = blockIdx.z;
  const __nv_bfloat16* src = in + static_cast<long long>(f) * rows * ld_in;
  __nv_bfloat16* dst = out + static_cast<long long>(f) * rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = src[static_cast<long long>(r) * ld_in + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[static_cast<long long>(c) * rows + r] = tile[threadIdx.x][i];
  }
}

template <typename T>
__device__ __forceinline__ float ldv(const T* p, long long i) {
  if constexpr (sizeof(T) == 2)
    return __bfloat162float(p[i]);
  else
    return p[i];
}

// latent tile: z [C][T][h][w] (batch 1) cropped at (y0, x0) -> [T][th][tw][Cpad] bf16
template <typename T>
__global__ void latent_tile_kernel(const T* __restrict__ z, __nv_bfloat16* __restrict__ out, int C,
                                   int Tn, int h, int w, int y0, int x0, int th, int tw, int Cpad) {
  const long long total = static_cast<long long>(Tn) * th * tw * Cpad;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = i % Cpad;
  long long r = i / Cpad;
  const int x = r % tw;
  r /= tw;
  const int y = r % th;
  const int t = r / th;
  float v = 0.f;
  if (c < C) v = ldv(z, ((static_cast<long long>(c) * Tn + t) * h + (y0 + y)) * w + (x0 + x));
  out[i] = __float2bfloat16(v);
}

// Tap gather of a causal 3x3x3 conv with <= 4 output channels (conv_out: 128 -> 3).  The GEMM
// wrote per-tap partial products P[tap][pixel][4] (EPI_TAPS); out[t][h][w][c] = bias[c] +
// sum_tap P[tap][(t+dt-2, h+dh-1, w+dw-1)][c], zero outside the frame / before the first frame
// (vae.py:225-252).  One thread per pixel: every tap plane is read with coalesced 16-byte loads.
__global__ void __launch_bounds__(256) tap_gather_kernel(const float4* __restrict__ P,
                                                         const float* __restrict__ bias,
                                                         __nv_bfloat16* __restrict__ out, int T, int H,
                                                         int W, int Cout, int t_first) {
  const long long npix = static_cast<long long>(T) * H * W;
  // frames [t_first, T) only; plane and output indices stay absolute
  const long long i = static_cast<long long>(t_first) * H * W + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const int w = static_cast<int>(i % W);
  const int h = static_cast<int>((i / W) % H);
  const int t = static_cast<int>(i / (static_cast<long long>(W) * H));
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
  for (int dt = 0; dt < 3; ++dt) {
    const int tt = t + dt - 2;
    if (tt < 0) continue;
#pragma unroll
    for (int dh = 0; dh < 3; ++dh) {
      const int hh = h + dh - 1;
      if (hh < 0 || hh >= H) continue;
#pragma unroll
      for (int dw = 0; dw < 3; ++dw) {
        const int ww = w + dw - 1;
        if (ww < 0 || ww >= W) continue;
        const int tap = (dt * 3 + dh) * 3 + dw;
        const float4 v = __ldg(P + static_cast<long long>(tap) * npix +
                               (static_cast<long long>(tt) * H + hh) * W + ww);
        a0 += v.x;
        a1 += v.y;
        a2 += v.z;
        a3 += v.w;
      }
    }
  }
  const float acc[4] = {a0, a1, a2, a3};
  __nv_bfloat16* o = out + i * Cout;
  for (int c = 0; c < Cout; ++c) o[c] = __float2bfloat16(acc[c] + __ldg(bias + c));
}

}  // namespace

int launch_tap_gather(const float* planes, const float* bias, __nv_bfloat16* out, int T, int H, int W,
                      int Cout, int t_first, cudaStream_t stream) {
  DV_REQUIRE(Cout >= 1 && Cout <= 4 && t_first >= 0 && t_first < T, "tap_gather: Cout=%d, first frame %d of %d", Cout,
             t_first, T);
  const long long npix = static_cast<long long>(T - t_first) * H * W;
  ProfScope ps(PROF_OTHER, 0.0, static_cast<double>(npix) * (27 * 16.0 + Cout * 2.0), stream, "tap_gather");
  tap_gather_kernel<<<static_cast<unsigned>((npix + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(planes), bias, out, T, H, W, Cout, t_first);
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

// pixels per CTA: every thread should see >= 8 vectors (two unrolled rounds) when the tensor is
// large, while small tensors still spread over >= 2 CTAs per SM
static int gn_pix_per_cta(int HW, int frames, int C) {
  const int pstep = kGnThreads / (C / 8);
  static const int mult = getenv("DV_GN_PPC") ? atoi(getenv("DV_GN_PPC")) : 16;
  int ppc = pstep * mult;
  while (ppc > pstep * 4 && static_cast<long long>((HW + ppc - 1) / ppc) * frames < 2 * sm_count()) ppc /= 2;
  return ppc;
}

int launch_gn_stats(const __nv_bfloat16* x, double* acc, int frames, int HW, int C, int G,
                    cudaStream_t stream) {
  DV_REQUIRE(C % 8 == 0 && G <= 64 && C % G == 0, "gn_stats: C=%d G=%d", C, G);
  const int cpg = C / G;
  DV_REQUIRE(cpg == 4 || cpg % 8 == 0, "gn_stats: %d channels per group unsupported", cpg);
  DV_REQUIRE(kGnThreads % (C / 8) == 0, "gn_stats: C=%d does not divide the CTA", C);
  const int ppc = gn_pix_per_cta(HW, frames, C);
  dim3 grid((HW + ppc - 1) / ppc, frames);
  ProfScope ps(PROF_OTHER, 0.0, static_cast<double>(frames) * HW * C * 2.0, stream, "gn_stats");
  gn_partial_kernel<<<grid, kGnThreads, 0, stream>>>(x, acc, HW, C, G, ppc);
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

int launch_gn_apply(const __nv_bfloat16* x, const double* acc, int replicas, long long replica_stride,
                    const float* gamma, const float* beta, __nv_bfloat16* y, int frames, int HW, int C,
                    int G, float eps, int silu_on, cudaStream_t stream) {
  DV_REQUIRE(C % 8 == 0 && C % G == 0 && kGnThreads % (C / 8) == 0, "gn_apply: C=%d G=%d", C, G);
  const int ppc = gn_pix_per_cta(HW, frames, C);
  dim3 grid((HW + ppc - 1) / ppc, frames);
  static const bool silu_exp = getenv("DV_GN_SILU_EXP") != nullptr;   // A/B switch: the two-MUFU form
  if (silu_on && !silu_exp) silu_on = 2;
  ProfScope ps(PROF_OTHER, 0.0, static_cast<double>(frames) * HW * C * 4.0, stream, "gn_apply");
  DV_CHECK_CUDA(launch_pdl(gn_apply_kernel, grid, dim3(kGnThreads), 0, stream, 1, x, acc, replicas,
                           replica_stride, 1.0 / (static_cast<double>(HW) * (C / G)), eps, gamma, beta, y,
                           HW, C, G, silu_on, ppc));
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

int launch_softmax_rows(const float* s, __nv_bfloat16* p, long long rows, int cols, float scale,
                        cudaStream_t stream) {
  const long long threads = rows * 32;
  softmax_rows_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, stream>>>(s, p, rows,
                                                                                       cols, scale);
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

int launch_transpose(const __nv_bfloat16* in, int ld_in, __nv_bfloat16* out, int frames, int rows,
                     int cols, cudaStream_t stream) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, frames);
  transpose_kernel<<<grid, dim3(32, 8), 0, stream>>>(in, ld_in, out, rows, cols);
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

int launch_latent_tile(const void* z, int is_bf16, __nv_bfloat16* out, int C, int T, int h, int w,
                       int y0, int x0, int th, int tw, int Cpad, cudaStream_t stream) {
  const long long total = static_cast<long long>(T) * th * tw * Cpad;
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  if (is_bf16)
    latent_tile_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(z), out, C, T, h, w, y0, x0, th, tw, Cpad);
  else
    latent_tile_kernel<float><<<blocks, 256, 0, stream>>>(reinterpret_cast<const float*>(z), out, C,
                                                          T, h, w, y0, x0, th, tw, Cpad);
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

}  // namespace dv

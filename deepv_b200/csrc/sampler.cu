// deepv_b200 — per-step sampler arithmetic: CFG combine + flow-matching Euler update, the
// pyramid stage transition, and the correlated 2x2 block noise.
//
// The bf16 variants reproduce the reference's ATen rounding sequence bit for bit (every
// ATen elementwise op on bf16 tensors computes in fp32 and rounds its result to bf16):
//   pipeline.py:504-513   u + s*(t - u) [+ h*(th - t)]
//   scheduler.py:278-286  bf16( fp32(x) + fp32( bf16( bf16(dsigma) * v ) ) )
//   pipeline.py:455-465   nearest x2, bf16(bf16(alpha*x) + bf16(beta*n))
// Explicit __f*_rn intrinsics keep nvcc from contracting mul+add into FMA, which would
// change the rounding.
#include "kernels.cuh"

namespace dv {
namespace {

__device__ __forceinline__ float rb(float x) {  // round-trip through bf16 (RNE)
  return __bfloat162float(__float2bfloat16_rn(x));
}

template <bool kBf16>
__device__ __forceinline__ float rnd(float x) {
  return kBf16 ? rb(x) : x;
}

template <typename T, bool kBf16>
__global__ void cfg_euler_kernel(const T* __restrict__ pred, int n_branch, const T* __restrict__ x,
                                 T* __restrict__ out, long long numel, float w_text, float w_hist,
                                 float dsigma) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= numel) return;
  auto ld = [&](const T* p, long long k) -> float {
    if constexpr (kBf16)
      return __bfloat162float(p[k]);
    else
      return p[k];
  };
  float v;
  if (n_branch == 1) {
    v = ld(pred, i);
  } else {
    const float u = ld(pred, i);
    const float t = ld(pred, numel + i);
    const float a1 = rnd<kBf16>(__fsub_rn(t, u));
    const float a2 = rnd<kBf16>(__fmul_rn(a1, w_text));
    v = rnd<kBf16>(__fadd_rn(u, a2));
    if (n_branch == 3) {
      const float h = ld(pred, 2 * numel + i);
      const float b1 = rnd<kBf16>(__fsub_rn(h, t));
      const float b2 = rnd<kBf16>(__fmul_rn(b1, w_hist));
      v = rnd<kBf16>(__fadd_rn(v, b2));
    }
  }
  const float prod = rnd<kBf16>(__fmul_rn(dsigma, v));
  const float r = __fadd_rn(ld(x, i), prod);
  if constexpr (kBf16)
    out[i] = __float2bfloat16_rn(r);
  else
    out[i] = r;
}

template <typename T, bool kBf16>
__global__ void stage_renoise_kernel(const T* __restrict__ lo, const T* __restrict__ noise,
                                     T* __restrict__ out, int planes, int h, int w, float alpha,
                                     float beta) {
  const int H2 = 2 * h, W2 = 2 * w;
  const long long total = static_cast<long long>(planes) * H2 * W2;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = i % W2;
  const int y = (i / W2) % H2;
  const long long p = i / (static_cast<long long>(W2) * H2);
  float src, nz;
  if constexpr (kBf16) {
    src = __bfloat162float(lo[(p * h + (y >> 1)) * w + (x >> 1)]);
    nz = __bfloat162float(noise[i]);
  } else {
    src = lo[(p * h + (y >> 1)) * w + (x >> 1)];
    nz = noise[i];
  }
  const float a = rnd<kBf16>(__fmul_rn(src, alpha));
  const float b = rnd<kBf16>(__fmul_rn(nz, beta));
  const float r = __fadd_rn(a, b);
  if constexpr (kBf16)
    out[i] = __float2bfloat16_rn(r);
  else
    out[i] = r;
}

// noise block (p,q in 2x2) = L z, L = chol((1+g) I - g 11^T); z iid N(0,1) [planes][h/2][w/2][4]
template <typename T>
__global__ void block_noise_kernel(const float* __restrict__ z, T* __restrict__ out, int planes,
                                   int h, int w, float l00, float l10, float l11, float l21,
                                   float l22, float l32, float l33) {
  const int bh = h / 2, bw = w / 2;
  const long long total = static_cast<long long>(planes) * bh * bw;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int bx = i % bw;
  const int by = (i / bw) % bh;
  const long long p = i / (static_cast<long long>(bw) * bh);
  const float4 zz = reinterpret_cast<const float4*>(z)[i];
  // L has the structure: column j constant below the diagonal (exchangeable covariance)
  const float n0 = l00 * zz.x;
  const float n1 = l10 * zz.x + l11 * zz.y;
  const float n2 = l10 * zz.x + l21 * zz.y + l22 * zz.z;
  const float n3 = l10 * zz.x + l21 * zz.y + l32 * zz.z + l33 * zz.w;
  T* o = out + (p * h + 2 * by) * w + 2 * bx;
  if constexpr (sizeof(T) == 2) {
    o[0] = __float2bfloat16_rn(n0);
    o[1] = __float2bfloat16_rn(n1);
    o[w] = __float2bfloat16_rn(n2);
    o[w + 1] = __float2bfloat16_rn(n3);
  } else {
    o[0] = n0;
    o[1] = n1;
    o[w] = n2;
    o[w + 1] = n3;
  }
}

}  // namespace

int launch_cfg_euler(const void* noise_pred, int n_branch, const void* sample, void* out,
                     long long numel, float w_text, float w_hist, double sigma, double sigma_next,
                     int is_bf16, cudaStream_t stream) {
  DV_REQUIRE(n_branch >= 1 && n_branch <= 3, "cfg_euler: n_branch=%d", n_branch);
  if (numel == 0) return 0;
  const int blocks = static_cast<int>((numel + 255) / 256);
  const double ds = sigma_next - sigma;  // fp64 like the 0-dim tensor in scheduler.py:280-283
  if (is_bf16) {
    // the 0-dim fp64 operand is cast to the bf16 result dtype before the multiply
    const float dsf = __bfloat162float(__float2bfloat16_rn(static_cast<float>(ds)));
    // NOTE: double -> bf16 is done through float; exact for the reference's sigma tables
    cfg_euler_kernel<__nv_bfloat16, true><<<blocks, 256, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(noise_pred), n_branch,
        reinterpret_cast<const __nv_bfloat16*>(sample), reinterpret_cast<__nv_bfloat16*>(out),
        numel, w_text, w_hist, dsf);
  } else {
    cfg_euler_kernel<float, false><<<blocks, 256, 0, stream>>>(
        reinterpret_cast<const float*>(noise_pred), n_branch,
        reinterpret_cast<const float*>(sample), reinterpret_cast<float*>(out), numel, w_text,
        w_hist, static_cast<float>(ds));
  }
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

int launch_stage_renoise(const void* lat_lo, const void* noise, void* out, int planes, int h, int w,
                         double alpha, double beta, int is_bf16, cudaStream_t stream) {
  const long long total = static_cast<long long>(planes) * 4 * h * w;
  if (total == 0) return 0;
  const int blocks = static_cast<int>((total + 255) / 256);
  if (is_bf16)
    stage_renoise_kernel<__nv_bfloat16, true><<<blocks, 256, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(lat_lo),
        reinterpret_cast<const __nv_bfloat16*>(noise), reinterpret_cast<__nv_bfloat16*>(out),
        planes, h, w, static_cast<float>(alpha), static_cast<float>(beta));
  else
    stage_renoise_kernel<float, false><<<blocks, 256, 0, stream>>>(
        reinterpret_cast<const float*>(lat_lo), reinterpret_cast<const float*>(noise),
        reinterpret_cast<float*>(out), planes, h, w, static_cast<float>(alpha),
        static_cast<float>(beta));
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

int launch_block_noise(const float* z, void* out, int planes, int h, int w, float gamma,
                       int is_bf16, cudaStream_t stream) {
  DV_REQUIRE(h % 2 == 0 && w % 2 == 0, "block_noise: h=%d w=%d must be even", h, w);
  // Cholesky of (1+g) I - g 11^T, 4x4, in double
  double a[4][4], l[4][4] = {{0}};
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) a[i][j] = (i == j ? 1.0 + gamma : 0.0) - gamma;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = a[i][j];
      for (int k = 0; k < j; ++k) s -= l[i][k] * l[j][k];
      if (i == j) {
        // gamma = 1/3 makes the covariance exactly singular (last pivot 0): semi-definite is fine
        DV_REQUIRE(s > -1e-6, "block_noise: covariance not positive semi-definite (gamma=%f)", gamma);
        l[i][j] = s > 0.0 ? sqrt(s) : 0.0;
      } else {
        l[i][j] = l[j][j] > 0.0 ? s / l[j][j] : 0.0;
      }
    }
  const long long total = static_cast<long long>(planes) * (h / 2) * (w / 2);
  const int blocks = static_cast<int>((total + 255) / 256);
  if (is_bf16)
    block_noise_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(
        z, reinterpret_cast<__nv_bfloat16*>(out), planes, h, w, (float)l[0][0], (float)l[1][0],
        (float)l[1][1], (float)l[2][1], (float)l[2][2], (float)l[3][2], (float)l[3][3]);
  else
    block_noise_kernel<float><<<blocks, 256, 0, stream>>>(
        z, reinterpret_cast<float*>(out), planes, h, w, (float)l[0][0], (float)l[1][0],
        (float)l[1][1], (float)l[2][1], (float)l[2][2], (float)l[3][2], (float)l[3][3]);
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

namespace {
// ---------------------------------------------------------------------------------------
// Bilinear down-sampling by exactly 2 (pipeline.py:226-240 get_pyramid_latent, :554-557 initial
// noise pyramid; F.interpolate(mode='bilinear', align_corners=False)): source index 2*o + 0.5, both
// lambdas 0.5, ATen's association  0.5*(0.5*a + 0.5*b) + 0.5*(0.5*c + 0.5*d)  in fp32 — every
// scaling by 0.5 is exact, so the three additions below round exactly where ATen's do.  `scale`
// is a separate multiply of the rounded result (the reference's `* 2` on the noise latents).
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void resize_half_kernel(const T* __restrict__ in, T* __restrict__ out, long long planes, int H,
                                   int W, float scale) {
  const int oh = H / 2, ow = W / 2;
  const long long total = planes * oh * ow;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int x = static_cast<int>(idx % ow);
  const int y = static_cast<int>((idx / ow) % oh);
  const long long pl = idx / (static_cast<long long>(ow) * oh);
  const T* src = in + (pl * H + 2 * y) * W + 2 * x;
  float a, b, c, d;
  if constexpr (sizeof(T) == 2) {
    a = __bfloat162float(src[0]);
    b = __bfloat162float(src[1]);
    c = __bfloat162float(src[W]);
    d = __bfloat162float(src[W + 1]);
  } else {
    a = src[0];
    b = src[1];
    c = src[W];
    d = src[W + 1];
  }
  const float top = __fmul_rn(0.5f, __fadd_rn(a, b));
  const float bot = __fmul_rn(0.5f, __fadd_rn(c, d));
  const float v = __fmul_rn(0.5f, __fadd_rn(top, bot));
  if constexpr (sizeof(T) == 2) {
    const __nv_bfloat16 r = __float2bfloat16(v);
    out[idx] = scale == 1.0f ? r : __float2bfloat16(__fmul_rn(__bfloat162float(r), scale));
  } else {
    out[idx] = scale == 1.0f ? v : __fmul_rn(v, scale);
  }
}

}  // namespace

int launch_resize_half(const void* in, void* out, long long planes, int H, int W, float scale, int is_bf16,
                       cudaStream_t stream) {
  DV_REQUIRE(H % 2 == 0 && W % 2 == 0 && H >= 2 && W >= 2, "resize_half: H=%d W=%d must be even", H, W);
  const long long total = planes * (H / 2) * (W / 2);
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  if (is_bf16)
    resize_half_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(in),
                                                                 reinterpret_cast<__nv_bfloat16*>(out), planes, H, W,
                                                                 scale);
  else
    resize_half_kernel<float><<<blocks, 256, 0, stream>>>(reinterpret_cast<const float*>(in),
                                                          reinterpret_cast<float*>(out), planes, H, W, scale);
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

namespace {
}  // namespace

}  // namespace dv

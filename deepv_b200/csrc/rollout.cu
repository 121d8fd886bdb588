// deepv_b200 — the feedback between rollout iterations on the device.
//
// The reference carries an iteration's results to the next one through the host: decoded frames
// go GPU -> numpy -> uint8 PIL images -> ToTensor/Normalize -> GPU (pipeline.py:339-344,564-568),
// disparity is post-processed and renormalised with a `.max()` (:311-313,346-350), generated ray
// maps are turned into camera poses on the CPU (:77-163,692) and poses back into ray maps
// (:29-75,362-368,406-410).  These kernels keep all of it in HBM with no host synchronisation:
//   dv_frames_requantise   the uint8 round trip as arithmetic (truncation included)
//   dv_disparity_post      mean over channels, square, running scale
//   dv_disparity_renorm    1/max of the first kept frame (device scalar), sqrt re-encoding
//   dv_raymap_to_pose      per-frame reductions of the ray map -> camera-to-world + intrinsics
//   dv_camera_raymap       cameras -> 8x8-averaged unit rays + origins, optionally normalised
// All are HBM/latency-bound elementwise or small-reduction kernels in fp32.
#include "../../include/deepv_b200.h"
#include "common.cuh"

using namespace dv;

namespace {

// pipeline.py:200-201
__constant__ float kRayMean[6] = {-0.0016f, -0.0010f, 0.9015f, 0.0313f, -0.0538f, 0.2079f};
__constant__ float kRayStd[6] = {0.3333f, 0.2567f, 0.0927f, 0.4338f, 0.1746f, 0.5802f};

template <typename T>
__device__ __forceinline__ float ldf(const T* p, long long i) {
  if constexpr (sizeof(T) == 2)
    return __bfloat162float(p[i]);
  else
    return p[i];
}
template <typename T>
__device__ __forceinline__ void stf(T* p, long long i, float v) {
  if constexpr (sizeof(T) == 2)
    p[i] = __float2bfloat16_rn(v);
  else
    p[i] = v;
}

// ---------------------------------------------------------------------------------------
// frames [3][T][HW] in [-1,1] -> q = trunc(clamp(x*0.5+0.5, 0, 1) * 255) (numpy astype(uint8) truncates,
// pipeline.py:341), then ToTensor (q/255) and Normalize ((v-0.5)/0.5) (pipeline.py:564-567).
template <typename TI, typename TO>
__global__ void requantise_kernel(const TI* __restrict__ in, int T, long long HW, int t0, int n,
                                  TO* __restrict__ out, unsigned char* __restrict__ u8) {
  const long long total = 3LL * n * HW;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long p = i % HW;
  const int f = static_cast<int>((i / HW) % n);
  const int c = static_cast<int>(i / (HW * n));
  const float x = ldf(in, (static_cast<long long>(c) * T + t0 + f) * HW + p);
  float v;
  if constexpr (sizeof(TI) == 2) {
    // the reference does `x * 0.5 + 0.5` in the decode dtype and only then `.to(float32)` (pipeline.py:341):
    // for bf16 frames ATen rounds after each op, so the sum sits on the bf16 grid before `* 255`
    v = __bfloat162float(__float2bfloat16_rn(__fmul_rn(x, 0.5f)));
    v = __bfloat162float(__float2bfloat16_rn(__fadd_rn(v, 0.5f)));
  } else {
    v = __fadd_rn(__fmul_rn(x, 0.5f), 0.5f);
  }
  v = fminf(fmaxf(v, 0.0f), 1.0f);
  const float q = truncf(__fmul_rn(v, 255.0f));
  if (u8) u8[(static_cast<long long>(f) * HW + p) * 3 + c] = static_cast<unsigned char>(q);
  const float y = __fdiv_rn(__fsub_rn(__fdiv_rn(q, 255.0f), 0.5f), 0.5f);
  stf(out, i, y);
}

// disparity = clamp(mean_c(raw) * 0.5 + 0.5, 0, 1)^2 / scale / 0.95, written to 3 channels (pipeline.py:311-313)
template <typename TI>
__global__ void disparity_post_kernel(const TI* __restrict__ raw, long long THW, const float* __restrict__ scale,
                                      float* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= THW) return;
  const float s = scale ? *scale : 1.0f;
  const float sum = __fadd_rn(__fadd_rn(ldf(raw, i), ldf(raw, THW + i)), ldf(raw, 2 * THW + i));
  float v = __fadd_rn(__fmul_rn(__fdiv_rn(sum, 3.0f), 0.5f), 0.5f);
  v = fminf(fmaxf(v, 0.0f), 1.0f);
  v = __fdiv_rn(__fdiv_rn(__fmul_rn(v, v), s), 0.95f);
  out[i] = v;
  out[THW + i] = v;
  out[2 * THW + i] = v;
}

// scale = 1 / max(disp[:, :, t0]) (pipeline.py:347); one CTA, fixed order -> deterministic
__global__ void disparity_scale_kernel(const float* __restrict__ disp, int T, long long HW, int t0,
                                       float* __restrict__ scale) {
  __shared__ float red[32];
  float m = -INFINITY;
  for (int c = 0; c < 3; ++c) {
    const float* src = disp + (static_cast<long long>(c) * T + t0) * HW;
    for (long long i = threadIdx.x; i < HW; i += blockDim.x) m = fmaxf(m, src[i]);
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -INFINITY;
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) *scale = __fdiv_rn(1.0f, m);
  }
}

// out = sqrt(disp * scale * 0.95) * 2 - 1 for frames [t0, t0+n) (pipeline.py:348-350; :399-401 adds the clamp)
template <typename TO>
__global__ void disparity_renorm_kernel(const float* __restrict__ disp, int T, long long HW, int t0, int n,
                                        const float* __restrict__ scale, int clamp, TO* __restrict__ out) {
  const long long total = 3LL * n * HW;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long p = i % HW;
  const int f = static_cast<int>((i / HW) % n);
  const int c = static_cast<int>(i / (HW * n));
  const float d = disp[(static_cast<long long>(c) * T + t0 + f) * HW + p];
  float v = __fsqrt_rn(__fmul_rn(__fmul_rn(d, *scale), 0.95f));
  v = __fsub_rn(__fmul_rn(v, 2.0f), 1.0f);
  if (clamp) v = fminf(fmaxf(v, -1.0f), 1.0f);
  stf(out, i, v);
}

// ---------------------------------------------------------------------------------------
// Ray map -> pose (pipeline.py:77-163).  One CTA per decoded frame f in [1, T).
struct V3 {
  float x, y, z;
};
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ float norm(V3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
__device__ __forceinline__ V3 unit(V3 a) {
  const float n = norm(a);
  return {a.x / n, a.y / n, a.z / n};
}

template <int N>
__device__ void block_sum(float (&v)[N], float* smem /* [N][32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    float s = v[k];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) smem[k * 32 + warp] = s;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      float s = lane < nw ? smem[k * 32 + lane] : 0.0f;
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) smem[k * 32] = s;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) v[k] = smem[k * 32];
  __syncthreads();
}

template <typename T>
__global__ void raymap_pose_kernel(const T* __restrict__ lat, int c0, int Tn, int h, int w, int ds,
                                   float* __restrict__ pose, float* __restrict__ intr) {
  __shared__ float red[18 * 32];
  const int f = blockIdx.x + 1;
  const int hw = h * w;
  const long long cs = static_cast<long long>(Tn) * hw;          // channel stride
  const T* base = lat + static_cast<long long>(c0) * cs + static_cast<long long>(f) * hw;
  auto ray = [&](int c, int p) { return ldf(base, c * cs + p) * kRayStd[c] + kRayMean[c]; };

  float s[3] = {0.f, 0.f, 0.f};
  for (int p = threadIdx.x; p < hw; p += blockDim.x)
    for (int c = 0; c < 3; ++c) s[c] += ray(c, p);
  block_sum<3>(s, red);
  const V3 ref = unit(V3{s[0] / hw, s[1] / hw, s[2] / hw});

  float a[18];
#pragma unroll
  for (int k = 0; k < 18; ++k) a[k] = 0.f;
  for (int p = threadIdx.x; p < hw; p += blockDim.x) {
    const int x = p % w, y = p / w;
    float d[3], o[3];
    for (int c = 0; c < 3; ++c) d[c] = ray(c, p);
    const float proj = d[0] * ref.x + d[1] * ref.y + d[2] * ref.z;
    for (int c = 0; c < 3; ++c) {
      d[c] = d[c] / proj;
      const float v = ray(3 + c, p);
      o[c] = copysignf(v * v, v);
      a[c] += o[c];
      a[3 + c] += o[c] + d[c];
      if (x == 0) a[6 + c] += d[c];
      if (x == w - 1) a[9 + c] += d[c];
      if (y == 0) a[12 + c] += d[c];
      if (y == h - 1) a[15 + c] += d[c];
    }
  }
  block_sum<18>(a, red);
  if (threadIdx.x != 0) return;
  const V3 loc{a[0] / hw, a[1] / hw, a[2] / hw};
  const V3 img{a[3] / hw, a[4] / hw, a[5] / hw};
  const V3 left{a[6] / h, a[7] / h, a[8] / h}, right{a[9] / h, a[10] / h, a[11] / h};
  const V3 up{a[12] / w, a[13] / w, a[14] / w}, down{a[15] / w, a[16] / w, a[17] / w};
  const V3 zd = img - loc;
  const float focal = norm(zd);
  const V3 wv = right - left, hv = up - down;
  const float w_real = norm(cross(wv, zd)) / (w - 1) * w;
  const float h_real = norm(cross(hv, zd)) / (h - 1) * h;
  const V3 yd = cross(zd, wv);
  const V3 xd = cross(yd, zd);
  const V3 X = unit(xd), Y = unit(yd), Z = unit(zd);
  float* P = pose + f * 16;
  P[0] = X.x, P[1] = Y.x, P[2] = Z.x, P[3] = loc.x;
  P[4] = X.y, P[5] = Y.y, P[6] = Z.y, P[7] = loc.y;
  P[8] = X.z, P[9] = Y.z, P[10] = Z.z, P[11] = loc.z;
  P[12] = 0.f, P[13] = 0.f, P[14] = 0.f, P[15] = 1.f;
  const float rescale = (w / w_real + h / h_real) / 2 * ds;
  for (int rep = 0; rep < (f == 1 ? 2 : 1); ++rep) {      // frame 0 reuses frame 1's intrinsics (pipeline.py:150-151)
    float* K = intr + (rep ? 0 : f) * 16;
    for (int k = 0; k < 16; ++k) K[k] = 0.f;
    K[0] = focal * rescale;
    K[5] = focal * rescale;
    K[2] = w / 2.0f * ds;
    K[6] = h / 2.0f * ds;
    K[10] = 1.f;
    K[15] = 1.f;
  }
}

// pose[0] = I, pose[i] = pose[i-1] @ pose[i] (pipeline.py:146-148,156-158): 16 threads, one per entry
__global__ void pose_chain_kernel(float* __restrict__ pose, int Tn) {
  __shared__ float prev[16], cur[16];
  const int r = threadIdx.x >> 2, c = threadIdx.x & 3;
  prev[threadIdx.x] = r == c ? 1.f : 0.f;
  pose[threadIdx.x] = prev[threadIdx.x];
  __syncthreads();
  for (int i = 1; i < Tn; ++i) {
    cur[threadIdx.x] = pose[i * 16 + threadIdx.x];
    __syncthreads();
    float v = 0.f;
    for (int k = 0; k < 4; ++k) v += prev[r * 4 + k] * cur[k * 4 + c];
    __syncthreads();
    prev[threadIdx.x] = v;
    pose[i * 16 + threadIdx.x] = v;
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// Cameras -> ray map (pipeline.py:29-75): per latent pixel the mean over its ds x ds pixel block of
// ((u-cu)/fu, (v-cv)/fv, 1), rotated by the camera-to-world rotation and normalised; origin = translation.
template <typename T>
__global__ void camera_raymap_kernel(const float* __restrict__ k2, const float* __restrict__ k3, int n, int h, int w,
                                     int ds, int normalise, T* __restrict__ out) {
  const long long total = static_cast<long long>(n) * h * w;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = static_cast<int>(i % w), y = static_cast<int>((i / w) % h), f = static_cast<int>(i / (static_cast<long long>(w) * h));
  const float* K = k2 + f * 16;
  const float* P = k3 + f * 16;
  const float fu = K[0], fv = K[5], cu = K[2], cv = K[6];
  float sx = 0.f, sy = 0.f;
  for (int d = 0; d < ds; ++d) {
    sx += (static_cast<float>(x * ds + d) - cu) / fu;
    sy += (static_cast<float>(y * ds + d) - cv) / fv;
  }
  const float rx = sx / ds, ry = sy / ds;
  V3 d{P[0] * rx + P[1] * ry + P[2], P[4] * rx + P[5] * ry + P[6], P[8] * rx + P[9] * ry + P[10]};
  d = unit(d);
  const float v[6] = {d.x, d.y, d.z, P[3], P[7], P[11]};
  const long long hw = static_cast<long long>(h) * w;
  for (int c = 0; c < 6; ++c) {
    const float r = normalise ? (v[c] - kRayMean[c]) / kRayStd[c] : v[c];
    stf(out, (static_cast<long long>(c) * n + f) * hw + static_cast<long long>(y) * w + x, r);
  }
}

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline bool ok_dtype(int d) { return d == DV_DTYPE_F32 || d == DV_DTYPE_BF16; }
inline unsigned blocks_for(long long total) { return static_cast<unsigned>((total + 255) / 256); }

}  // namespace

extern "C" int dv_frames_requantise(const void* frames_dev, int dtype, int T, int H, int W, int t0, int n,
                                    void* out_dev, int out_dtype, unsigned char* u8_dev, void* stream) {
  DV_REQUIRE(frames_dev && out_dev, "dv_frames_requantise: null pointer");
  DV_REQUIRE(ok_dtype(dtype) && ok_dtype(out_dtype), "dv_frames_requantise: dtype %d -> %d", dtype, out_dtype);
  DV_REQUIRE(T > 0 && H > 0 && W > 0 && t0 >= 0 && n > 0 && t0 + n <= T,
             "dv_frames_requantise: frames [%d, %d) of %d", t0, t0 + n, T);
  const long long HW = static_cast<long long>(H) * W;
  const unsigned g = blocks_for(3LL * n * HW);
  using bf = __nv_bfloat16;
  if (dtype == DV_DTYPE_BF16 && out_dtype == DV_DTYPE_BF16)
    requantise_kernel<bf, bf><<<g, 256, 0, S(stream)>>>(static_cast<const bf*>(frames_dev), T, HW, t0, n, static_cast<bf*>(out_dev), u8_dev);
  else if (dtype == DV_DTYPE_BF16)
    requantise_kernel<bf, float><<<g, 256, 0, S(stream)>>>(static_cast<const bf*>(frames_dev), T, HW, t0, n, static_cast<float*>(out_dev), u8_dev);
  else if (out_dtype == DV_DTYPE_BF16)
    requantise_kernel<float, bf><<<g, 256, 0, S(stream)>>>(static_cast<const float*>(frames_dev), T, HW, t0, n, static_cast<bf*>(out_dev), u8_dev);
  else
    requantise_kernel<float, float><<<g, 256, 0, S(stream)>>>(static_cast<const float*>(frames_dev), T, HW, t0, n, static_cast<float*>(out_dev), u8_dev);
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

extern "C" int dv_disparity_post(const void* raw_dev, int dtype, int T, int H, int W, const float* scale_dev,
                                 float* out_dev, void* stream) {
  DV_REQUIRE(raw_dev && out_dev, "dv_disparity_post: null pointer");
  DV_REQUIRE(ok_dtype(dtype), "dv_disparity_post: dtype %d", dtype);
  DV_REQUIRE(T > 0 && H > 0 && W > 0, "dv_disparity_post: T=%d H=%d W=%d", T, H, W);
  const long long THW = static_cast<long long>(T) * H * W;
  if (dtype == DV_DTYPE_BF16)
    disparity_post_kernel<__nv_bfloat16><<<blocks_for(THW), 256, 0, S(stream)>>>(static_cast<const __nv_bfloat16*>(raw_dev), THW, scale_dev, out_dev);
  else
    disparity_post_kernel<float><<<blocks_for(THW), 256, 0, S(stream)>>>(static_cast<const float*>(raw_dev), THW, scale_dev, out_dev);
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

extern "C" int dv_disparity_renorm(const float* disp_dev, int T, int H, int W, int t0, int n, float* scale_dev,
                                   int compute_scale, int clamp, void* out_dev, int out_dtype, void* stream) {
  DV_REQUIRE(disp_dev && scale_dev && out_dev, "dv_disparity_renorm: null pointer");
  DV_REQUIRE(ok_dtype(out_dtype), "dv_disparity_renorm: dtype %d", out_dtype);
  DV_REQUIRE(T > 0 && H > 0 && W > 0 && t0 >= 0 && n > 0 && t0 + n <= T,
             "dv_disparity_renorm: frames [%d, %d) of %d", t0, t0 + n, T);
  const long long HW = static_cast<long long>(H) * W;
  if (compute_scale) {
    disparity_scale_kernel<<<1, 1024, 0, S(stream)>>>(disp_dev, T, HW, t0, scale_dev);
    DV_CHECK_CUDA(cudaGetLastError());
    note_launch();
  }
  const unsigned g = blocks_for(3LL * n * HW);
  if (out_dtype == DV_DTYPE_BF16)
    disparity_renorm_kernel<__nv_bfloat16><<<g, 256, 0, S(stream)>>>(disp_dev, T, HW, t0, n, scale_dev, clamp, static_cast<__nv_bfloat16*>(out_dev));
  else
    disparity_renorm_kernel<float><<<g, 256, 0, S(stream)>>>(disp_dev, T, HW, t0, n, scale_dev, clamp, static_cast<float*>(out_dev));
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

extern "C" int dv_raymap_to_pose(const void* latents_dev, int dtype, int C, int c0, int T, int h, int w, int ds,
                                 float* trans3d_dev, float* trans2d_dev, void* stream) {
  DV_REQUIRE(latents_dev && trans3d_dev && trans2d_dev, "dv_raymap_to_pose: null pointer");
  DV_REQUIRE(ok_dtype(dtype), "dv_raymap_to_pose: dtype %d", dtype);
  DV_REQUIRE(c0 >= 0 && c0 + 6 <= C && T >= 2 && h >= 2 && w >= 2 && ds > 0,
             "dv_raymap_to_pose: C=%d c0=%d T=%d h=%d w=%d", C, c0, T, h, w);
  if (dtype == DV_DTYPE_BF16)
    raymap_pose_kernel<__nv_bfloat16><<<T - 1, 256, 0, S(stream)>>>(static_cast<const __nv_bfloat16*>(latents_dev), c0, T, h, w, ds, trans3d_dev, trans2d_dev);
  else
    raymap_pose_kernel<float><<<T - 1, 256, 0, S(stream)>>>(static_cast<const float*>(latents_dev), c0, T, h, w, ds, trans3d_dev, trans2d_dev);
  DV_CHECK_CUDA(cudaGetLastError());
  pose_chain_kernel<<<1, 16, 0, S(stream)>>>(trans3d_dev, T);
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch(2);
  return 0;
}

extern "C" int dv_camera_raymap(const float* trans2d_dev, const float* trans3d_dev, int n, int H, int W, int ds,
                                int normalise, void* out_dev, int out_dtype, void* stream) {
  DV_REQUIRE(trans2d_dev && trans3d_dev && out_dev, "dv_camera_raymap: null pointer");
  DV_REQUIRE(ok_dtype(out_dtype), "dv_camera_raymap: dtype %d", out_dtype);
  DV_REQUIRE(n > 0 && ds > 0 && H >= ds && W >= ds, "dv_camera_raymap: n=%d H=%d W=%d ds=%d", n, H, W, ds);
  const int h = H / ds, w = W / ds;
  const unsigned g = blocks_for(static_cast<long long>(n) * h * w);
  if (out_dtype == DV_DTYPE_BF16)
    camera_raymap_kernel<__nv_bfloat16><<<g, 256, 0, S(stream)>>>(trans2d_dev, trans3d_dev, n, h, w, ds, normalise, static_cast<__nv_bfloat16*>(out_dev));
  else
    camera_raymap_kernel<float><<<g, 256, 0, S(stream)>>>(trans2d_dev, trans3d_dev, n, h, w, ds, normalise, static_cast<float*>(out_dev));
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

// deepv_b200 — C ABI entry points that are thin shims over the kernel launchers
// (sampler step and the exported building blocks).  dv_mmdit_* live in mmdit.cu and
// dv_vae_* in vae.cu.  See include/deepv_b200.h for the contract.
#include "../../include/deepv_b200.h"
#include "gemm.cuh"
#include "kernels.cuh"

using namespace dv;

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" const char* dv_last_error(void) { return last_error(); }
extern "C" int dv_version(void) { return 100; }
extern "C" long long dv_launch_count(void) { return launch_count(); }
extern "C" void dv_launch_count_reset(void) { launch_count_reset(); }

extern "C" void dv_profile_enable(int on) { prof_enable(on != 0); }
extern "C" void dv_profile_reset(void) { prof_reset(); }
extern "C" int dv_profile_dump(const char* path) {
  DV_REQUIRE(path, "dv_profile_dump: null path");
  return prof_dump(path);
}
extern "C" int dv_profile_summary(int kind, long long* count, double* ms, double* flops,
                                  double* bytes) {
  DV_REQUIRE(count && ms && flops && bytes, "dv_profile_summary: null pointer");
  return prof_summary(kind, count, ms, flops, bytes);
}

extern "C" int dv_cfg_euler_step(const void* noise_pred_dev, int n_branch, const void* sample_dev,
                                 void* out_dev, long long numel, float w_text, float w_hist,
                                 double sigma, double sigma_next, int dtype, void* stream) {
  DV_REQUIRE(noise_pred_dev && sample_dev && out_dev, "dv_cfg_euler_step: null pointer");
  DV_REQUIRE(dtype == DV_DTYPE_F32 || dtype == DV_DTYPE_BF16, "dv_cfg_euler_step: dtype %d", dtype);
  return launch_cfg_euler(noise_pred_dev, n_branch, sample_dev, out_dev, numel, w_text, w_hist,
                          sigma, sigma_next, dtype == DV_DTYPE_BF16, S(stream));
}

extern "C" int dv_stage_renoise(const void* lat_lo_dev, const void* noise_dev, void* out_dev,
                                int planes, int h, int w, double alpha, double beta, int dtype,
                                void* stream) {
  DV_REQUIRE(lat_lo_dev && noise_dev && out_dev, "dv_stage_renoise: null pointer");
  DV_REQUIRE(dtype == DV_DTYPE_F32 || dtype == DV_DTYPE_BF16, "dv_stage_renoise: dtype %d", dtype);
  return launch_stage_renoise(lat_lo_dev, noise_dev, out_dev, planes, h, w, alpha, beta,
                              dtype == DV_DTYPE_BF16, S(stream));
}

extern "C" int dv_resize_half(const void* in_dev, void* out_dev, long long planes, int H, int W, float scale,
                              int dtype, void* stream) {
  DV_REQUIRE(in_dev && out_dev && planes > 0, "dv_resize_half: bad argument");
  return launch_resize_half(in_dev, out_dev, planes, H, W, scale, dtype == DV_DTYPE_BF16, S(stream));
}

extern "C" int dv_block_noise(const float* z_dev, void* out_dev, int planes, int h, int w,
                              float gamma, int dtype, void* stream) {
  DV_REQUIRE(z_dev && out_dev, "dv_block_noise: null pointer");
  return launch_block_noise(z_dev, out_dev, planes, h, w, gamma, dtype == DV_DTYPE_BF16, S(stream));
}

extern "C" int dv_gemm_bf16(const void* A_dev, const void* W_dev, const float* bias_dev,
                            void* C_dev, int batch, int M, int N, int K, int epi, void* stream) {
  DV_REQUIRE(A_dev && W_dev && C_dev, "dv_gemm_bf16: null pointer");
  GemmDesc d = {};
  d.batch = batch;
  d.M = M;
  d.N = N;
  d.K = K;
  d.A = A_dev;
  d.a_batch_stride = static_cast<long long>(M) * K;
  d.lda = K;
  d.W = W_dev;
  d.w_rows = N;
  d.bias = bias_dev;
  d.mode = epi == 1 ? EPI_GELU : EPI_BF16;
  d.out = C_dev;
  d.out_batch_stride = static_cast<long long>(M) * N;
  d.ldo = N;
  return launch_gemm(d, S(stream));
}

extern "C" int dv_attention(const void* qkv_dev, void* out_dev, const int* kv_end_dev,
                            const float* key_bias_dev, int B, int L, int Lpad, int H,
                            void* stream) {
  DV_REQUIRE(qkv_dev && out_dev && kv_end_dev && key_bias_dev, "dv_attention: null pointer");
  return launch_attention(qkv_dev, out_dev, kv_end_dev, key_bias_dev, nullptr, B, L, Lpad, H,
                          S(stream));
}

extern "C" int dv_conv3d_strided_cl(const void* x_dev, const void* w_dev, const float* bias_dev,
                                    void* out_dev, int B, int T, int H, int W, int Cin, int Cout,
                                    int w_rows, int sT, int sH, int sW, void* stream) {
  DV_REQUIRE(x_dev && w_dev && out_dev, "dv_conv3d_strided_cl: null pointer");
  GemmDesc d = {};
  d.batch = B;
  d.N = Cout;
  d.A = x_dev;
  d.a_mode = 1;
  d.cT = T;
  d.cH = H;
  d.cW = W;
  d.cC = Cin;
  d.kt = d.kh = d.kw = 3;
  d.sT = sT;
  d.sH = sH;
  d.sW = sW;
  d.W = w_dev;
  d.w_rows = w_rows;
  d.bias = bias_dev;
  d.mode = EPI_CONV;
  d.out = out_dev;
  d.conv_store = CONV_PLAIN;
  d.out_C = Cout;
  return launch_gemm(d, S(stream));
}

extern "C" int dv_conv3d_cl(const void* x_dev, const void* w_dev, const float* bias_dev,
                            const void* residual_dev, void* out_dev, int B, int T, int H, int W,
                            int Cin, int Cout, int w_rows, int ksize, int store, int drop_first,
                            void* stream) {
  DV_REQUIRE(x_dev && w_dev && out_dev, "dv_conv3d_cl: null pointer");
  DV_REQUIRE(ksize == 1 || ksize == 3, "dv_conv3d_cl: ksize %d", ksize);
  GemmDesc d = {};
  d.batch = B;
  d.N = Cout;
  d.A = x_dev;
  d.a_mode = 1;
  d.cT = T;
  d.cH = H;
  d.cW = W;
  d.cC = Cin;
  d.kt = d.kh = d.kw = ksize;
  d.W = w_dev;
  d.w_rows = w_rows;
  d.bias = bias_dev;
  d.mode = EPI_CONV;
  d.out = out_dev;
  d.conv_store = store;
  d.conv_drop_first = drop_first;
  d.residual = residual_dev;
  d.out_C = store == CONV_SHUFFLE_HW ? Cout / 4 : (store == CONV_INTERLEAVE_T ? Cout / 2 : Cout);
  DV_REQUIRE(residual_dev == nullptr || store == CONV_PLAIN,
             "dv_conv3d_cl: residual only with the plain store");
  return launch_gemm(d, S(stream));
}

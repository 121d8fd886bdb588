// Probe (experiment, not part of the library): can a tcgen05.mma B operand be a ROW-SHIFTED view of a
// 128B-swizzled K-major tile that TMA wrote once?  That is what a conv needs to serve the three kw
// taps from one pixel tile + halo.
//   mode 0: rows [shift, shift+256) of a 272-row tile, 8-row groups 1024 B apart (contiguous rows)
//   mode 1: rows h*10 + w + shift (h < 32, w < 8) of a 320-row tile = an 8-wide pixel tile with a 2-pixel
//           halo, 8-row groups 1280 B apart
// base_off: 0 = leave the descriptor's base-offset field 0, 1 = (start >> 7) & 7.
#include "../../deepv_b200/csrc/common.cuh"

using namespace dv;

namespace {
constexpr int kRowsB = 320;
struct Args {
  CUtensorMap tmA, tmB;
  float* out;   // [128][256]
  int shift, mode, base_off;
};

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ Args a) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sa = smem;                  // 128 rows x 128 B
  uint8_t* sb = smem + 16384;          // 320 rows x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(sb + kRowsB * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<256>(slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[0], 16384 + kRowsB * 128);
    tma_load_2d(&a.tmA, &bars[0], sa, 0, 0);
    for (int r = 0; r < kRowsB; r += 64) tma_load_2d(&a.tmB, &bars[0], sb + r * 128, 0, r);   // 64-row boxes
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t start = smem_u32(sb) + a.shift * 128;
    uint64_t db = umma_desc_sw128(start, 16, a.mode == 0 ? 1024 : 1280);
    if (a.base_off) db |= static_cast<uint64_t>((start >> 7) & 7) << 49;
    const uint64_t da = umma_desc_sw128(smem_u32(sa), 16, 1024);
    constexpr uint32_t idesc = umma_idesc_bf16(128, 256, 0, 0);
    for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem, da + 2 * k, db + 2 * k, idesc, k != 0);
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
  const int row = warp * 32 + lane;
  for (int c = 0; c < 8; ++c) {
    uint32_t v[32];
    tmem_ld_32x32(tmem + lane_addr + c * 32, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) a.out[row * 256 + c * 32 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}
}  // namespace

extern "C" int probe_desc_shift(const void* A, const void* B, float* out, int shift, int mode, int base_off) {
  Args a;
  uint64_t dimsA[2] = {64, 128}, dimsB[2] = {64, kRowsB};
  uint64_t strides[1] = {128};
  uint32_t boxA[2] = {64, 128}, boxB[2] = {64, 64};
  if (make_tensor_map_bf16(&a.tmA, A, 2, dimsA, strides, boxA, 1)) return -1;
  if (make_tensor_map_bf16(&a.tmB, B, 2, dimsB, strides, boxB, 1)) return -2;
  a.out = out;
  a.shift = shift;
  a.mode = mode;
  a.base_off = base_off;
  const int smem = 16384 + kRowsB * 128 + 64 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<<<1, 128, smem>>>(a);
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : -3;
}

// deepv_b200 — joint (context ‖ video) attention, head_dim 64, on tcgen05/TMEM.
//
// Replaces reference model/mmdit.py:138-180 (VarlenSelfAttentionWithT5Mask) +
// the [B,1,L,L] bool mask built in mmdit.py:1414-1434.  The mask is never
// materialised: tokens are ordered (context | clip_0 | clip_1 | ...), frame ids are
// non-decreasing along the sequence, so "same sample  AND  frame(q) >= frame(k)"
// collapses to  k < kv_end[q]  AND  key_live[b][k]   (dead keys = padded text /
// masked-history tokens, sample id 0 in the reference).
//
// One CTA = (128 queries, one head, one batch row); 192 threads:
//   warp 0     TMA producer: Q once, then (K_j, V_j) pairs into a 2-stage ring
//   warp 1     MMA issuer:   S = Q K_j^T (M128 N128 K64) -> TMEM;  PV_j = P_j V_j (M128 N64 K128)
//   warps 2-5  softmax: one query row per thread; online softmax in fp32 registers,
//              P_j written as bf16 into a 128B-swizzled smem A-operand, O kept in registers.
// q, k, v are read straight out of the fused QKV activation [B][L][3*H*64] with
// strided TMA boxes; the output is token-major [B][L][H*64].
#include "common.cuh"
#include "kernels.cuh"

namespace dv {
namespace {

constexpr int kQ = 128;    // queries per CTA
constexpr int kKV = 128;   // keys per inner tile
constexpr int kD = 64;     // head dim
constexpr int kTileBytes = 128 * 128;  // 128 rows x 64 bf16
constexpr int kKvStages = 2;
constexpr int kAttnThreads = 192;
constexpr uint32_t kTmemCols = 256;  // S: 128 cols, PV: 64 cols (power-of-two allocation)
constexpr int kSmemBytes = kTileBytes /*Q*/ + kKvStages * 2 * kTileBytes /*K,V*/ +
                           2 * kTileBytes /*P: two 64-wide k-blocks*/ + 256 + 1024;

struct AttnArgs {
  alignas(64) CUtensorMap tmQKV;  // (3*H*64, L, B) bf16, box {64, 128, 1}
  __nv_bfloat16* out;             // [B][L][H*64]
  const int* kv_end;              // [L]
  const float* key_bias;          // [B][Lpad]: 0 live, -inf dead / beyond L (Lpad % 128 == 0)
  int L, Lpad, H, B;
  float scale_log2;               // head_dim^-0.5 * log2(e)
};

__global__ void __launch_bounds__(kAttnThreads, 2) attn_kernel(const __grid_constant__ AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + kTileBytes;                      // [stage][K | V]
  uint8_t* sP = sKV + kKvStages * 2 * kTileBytes;      // [kblock 0 | kblock 1]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kTileBytes);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                // [2]
  uint64_t* kv_empty = bars + 3;               // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kQ;
  const int head = blockIdx.y;
  const int b = blockIdx.z;

  // keys needed by this query tile: kv_end is non-decreasing in q
  const int q_last = min(q0 + kQ, a.L) - 1;
  const int kv_len = __ldg(a.kv_end + q_last);
  const int n_kv = (kv_len + kKV - 1) / kKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.tmQKV);
    mbar_init(q_full, 1);
    for (int i = 0; i < kKvStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_PV = tmem_base + 128;

  const int HD = a.H * kD;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, kTileBytes);
      tma_load_3d(&a.tmQKV, q_full, sQ, head * kD, q0, b);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&kv_empty[s], ph ^ 1);
        uint8_t* sK = sKV + s * 2 * kTileBytes;
        uint8_t* sV = sK + kTileBytes;
        mbar_expect_tx(&kv_full[s], 2 * kTileBytes);
        tma_load_3d(&a.tmQKV, &kv_full[s], sK, HD + head * kD, j * kKV, b);
        tma_load_3d(&a.tmQKV, &kv_full[s], sV, 2 * HD + head * kD, j * kKV, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);   // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);   // P (K-major) x V (MN-major)
      const uint32_t aQ = smem_u32(sQ);
      const uint32_t aP = smem_u32(sP);
      mbar_wait(q_full, 0);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        const uint32_t aK = smem_u32(sKV + s * 2 * kTileBytes);
        const uint32_t aV = aK + kTileBytes;
        mbar_wait(&kv_full[s], ph);
        // (the S buffer is free: p_full(j-1) was waited on in the previous iteration)
        tc_fence_after();
        {
          const uint64_t dq = umma_desc_sw128(aQ, 16, 1024);
          const uint64_t dk = umma_desc_sw128(aK, 16, 1024);
#pragma unroll
          for (int k = 0; k < kD / 16; ++k)
            umma_bf16_ss(tmem_S, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
          umma_commit(s_full);
        }
        // P_j in smem (and S_j drained from TMEM)
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        {
#pragma unroll
          for (int k = 0; k < kKV / 16; ++k) {
            // A = P: k-block (k / 4) of 16 KB, +32 B per K=16 slice inside the swizzle row
            const uint64_t dp =
                umma_desc_sw128(aP + (k >> 2) * kTileBytes + (k & 3) * 32, 16, 1024);
            // B = V, MN-major: 16 kv rows (two 8-row groups, SBO = 1024 B) per K=16 slice
            const uint64_t dv = umma_desc_sw128(aV + k * 2048, 1024, 1024);
            umma_bf16_ss(tmem_PV, dp, dv, idesc_pv, k != 0);
          }
          umma_commit(o_full);
          umma_commit(&kv_empty[s]);
        }
      }
    }
  } else {
    // ------------------------------ softmax warps --------------------------------
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // row inside the tile == TMEM lane
    const int qi = q0 + r;
    const bool row_ok = qi < a.L;
    const int kv_end = row_ok ? __ldg(a.kv_end + qi) : 0;
    const float* kb = a.key_bias + static_cast<long long>(b) * a.Lpad;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;

    float o[kD];
#pragma unroll
    for (int i = 0; i < kD; ++i) o[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;

    uint8_t* prow = sP + (r >> 3) * 1024 + (r & 7) * 128;
    const int sw = r & 7;

    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // pass 1: scores -> scale + mask -> row max (S stays in TMEM; it is re-read in
      // pass 2 instead of being held in 128 registers)
      float m_tile = -INFINITY;
#pragma unroll
      for (int c = 0; c < kKV / 32; ++c) {
        uint32_t raw[32];
        tmem_ld_32x32(tmem_S + lane_addr + c * 32, raw);
        tmem_ld_wait();
        const int k0 = j * kKV + c * 32;
        const float4* kb4 = reinterpret_cast<const float4*>(kb + k0);
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 bb = __ldg(kb4 + i4);
          const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = 4 * i4 + e;
            float x = __uint_as_float(raw[i]) * a.scale_log2 + bv[e];
            x = (k0 + i < kv_end) ? x : -INFINITY;
            m_tile = fmaxf(m_tile, x);
          }
        }
      }
      const float m_new = fmaxf(m_run, m_tile);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;  // fully-masked-so-far rows
      const float alpha = exp2f(m_run - m_use);                // m_run = -inf -> 0
      float l_tile = 0.f;
      // pass 2: exponentiate, pack to bf16, store into the swizzled A-operand tile
#pragma unroll
      for (int c = 0; c < kKV / 32; ++c) {
        uint32_t raw[32];
        tmem_ld_32x32(tmem_S + lane_addr + c * 32, raw);
        tmem_ld_wait();
        const int k0 = j * kKV + c * 32;
        const float4* kb4 = reinterpret_cast<const float4*>(kb + k0);
        float p[32];
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 bb = __ldg(kb4 + i4);
          const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = 4 * i4 + e;
            float x = __uint_as_float(raw[i]) * a.scale_log2 + bv[e];
            x = (k0 + i < kv_end) ? x : -INFINITY;
            p[i] = exp2f(x - m_use);
            l_tile += p[i];
          }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {  // 16-byte chunks of 8 keys
          uint4 q;
          q.x = pack_bf16x2(p[8 * g + 0], p[8 * g + 1]);
          q.y = pack_bf16x2(p[8 * g + 2], p[8 * g + 3]);
          q.z = pack_bf16x2(p[8 * g + 4], p[8 * g + 5]);
          q.w = pack_bf16x2(p[8 * g + 6], p[8 * g + 7]);
          const int c16 = c * 4 + g;
          const int kblk = c16 >> 3;  // which 64-key k-block
          const int chunk = c16 & 7;  // 16 B chunk inside the 128 B row
          *reinterpret_cast<uint4*>(prow + kblk * kTileBytes + ((chunk ^ sw) << 4)) = q;
        }
      }
      l_run = l_run * alpha + l_tile;
      m_run = m_new;
      // make the generic-proxy smem writes visible to the tensor-core (async) proxy
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);

      // O <- O * alpha + P_j V_j
      mbar_wait(o_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < kD / 32; ++c) {
        uint32_t raw[32];
        tmem_ld_32x32(tmem_PV + lane_addr + c * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c * 32 + i] = o[c * 32 + i] * alpha + __uint_as_float(raw[i]);
      }
      tc_fence_before();
    }

    if (row_ok) {
      const float inv = (l_run > 0.f) ? 1.0f / l_run : 0.f;
      __nv_bfloat16* dst = a.out + (static_cast<long long>(b) * a.L + qi) * HD + head * kD;
      uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint4 q;
        q.x = pack_bf16x2(o[8 * i + 0] * inv, o[8 * i + 1] * inv);
        q.y = pack_bf16x2(o[8 * i + 2] * inv, o[8 * i + 3] * inv);
        q.z = pack_bf16x2(o[8 * i + 4] * inv, o[8 * i + 5] * inv);
        q.w = pack_bf16x2(o[8 * i + 6] * inv, o[8 * i + 7] * inv);
        d4[i] = q;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

}  // namespace

int launch_attention(const void* qkv, void* out, const int* kv_end, const float* key_bias, int B,
                     int L, int Lpad, int H, cudaStream_t stream, double flops) {
  DV_REQUIRE(Lpad % 128 == 0 && Lpad >= L, "attention: Lpad=%d must be a multiple of 128 >= L=%d",
             Lpad, L);
  DV_REQUIRE(B > 0 && L > 0 && H > 0, "attention: empty problem B=%d L=%d H=%d", B, L, H);
  AttnArgs a;
  uint64_t dims[3] = {(uint64_t)(3 * H * kD), (uint64_t)L, (uint64_t)B};
  uint64_t strides[2] = {(uint64_t)(3 * H * kD) * 2, (uint64_t)(3 * H * kD) * 2 * (uint64_t)L};
  uint32_t box[3] = {64, 128, 1};
  int rc = make_tensor_map_bf16(&a.tmQKV, qkv, 3, dims, strides, box, 1);
  if (rc) return rc;
  a.out = reinterpret_cast<__nv_bfloat16*>(out);
  a.kv_end = kv_end;
  a.key_bias = key_bias;
  a.L = L;
  a.Lpad = Lpad;
  a.H = H;
  a.B = B;
  a.scale_log2 = 0.125f * 1.4426950408889634f;
  static bool attr_set = false;
  if (!attr_set) {
    DV_CHECK_CUDA(cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kSmemBytes));
    attr_set = true;
  }
  dim3 grid((L + kQ - 1) / kQ, H, B);
  const int pid = prof_begin(PROF_ATTN, flops, 2.0 * B * L * 4.0 * H * kD, stream);
  attn_kernel<<<grid, kAttnThreads, kSmemBytes, stream>>>(a);
  prof_end(pid, stream);
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

}  // namespace dv

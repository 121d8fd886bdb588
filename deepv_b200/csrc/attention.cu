// deepv_b200 — joint (context ‖ video) attention, head_dim 64, on tcgen05/TMEM.
//
// Replaces reference model/mmdit.py:138-180 (VarlenSelfAttentionWithT5Mask) + the [B,1,L,L] bool
// mask built in mmdit.py:1414-1434.  The mask is never materialised: tokens are ordered
// (context | clip_0 | clip_1 | ...), frame ids are non-decreasing along the sequence, so
// "same sample AND frame(q) >= frame(k)" collapses to  k < kv_end[q]  AND  key_bias[b][k] == 0
// (dead keys = padded text / masked-history tokens, sample id 0 in the reference).
//
// One CTA = 256 queries (two 128-row tiles) x one head x one batch row; 320 threads:
//   warp 0      TMA producer: Q0,Q1 once, then (K_j, V_j) into a 3-stage ring
//   warp 1      MMA issuer: S_w = Q_w K_j^T (M128 N128 K64) into TMEM, O_w += P_w V_j (M128 N64
//               K128, V as the MN-major B operand straight from the token-major qkv rows)
//   warps 2-5   softmax group 0 (query tile 0), one row per thread
//   warps 6-9   softmax group 1 (query tile 1)
// The two groups ping-pong: while group 0 exponentiates S_0(j), the tensor pipe runs
// O_1 += P_1 V_(j-1) and S_1(j).  O lives in TMEM for the whole KV sweep; the running max is
// only refreshed (and O rescaled in TMEM) when it grows by more than 2^8, so the rescale is
// rare after the first tiles.  Masking is applied only on tiles that need it (frame boundary
// or dead keys).
#include <cstdio>

#include "common.cuh"
#include "kernels.cuh"

namespace dv {
namespace {

constexpr int kTile = 128;             // query rows per group, keys per inner tile
constexpr int kD = 64;                 // head dim
constexpr int kTileBytes = 128 * 128;  // 128 rows x 64 bf16
constexpr int kKvStages = 3;
constexpr int kAttnThreads = 320;
constexpr uint32_t kTmemCols = 512;    // S0 | S1 (128 each) | O0 | O1 (64 each)
constexpr int kSmemBytes = 2 * kTileBytes /*Q0,Q1*/ + kKvStages * 2 * kTileBytes /*K,V*/ +
                           2 * 2 * kTileBytes /*P0,P1: two 64-key k-blocks each*/ + 256 + 1024;
constexpr float kRescaleThreshold = 8.0f;  // log2 domain

struct AttnArgs {
  alignas(64) CUtensorMap tmQKV;  // (3*H*64, L, B) bf16, box {64, 128, 1}
  __nv_bfloat16* out;             // [B][L][H*64]
  const int* kv_end;              // [L]
  const float* key_bias;          // [B][Lpad]: 0 live, -inf dead / beyond L (Lpad % 128 == 0)
  const int* tile_dead;           // [B][Lpad/128]: tile holds a dead key (nullptr: assume yes)
  int L, Lpad, H, B;
  float scale_log2;               // head_dim^-0.5 * log2(e)
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
        "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]),
        "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(kAttnThreads, 1) attn_kernel(const __grid_constant__ AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                                  // [2][16 KB]
  uint8_t* sKV = sQ + 2 * kTileBytes;                  // [stage][K | V]
  uint8_t* sP = sKV + kKvStages * 2 * kTileBytes;      // [group][kblock 0 | kblock 1]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * kTileBytes);
  uint64_t* q_full = bars;              // 1
  uint64_t* kv_full = bars + 1;         // [3]
  uint64_t* kv_empty = bars + 4;        // [3]
  uint64_t* s_full = bars + 7;          // [2]
  uint64_t* p_full = bars + 9;          // [2]
  uint64_t* o_done = bars + 11;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 2 * kTile;
  const int head = blockIdx.y;
  const int b = blockIdx.z;

  // keys needed by each query tile (kv_end is non-decreasing in q)
  int n_kv_w[2];
#pragma unroll
  for (int w = 0; w < 2; ++w) {
    const int first = q0 + w * kTile;
    if (first >= a.L) {
      n_kv_w[w] = 0;
    } else {
      const int last = min(first + kTile, a.L) - 1;
      n_kv_w[w] = (__ldg(a.kv_end + last) + kTile - 1) / kTile;
    }
  }
  const int n_kv = max(n_kv_w[0], n_kv_w[1]);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.tmQKV);
    mbar_init(q_full, 1);
    for (int i = 0; i < kKvStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int w = 0; w < 2; ++w) {
      mbar_init(&s_full[w], 1);
      mbar_init(&p_full[w], 128);
      mbar_init(&o_done[w], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int HD = a.H * kD;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * kTileBytes);
      tma_load_3d(&a.tmQKV, q_full, sQ, head * kD, q0, b);
      tma_load_3d(&a.tmQKV, q_full, sQ + kTileBytes, head * kD, q0 + kTile, b);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % kKvStages;
        const uint32_t ph = (j / kKvStages) & 1;
        mbar_wait(&kv_empty[s], ph ^ 1);
        uint8_t* sK = sKV + s * 2 * kTileBytes;
        mbar_expect_tx(&kv_full[s], 2 * kTileBytes);
        tma_load_3d(&a.tmQKV, &kv_full[s], sK, HD + head * kD, j * kTile, b);
        tma_load_3d(&a.tmQKV, &kv_full[s], sK + kTileBytes, 2 * HD + head * kD, j * kTile, b);
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (lane == 0 && n_kv > 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);   // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);   // P (K-major) x V (MN-major)
      auto issue_s = [&](int w, int j) {
        const uint32_t aQ = smem_u32(sQ + w * kTileBytes);
        const uint32_t aK = smem_u32(sKV + (j % kKvStages) * 2 * kTileBytes);
        const uint64_t dq = umma_desc_sw128(aQ, 16, 1024);
        const uint64_t dk = umma_desc_sw128(aK, 16, 1024);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k)
          umma_bf16_ss(tmem_base + w * 128, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
        umma_commit(&s_full[w]);
      };
      auto issue_pv = [&](int w, int j) {
        const uint32_t aP = smem_u32(sP + w * 2 * kTileBytes);
        const uint32_t aV = smem_u32(sKV + (j % kKvStages) * 2 * kTileBytes) + kTileBytes;
#pragma unroll
        for (int k = 0; k < kTile / 16; ++k) {
          // A = P: k-block (k / 4) of 16 KB, +32 B per K=16 slice inside the swizzle row;
          // B = V, MN-major: 16 kv rows (two 8-row groups, SBO = 1024 B) per K=16 slice
          const uint64_t dp = umma_desc_sw128(aP + (k >> 2) * kTileBytes + (k & 3) * 32, 16, 1024);
          const uint64_t dv = umma_desc_sw128(aV + k * 2048, 1024, 1024);
          umma_bf16_ss(tmem_base + 256 + w * 64, dp, dv, idesc_pv, (j | k) != 0);
        }
      };
      mbar_wait(q_full, 0);
      int kv_waited = -1;  // highest KV tile whose kv_full has been observed
      auto need_kv = [&](int j) {
        if (j > kv_waited) {
          mbar_wait(&kv_full[j % kKvStages], (j / kKvStages) & 1);
          kv_waited = j;
        }
      };
      need_kv(0);
      tc_fence_after();
      for (int w = 0; w < 2; ++w)
        if (n_kv_w[w] > 0) issue_s(w, 0);
      for (int j = 0; j < n_kv; ++j) {
        for (int w = 0; w < 2; ++w) {
          if (j >= n_kv_w[w]) continue;
          mbar_wait(&p_full[w], j & 1);  // P_w(j) in smem, S_w(j) drained, O_w rescaled
          tc_fence_after();
          issue_pv(w, j);
          if (j + 1 < n_kv_w[w]) {
            need_kv(j + 1);
            tc_fence_after();
            issue_s(w, j + 1);
          } else {
            umma_commit(&o_done[w]);
          }
        }
        umma_commit(&kv_empty[j % kKvStages]);
      }
    }
  } else {
    // ================================ softmax groups ==============================
    const int w = (warp - 2) >> 2;       // group / query tile
    const int quarter = warp & 3;        // TMEM lane quarter this warp may touch
    const int r = quarter * 32 + lane;   // row inside the tile == TMEM lane
    const int qi = q0 + w * kTile + r;
    const bool row_ok = qi < a.L;
    const int my_n_kv = n_kv_w[w];
    if (my_n_kv > 0) {
      const int kv_end = row_ok ? __ldg(a.kv_end + qi) : 0;
      const int kv_end_min = __ldg(a.kv_end + q0 + w * kTile);  // first row of the tile
      const float* kb = a.key_bias + static_cast<long long>(b) * a.Lpad;
      const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
      const uint32_t tS = tmem_base + w * 128 + lane_addr;
      const uint32_t tO = tmem_base + 256 + w * 64 + lane_addr;
      uint8_t* prow = sP + w * 2 * kTileBytes + (r >> 3) * 1024 + (r & 7) * 128;
      const int sw = r & 7;
      float m_ref = -INFINITY, l_run = 0.f;

      for (int j = 0; j < my_n_kv; ++j) {
        mbar_wait(&s_full[w], j & 1);
        tc_fence_after();
        float s[kTile];
#pragma unroll
        for (int c = 0; c < kTile / 32; ++c) {
          uint32_t raw[32];
          tmem_ld_32x32(tS + c * 32, raw);
#pragma unroll
          for (int i = 0; i < 32; ++i) s[c * 32 + i] = __uint_as_float(raw[i]);
        }
        tmem_ld_wait();
        const int k0 = j * kTile;
        const bool need_mask = (k0 + kTile > kv_end_min) ||
                               (a.tile_dead == nullptr) ||
                               (__ldg(a.tile_dead + b * (a.Lpad / kTile) + j) != 0);
        float m_tile = -INFINITY;
        if (need_mask) {
#pragma unroll
          for (int i4 = 0; i4 < kTile / 4; ++i4) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(kb + k0) + i4);
            const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = 4 * i4 + e;
              float x = s[i] * a.scale_log2 + bv[e];
              x = (k0 + i < kv_end) ? x : -INFINITY;
              s[i] = x;
              m_tile = fmaxf(m_tile, x);
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < kTile; ++i) {
            s[i] *= a.scale_log2;
            m_tile = fmaxf(m_tile, s[i]);
          }
        }
        // running-max policy: keep a stale reference unless the new max exceeds it by > 2^8
        const bool had = m_ref > -INFINITY;
        const bool grow = had ? (m_tile > m_ref + kRescaleThreshold) : (m_tile > -INFINITY);
        const float m_new = grow ? m_tile : m_ref;
        if (j > 0 && __any_sync(0xffffffffu, grow && had)) {
          // rescale O_w (in TMEM) and l by 2^(m_ref - m_new); rows that do not grow use 1
          const float alpha = (grow && had) ? ex2(m_ref - m_new) : 1.0f;
#pragma unroll
          for (int c = 0; c < kD / 32; ++c) {
            uint32_t o[32];
            tmem_ld_32x32(tO + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32(tO + c * 32, o);
          }
          tmem_st_wait();
          l_run *= alpha;
        }
        m_ref = m_new;
        const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;  // fully-masked-so-far rows -> p = 0
        float l_tile = 0.f;
#pragma unroll
        for (int g = 0; g < kTile / 8; ++g) {  // 16-byte chunks of 8 keys
          float p[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            p[i] = ex2(s[g * 8 + i] - m_use);
            l_tile += p[i];
          }
          uint4 q;
          q.x = pack_bf16x2(p[0], p[1]);
          q.y = pack_bf16x2(p[2], p[3]);
          q.z = pack_bf16x2(p[4], p[5]);
          q.w = pack_bf16x2(p[6], p[7]);
          const int kblk = g >> 3;   // which 64-key k-block
          const int chunk = g & 7;   // 16 B chunk inside the 128 B row
          *reinterpret_cast<uint4*>(prow + kblk * kTileBytes + ((chunk ^ sw) << 4)) = q;
        }
        l_run += l_tile;
        fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor-core proxy
        tc_fence_before();
        mbar_arrive(&p_full[w]);
      }

      // epilogue: O_w / l -> bf16, token-major store
      mbar_wait(&o_done[w], 0);
      tc_fence_after();
      const float inv = (l_run > 0.f) ? 1.0f / l_run : 0.f;
      float o[kD];
#pragma unroll
      for (int c = 0; c < kD / 32; ++c) {
        uint32_t raw[32];
        tmem_ld_32x32(tO + c * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c * 32 + i] = __uint_as_float(raw[i]) * inv;
      }
      if (row_ok) {
        uint4* d4 = reinterpret_cast<uint4*>(a.out + (static_cast<long long>(b) * a.L + qi) * HD +
                                             head * kD);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint4 q;
          q.x = pack_bf16x2(o[8 * i + 0], o[8 * i + 1]);
          q.y = pack_bf16x2(o[8 * i + 2], o[8 * i + 3]);
          q.z = pack_bf16x2(o[8 * i + 4], o[8 * i + 5]);
          q.w = pack_bf16x2(o[8 * i + 6], o[8 * i + 7]);
          d4[i] = q;
        }
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

int launch_attention(const void* qkv, void* out, const int* kv_end, const float* key_bias,
                     const int* tile_dead, int B, int L, int Lpad, int H, cudaStream_t stream,
                     double flops) {
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
  a.tile_dead = tile_dead;
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
  dim3 grid((L + 2 * kTile - 1) / (2 * kTile), H, B);
  char tag[56] = "";
  if (prof_on()) snprintf(tag, sizeof(tag), "attn B%d L%d H%d", B, L, H);
  const int pid = prof_begin(PROF_ATTN, flops, 2.0 * B * L * 4.0 * H * kD, stream, tag);
  attn_kernel<<<grid, kAttnThreads, kSmemBytes, stream>>>(a);
  prof_end(pid, stream);
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

}  // namespace dv

// deepv_b200 — joint (context ‖ video) attention, head_dim 64, on tcgen05/TMEM.
//
// Replaces reference model/mmdit.py:138-180 (VarlenSelfAttentionWithT5Mask) + the [B,1,L,L] bool
// mask built in mmdit.py:1414-1434.  The mask is never materialised: tokens are ordered
// (context | clip_0 | clip_1 | ...), frame ids are non-decreasing along the sequence, so
// "same sample AND frame(q) >= frame(k)" collapses to  k < kv_end[q]  AND  key_bias[b][k] == 0
// (dead keys = padded text / masked-history tokens, sample id 0 in the reference).
//
// Four kernels share the scheme (DV_ATTN_PIPE selects; launch_attention at the end of the file):
//   attn_pipe_kernel<64>   DEFAULT.  CTA = 128 queries x one head x one batch row, 192 threads, 64-key tiles, S double-
//                          buffered in tensor memory, TWO CTAs per SM (256 TMEM columns, 82 KB shared memory, 158 registers)
//   attn_pipe_kernel<128>  the same with 128-key tiles: 512 TMEM columns, one CTA per SM
//   attn_pair_kernel       two query tiles per CTA sharing K / V, one softmax warpgroup each, setmaxnreg 224 / 56
//   attn_kernel            round 1: S overwritten by P, row sums on the tensor pipe, two CTAs per SM at the 168-register cap
// Roles in a CTA:
//   TMA warp    Q once, then (K_j, V_j) into a ring; MMA warp: both run their loops as whole converged warps and elect
//               the issuing lane inside the instruction (common.cuh w_* forms) — issued from `if (lane == 0)` every
//               tcgen05.mma / TMA costs ~80 cycles of ELECT / R2UR.BROADCAST loop, which paced the round-1 kernel.
//               Every MMA takes its A operand FROM TENSOR MEMORY: measured on B200 an M128 MMA with A in shared memory
//               costs >= 128 cycles whatever N is (the 128-row A read), which would hold S = Q K^T to half rate and
//               P V (N = 64) to a quarter; with A in TMEM they run at their math rate.
//                 S  = Q K_j^T            (M128 N=KT K64)   A = Q (bf16, parked in TMEM once)
//                 O += P V_j              (M128 N64 K=KT)   A = P (bf16, written over S by the softmax warps), V MN-major
//   4 softmax warps, one query row per thread (softmax_tile): S from TMEM, four independent max chains, running max with
//               a lazy rescale (O is rescaled in TMEM only when the max grows by more than 2^8), scale-and-shift / row
//               sums as packed fp32x2, one exponential pair in four as a polynomial on the FMA pipe, P back into TMEM.
//               Masking runs only on tiles that hold a frame boundary or dead keys.
// Query tiles are issued heaviest-first (late frames see the most keys).  Measurements: DESIGN.md §3.3,
// profiles/r02p_attn_modes.txt, scripts/probe/attn_trace.py (-DDV_ATTN_TRACE timeline build).
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "kernels.cuh"

namespace dv {
namespace {

constexpr int kTile = 128;             // query rows per CTA, keys per inner tile
constexpr int kD = 64;                 // head dim
constexpr int kTileBytes = 128 * 128;  // 128 rows x 64 bf16
constexpr int kKvStages = 2;
constexpr int kAttnThreads = 192;
constexpr int kOnesBytes = 2048;       // 16 rows x 128 B of bf16 ones (B operand of the row-sum MMA)
constexpr uint32_t kTmemCols = 256;
// TMEM column map
constexpr uint32_t kColS = 0;     // S fp32 [128]; P bf16x2 aliases [0, 64)
constexpr uint32_t kColO = 128;   // O fp32 [64]
constexpr uint32_t kColL = 192;   // row sums fp32 [16] (all columns equal)
constexpr uint32_t kColQ = 208;   // Q bf16x2 [32]
constexpr int kSmemBytes = kTileBytes /*Q*/ + kKvStages * 2 * kTileBytes /*K,V*/ + kOnesBytes + 256 + 1024;
constexpr float kRescaleThreshold = 8.0f;  // log2 domain
#ifndef DV_ATTN_POLY
#define DV_ATTN_POLY 0
#endif
constexpr int kPolyPer8 = DV_ATTN_POLY;    // exponentials per 8 computed on the FMA pipe

struct AttnArgs {
  alignas(64) CUtensorMap tmQKV;  // (3*H*64, L, B) bf16, box {64, 128, 1}
  alignas(64) CUtensorMap tmKV64; // the same tensor, box {64, 64, 1}: K / V tiles of the 64-key kernel
  __nv_bfloat16* out;             // [B][L][H*64]
  const int* kv_end;              // [L]
  const float* key_bias;          // [B][Lpad]: 0 live, -inf dead / beyond L (Lpad % 128 == 0)
  const int* tile_dead;           // [B][Lpad/128]: tile holds a dead key (nullptr: assume yes)
  int L, Lpad, H, B;              // H = heads in the qkv / out rows
  int head0;                      // first head this launch handles (Ulysses: a rank's head slice)
  // Ulysses over peer memory: context rows (< peer_Lc) go to every rank's buffer, video row q to
  // out_peer[(q - peer_Lc) / peer_Lw]; n_peers == 0: everything to `out`
  __nv_bfloat16* out_peer[8];
  int n_peers, peer_Lc, peer_Lw;
  float scale_log2;               // head_dim^-0.5 * log2(e)
  int pdl_early;
#ifdef DV_ATTN_TRACE
  long long* trace;               // probe build only (scripts/probe/attn_trace.py): clock64 stamps of CTA (0,0,0)
#endif
};

// probe build: stamp (role, tile j, event k) of the heaviest CTA; role 0/1 = first warp of softmax group 0/1, 2/3 = the
// MMA thread's work for group 0/1
#ifdef DV_ATTN_TRACE
#define DV_TR(on, role, j, k)                                                              \
  do {                                                                                     \
    if ((on) && (j) < 32) a.trace[((role) * 32 + (j)) * 8 + (k)] = clock64();              \
  } while (0)
// per-CTA record (pipe kernel): slot k of CTA `cta` <- globaltimer; slot 0 holds the SM id
#define DV_TC(on, cta, k)                                                                                     \
  do {                                                                                                        \
    if (on) a.trace[1024 + (cta) * 8 + (k)] = static_cast<long long>(global_ns());                            \
  } while (0)
#else
#define DV_TR(on, role, j, k) \
  do {                        \
  } while (0)
#define DV_TC(on, cta, k) \
  do {                    \
  } while (0)
#endif

// 2^x on the FMA pipe: round-to-nearest split x = n + f, degree-3 minimax of 2^f on [-0.5, 0.5]
// (relative error 2e-4, below the bf16 rounding of P), exponent added as integer bits.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float magic = 12582912.0f;  // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float xr = x + magic;
  const float f = x - (xr - magic);
  float p = fmaf(f, 0.053027521818876266f, 0.24221394956111908f);
  p = fmaf(p, f, 0.6935725808143616f);
  p = fmaf(p, f, 0.9999590516090393f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(xr) << 23));
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
        "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]),
        "r"(r[14]), "r"(r[15])
      : "memory");
}

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
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
// one 32-bit column of this thread's lane
__device__ __forceinline__ uint32_t tmem_ld_1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
  return v;
}
__device__ __forceinline__ void tmem_st_1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}

__global__ void __launch_bounds__(kAttnThreads, 2) attn_kernel(const __grid_constant__ AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                                  // [16 KB]
  uint8_t* sKV = sQ + kTileBytes;                      // [stage][K | V]
  uint8_t* sOnes = sKV + kKvStages * 2 * kTileBytes;   // 2 KB of bf16 1.0
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + kOnesBytes);
  uint64_t* q_full = bars;                       // TMA: Q landed in smem
  uint64_t* q_ready = bars + 1;                  // softmax warps: Q parked in TMEM (128 arrivals)
  uint64_t* kv_full = bars + 2;                  // [kKvStages]
  uint64_t* kv_empty = kv_full + kKvStages;      // [kKvStages]
  uint64_t* s_full = kv_empty + kKvStages;       // S_j ready (and P_{j-1} V consumed)
  uint64_t* p_full = s_full + 1;                 // P_j in TMEM (128 arrivals)
  uint64_t* o_done = p_full + 1;                 // last P V retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // heaviest query tiles first: later frames see more keys
  // longest work first, chip-wide: the query tile is the slowest grid index and runs from the last tile (most
  // key tiles under the frame-causal mask) to the first, so the short items fill the tail of the launch
  const int qt = static_cast<int>(gridDim.z) - 1 - static_cast<int>(blockIdx.z);
  const int q0 = qt * kTile;
  const int head = a.head0 + static_cast<int>(blockIdx.x);
  const int b = blockIdx.y;

  // keys needed by this query tile (kv_end is non-decreasing in q)
  const int last_q = min(q0 + kTile, a.L) - 1;
  const int n_kv = (__ldg(a.kv_end + last_q) + kTile - 1) / kTile;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.tmQKV);
    mbar_init(q_full, 1);
    mbar_init(q_ready, 128);
    for (int i = 0; i < kKvStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  // the tile of ones read by the row-sum MMA (generic-proxy writes -> async proxy)
  for (int i = threadIdx.x; i < kOnesBytes / 4; i += kAttnThreads)
    reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3F803F80u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int HD = a.H * kD;
  pdl_wait();  // (kv_end is plan data written long before; everything above overlapped the predecessor)
  if (a.pdl_early) pdl_trigger();

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      mbar_expect_tx(q_full, kTileBytes);
      tma_load_3d(&a.tmQKV, q_full, sQ, head * kD, q0, b);
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
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);   // Q (TMEM) x K (K-major)
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);   // P (TMEM) x V (MN-major)
      constexpr uint32_t idesc_l = umma_idesc_bf16(128, 16, 0, 0);    // P (TMEM) x ones (K-major)
      const uint32_t tS = tmem_base + kColS, tO = tmem_base + kColO, tL = tmem_base + kColL;
      const uint32_t tQ = tmem_base + kColQ;
      const uint64_t d_ones = umma_desc_sw128(smem_u32(sOnes), 16, 1024);
      auto issue_s = [&](int j) {
        const uint32_t aK = smem_u32(sKV + (j % kKvStages) * 2 * kTileBytes);
        const uint64_t dk = umma_desc_sw128(aK, 16, 1024);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) umma_bf16_ts(tS, tQ + k * 8, dk + 2 * k, idesc_s, k != 0);
        umma_commit(s_full);
      };
      mbar_wait(q_ready, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(p_full, j & 1);  // P_j in TMEM, S_j drained, O / l rescaled
        tc_fence_after();
        const uint32_t aV = smem_u32(sKV + (j % kKvStages) * 2 * kTileBytes) + kTileBytes;
#pragma unroll
        for (int k = 0; k < kTile / 16; ++k) {
          // A = P in TMEM: 16 keys = 8 packed columns per K=16 slice;
          // B = V, MN-major: 16 kv rows (two 8-row groups, SBO = 1024 B) per K=16 slice
          const uint64_t dv = umma_desc_sw128(aV + k * 2048, 1024, 1024);
          umma_bf16_ts(tO, tS + k * 8, dv, idesc_pv, (j | k) != 0);
          umma_bf16_ts(tL, tS + k * 8, d_ones, idesc_l, (j | k) != 0);
        }
        umma_commit(&kv_empty[j % kKvStages]);
        if (j + 1 < n_kv) {
          mbar_wait(&kv_full[(j + 1) % kKvStages], ((j + 1) / kKvStages) & 1);
          tc_fence_after();
          issue_s(j + 1);  // in order behind P_j V: S may overwrite P
        } else {
          umma_commit(o_done);
        }
      }
    }
  } else {
    // ================================ softmax =====================================
    const int quarter = warp & 3;        // TMEM lane quarter this warp may touch
    const int r = quarter * 32 + lane;   // row inside the tile == TMEM lane
    const int qi = q0 + r;
    const bool row_ok = qi < a.L;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem_base + kColS + lane_addr;
    const uint32_t tO = tmem_base + kColO + lane_addr;
    const uint32_t tL = tmem_base + kColL + lane_addr;
    if (n_kv > 0) {
      // park this row of Q (128 B, 128B-swizzled by TMA) in TMEM as the A operand of S = Q K^T
      mbar_wait(q_full, 0);
      {
        uint32_t qr[32];
        const uint8_t* row = sQ + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 t = *reinterpret_cast<const uint4*>(row + ((c ^ (r & 7)) << 4));
          qr[4 * c + 0] = t.x;
          qr[4 * c + 1] = t.y;
          qr[4 * c + 2] = t.z;
          qr[4 * c + 3] = t.w;
        }
        tmem_st_32x32(tmem_base + kColQ + lane_addr, qr);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(q_ready);
      }
      const int kv_end = row_ok ? __ldg(a.kv_end + qi) : 0;
      const int kv_end_min = __ldg(a.kv_end + q0);  // first row of the tile
      const float* kb = a.key_bias + static_cast<long long>(b) * a.Lpad;
      const float sc = a.scale_log2;
      float m_ref = -INFINITY;  // running max of the scaled scores (log2 domain)

      const int* dead_row = a.tile_dead ? a.tile_dead + b * (a.Lpad / kTile) : nullptr;
      for (int j = 0; j < n_kv; ++j) {
        const int k0 = j * kTile;
        // (issued before the wait: the flag load overlaps the MMA)
        const bool need_mask = (k0 + kTile > kv_end_min) || (dead_row == nullptr) || (__ldg(dead_row + j) != 0);
        mbar_wait(s_full, j & 1);
        tc_fence_after();
        // One pass over S: all 128 scores of the row come out of TMEM with a single wait (a
        // tcgen05.ld queues behind the MMAs in flight — measured ~600 cycles per round trip with
        // two CTAs per SM — so the number of round trips per tile is what matters).
        float s[kTile];
#pragma unroll
        for (int c = 0; c < kTile / 32; ++c) {
          uint32_t raw[32];
          tmem_ld_32x32(tS + c * 32, raw);
#pragma unroll
          for (int i = 0; i < 32; ++i) s[c * 32 + i] = __uint_as_float(raw[i]);
        }
        tmem_ld_wait();
        float m_tile = -INFINITY;
        if (need_mask) {  // frame boundary / dead keys in this tile
#pragma unroll
          for (int i4 = 0; i4 < kTile / 4; ++i4) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(kb + k0) + i4);
            const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = 4 * i4 + e;
              float x = s[i] + bv[e];  // -inf for dead keys
              x = (k0 + i < kv_end) ? x : -INFINITY;
              s[i] = x;
              m_tile = fmaxf(m_tile, x);
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < kTile; ++i) m_tile = fmaxf(m_tile, s[i]);
        }
        m_tile *= sc;  // sc > 0: max commutes with the scale
        // running-max policy: keep a stale reference unless the new max exceeds it by > 2^8
        const bool had = m_ref > -INFINITY;
        const bool grow = had ? (m_tile > m_ref + kRescaleThreshold) : (m_tile > -INFINITY);
        const float m_new = grow ? m_tile : m_ref;
        if (j > 0 && __any_sync(0xffffffffu, grow && had)) {
          // rescale O and l (in TMEM) by 2^(m_ref - m_new); rows that do not grow use 1
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
          const uint32_t lv = tmem_ld_1(tL);
          tmem_ld_wait();
          tmem_st_1(tL, __float_as_uint(__uint_as_float(lv) * alpha));
          tmem_st_wait();
        }
        m_ref = m_new;
        const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;  // fully-masked-so-far rows -> p = 0
        // P = 2^(s*scale - m) as bf16, two keys per 32-bit column, written back into TMEM over the
        // first 64 columns of S: the A operand of the P V and row-sum MMAs.  The XU pipe (MUFU, 16
        // lanes per SM on B200) is the busiest unit of this kernel (ncu: 74 %), so kPolyPer8 of
        // every 8 exponentials run as a polynomial on the FMA pipe instead.
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
          uint32_t pk[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float x0 = fmaf(s[hlf * 64 + 2 * i], sc, neg_m);
            const float x1 = fmaf(s[hlf * 64 + 2 * i + 1], sc, neg_m);
            const float p0 = (((2 * i) & 7) < kPolyPer8) ? ex2_poly(x0) : ex2(x0);
            const float p1 = (((2 * i + 1) & 7) < kPolyPer8) ? ex2_poly(x1) : ex2(x1);
            pk[i] = pack_bf16x2(p0, p1);
          }
          tmem_st_32x32(tS + hlf * 32, pk);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(p_full);
      }

      // epilogue: O / l -> bf16, token-major store
      mbar_wait(o_done, 0);
      tc_fence_after();
      const float l_run = __uint_as_float(tmem_ld_1(tL));
      float o[kD];
#pragma unroll
      for (int c = 0; c < kD / 32; ++c) {
        uint32_t raw[32];
        tmem_ld_32x32(tO + c * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c * 32 + i] = __uint_as_float(raw[i]);
      }
      const float inv = (l_run > 0.f) ? 1.0f / l_run : 0.f;
      if (row_ok) {
        uint4 q[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          q[i].x = pack_bf16x2(o[8 * i + 0] * inv, o[8 * i + 1] * inv);
          q[i].y = pack_bf16x2(o[8 * i + 2] * inv, o[8 * i + 3] * inv);
          q[i].z = pack_bf16x2(o[8 * i + 4] * inv, o[8 * i + 5] * inv);
          q[i].w = pack_bf16x2(o[8 * i + 6] * inv, o[8 * i + 7] * inv);
        }
        const long long off = (static_cast<long long>(b) * a.L + qi) * HD + head * kD;
        if (a.n_peers == 0) {
          uint4* d4 = reinterpret_cast<uint4*>(a.out + off);
#pragma unroll
          for (int i = 0; i < 8; ++i) d4[i] = q[i];
        } else if (qi >= a.peer_Lc) {  // a video row: straight into its owner's buffer (NVLink store)
          uint4* d4 = reinterpret_cast<uint4*>(a.out_peer[(qi - a.peer_Lc) / a.peer_Lw] + off);
#pragma unroll
          for (int i = 0; i < 8; ++i) d4[i] = q[i];
        } else {                       // a context row: every rank continues the replicated stream
          for (int p = 0; p < a.n_peers; ++p) {
            uint4* d4 = reinterpret_cast<uint4*>(a.out_peer[p] + off);
#pragma unroll
            for (int i = 0; i < 8; ++i) d4[i] = q[i];
          }
        }
      }
    }
  }

  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ---- one key tile of one query row (shared by attn_pipe_kernel and attn_pair_kernel) ------------------------------
// kPolyPairs of every 4 (pairs of) exponentials run on the FMA pipe (ex2_poly2) instead of the XU pipe: with one row per
// thread the 128 MUFU.EX2 of a tile hold the XU pipe (4 lanes / clock / SM sub-partition) for 1024 cycles, which is
// what bounds the tile once the MMA issue is out of the way (scripts/probe/attn_trace.py: exp + pack 1240 of 1900
// cycles).  Scale-and-shift, the polynomial and the row sum use the packed fp32x2 forms (FFMA2 / FADD2).
#ifndef DV_ATTN_POLY_PAIRS
#define DV_ATTN_POLY_PAIRS 1
#endif
constexpr int kPolyPairs = DV_ATTN_POLY_PAIRS;

__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  x.x = fmaxf(x.x, -125.0f);
  x.y = fmaxf(x.y, -125.0f);
  const float2 magic = make_float2(12582912.0f, 12582912.0f);   // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float2 xr = __fadd2_rn(x, magic);
  const float2 t = __fadd2_rn(xr, make_float2(-12582912.0f, -12582912.0f));
  const float2 f = __ffma2_rn(t, make_float2(-1.0f, -1.0f), x);
  float2 p = __ffma2_rn(f, make_float2(0.053027521818876266f, 0.053027521818876266f),
                        make_float2(0.24221394956111908f, 0.24221394956111908f));
  p = __ffma2_rn(p, f, make_float2(0.6935725808143616f, 0.6935725808143616f));
  p = __ffma2_rn(p, f, make_float2(0.9999590516090393f, 0.9999590516090393f));
  return make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(xr.x) << 23)),
                     __int_as_float(__float_as_int(p.y) + (__float_as_int(xr.y) << 23)));
}

struct RowState {
  float m_ref;   // running maximum the stored exponentials are relative to (log2 domain, lazily raised)
  float l;       // running row sum
};

// S (fp32, this thread's KT columns at tS) -> P = 2^(s * sc - m) as bf16 over S[0, KT / 2); keeps the running maximum and
// row sum in `st`; rescales O (64 fp32 columns at tO) when the maximum grew by more than 2^kRescaleThreshold — only
// after `pv_bar` says the previous P V has retired (first == no previous tile).
template <int KT>
__device__ __forceinline__ void softmax_tile(RowState& st, uint32_t tS, uint32_t tO, bool need_mask, const float* kb_tile,
                                             int k0, int kv_end, float sc, bool first, uint64_t* pv_bar,
                                             uint32_t pv_parity) {
  float s[KT];
#pragma unroll
  for (int c = 0; c < KT / 32; ++c) {
    uint32_t raw[32];
    tmem_ld_32x32(tS + c * 32, raw);
#pragma unroll
    for (int i = 0; i < 32; ++i) s[c * 32 + i] = __uint_as_float(raw[i]);
  }
  tmem_ld_wait();
  if (need_mask) {
#pragma unroll
    for (int i4 = 0; i4 < KT / 4; ++i4) {
      const float4 bb = __ldg(reinterpret_cast<const float4*>(kb_tile) + i4);
      const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = 4 * i4 + e;
        const float x = s[i] + bv[e];
        s[i] = (k0 + i < kv_end) ? x : -INFINITY;
      }
    }
  }
  // four independent maximum chains (one dependent chain of 64 three-input maxima costs ~450 cycles)
  float mx[4];
  constexpr int W = KT / 4;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float m = s[W * c];
#pragma unroll
    for (int i = 1; i < W - 1; i += 2) m = fmaxf(fmaxf(m, s[W * c + i]), s[W * c + i + 1]);
    mx[c] = fmaxf(m, s[W * c + W - 1]);
  }
  const float m_tile = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * sc;
  const bool had = st.m_ref > -INFINITY;
  const bool grow = had ? (m_tile > st.m_ref + kRescaleThreshold) : (m_tile > -INFINITY);
  const float m_new = grow ? m_tile : st.m_ref;
  if (!first && __any_sync(0xffffffffu, grow && had)) {
    // O still receives the previous P V: wait until it has retired before touching it
    mbar_wait(pv_bar, pv_parity);
    tc_fence_after();
    const float alpha = (grow && had) ? ex2(st.m_ref - m_new) : 1.0f;
#pragma unroll
    for (int c = 0; c < kD / 32; ++c) {
      uint32_t o[32];
      tmem_ld_32x32(tO + c * 32, o);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
      tmem_st_32x32(tO + c * 32, o);
    }
    st.l *= alpha;
    tmem_st_wait();
  }
  st.m_ref = m_new;
  const float neg_m = (m_new == -INFINITY) ? 0.f : -m_new;
  const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(neg_m, neg_m);
  float2 acc[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
  for (int hlf = 0; hlf < KT / 64; ++hlf) {
    uint32_t pk[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float2 x = __ffma2_rn(make_float2(s[hlf * 64 + 2 * i], s[hlf * 64 + 2 * i + 1]), sc2, nm2);
      float2 e;
      if ((i & 3) < kPolyPairs) {
        e = ex2_poly2(x);
      } else {
        e.x = ex2(x.x);
        e.y = ex2(x.y);
      }
      acc[i & 3] = __fadd2_rn(acc[i & 3], e);
      pk[i] = pack_bf16x2(e.x, e.y);
    }
    tmem_st_32x32(tS + hlf * 32, pk);
  }
  const float2 a01 = __fadd2_rn(acc[0], acc[1]), a23 = __fadd2_rn(acc[2], acc[3]);
  st.l += (a01.x + a01.y) + (a23.x + a23.y);
}

// normalise this thread's output row (64 fp32 columns at tO) and store it (Ulysses: to the owning ranks' buffers)
__device__ __forceinline__ void store_row(const AttnArgs& a, uint32_t tO, float l_run, bool row_ok, int b, int qi, int head) {
  float o[kD];
#pragma unroll
  for (int c = 0; c < kD / 32; ++c) {
    uint32_t raw[32];
    tmem_ld_32x32(tO + c * 32, raw);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) o[c * 32 + i] = __uint_as_float(raw[i]);
  }
  const float inv = (l_run > 0.f) ? 1.0f / l_run : 0.f;
  if (!row_ok) return;
  uint4 q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    q[i].x = pack_bf16x2(o[8 * i + 0] * inv, o[8 * i + 1] * inv);
    q[i].y = pack_bf16x2(o[8 * i + 2] * inv, o[8 * i + 3] * inv);
    q[i].z = pack_bf16x2(o[8 * i + 4] * inv, o[8 * i + 5] * inv);
    q[i].w = pack_bf16x2(o[8 * i + 6] * inv, o[8 * i + 7] * inv);
  }
  const int HD = a.H * kD;
  const long long off = (static_cast<long long>(b) * a.L + qi) * HD + head * kD;
  if (a.n_peers == 0) {
    uint4* d4 = reinterpret_cast<uint4*>(a.out + off);
#pragma unroll
    for (int i = 0; i < 8; ++i) d4[i] = q[i];
  } else if (qi >= a.peer_Lc) {  // a video row: straight into its owner's buffer (NVLink store)
    uint4* d4 = reinterpret_cast<uint4*>(a.out_peer[(qi - a.peer_Lc) / a.peer_Lw] + off);
#pragma unroll
    for (int i = 0; i < 8; ++i) d4[i] = q[i];
  } else {                       // a context row: every rank continues the replicated stream
    for (int p = 0; p < a.n_peers; ++p) {
      uint4* d4 = reinterpret_cast<uint4*>(a.out_peer[p] + off);
#pragma unroll
      for (int i = 0; i < 8; ++i) d4[i] = q[i];
    }
  }
}

// =====================================================================================================
// attn_pipe_kernel — the DEFAULT since round 2 (DV_ATTN_PIPE=0 selects attn_kernel above): all GPU suites (kernel,
// model-level parity, sharding, rollout, boundary) pass with it; attention time of a rollout 306 -> 265 ms, B3 / L2237
// 209 -> 184 us (profiles/r02a_summary.txt, r02b_launch_table_rollout.txt).
// one CTA per SM with TWO S buffers in tensor memory.  The MMA warp issues S_{j+1} = Q K_{j+1}^T before
// O += P_j V_j, so the softmax warps find the next scores ready when they finish a tile and have the XU
// pipe to themselves; the price is 512 TMEM columns (one CTA per SM) and a 4-stage K/V ring.
//   TMEM: S0 [0,128)  S1 [128,256)  O [256,320)  Q [320,352);  P_j aliases S_{j&1}[0,64).  Row sums stay in registers.
//   order on the tensor pipe: S0, S1, (P0V0, S2), (P1V1, S3), ...  — in order, so S_{j+2} may overwrite P_j
//   lazy rescale of O / l (rare path) first waits for pv_done(j-1): P_{j-1} V_{j-1} has retired.
// =====================================================================================================
constexpr int kPipeStages = 4;
constexpr uint32_t kPipeTmemCols = 512;   // (the pair kernel's allocation; the pipe kernel takes 4 * KT)
constexpr int pipe_smem_bytes(int KT) { return kTileBytes + kPipeStages * 2 * KT * 128 + 256 + 1024; }

// KT = keys per inner tile.  KT = 128: one CTA per SM (512 TMEM columns).  KT = 64: S0 [0,64) S1 [64,128) O [128,192)
// Q [192,224) = 256 columns and 82 KB of shared memory, so TWO CTAs share an SM: a second softmax warp on every
// scheduler to fill the dependency stalls of the first (one warp per scheduler issues at ~0.4 IPC), and a CTA's
// prologue / epilogue hidden behind its neighbour's main loop.
template <int KT>
__global__ void __launch_bounds__(kAttnThreads, KT == 64 ? 2 : 1) attn_pipe_kernel(const __grid_constant__ AttnArgs a) {
  constexpr uint32_t kPColS = 0, kPColO = 2 * KT, kPColQ = 2 * KT + 64, kCols = 4 * KT;
  constexpr int kStageBytes = 2 * KT * 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + kPipeStages * kStageBytes);
  uint64_t* q_full = bars;
  uint64_t* q_ready = bars + 1;
  uint64_t* kv_full = bars + 2;                  // [kPipeStages]
  uint64_t* kv_empty = kv_full + kPipeStages;    // [kPipeStages]
  uint64_t* s_full = kv_empty + kPipeStages;     // [2]: S_j ready in buffer j & 1
  uint64_t* p_full = s_full + 2;                 // [2]: P_j written over buffer j & 1 (128 arrivals)
  uint64_t* pv_done = p_full + 2;                // P_j V_j retired (one phase per tile)
  uint64_t* o_done = pv_done + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = static_cast<int>(gridDim.z) - 1 - static_cast<int>(blockIdx.z);
  const int q0 = qt * kTile;
  const int head = a.head0 + static_cast<int>(blockIdx.x);
  const int b = blockIdx.y;
  const int last_q = min(q0 + kTile, a.L) - 1;
  const int n_kv = (__ldg(a.kv_end + last_q) + KT - 1) / KT;
#ifdef DV_ATTN_TRACE
  const bool tr_cta = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0;
  const bool tc = threadIdx.x == 64;
  const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  if (tc && cta < 8192) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    a.trace[1024 + cta * 8] = smid;
  }
  DV_TC(tc, cta, 1);   // entry
#endif

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.tmQKV);
    if (KT != 128) tma_prefetch_desc(&a.tmKV64);
    mbar_init(q_full, 1);
    mbar_init(q_ready, 128);
    for (int i = 0; i < kPipeStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
    }
    mbar_init(pv_done, 1);
    mbar_init(o_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int HD = a.H * kD;
  pdl_wait();
  if (a.pdl_early) pdl_trigger();
  DV_TC(tc, cta, 2);   // set up (barriers, tensor memory), predecessor's data visible

  if (warp == 0) {
    // whole warp, converged: the w_ forms elect one lane and keep their operands in uniform registers
    w_mbar_expect_tx(q_full, kTileBytes);
    w_tma_load_3d(&a.tmQKV, q_full, sQ, head * kD, q0, b);
    for (int j = 0; j < n_kv; ++j) {
      const int s = j % kPipeStages;
      const uint32_t ph = (j / kPipeStages) & 1;
      mbar_wait(&kv_empty[s], ph ^ 1);
      uint8_t* sK = sKV + s * kStageBytes;
      const CUtensorMap* tm = KT == 128 ? &a.tmQKV : &a.tmKV64;
      w_mbar_expect_tx(&kv_full[s], kStageBytes);
      w_tma_load_3d(tm, &kv_full[s], sK, HD + head * kD, j * KT, b);
      w_tma_load_3d(tm, &kv_full[s], sK + kStageBytes / 2, 2 * HD + head * kD, j * KT, b);
    }
  } else if (warp == 1) {
    if (n_kv > 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, KT, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);
      const uint32_t tO = tmem_base + kPColO, tQ = tmem_base + kPColQ;
      auto issue_s = [&](int j) {
        mbar_wait(&kv_full[j % kPipeStages], (j / kPipeStages) & 1);
        tc_fence_after();
        const uint32_t tS = tmem_base + kPColS + (j & 1) * KT;
        const uint32_t aK = smem_u32(sKV + (j % kPipeStages) * kStageBytes);
        const uint64_t dk = umma_desc_sw128(aK, 16, 1024);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) w_umma_bf16_ts(tS, tQ + k * 8, dk + 2 * k, idesc_s, k != 0);
        w_umma_commit(&s_full[j & 1]);
      };
      mbar_wait(q_ready, 0);
      issue_s(0);
      if (n_kv > 1) issue_s(1);
      for (int j = 0; j < n_kv; ++j) {
        DV_TR(tr_cta, 2, j, 0);
        mbar_wait(&p_full[j & 1], (j >> 1) & 1);  // P_j in TMEM (and O / l rescaled if the max grew)
        tc_fence_after();
        DV_TR(tr_cta, 2, j, 1);
        const uint32_t tP = tmem_base + kPColS + (j & 1) * KT;
        const uint32_t aV = smem_u32(sKV + (j % kPipeStages) * kStageBytes) + kStageBytes / 2;
#pragma unroll
        for (int k = 0; k < KT / 16; ++k) {
          const uint64_t dv = umma_desc_sw128(aV + k * 2048, 1024, 1024);
          w_umma_bf16_ts(tO, tP + k * 8, dv, idesc_pv, (j | k) != 0);
        }
        w_umma_commit(&kv_empty[j % kPipeStages]);
        w_umma_commit(pv_done);
        DV_TR(tr_cta, 2, j, 2);
        if (j + 2 < n_kv) issue_s(j + 2);  // in order behind P_j V_j: S_{j+2} may overwrite P_j
        if (j + 1 == n_kv) w_umma_commit(o_done);
        DV_TR(tr_cta, 2, j, 3);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int qi = q0 + r;
    const bool row_ok = qi < a.L;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tO = tmem_base + kPColO + lane_addr;
    if (n_kv > 0) {
      mbar_wait(q_full, 0);
      {
        uint32_t qr[32];
        const uint8_t* row = sQ + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 t = *reinterpret_cast<const uint4*>(row + ((c ^ (r & 7)) << 4));
          qr[4 * c + 0] = t.x;
          qr[4 * c + 1] = t.y;
          qr[4 * c + 2] = t.z;
          qr[4 * c + 3] = t.w;
        }
        tmem_st_32x32(tmem_base + kPColQ + lane_addr, qr);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(q_ready);
      }
      const int kv_end = row_ok ? __ldg(a.kv_end + qi) : 0;
      const int kv_end_min = __ldg(a.kv_end + q0);
      const float* kb = a.key_bias + static_cast<long long>(b) * a.Lpad;
      const float sc = a.scale_log2;
      RowState st = {-INFINITY, 0.f};
      DV_TC(tc, cta, 3);   // Q parked in tensor memory
      const int* dead_row = a.tile_dead ? a.tile_dead + b * (a.Lpad / kTile) : nullptr;
      for (int j = 0; j < n_kv; ++j) {
        const int k0 = j * KT;
        const uint32_t tS = tmem_base + kPColS + (j & 1) * KT + lane_addr;
        const bool need_mask = (k0 + KT > kv_end_min) || (dead_row == nullptr) || (__ldg(dead_row + k0 / kTile) != 0);
#ifdef DV_ATTN_TRACE
        const bool tr = tr_cta && warp == 2;
#endif
        DV_TR(tr, 0, j, 0);
        mbar_wait(&s_full[j & 1], (j >> 1) & 1);
        tc_fence_after();
        DV_TR(tr, 0, j, 1);
        if (j == 0) DV_TC(tc, cta, 4);   // first scores ready
        softmax_tile<KT>(st, tS, tO, need_mask, kb + k0, k0, kv_end, sc, j == 0, pv_done, (j - 1) & 1);
        DV_TR(tr, 0, j, 4);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[j & 1]);
        DV_TR(tr, 0, j, 5);
      }

      DV_TC(tc, cta, 5);   // last P handed over
      mbar_wait(o_done, 0);
      tc_fence_after();
      store_row(a, tO, st.l, row_ok, b, qi, head);
      DV_TC(tc, cta, 6);   // row stored
    }
  }

  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kCols>(tmem_base);
  }
  DV_TC(tc, cta, 7);   // exit
}

// =====================================================================================================
// attn_pair_kernel (round 2, DV_ATTN_PIPE=2): TWO query tiles (256 rows) of one head per CTA, one softmax warpgroup
// each, sharing every K/V tile.  While group g waits for O_g += P_g V_j and S_g = Q_g K_{j+1}^T on the tensor pipe the
// other group has the XU pipe (ex2) to itself, and each K/V tile is fetched once for 256 queries instead of 128.  The
// register file is re-partitioned with setmaxnreg: the two softmax warpgroups (one row of S = 128 fp32 per thread) grow
// to 232 registers, the warpgroup that holds the TMA and MMA warps shrinks to 40 — the first two-group attempt
// (attn_pipe2, groups on alternate key tiles of the SAME query tile, 320 threads at the 168-register cap, partial
// results merged at the end) spilled in the softmax loop and measured 28 % slower than one group (profiles/r02h_summary.txt).
//   warps 0-3   softmax of rows q0 + [0, 128)       warps 4-7   softmax of rows q0 + [128, 256)
//   warp 8      TMA producer (Q0, Q1, then K_j / V_j into a 4-stage ring)      warp 9   MMA issuer     warps 10, 11 idle
//   TMEM: S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384) Q0 [384,416) Q1 [416,448); P_g aliases S_g[0,64); the row
//   sums stay in registers.  Tensor-pipe order: S0(0) S1(0) | PV0(j) S0(j+1) PV1(j) S1(j+1) | ... — in order, so
//   S_g(j+1) may overwrite P_g(j).  The second tile sees at least as many keys as the first (kv_end is non-decreasing);
//   a tile beyond L takes no part.
// =====================================================================================================
constexpr int kPairThreads = 384;
constexpr uint32_t kQColS = 0, kQColO = 256, kQColQ = 384;
constexpr int kPairSmemBytes = 2 * kTileBytes + kPipeStages * 2 * kTileBytes + 256 + 1024;

__global__ void __launch_bounds__(kPairThreads, 1) attn_pair_kernel(const __grid_constant__ AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                            // two tiles
  uint8_t* sKV = sQ + 2 * kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + kPipeStages * 2 * kTileBytes);
  uint64_t* q_full = bars;                       // [2]
  uint64_t* q_ready = bars + 2;                  // [2] (128 arrivals)
  uint64_t* kv_full = bars + 4;                  // [kPipeStages]
  uint64_t* kv_empty = kv_full + kPipeStages;    // [kPipeStages]
  uint64_t* s_full = kv_empty + kPipeStages;     // [2]: S_g(j) ready
  uint64_t* p_full = s_full + 2;                 // [2]: P_g(j) written over S_g (128 arrivals)
  uint64_t* pv_done = p_full + 2;                // [2]: P_g(j) V_j retired (one phase per tile)
  uint64_t* o_done = pv_done + 2;                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int wg = warp >> 2;
  const int qt = static_cast<int>(gridDim.z) - 1 - static_cast<int>(blockIdx.z);
  const int q0 = qt * 2 * kTile;
  const int head = a.head0 + static_cast<int>(blockIdx.x);
  const int b = blockIdx.y;
  int n_kv[2];
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int first = q0 + g * kTile;
    n_kv[g] = first < a.L ? (__ldg(a.kv_end + min(first + kTile, a.L) - 1) + kTile - 1) / kTile : 0;
  }
  const int n_max = max(n_kv[0], n_kv[1]);
#ifdef DV_ATTN_TRACE
  const bool tr_cta = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0;
#endif

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&a.tmQKV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_ready[i], 128);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
      mbar_init(&pv_done[i], 1);
      mbar_init(&o_done[i], 1);
    }
    for (int i = 0; i < kPipeStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<kPipeTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int HD = a.H * kD;
  pdl_wait();
  if (a.pdl_early) pdl_trigger();

  if (wg == 2) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 8) {
      // whole warp, converged: the w_ forms elect one lane and keep their operands in uniform registers
      for (int g = 0; g < 2; ++g)
        if (n_kv[g] > 0) {
          w_mbar_expect_tx(&q_full[g], kTileBytes);
          w_tma_load_3d(&a.tmQKV, &q_full[g], sQ + g * kTileBytes, head * kD, q0 + g * kTile, b);
        }
      for (int j = 0; j < n_max; ++j) {
        const int s = j % kPipeStages;
        const uint32_t ph = (j / kPipeStages) & 1;
        mbar_wait(&kv_empty[s], ph ^ 1);
        uint8_t* sK = sKV + s * 2 * kTileBytes;
        w_mbar_expect_tx(&kv_full[s], 2 * kTileBytes);
        w_tma_load_3d(&a.tmQKV, &kv_full[s], sK, HD + head * kD, j * kTile, b);
        w_tma_load_3d(&a.tmQKV, &kv_full[s], sK + kTileBytes, 2 * HD + head * kD, j * kTile, b);
      }
    } else if (warp == 9) {
      if (n_max > 0) {
        constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
        constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);
        auto issue_s = [&](int g, int j) {
          mbar_wait(&kv_full[j % kPipeStages], (j / kPipeStages) & 1);
          tc_fence_after();
          const uint32_t tS = tmem_base + kQColS + g * 128;
          const uint32_t tQ = tmem_base + kQColQ + g * 32;
          const uint64_t dk = umma_desc_sw128(smem_u32(sKV + (j % kPipeStages) * 2 * kTileBytes), 16, 1024);
#pragma unroll
          for (int k = 0; k < kD / 16; ++k) w_umma_bf16_ts(tS, tQ + k * 8, dk + 2 * k, idesc_s, k != 0);
          w_umma_commit(&s_full[g]);
        };
        for (int g = 0; g < 2; ++g)
          if (n_kv[g] > 0) {
            mbar_wait(&q_ready[g], 0);
            issue_s(g, 0);
          }
        for (int j = 0; j < n_max; ++j) {
          for (int g = 0; g < 2; ++g) {
            if (j >= n_kv[g]) continue;
            DV_TR(tr_cta, 2 + g, j, 0);
            mbar_wait(&p_full[g], j & 1);  // P_g(j) in TMEM (and O_g / l_g rescaled if the max grew)
            tc_fence_after();
            DV_TR(tr_cta, 2 + g, j, 1);
            const uint32_t tP = tmem_base + kQColS + g * 128;
            const uint32_t tO = tmem_base + kQColO + g * 64;
            const uint32_t aV = smem_u32(sKV + (j % kPipeStages) * 2 * kTileBytes) + kTileBytes;
#pragma unroll
            for (int k = 0; k < kTile / 16; ++k) {
              const uint64_t dv = umma_desc_sw128(aV + k * 2048, 1024, 1024);
              w_umma_bf16_ts(tO, tP + k * 8, dv, idesc_pv, (j | k) != 0);
            }
            if (g == 1 || j >= n_kv[1]) w_umma_commit(&kv_empty[j % kPipeStages]);  // the last reader of K_j / V_j
            w_umma_commit(&pv_done[g]);
            DV_TR(tr_cta, 2 + g, j, 2);
            if (j + 1 < n_kv[g]) issue_s(g, j + 1);   // in order behind P_g(j) V_j: S_g(j+1) may overwrite P_g(j)
            else w_umma_commit(&o_done[g]);
            DV_TR(tr_cta, 2 + g, j, 3);
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int g = wg;
    const int n_g = n_kv[g];
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int qbase = q0 + g * kTile;
    const int qi = qbase + r;
    const bool row_ok = qi < a.L;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem_base + kQColS + g * 128 + lane_addr;
    const uint32_t tO = tmem_base + kQColO + g * 64 + lane_addr;
    if (n_g > 0) {
      mbar_wait(&q_full[g], 0);
      {
        uint32_t qr[32];
        const uint8_t* row = sQ + g * kTileBytes + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 t = *reinterpret_cast<const uint4*>(row + ((c ^ (r & 7)) << 4));
          qr[4 * c + 0] = t.x;
          qr[4 * c + 1] = t.y;
          qr[4 * c + 2] = t.z;
          qr[4 * c + 3] = t.w;
        }
        tmem_st_32x32(tmem_base + kQColQ + g * 32 + lane_addr, qr);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&q_ready[g]);
      }
      const int kv_end = row_ok ? __ldg(a.kv_end + qi) : 0;
      const int kv_end_min = __ldg(a.kv_end + qbase);
      const float* kb = a.key_bias + static_cast<long long>(b) * a.Lpad;
      const float sc = a.scale_log2;
      RowState st = {-INFINITY, 0.f};
      const int* dead_row = a.tile_dead ? a.tile_dead + b * (a.Lpad / kTile) : nullptr;
      for (int j = 0; j < n_g; ++j) {
        const int k0 = j * kTile;
        const bool need_mask = (k0 + kTile > kv_end_min) || (dead_row == nullptr) || (__ldg(dead_row + j) != 0);
#ifdef DV_ATTN_TRACE
        const bool tr = tr_cta && quarter == 0;
#endif
        DV_TR(tr, g, j, 0);
        mbar_wait(&s_full[g], j & 1);
        tc_fence_after();
        DV_TR(tr, g, j, 1);
        softmax_tile<kTile>(st, tS, tO, need_mask, kb + k0, k0, kv_end, sc, j == 0, &pv_done[g], (j - 1) & 1);
        DV_TR(tr, g, j, 4);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[g]);
        DV_TR(tr, g, j, 5);
      }

      mbar_wait(&o_done[g], 0);
      tc_fence_after();
      store_row(a, tO, st.l, row_ok, b, qi, head);
    }
  }

  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<kPipeTmemCols>(tmem_base);
  }
}

}  // namespace

#ifdef DV_ATTN_TRACE
long long* g_attn_trace = nullptr;
extern "C" long long* dv_attn_trace_buffer() {
  if (!g_attn_trace) {
    cudaMalloc(&g_attn_trace, (1024 + 8 * 8192) * sizeof(long long));
    cudaMemset(g_attn_trace, 0, (1024 + 8 * 8192) * sizeof(long long));
  }
  return g_attn_trace;
}
#endif

int launch_attention(const void* qkv, void* out, const int* kv_end, const float* key_bias,
                     const int* tile_dead, int B, int L, int Lpad, int H, cudaStream_t stream,
                     double flops, int head0, int n_heads, void* const* out_peers, int n_peers,
                     int peer_Lc, int peer_Lw) {
  if (n_heads <= 0) n_heads = H;
  DV_REQUIRE(head0 >= 0 && head0 + n_heads <= H, "attention: heads [%d, %d) of %d", head0, head0 + n_heads, H);
  DV_REQUIRE(Lpad % 128 == 0 && Lpad >= L, "attention: Lpad=%d must be a multiple of 128 >= L=%d",
             Lpad, L);
  DV_REQUIRE(B > 0 && L > 0 && H > 0, "attention: empty problem B=%d L=%d H=%d", B, L, H);
  AttnArgs a;
  uint64_t dims[3] = {(uint64_t)(3 * H * kD), (uint64_t)L, (uint64_t)B};
  uint64_t strides[2] = {(uint64_t)(3 * H * kD) * 2, (uint64_t)(3 * H * kD) * 2 * (uint64_t)L};
  uint32_t box[3] = {64, 128, 1};
  int rc = make_tensor_map_bf16(&a.tmQKV, qkv, 3, dims, strides, box, 1);
  if (rc) return rc;
  // DV_ATTN_PIPE: 0 = round-1 kernel, 1 = pipelined kernel, 128-key tiles, one CTA per SM, 2 = two query tiles per CTA
  // (attn_pair_kernel), 3 = pipelined kernel, 64-key tiles, two CTAs per SM
  // (the default: 291 us over four rollout layouts against 367 us for mode 1 and 352 us for mode 2, scripts/probe/attn_time.py,
  // profiles/r02p_attn_modes.txt)
  static const int pipe_mode = getenv("DV_ATTN_PIPE") != nullptr ? atoi(getenv("DV_ATTN_PIPE")) : 3;
  if (pipe_mode == 3) {
    uint32_t box64[3] = {64, 64, 1};
    rc = make_tensor_map_bf16(&a.tmKV64, qkv, 3, dims, strides, box64, 1);
    if (rc) return rc;
  } else {
    a.tmKV64 = a.tmQKV;
  }
  a.out = reinterpret_cast<__nv_bfloat16*>(out);
  a.kv_end = kv_end;
  a.key_bias = key_bias;
  a.tile_dead = tile_dead;
  a.L = L;
  a.Lpad = Lpad;
  a.H = H;
  a.B = B;
  a.head0 = head0;
  DV_REQUIRE(n_peers >= 0 && n_peers <= 8, "attention: %d peers", n_peers);
  a.n_peers = out_peers ? n_peers : 0;
  a.peer_Lc = peer_Lc;
  a.peer_Lw = peer_Lw > 0 ? peer_Lw : 1;
  for (int i = 0; i < 8; ++i)
    a.out_peer[i] = (out_peers && i < n_peers) ? reinterpret_cast<__nv_bfloat16*>(out_peers[i]) : nullptr;
  a.scale_log2 = 0.125f * 1.4426950408889634f;
  a.pdl_early = pdl_early() ? 1 : 0;
#ifdef DV_ATTN_TRACE
  a.trace = dv_attn_trace_buffer();
  if (a.tile_dead == nullptr) {   // the probe calls dv_attention (no dead-tile table): use an all-live one, as a forward has
    static int* zeros = nullptr;
    if (!zeros) {
      cudaMalloc(&zeros, 65536 * sizeof(int));
      cudaMemset(zeros, 0, 65536 * sizeof(int));
    }
    a.tile_dead = zeros;
  }
#endif
  const bool pipe = pipe_mode != 0;
  static bool attr_set = false;
  if (!attr_set) {
    DV_CHECK_CUDA(cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kSmemBytes));
    DV_CHECK_CUDA(cudaFuncSetAttribute(attn_pipe_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       pipe_smem_bytes(128)));
    DV_CHECK_CUDA(cudaFuncSetAttribute(attn_pipe_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       pipe_smem_bytes(64)));
    DV_CHECK_CUDA(cudaFuncSetAttribute(attn_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kPairSmemBytes));
    attr_set = true;
  }
  dim3 grid(n_heads, B, (L + kTile - 1) / kTile);
  const dim3 grid_pair(n_heads, B, (L + 2 * kTile - 1) / (2 * kTile));
  char tag[56] = "";
  if (prof_on()) snprintf(tag, sizeof(tag), "attn B%d L%d H%d%s", B, L, H, pipe_mode == 3 ? " k64" : pipe_mode == 2 ? " pair" : (pipe ? " pipe" : ""));
  const int pid = prof_begin(PROF_ATTN, flops, 2.0 * B * L * 4.0 * H * kD, stream, tag);
  if (pipe_mode == 3)
    DV_CHECK_CUDA(launch_pdl(attn_pipe_kernel<64>, grid, dim3(kAttnThreads), pipe_smem_bytes(64), stream, 1, a));
  else if (pipe_mode == 2)
    DV_CHECK_CUDA(launch_pdl(attn_pair_kernel, grid_pair, dim3(kPairThreads), kPairSmemBytes, stream, 1, a));
  else if (pipe)
    DV_CHECK_CUDA(launch_pdl(attn_pipe_kernel<128>, grid, dim3(kAttnThreads), pipe_smem_bytes(128), stream, 1, a));
  else
    DV_CHECK_CUDA(launch_pdl(attn_kernel, grid, dim3(kAttnThreads), kSmemBytes, stream, 1, a));
  prof_end(pid, stream);
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

}  // namespace dv

// deepv_b200 — persistent, warp-specialised tcgen05 GEMM / implicit-GEMM conv3d.
//
// One CTA per SM (grid = min(#tiles, #SMs)), 192 threads:
//   warp 0      TMA producer   (one lane): A/B k-blocks -> 128B-swizzled smem ring
//   warp 1      MMA issuer     (one lane): tcgen05.mma M=128, N=BN, K=16, fp32 accum in TMEM;
//               also owns TMEM alloc/dealloc
//   warps 2..5  epilogue       (128 threads, one accumulator row each): tcgen05.ld ->
//               fused epilogue -> vectorised global stores
// Two TMEM accumulator stages let the epilogue of tile i overlap the mainloop of
// tile i+1.  See gemm.cuh for the operand / epilogue contract.
#include "gemm.cuh"

namespace dv {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int kThreads = 192;

template <int BN>
struct Cfg {
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarBytes = 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;  // + align slack
  static constexpr uint32_t kTmemCols = 2 * BN;  // 128, 256 or 512 (power of two)
};

struct KArgs {
  alignas(64) CUtensorMap tmA;
  alignas(64) CUtensorMap tmB;
  GemmDesc d;
  int m_tiles, n_tiles, k_blocks, total_tiles;
  int splits, kb_per_split;  // split-K (gate-residual epilogue only: fp32 atomics into x)
  int tiles_w, tiles_h;      // conv: M-tile grid inside one frame
  int c_blocks;              // conv: Cin / 64
};

struct TileCoord {
  int b, m_tile, n_tile, split, kb0, kb1;
};

__device__ __forceinline__ TileCoord decode_tile(const KArgs& a, int tile) {
  TileCoord tc;
  tc.n_tile = tile % a.n_tiles;
  int rest = tile / a.n_tiles;
  tc.split = rest % a.splits;
  rest /= a.splits;
  tc.m_tile = rest % a.m_tiles;
  tc.b = rest / a.m_tiles;
  tc.kb0 = tc.split * a.kb_per_split;
  tc.kb1 = min(tc.kb0 + a.kb_per_split, a.k_blocks);
  return tc;
}

// ------------------------------------------------------------------------------
// epilogue helpers: one thread = one output row, 32 columns at a time
// ------------------------------------------------------------------------------
__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const float (&v)[32]) {
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

__device__ __forceinline__ void load_bias32(const float* bias, int n, float (&v)[32]) {
  if (bias != nullptr) {
    const float4* bp = reinterpret_cast<const float4*>(bias + n);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 t = __ldg(bp + i);
      v[4 * i + 0] += t.x;
      v[4 * i + 1] += t.y;
      v[4 * i + 2] += t.z;
      v[4 * i + 3] += t.w;
    }
  }
}

template <int BN>
__device__ __forceinline__ void epilogue_tile(const KArgs& a, const TileCoord& tc,
                                              uint32_t tmem_acc, int row_in_tile, int quarter) {
  const GemmDesc& d = a.d;
  const int n0 = tc.n_tile * BN;
  const uint32_t taddr_row = tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16);

  // ---- row geometry ----------------------------------------------------------
  int m = tc.m_tile * BM + row_in_tile;  // dense row index inside batch
  bool row_ok;
  // conv geometry
  int ct = 0, ch = 0, cw = 0;
  if (d.a_mode == 1) {
    int per_frame = a.tiles_w * a.tiles_h;
    ct = tc.m_tile / per_frame;
    int r = tc.m_tile % per_frame;
    int ty = r / a.tiles_w, tx = r % a.tiles_w;
    ch = ty * 8 + (row_in_tile >> 4);
    cw = tx * 16 + (row_in_tile & 15);
    row_ok = (ct < d.cT) && (ch < d.cH) && (cw < d.cW);
  } else {
    row_ok = m < d.M;
  }

  uint32_t raw[32];
  float v[32];

  if (d.mode == EPI_QKV) {
    // 64 columns (one head) at a time: RMSNorm over the head, then RoPE pairs.
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(d.out) +
                         static_cast<long long>(tc.b) * d.out_batch_stride +
                         static_cast<long long>(m + d.out_row_offset) * d.ldo;
    const int fid = row_ok ? __ldg(d.frame_id + m) : 0;
    const float2* cs = reinterpret_cast<const float2*>(d.rope_cs) + fid * 32;
    for (int c = 0; c < BN / 64; ++c) {
      float h[64];
      const int n = n0 + c * 64;
      tmem_ld_32x32(taddr_row + c * 64, raw);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) h[i] = __uint_as_float(raw[i]);
      tmem_ld_32x32(taddr_row + c * 64 + 32, raw);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) h[32 + i] = __uint_as_float(raw[i]);
      if (n >= d.N || !row_ok) continue;
      if (d.bias != nullptr) {
#pragma unroll
        for (int i = 0; i < 64; ++i) h[i] += __ldg(d.bias + n + i);
      }
      const int region = n / d.heads_dim;  // 0 q, 1 k, 2 v
      if (region < 2) {
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 64; ++i) ss += h[i] * h[i];
        const float rs = rsqrtf(ss * (1.0f / 64.0f) + 1e-5f);  // RMSNorm eps, mmdit.py:195,453-454
        const float* w = d.qk_norm_w + region * 64;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float x0 = h[2 * i] * rs * __ldg(w + 2 * i);
          float x1 = h[2 * i + 1] * rs * __ldg(w + 2 * i + 1);
          float2 t = __ldg(cs + i);  // (cos, sin) of frame * 10000^(-2i/64)
          h[2 * i] = t.x * x0 - t.y * x1;
          h[2 * i + 1] = t.y * x0 + t.x * x1;
        }
      }
      float lo[32], hi[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        lo[i] = h[i];
        hi[i] = h[32 + i];
      }
      store_bf16x32(out + n, lo);
      store_bf16x32(out + n + 32, hi);
    }
    return;
  }

  for (int c = 0; c < BN / 32; ++c) {
    const int n = n0 + c * 32;
    tmem_ld_32x32(taddr_row + c * 32, raw);
    tmem_ld_wait();
    if (n >= d.N || !row_ok) continue;  // warp-uniform in n; row predicate only skips stores
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);

    switch (d.mode) {
      case EPI_BF16:
      case EPI_GELU: {
        load_bias32(d.bias, n, v);
        if (d.mode == EPI_GELU) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = gelu_tanh(v[i]);
        }
        const long long off = static_cast<long long>(tc.b) * d.out_batch_stride +
                              static_cast<long long>(m + d.out_row_offset) * d.ldo + n;
        if (d.residual != nullptr) {
          const uint4* r4 = reinterpret_cast<const uint4*>(
              reinterpret_cast<const __nv_bfloat16*>(d.residual) + off);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 t = __ldg(r4 + i);
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
        store_bf16x32(reinterpret_cast<__nv_bfloat16*>(d.out) + off, v);
      } break;
      case EPI_BF16_ROWBIAS: {
        const float bm = d.bias ? __ldg(d.bias + m) : 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += bm;
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(d.out) +
                             static_cast<long long>(tc.b) * d.out_batch_stride +
                             static_cast<long long>(m + d.out_row_offset) * d.ldo + n;
        store_bf16x32(out, v);
      } break;
      case EPI_RESID_GATE: {
        if (tc.split == 0) load_bias32(d.bias, n, v);
        float* x = reinterpret_cast<float*>(d.out) +
                   static_cast<long long>(tc.b) * d.out_batch_stride +
                   static_cast<long long>(m + d.out_row_offset) * d.ldo + n;
        const float4* g4 =
            reinterpret_cast<const float4*>(d.gate + tc.b * d.gate_batch_stride + n);
        if (a.splits > 1) {
          // split-K partial sums: x += gate * partial, merged with fp32 reductions in L2
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 g = __ldg(g4 + i);
            atomicAdd(x + 4 * i + 0, g.x * v[4 * i + 0]);
            atomicAdd(x + 4 * i + 1, g.y * v[4 * i + 1]);
            atomicAdd(x + 4 * i + 2, g.z * v[4 * i + 2]);
            atomicAdd(x + 4 * i + 3, g.w * v[4 * i + 3]);
          }
          break;
        }
        float4* x4 = reinterpret_cast<float4*>(x);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 g = __ldg(g4 + i);
          float4 r = x4[i];
          r.x += g.x * v[4 * i + 0];
          r.y += g.y * v[4 * i + 1];
          r.z += g.z * v[4 * i + 2];
          r.w += g.w * v[4 * i + 3];
          x4[i] = r;
        }
      } break;
      case EPI_F32_ADD: {
        load_bias32(d.bias, n, v);
        if (d.addend != nullptr) {
          const int ar = d.row_map ? __ldg(d.row_map + m) : m;
          const float4* a4 =
              reinterpret_cast<const float4*>(d.addend + static_cast<long long>(ar) * d.N + n);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 t = __ldg(a4 + i);
            v[4 * i + 0] += t.x;
            v[4 * i + 1] += t.y;
            v[4 * i + 2] += t.z;
            v[4 * i + 3] += t.w;
          }
        }
        float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(d.out) +
                                               static_cast<long long>(tc.b) * d.out_batch_stride +
                                               static_cast<long long>(m + d.out_row_offset) * d.ldo +
                                               n);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          o4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      } break;
      case EPI_UNPATCH: {
        // row m = (gy, gx) of the noisy clip's token grid; column n = (p1, p2, c)
        // out[b][c][0][2*gy+p1][2*gx+p2]   (mmdit.py:1453-1457, patch 2)
        const int gy = m / d.up_gw, gx = m % d.up_gw;
        const int H2 = 2 * d.up_gh, W2 = 2 * d.up_gw;
        const long long ob = static_cast<long long>(tc.b) * d.out_batch_stride;
#pragma unroll 4
        for (int i = 0; i < 32; ++i) {
          const int nn = n + i;
          if (nn < d.N) {
            float val = v[i] + (d.bias ? __ldg(d.bias + nn) : 0.f);
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
      } break;
      case EPI_CONV: {
        load_bias32(d.bias, n, v);
        int ot = ct, oh = ch, ow = cw, oc = n, oH = d.cH, oW = d.cW;
        if (d.conv_store == CONV_SHUFFLE_HW) {
          // packed weight rows are ordered (p1, p2, c): vae.py:382
          const int q = n / d.out_C;
          oc = n % d.out_C;
          oh = 2 * ch + (q >> 1);
          ow = 2 * cw + (q & 1);
          oH = 2 * d.cH;
          oW = 2 * d.cW;
        } else if (d.conv_store == CONV_INTERLEAVE_T) {
          // packed weight rows are ordered (p, c): vae.py:407-409
          const int p = n / d.out_C;
          oc = n % d.out_C;
          ot = 2 * ct + p - (d.conv_drop_first ? 1 : 0);
          if (ot < 0) break;
        }
        const int oT = (d.conv_store == CONV_INTERLEAVE_T)
                           ? (2 * d.cT - (d.conv_drop_first ? 1 : 0))
                           : d.cT;
        const long long off =
            (((static_cast<long long>(tc.b) * oT + ot) * oH + oh) * oW + ow) * d.out_C + oc;
        if (d.residual != nullptr) {
          const uint4* r4 =
              reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(d.residual) + off);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 t = __ldg(r4 + i);
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
        if (d.out_C >= 32) {
          store_bf16x32(reinterpret_cast<__nv_bfloat16*>(d.out) + off, v);
        } else {
          // narrow outputs (conv_out: 3 channels padded to 16 weight rows)
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(d.out) + off;
          for (int i = 0; i < 32 && (n + i) < d.out_C; ++i) o[i] = __float2bfloat16(v[i]);
        }
      } break;
      default:
        break;
    }
  }
}

// ------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------
template <int BN>
__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(const __grid_constant__ KArgs a) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle needs 1024-byte aligned tiles.
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::kStages;
  uint64_t* tmem_full = bars + 2 * C::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.tmA);
    tma_prefetch_desc(&a.tmB);
    for (int i = 0; i < C::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc<C::kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const GemmDesc& d = a.d;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(a, tile);
        int ct = 0, h0 = 0, w0 = 0;
        if (d.a_mode == 1) {
          int per_frame = a.tiles_w * a.tiles_h;
          ct = tc.m_tile / per_frame;
          int r = tc.m_tile % per_frame;
          h0 = (r / a.tiles_w) * 8;
          w0 = (r % a.tiles_w) * 16;
        }
        for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::kStageBytes;
          uint8_t* sb = sa + C::kABytes;
          mbar_expect_tx(&full_bar[stage], C::kStageBytes);
          if (d.a_mode == 1) {
            const int tap = kb / a.c_blocks;
            const int cb = kb - tap * a.c_blocks;
            const int dt = tap / (d.kh * d.kw);
            const int dh = (tap / d.kw) % d.kh;
            const int dw = tap % d.kw;
            // causal in time (kt-1 frames of zero history), centred in space
            tma_load_5d(&a.tmA, &full_bar[stage], sa, cb * 64, w0 + dw - d.kw / 2,
                        h0 + dh - d.kh / 2, ct + dt - (d.kt - 1), tc.b);
          } else {
            tma_load_3d(&a.tmA, &full_bar[stage], sa, kb * BK, tc.m_tile * BM, tc.b);
          }
          if (d.w_batch_stride != 0)
            tma_load_3d(&a.tmB, &full_bar[stage], sb, kb * BK, tc.n_tile * BN, tc.b);
          else
            tma_load_2d(&a.tmB, &full_bar[stage], sb, kb * BK, tc.n_tile * BN);
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ==============================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + as * BN;
        const TileCoord tc = decode_tile(a, tile);
        for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
          const uint32_t sb = sa + C::kABytes;
          const uint64_t da = umma_desc_sw128(sa, 16, 1024);
          const uint64_t db = umma_desc_sw128(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // +32 B per K=16 slice inside the 128 B swizzle row (address field is >>4)
            umma_bf16_ss(tmem_acc, da + 2 * k, db + 2 * k, idesc, ((kb - tc.kb0) | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[as]);
      }
    }
  } else {
    // ============================ epilogue ================================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may read
    const int row_in_tile = quarter * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      const TileCoord tc = decode_tile(a, tile);
      mbar_wait(&tmem_full[as], aph);
      tc_fence_after();
      epilogue_tile<BN>(a, tc, tmem_base + as * BN, row_in_tile, quarter);
      tc_fence_before();
      mbar_arrive(&tmem_empty[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

template <int BN>
int launch_bn(const KArgs& ka, cudaStream_t stream) {
  using C = Cfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    DV_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    attr_set = true;
  }
  int grid = ka.total_tiles < sm_count() ? ka.total_tiles : sm_count();
  gemm_tc_kernel<BN><<<grid, kThreads, C::kSmemBytes, stream>>>(ka);
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

}  // namespace

int launch_gemm(const GemmDesc& d, cudaStream_t stream) {
  KArgs ka;
  ka.d = d;
  DV_REQUIRE(d.batch > 0 && d.N > 0, "gemm: empty problem (batch=%d N=%d)", d.batch, d.N);

  // ---- tile shape: wide tiles when there is enough work to fill the SMs -----
  int K;
  if (d.a_mode == 1) {
    DV_REQUIRE(d.cC % 64 == 0, "conv: Cin=%d must be a multiple of 64 (pad at pack time)", d.cC);
    DV_REQUIRE(d.cW % 16 == 0 && d.cH % 8 == 0, "conv: H=%d W=%d must be multiples of 8/16", d.cH,
               d.cW);
    ka.c_blocks = d.cC / 64;
    K = d.kt * d.kh * d.kw * d.cC;
    ka.tiles_w = d.cW / 16;
    ka.tiles_h = d.cH / 8;
    ka.m_tiles = d.cT * ka.tiles_w * ka.tiles_h;
  } else {
    DV_REQUIRE(d.K % 64 == 0, "gemm: K=%d must be a multiple of 64", d.K);
    DV_REQUIRE(d.M > 0, "gemm: M=%d", d.M);
    K = d.K;
    ka.c_blocks = 1;
    ka.tiles_w = ka.tiles_h = 1;
    ka.m_tiles = (d.M + BM - 1) / BM;
  }
  ka.k_blocks = K / BK;

  // ---- tile width and split-K from a small cost model --------------------------------------
  // A CTA streams (128 + BN) x K_split x 2 bytes through its SM's L2 port (~80 GB/s per SM, ~10
  // TB/s aggregate) and spends 128 x BN x K_split / 4096 tensor cycles; small-M problems are
  // port-bound unless the reduction is split across more SMs (only the gate-residual epilogue
  // can merge partial sums: fp32 atomics into the residual stream).
  const long long mt_total = static_cast<long long>(ka.m_tiles) * d.batch;
  const int nsm = sm_count();
  int bn = 64, splits = 1;
  {
    double best = 1e30;
    const int cand_bn[3] = {256, 128, 64};
    const int cand_s[6] = {1, 2, 3, 4, 6, 8};
    for (int bi = 0; bi < 3; ++bi) {
      const int b_ = cand_bn[bi];
      if (b_ > 64 && (d.N % b_ != 0)) continue;
      const long long tiles = mt_total * ((d.N + b_ - 1) / b_);
      for (int si = 0; si < 6; ++si) {
        const int s_ = cand_s[si];
        if (s_ > 1 && (d.mode != EPI_RESID_GATE || ka.k_blocks / s_ < 4)) break;
        const int kbs = (ka.k_blocks + s_ - 1) / s_;
        const long long ctas = tiles * s_;
        const double waves = static_cast<double>((ctas + nsm - 1) / nsm);
        const double cta_bytes = (128.0 + b_) * kbs * 64 * 2;
        const double t_port = waves * cta_bytes / 80e9;
        const double t_l2 = ctas * cta_bytes / 10e12;
        const double t_mma = waves * (128.0 * b_ * kbs * 64 / 4096.0) / 1.8e9;
        const double t_epi = waves * (s_ > 1 ? 3.0e-6 : 1.0e-6) * (b_ / 64.0) * 0.5;
        double t = t_port > t_l2 ? t_port : t_l2;
        t = t > t_mma ? t : t_mma;
        t += t_epi + 2.0e-6;
        if (t < best) {
          best = t;
          bn = b_;
          splits = s_;
        }
      }
    }
  }
  DV_REQUIRE(d.mode == EPI_UNPATCH || d.mode == EPI_CONV || d.N % 32 == 0,
             "gemm: N=%d must be a multiple of 32 for epilogue mode %d", d.N, d.mode);
  ka.n_tiles = (d.N + bn - 1) / bn;
  ka.kb_per_split = (ka.k_blocks + splits - 1) / splits;
  splits = (ka.k_blocks + ka.kb_per_split - 1) / ka.kb_per_split;  // no empty split
  ka.splits = splits;
  ka.total_tiles = static_cast<int>(mt_total) * ka.n_tiles * splits;

  // ---- tensor maps ------------------------------------------------------------
  if (d.a_mode == 1) {
    uint64_t dims[5] = {(uint64_t)d.cC, (uint64_t)d.cW, (uint64_t)d.cH, (uint64_t)d.cT,
                        (uint64_t)d.batch};
    uint64_t strides[4] = {(uint64_t)d.cC * 2, (uint64_t)d.cC * d.cW * 2,
                           (uint64_t)d.cC * d.cW * d.cH * 2,
                           (uint64_t)d.cC * d.cW * d.cH * d.cT * 2};
    uint32_t box[5] = {64, 16, 8, 1, 1};
    int rc = make_tensor_map_bf16(&ka.tmA, d.A, 5, dims, strides, box, 1);
    if (rc) return rc;
  } else {
    uint64_t dims[3] = {(uint64_t)d.K, (uint64_t)d.M, (uint64_t)d.batch};
    uint64_t strides[2] = {(uint64_t)d.lda * 2, (uint64_t)d.a_batch_stride * 2};
    if (d.batch == 1) strides[1] = (uint64_t)d.lda * 2 * (uint64_t)d.M;
    uint32_t box[3] = {64, 128, 1};
    int rc = make_tensor_map_bf16(&ka.tmA, d.A, 3, dims, strides, box, 1);
    if (rc) return rc;
  }
  const uint64_t ldw_bytes = (uint64_t)(d.ldw > 0 ? d.ldw : K) * 2;
  if (d.w_batch_stride != 0) {
    uint64_t dims[3] = {(uint64_t)K, (uint64_t)d.w_rows, (uint64_t)d.batch};
    uint64_t strides[2] = {ldw_bytes, (uint64_t)d.w_batch_stride * 2};
    uint32_t box[3] = {64, (uint32_t)bn, 1};
    int rc = make_tensor_map_bf16(&ka.tmB, d.W, 3, dims, strides, box, 1);
    if (rc) return rc;
  } else {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)d.w_rows};
    uint64_t strides[1] = {ldw_bytes};
    uint32_t box[2] = {64, (uint32_t)bn};
    int rc = make_tensor_map_bf16(&ka.tmB, d.W, 2, dims, strides, box, 1);
    if (rc) return rc;
  }

  const double rows = static_cast<double>(ka.m_tiles ? (d.a_mode == 1 ? (double)d.cT * d.cH * d.cW : d.M) : 0) * d.batch;
  const int pid = prof_begin(d.a_mode == 1 ? PROF_CONV : PROF_GEMM, 2.0 * rows * d.N * K,
                             2.0 * (rows * K / (d.a_mode == 1 ? d.kt * d.kh * d.kw : 1) + (double)d.N * K + rows * d.N),
                             stream);
  int rc;
  switch (bn) {
    case 256:
      rc = launch_bn<256>(ka, stream);
      break;
    case 128:
      rc = launch_bn<128>(ka, stream);
      break;
    default:
      rc = launch_bn<64>(ka, stream);
      break;
  }
  prof_end(pid, stream);
  return rc;
}

}  // namespace dv

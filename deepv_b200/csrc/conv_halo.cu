// deepv_b200 — 3x3x3 stride-1 conv3d: one pixel tile + halo per nine taps, 128 output channels per tile.
//
// The generic implicit-GEMM path (gemm.cu) pulls a fresh 32 KB pixel tile AND a 16 KB weight tile from L2
// for every (tap, 64-channel) k-block: 85 FLOP per L2 byte, which at the chip's L2 throughput caps these
// convs near 1.1 PFLOP/s.  The nine spatial taps of one frame read the same pixels shifted
// by at most one row / column, so here ONE TMA box of (32+2) x (8+2) pixels x 64 channels lands in shared
// memory (row r = y * 10 + x, 128 B per row, 128B-swizzled) and every tap is a ROW-SHIFTED view of it:
//   B descriptor start = tile + (dy * 10 + dx) * 128 B, 8-row groups 1280 B apart (one image row each).
// tcgen05.mma applies the swizzle to absolute shared-memory address bits, so the shifted view reads what TMA
// wrote (checked on B200 by scripts/probe/desc_shift.cu for every shift and for the 1280 B group stride).
// L2 traffic per nine k-blocks: 43.5 KB pixels + 9 x 16 KB weights instead of 9 x 48 KB.
//
// Roles as in gemm_tc_kernel: warp 0 TMA producer (pixel ring + weight ring), warp 1 MMA issuer
// (D[128 output channels][256 pixels] in TMEM, double-buffered), warps 2-9 epilogue (bias, residual, bf16
// NDHWC store with the pixel-shuffle / frame-interleave address maps of vae.py:382,407-409, fused GroupNorm
// statistics).  Convs with more than 128 output channels run one tile per 128-channel chunk; the chunks of
// one pixel tile are neighbours in the tile order, so their pixel loads hit L2 together.  Reference op: CausalConv3d, vae.py:225-252 (zero causal / spatial padding =
// TMA out-of-bounds fill).
#include <cstdio>
#include <cstdlib>

#include "gemm.cuh"

namespace dv {
namespace {

constexpr int kTileW = 8, kTileH = 32;                 // output pixels per tile (N = 256)
constexpr int kBoxW = kTileW + 2, kBoxH = kTileH + 2;  // with the 1-pixel halo
constexpr int kPixBytes = kBoxW * kBoxH * 128;         // 43520
constexpr int kPixStride = 44 * 1024;                  // ring slot, 1024-aligned
constexpr int kPixStages = 2;
constexpr int kWBytes = 128 * 128;                     // 128 weight rows x 64 channels bf16
constexpr int kWStages = 6;
constexpr int kHaloThreads = 320;
constexpr int kHaloEpiWarps = 8;
constexpr int kHaloBars = 2 * kPixStages + 2 * kWStages + 4;
constexpr int kHaloSmem = kPixStages * kPixStride + kWStages * kWBytes + kHaloBars * 8 + 16 + 1024;

struct HaloArgs {
  CUtensorMap tmPix, tmW;
  GemmDesc d;
  int T, H, W, tiles_w, tiles_h, c_blocks, tiles, n_chunks, pdl_early;
  int t0;   // first output frame computed (GemmDesc::conv_t0)
};

__device__ __forceinline__ void decode(const HaloArgs& a, int tile, int* b, int* t, int* h0, int* w0, int* chunk) {
  *chunk = tile % a.n_chunks;
  tile /= a.n_chunks;
  const int per_frame = a.tiles_w * a.tiles_h;
  const int r = tile % per_frame;
  const int f = tile / per_frame;
  *w0 = (r % a.tiles_w) * kTileW;
  *h0 = (r / a.tiles_w) * kTileH;
  const int nt = a.T - a.t0;
  *t = a.t0 + f % nt;
  *b = f / nt;
}

__global__ void __launch_bounds__(kHaloThreads, 1) conv_halo_kernel(const __grid_constant__ HaloArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* pix = smem;
  uint8_t* wgt = smem + kPixStages * kPixStride;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wgt + kWStages * kWBytes);
  uint64_t* full_p = bars;
  uint64_t* empty_p = full_p + kPixStages;
  uint64_t* full_w = empty_p + kPixStages;
  uint64_t* empty_w = full_w + kWStages;
  uint64_t* tmem_full = empty_w + kWStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  const GemmDesc& d = a.d;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.tmPix);
    tma_prefetch_desc(&a.tmW);
    for (int i = 0; i < kPixStages; ++i) {
      mbar_init(&full_p[i], 1);
      mbar_init(&empty_p[i], 1);
    }
    for (int i = 0; i < kWStages; ++i) {
      mbar_init(&full_w[i], 1);
      mbar_init(&empty_w[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 32 * kHaloEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  if (a.pdl_early) pdl_trigger();

  const int tile0 = static_cast<int>(blockIdx.x), tile_step = static_cast<int>(gridDim.x);
  if (warp == 0) {
    // ============================ TMA producer ============================
    // (the whole converged warp runs the loop; the w_ forms elect the issuing lane and keep operands in uniform registers)
    {
      int ps = 0, ws = 0;
      uint32_t pph = 0, wph = 0;
      for (int tile = tile0; tile < a.tiles; tile += tile_step) {
        int b, t, h0, w0, chunk;
        decode(a, tile, &b, &t, &h0, &w0, &chunk);
        for (int dt = 0; dt < 3; ++dt) {
          for (int cb = 0; cb < a.c_blocks; ++cb) {
            mbar_wait(&empty_p[ps], pph ^ 1);
            w_mbar_expect_tx(&full_p[ps], kPixBytes);
            // causal in time (two frames of zero history), centred in space: out-of-bounds = zero padding
            w_tma_load_5d(&a.tmPix, &full_p[ps], pix + ps * kPixStride, cb * 64, w0 - 1, h0 - 1, t + dt - 2, b);
            if (++ps == kPixStages) {
              ps = 0;
              pph ^= 1;
            }
            for (int s = 0; s < 9; ++s) {
              const int kb = (dt * 9 + s) * a.c_blocks + cb;   // weight K index: tap-major, then channel block
              mbar_wait(&empty_w[ws], wph ^ 1);
              w_mbar_expect_tx(&full_w[ws], kWBytes);
              w_tma_load_2d(&a.tmW, &full_w[ws], wgt + ws * kWBytes, kb * 64, chunk * 128);
              if (++ws == kWStages) {
                ws = 0;
                wph ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ==============================
    {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 256, 0, 0);
      int ps = 0, ws = 0, it = 0;
      uint32_t pph = 0, wph = 0;
      for (int tile = tile0; tile < a.tiles; tile += tile_step, ++it) {
        const int as = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem_base + as * 256;
        bool first = true;
        for (int g = 0; g < 3 * a.c_blocks; ++g) {
          mbar_wait(&full_p[ps], pph);
          tc_fence_after();
          const uint32_t sp = smem_u32(pix + ps * kPixStride);
          for (int s = 0; s < 9; ++s) {
            mbar_wait(&full_w[ws], wph);
            tc_fence_after();
            const int dy = s / 3, dx = s - dy * 3;
            const uint64_t da = umma_desc_sw128(smem_u32(wgt + ws * kWBytes), 16, 1024);
            const uint64_t db = umma_desc_sw128(sp + (dy * kBoxW + dx) * 128, 16, kBoxW * 128);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              w_umma_bf16_ss(acc, da + 2 * k, db + 2 * k, idesc, !(first && k == 0));
            }
            first = false;
            w_umma_commit(&empty_w[ws]);
            if (++ws == kWStages) {
              ws = 0;
              wph ^= 1;
            }
          }
          w_umma_commit(&empty_p[ps]);
          if (++ps == kPixStages) {
            ps = 0;
            pph ^= 1;
          }
        }
        w_umma_commit(&tmem_full[as]);
      }
    }
  } else {
    // ============================ epilogue ================================
    const int quarter = warp & 3;       // TMEM lane quarter = 32 output channels
    const int half = (warp - 2) >> 2;   // which 128 pixels (16 image rows) of the tile
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(d.out);
    const __nv_bfloat16* res = reinterpret_cast<const __nv_bfloat16*>(d.residual);
    const int drop = d.conv_drop_first ? 1 : 0;
    int it = 0;
    for (int tile = tile0; tile < a.tiles; tile += tile_step, ++it) {
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      int b, t, h0, w0, chunk;
      decode(a, tile, &b, &t, &h0, &w0, &chunk);
      mbar_wait(&tmem_full[as], aph);
      tc_fence_after();
      const int n = chunk * 128 + quarter * 32 + lane;   // accumulator row = packed weight row
      if (chunk * 128 + quarter * 32 < d.N) {            // warp-uniform
        const bool n_ok = n < d.N;
        const float bias = (n_ok && d.bias != nullptr) ? __ldg(d.bias + n) : 0.f;
        // store map (vae.py:382,407-409): plain [b][t][h][w][n]; pixel shuffle n = (p1, p2, c) ->
        // [b][t][2h+p1][2w+p2][c]; frame interleave n = (p, c) -> [b][2t+p-drop][h][w][c]
        int oc = n, ot = t, oT = a.T, oH = a.H, oW = a.W, sh = 1, p1 = 0, p2 = 0;
        if (d.conv_store == CONV_SHUFFLE_HW) {
          const int q = n / d.out_C;
          oc = n % d.out_C;
          p1 = q >> 1;
          p2 = q & 1;
          sh = 2;
          oH = 2 * a.H;
          oW = 2 * a.W;
        } else if (d.conv_store == CONV_INTERLEAVE_T) {
          oc = n % d.out_C;
          ot = 2 * t + n / d.out_C - drop;
          oT = 2 * a.T - drop;
        }
        const bool live = n_ok && ot >= 0;
        const uint32_t taddr = tmem_base + as * 256 + (static_cast<uint32_t>(quarter * 32) << 16) + half * 128;
        const long long frame_off = (static_cast<long long>(b) * oT + ot) * oH;
        const long long px = static_cast<long long>(sh) * d.out_C;   // stride between neighbouring tile pixels
        float gsum = 0.f, gsq = 0.f;
        uint32_t buf[2][32];
        tmem_ld_32x32(taddr, buf[0]);
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {  // 32 columns = 4 image rows x 8 pixels
          tmem_ld_wait();
          if (ci + 1 < 4) tmem_ld_32x32(taddr + (ci + 1) * 32, buf[(ci + 1) & 1]);
#pragma unroll
          for (int hr = 0; hr < 4; ++hr) {
            const int oh = h0 + half * 16 + ci * 4 + hr;
            if (oh >= a.H || !live) continue;
            const long long row_off = ((frame_off + sh * oh + p1) * oW + sh * w0 + p2) * d.out_C + oc;
            float v[kTileW];
#pragma unroll
            for (int i = 0; i < kTileW; ++i) v[i] = __uint_as_float(buf[ci & 1][hr * kTileW + i]) + bias;
            if (res != nullptr) {
#pragma unroll
              for (int i = 0; i < kTileW; ++i) v[i] += __bfloat162float(res[row_off + i * px]);
            }
#pragma unroll
            for (int i = 0; i < kTileW; ++i) {
              const __nv_bfloat16 q = __float2bfloat16(v[i]);
              out[row_off + i * px] = q;
              const float x = __bfloat162float(q);
              gsum += x;
              gsq += x * x;
            }
          }
        }
        if (d.gn_acc != nullptr) {
          // channels of a group sit on adjacent lanes: fold them, one fp64 atomic pair per group
          for (int o = 1; o < d.gn_cpg; o <<= 1) {
            gsum += __shfl_xor_sync(0xffffffffu, gsum, o);
            gsq += __shfl_xor_sync(0xffffffffu, gsq, o);
          }
          if (live && (oc % d.gn_cpg) == 0) {
            double* accp = d.gn_acc + static_cast<long long>(blockIdx.x % d.gn_replicas) * d.gn_replica_stride +
                           (static_cast<long long>(ot) * (d.out_C / d.gn_cpg) + oc / d.gn_cpg) * 2;
            atomicAdd(accp, static_cast<double>(gsum));
            atomicAdd(accp + 1, static_cast<double>(gsq));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[as]);
    }
  }

  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

// Returns 1 when the problem is not one this kernel takes (the caller falls back to the generic path).
int launch_conv_halo(const GemmDesc& d, cudaStream_t stream) {
  static const bool off = getenv("DV_CONV_NOHALO") != nullptr;
  if (off) return 1;
  if (d.a_mode != 1 || d.mode != EPI_CONV || d.w_batch_stride != 0) return 1;
  if (d.N > 128 && (d.N % 32 != 0 || d.out_C % 32 != 0)) return 1;
  if (d.conv_store != CONV_PLAIN && (d.out_C % 32 != 0 || d.residual != nullptr)) return 1;
  if (d.kt != 3 || d.kh != 3 || d.kw != 3 || d.sT > 1 || d.sH > 1 || d.sW > 1) return 1;
  if (d.cC % 64 != 0 || d.cW % kTileW != 0 || d.batch <= 0) return 1;
  if (d.gn_acc != nullptr && d.batch != 1) return 1;   // (the statistics are indexed by frame only)
  HaloArgs a;
  a.d = d;
  a.T = d.cT;
  a.H = d.cH;
  a.W = d.cW;
  a.tiles_w = d.cW / kTileW;
  a.tiles_h = (d.cH + kTileH - 1) / kTileH;
  a.c_blocks = d.cC / 64;
  a.n_chunks = (d.N + 127) / 128;
  a.pdl_early = pdl_early() ? 1 : 0;
  a.t0 = d.conv_t0;
  DV_REQUIRE(a.t0 >= 0 && a.t0 < d.cT, "conv: first frame %d of %d", a.t0, d.cT);
  const int nT = d.cT - a.t0;   // frames computed
  const long long tiles = static_cast<long long>(d.batch) * nT * a.tiles_w * a.tiles_h * a.n_chunks;
  if (tiles < sm_count() || tiles >= (1ll << 30)) return 1;   // small problems: the generic path splits K
  {
    // rows padded up to the 32-row tile and the last, partly filled wave are wasted work here, while the
    // generic path has 8/16-row tiles and split-K: measured break-even near 60 % useful work
    const int nsm = sm_count();
    const double waves = static_cast<double>((tiles + nsm - 1) / nsm);
    const double useful = (static_cast<double>(d.cH) / (a.tiles_h * kTileH)) * (static_cast<double>(tiles) / (waves * nsm));
    if (useful < 0.6) return 1;
  }
  a.tiles = static_cast<int>(tiles);
  {
    uint64_t dims[5] = {(uint64_t)d.cC, (uint64_t)d.cW, (uint64_t)d.cH, (uint64_t)d.cT, (uint64_t)d.batch};
    uint64_t strides[4] = {(uint64_t)d.cC * 2, (uint64_t)d.cC * d.cW * 2, (uint64_t)d.cC * d.cW * d.cH * 2,
                           (uint64_t)d.cC * d.cW * d.cH * d.cT * 2};
    uint32_t box[5] = {64, (uint32_t)kBoxW, (uint32_t)kBoxH, 1, 1};
    int rc = make_tensor_map_bf16(&a.tmPix, d.A, 5, dims, strides, box, 1);
    if (rc) return rc;
  }
  {
    const int K = 27 * d.cC;
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)d.w_rows};
    uint64_t strides[1] = {(uint64_t)(d.ldw > 0 ? d.ldw : K) * 2};
    uint32_t box[2] = {64, 128};
    int rc = make_tensor_map_bf16(&a.tmW, d.W, 2, dims, strides, box, 1);
    if (rc) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    DV_CHECK_CUDA(cudaFuncSetAttribute(conv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHaloSmem));
    attr_set = true;
  }
  const int grid = a.tiles < sm_count() ? a.tiles : sm_count();
  const double pixels = static_cast<double>(d.batch) * nT * d.cH * d.cW;
  const double flops = 2.0 * pixels * 27.0 * d.cC * d.N;
  const double bytes = 2.0 * (pixels * d.cC + 27.0 * d.cC * d.N + pixels * d.N);
  char tag[56] = "";
  if (prof_on())
    snprintf(tag, sizeof(tag), "conv T%d H%d W%d Ci%d N%d k3 halo e%d", nT, d.cH, d.cW, d.cC, d.N, d.conv_store);
  const int pid = prof_begin(PROF_CONV, flops, bytes, stream, tag);
  DV_CHECK_CUDA(launch_pdl(conv_halo_kernel, dim3(grid), dim3(kHaloThreads), kHaloSmem, stream, 1, a));
  prof_end(pid, stream);
  DV_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return 0;
}

}  // namespace dv

// deepv_b200 — host-side runtime helpers: error state, TMA descriptor encoding, device info.
#include "common.cuh"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace dv {

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }
void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }
void launch_count_reset() { g_launches.store(0, std::memory_order_relaxed); }

// ---------------------------------------------------------------------------------------
// Optional per-launch profiler: CUDA event pairs on the launching stream around every kernel
// of a class, with the algorithmic FLOPs / bytes of that launch.  Off by default (zero cost);
// bench.py switches it on for one step to get roofline.achieved of the dominant kernel.
// ---------------------------------------------------------------------------------------
namespace {
struct ProfRec {
  cudaEvent_t e0, e1;
  int kind;
  double flops, bytes;
  char tag[56];
};
constexpr int kProfMax = 1 << 16;
bool g_prof_on = false;
std::vector<ProfRec>* g_prof = nullptr;
std::vector<cudaEvent_t>* g_prof_pool = nullptr;
cudaEvent_t prof_event() {
  if (g_prof_pool && !g_prof_pool->empty()) {
    cudaEvent_t e = g_prof_pool->back();
    g_prof_pool->pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

void prof_enable(bool on) {
  if (!g_prof) g_prof = new std::vector<ProfRec>();
  if (!g_prof_pool) g_prof_pool = new std::vector<cudaEvent_t>();
  g_prof_on = on;
}
void prof_reset() {
  if (!g_prof) return;
  for (auto& r : *g_prof) {
    g_prof_pool->push_back(r.e0);
    g_prof_pool->push_back(r.e1);
  }
  g_prof->clear();
}
bool prof_on() { return g_prof_on; }
int prof_begin(int kind, double flops, double bytes, cudaStream_t st, const char* tag) {
  if (!g_prof_on || static_cast<int>(g_prof->size()) >= kProfMax) return -1;
  ProfRec r;
  snprintf(r.tag, sizeof(r.tag), "%s", tag ? tag : "");
  r.e0 = prof_event();
  r.e1 = prof_event();
  r.kind = kind;
  r.flops = flops;
  r.bytes = bytes;
  cudaEventRecord(r.e0, st);
  g_prof->push_back(r);
  return static_cast<int>(g_prof->size()) - 1;
}
void prof_end(int id, cudaStream_t st) {
  if (id >= 0) cudaEventRecord((*g_prof)[id].e1, st);
}
// caller must have synchronised the stream(s)
int prof_summary(int kind, long long* count, double* ms, double* flops, double* bytes) {
  *count = 0;
  *ms = *flops = *bytes = 0;
  if (!g_prof) return 0;
  for (auto& r : *g_prof) {
    if (r.kind != kind) continue;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.e0, r.e1) != cudaSuccess) {
      set_error("prof_summary: events not complete (synchronise first)");
      return -2;
    }
    *count += 1;
    *ms += t;
    *flops += r.flops;
    *bytes += r.bytes;
  }
  return 0;
}

// one CSV line per recorded launch: kind,tag,flops,bytes,ms (caller synchronised the stream)
int prof_dump(const char* path) {
  FILE* f = fopen(path, "w");
  if (!f) {
    set_error("prof_dump: cannot open %s", path);
    return -1;
  }
  fprintf(f, "kind,tag,flops,bytes,ms\n");
  if (g_prof) {
    for (auto& r : *g_prof) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, r.e0, r.e1) != cudaSuccess) t = -1.f;
      fprintf(f, "%d,%s,%.0f,%.0f,%.6f\n", r.kind, r.tag, r.flops, r.bytes, t);
    }
  }
  fclose(f);
  return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Tensor maps are pure functions of (base, shape, strides, box): plans call the same GEMMs with the
// same buffers thousands of times per step, so the encoded descriptors are memoised (an encode
// through the driver costs ~1 us, a launch carries up to six of them).
namespace {
struct TmKey {
  const void* base;
  int rank, swz;
  uint64_t dims[5], strides[4];
  uint32_t box[5];
  bool operator==(const TmKey& o) const { return memcmp(this, &o, sizeof(TmKey)) == 0; }
};
struct TmHash {
  size_t operator()(const TmKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(TmKey) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
    return static_cast<size_t>(h);
  }
};
static_assert(sizeof(TmKey) % 8 == 0, "TmKey is hashed as 64-bit words");
std::mutex g_tm_mutex;
std::unordered_map<TmKey, CUtensorMap, TmHash>* g_tm_cache = nullptr;
}  // namespace

static int encode_tensor_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                             const uint64_t* strides_bytes, const uint32_t* box, int swizzle, int f32);

int make_tensor_map_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box, int swizzle128) {
  return make_tensor_map(out, base, 0, rank, dims, strides_bytes, box, swizzle128 ? 1 : 0);
}

// swizzle: 0 none, 1 = 128 B, 2 = 64 B; f32: 0 = bf16 elements, 1 = fp32 elements
int make_tensor_map(CUtensorMap* out, const void* base, int f32, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, int swizzle) {
  TmKey key;
  memset(&key, 0, sizeof(key));
  key.base = base;
  key.rank = rank;
  key.swz = swizzle | (f32 << 8);
  for (int i = 0; i < rank && i < 5; ++i) {
    key.dims[i] = dims[i];
    key.box[i] = box[i];
    if (i + 1 < rank) key.strides[i] = strides_bytes[i];
  }
  {
    std::lock_guard<std::mutex> lock(g_tm_mutex);
    if (!g_tm_cache) g_tm_cache = new std::unordered_map<TmKey, CUtensorMap, TmHash>();
    auto it = g_tm_cache->find(key);
    if (it != g_tm_cache->end()) {
      memcpy(out, &it->second, sizeof(CUtensorMap));
      return 0;
    }
  }
  int rc = encode_tensor_map(out, base, rank, dims, strides_bytes, box, swizzle, f32);
  if (rc) return rc;
  std::lock_guard<std::mutex> lock(g_tm_mutex);
  if (g_tm_cache->size() > (1u << 16)) g_tm_cache->clear();  // unbounded callers (sweeps): start over
  (*g_tm_cache)[key] = *out;
  return 0;
}

static int encode_tensor_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                             const uint64_t* strides_bytes, const uint32_t* box, int swizzle, int f32) {
  EncodeTiledFn enc = get_encode();
  DV_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver / device?)");
  DV_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base %p not 16-byte aligned", base);
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    DV_REQUIRE(box[i] >= 1 && box[i] <= 256, "TMA box[%d]=%u out of range", i, box[i]);
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides_bytes[i];
    DV_REQUIRE(strides_bytes[i] % 16 == 0, "TMA stride[%d]=%llu not a multiple of 16 bytes", i,
               (unsigned long long)strides_bytes[i]);
  }
  const CUtensorMapSwizzle sw = swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : (swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE);
  CUresult r = enc(out, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank,
                   const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DV_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r,
             rank);
  return 0;
}

bool pdl_enabled() {
  static const bool on = getenv("DV_NO_PDL") == nullptr;
  return on;
}

// griddepcontrol.launch_dependents right after griddepcontrol.wait (DV_PDL_LATE=1: at the end of the kernel, the
// round-1 placement).  The trigger only lets the successor's CTAs be SCHEDULED once every CTA of this grid has started;
// they run their prologue (barrier init, TMEM allocation, descriptor prefetch) on idle / finished SMs and still block in
// their own griddepcontrol.wait until this grid has completed and flushed.
bool pdl_early() {
  static const bool on = getenv("DV_PDL_LATE") == nullptr;
  return on;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

}  // namespace dv

// Host-pointer entry points of libmmsbm_b200: what a ctypes stub in the reference would bind
// (INTEGRATION.md).  They own their device memory, copy in, run the device-level calls of
// this library on a private stream, copy out and synchronise.  No CPU arithmetic happens
// here: even the int64 -> int32 split and the row padding run on the device.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace mmsbm {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }

__global__ void split_triples_kernel(const int64_t* data, int64_t n, int U, int I, int R,
                                     int32_t* u, int32_t* i, int32_t* r, int* bad) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int64_t a = data[3 * t], b = data[3 * t + 1], c = data[3 * t + 2];
  if (a < 0 || a >= U || b < 0 || b >= I || (R > 0 && (c < 0 || c >= R))) atomicExch(bad, 1);
  u[t] = (int32_t)a; i[t] = (int32_t)b; r[t] = (int32_t)c;
}

// [rows][w] compact <-> [rows][ld] zero padded
__global__ void pad_rows_kernel(const double* src, double* dst, size_t rows, int w, int ld) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * ld) return;
  size_t row = t / ld; int c = (int)(t - row * ld);
  dst[t] = (c < w) ? src[row * w + c] : 0.0;
}
__global__ void unpad_rows_kernel(const double* src, double* dst, size_t rows, int w, int ld) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * w) return;
  size_t row = t / w; int c = (int)(t - row * w);
  dst[t] = src[row * ld + c];
}

// A memory pool OWNED by this library, one per device: the host calls allocate from it
// (stream-ordered) and it stays warm between calls, so a loop of plugin calls does not pay
// cudaMalloc every time.  The device's default pool -- which other users of the process share,
// e.g. PyTorch -- is left untouched.  At the end of a call the pool is trimmed to
// MMSBM_POOL_KEEP_MB (default 4096): what a call of the largest shapes allocated goes back to the
// driver instead of being retained for ever.
static std::mutex g_pool_mutex;
static cudaMemPool_t g_pools[64] = {nullptr};

static int library_pool(int dev, cudaMemPool_t* out) {
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  MMSBM_REQUIRE(dev >= 0 && dev < 64, MMSBM_EINVAL, "device ordinal %d out of range", dev);
  if (!g_pools[dev]) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    MMSBM_CUDA(cudaMemPoolCreate(&g_pools[dev], &props));
    uint64_t keep = UINT64_MAX;                  // retained while a call runs; trimmed at its end
    MMSBM_CUDA(cudaMemPoolSetAttribute(g_pools[dev], cudaMemPoolAttrReleaseThreshold, &keep));
  }
  *out = g_pools[dev];
  return 0;
}

// everything one host call allocates, freed on scope exit
struct Scope {
  std::vector<void*> ptrs;
  cudaStream_t st = nullptr;
  cudaMemPool_t pool = nullptr;
  ~Scope() {
    if (st) {
      for (void* p : ptrs) cudaFreeAsync(p, st);
      cudaStreamSynchronize(st);
      if (pool) {
        const char* e = getenv("MMSBM_POOL_KEEP_MB");
        const size_t keep_mb = (e && *e) ? (size_t)atoll(e) : 4096;
        cudaMemPoolTrimTo(pool, keep_mb << 20);
      }
      cudaStreamDestroy(st);
    }
  }
  template <typename T>
  int alloc(T** out, size_t count) {
    void* p = nullptr;
    MMSBM_CUDA(cudaMallocFromPoolAsync(&p, count ? count * sizeof(T) : 16, pool, st));
    ptrs.push_back(p);
    *out = static_cast<T*>(p);
    return 0;
  }
  int open() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
      set_error("no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
      return MMSBM_ENODEV;
    }
    int dev = 0;
    MMSBM_CUDA(cudaGetDevice(&dev));
    int rc = library_pool(dev, &pool);
    if (rc) return rc;
    MMSBM_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    return 0;
  }
};

#define TRY(expr) do { int rc__ = (expr); if (rc__) return rc__; } while (0)

// MMSBM_TRACE=1: wall-clock of the stages of a host call on stderr (synchronises between stages)
struct Trace {
  bool on;
  cudaStream_t st;
  std::chrono::steady_clock::time_point t0;
  explicit Trace(cudaStream_t s) : on(getenv("MMSBM_TRACE") != nullptr), st(s), t0(std::chrono::steady_clock::now()) {}
  void mark(const char* what) {
    if (!on) return;
    cudaStreamSynchronize(st);
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[mmsbm] %-22s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

struct DevTriples { int32_t *u, *i, *r; };

static int upload_triples(Scope& sc, const int64_t* data, int64_t N, int U, int I, int R, DevTriples* out) {
  int64_t* raw; int* bad;
  TRY(sc.alloc(&raw, (size_t)N * 3));
  TRY(sc.alloc(&out->u, (size_t)N)); TRY(sc.alloc(&out->i, (size_t)N)); TRY(sc.alloc(&out->r, (size_t)N));
  TRY(sc.alloc(&bad, 1));
  MMSBM_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), sc.st));
  if (N > 0) {
    MMSBM_CUDA(cudaMemcpyAsync(raw, data, (size_t)N * 24, cudaMemcpyHostToDevice, sc.st));
    split_triples_kernel<<<(unsigned)((N + 255) / 256), 256, 0, sc.st>>>(raw, N, U, I, R, out->u, out->i, out->r, bad);
    MMSBM_LAUNCH_CHECK("split_triples_kernel");
  }
  int h = 0;
  MMSBM_CUDA(cudaMemcpyAsync(&h, bad, sizeof(int), cudaMemcpyDeviceToHost, sc.st));
  MMSBM_CUDA(cudaStreamSynchronize(sc.st));
  MMSBM_REQUIRE(h == 0, MMSBM_EINVAL, "data holds an id outside [0,U) x [0,I) x [0,R)");
  return 0;
}

// host compact [rows][w] -> device padded [rows][ld]
static int upload_rows(Scope& sc, const double* src, size_t rows, int w, double** out) {
  const int ld = row_stride(w);
  TRY(sc.alloc(out, rows * ld));
  if (ld == w) {
    MMSBM_CUDA(cudaMemcpyAsync(*out, src, rows * w * 8, cudaMemcpyHostToDevice, sc.st));
    return 0;
  }
  double* tmp;
  TRY(sc.alloc(&tmp, rows * w));
  MMSBM_CUDA(cudaMemcpyAsync(tmp, src, rows * w * 8, cudaMemcpyHostToDevice, sc.st));
  pad_rows_kernel<<<(unsigned)((rows * ld + 255) / 256), 256, 0, sc.st>>>(tmp, *out, rows, w, ld);
  MMSBM_LAUNCH_CHECK("pad_rows_kernel");
  return 0;
}

static int download_rows(Scope& sc, const double* dev, size_t rows, int w, double* dst) {
  const int ld = row_stride(w);
  if (ld == w) {
    MMSBM_CUDA(cudaMemcpyAsync(dst, dev, rows * w * 8, cudaMemcpyDeviceToHost, sc.st));
    return 0;
  }
  double* tmp;
  TRY(sc.alloc(&tmp, rows * w));
  unpad_rows_kernel<<<(unsigned)((rows * w + 255) / 256), 256, 0, sc.st>>>(dev, tmp, rows, w, ld);
  MMSBM_LAUNCH_CHECK("unpad_rows_kernel");
  MMSBM_CUDA(cudaMemcpyAsync(dst, tmp, rows * w * 8, cudaMemcpyDeviceToHost, sc.st));
  return 0;
}

struct DevGraph { int32_t *useg, *uadj, *uperm, *udeg, *iseg, *iadj, *iperm, *ideg, *usched, *isched; };

static int build_graph(Scope& sc, const DevTriples& t, int64_t N, int U, int I, int R, DevGraph* g) {
  TRY(sc.alloc(&g->useg, (size_t)U * R + 1)); TRY(sc.alloc(&g->uadj, (size_t)N));
  TRY(sc.alloc(&g->uperm, (size_t)N)); TRY(sc.alloc(&g->udeg, (size_t)U));
  TRY(sc.alloc(&g->iseg, (size_t)I * R + 1)); TRY(sc.alloc(&g->iadj, (size_t)N));
  TRY(sc.alloc(&g->iperm, (size_t)N)); TRY(sc.alloc(&g->ideg, (size_t)I));
  int64_t su = 0, si = 0;
  TRY(mmsbm_sched_elems(N, U, &su)); TRY(mmsbm_sched_elems(N, I, &si));
  TRY(sc.alloc(&g->usched, (size_t)su)); TRY(sc.alloc(&g->isched, (size_t)si));
  size_t wsb = 0;
  TRY(mmsbm_graph_workspace_bytes(N, U, I, R, &wsb));
  char* ws;
  TRY(sc.alloc(&ws, wsb));
  return mmsbm_graph_build(t.u, t.i, t.r, N, U, I, R, g->useg, g->uadj, g->uperm, g->udeg, g->iseg,
                           g->iadj, g->iperm, g->ideg, g->usched, g->isched, ws, wsb, sc.st);
}

// ---- index cache of the plugin calls ------------------------------------------------------------
// The reference's loop (src/mmsbm.py:243-250) calls update_coefficients with the SAME encoded
// array hundreds of times, and the plugin contract (src/backend.py:22) is stateless.  The last
// index structure is therefore kept on the device, keyed on (device, N, U, I, R, a 64-bit
// checksum of all 24*N bytes): a hit skips the H2D copy of the rows and both sorts.  The key is
// the CONTENT: an array modified in place misses, a copy of the same rows hits.  One entry per process (the
// reference's pool workers are processes); MMSBM_INDEX_CACHE=0 disables it.
struct IndexCache {
  bool valid = false;
  int dev = -1;
  const void* data = nullptr;
  int64_t N = 0;
  int U = 0, I = 0, R = 0;
  uint64_t sum = 0;
  DevGraph g{};
  std::vector<void*> owned;
  int64_t hits = 0, misses = 0;
};
static std::mutex g_cache_mutex;
static IndexCache g_cache;

static uint64_t checksum64(const int64_t* p, int64_t words) {
  uint64_t a = 0x9e3779b97f4a7c15ull, b = 0, c = 0, d = 0;
  int64_t k = 0;
  for (; k + 4 <= words; k += 4) {               // four independent lanes: vectorises, ~memory speed
    a = (a ^ (uint64_t)p[k]) * 0x100000001b3ull;
    b = (b ^ (uint64_t)p[k + 1]) * 0x100000001b3ull + 0x632be59bd9b4e019ull;
    c = (c ^ (uint64_t)p[k + 2]) * 0xff51afd7ed558ccdull;
    d = (d ^ (uint64_t)p[k + 3]) * 0xc4ceb9fe1a85ec53ull + 1;
  }
  for (; k < words; ++k) a = (a ^ (uint64_t)p[k]) * 0x100000001b3ull;
  return a ^ (b << 1) ^ (c << 2) ^ (d << 3);
}

static void cache_drop_locked() {
  for (void* p : g_cache.owned) cudaFree(p);
  g_cache.owned.clear();
  g_cache.valid = false;
}

// the index of `data` on the current device: from the cache, or built (and then cached)
static int cached_graph(Scope& sc, const int64_t* data, int64_t N, int U, int I, int R, DevGraph* out,
                        bool* from_cache) {
  *from_cache = false;
  const char* e = getenv("MMSBM_INDEX_CACHE");
  const bool enabled = !(e && *e == '0') && N > 0;
  int dev = 0;
  MMSBM_CUDA(cudaGetDevice(&dev));
  uint64_t sum = 0;
  if (enabled) {
    sum = checksum64(data, 3 * N);
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    if (g_cache.valid && g_cache.dev == dev && g_cache.N == N && g_cache.U == U &&
        g_cache.I == I && g_cache.R == R && g_cache.sum == sum) {
      *out = g_cache.g;
      *from_cache = true;
      ++g_cache.hits;
      return 0;
    }
  }
  DevTriples t;
  TRY(upload_triples(sc, data, N, U, I, R, &t));
  if (!enabled) return build_graph(sc, t, N, U, I, R, out);
  // build into memory the cache owns (plain cudaMalloc: it outlives the call's scope)
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  cache_drop_locked();
  ++g_cache.misses;
  DevGraph g{};
  int64_t su = 0, si = 0;
  TRY(mmsbm_sched_elems(N, U, &su)); TRY(mmsbm_sched_elems(N, I, &si));
  struct Want { int32_t** p; size_t n; } want[] = {
      {&g.useg, (size_t)U * R + 1}, {&g.uadj, (size_t)N}, {&g.uperm, (size_t)N}, {&g.udeg, (size_t)U},
      {&g.iseg, (size_t)I * R + 1}, {&g.iadj, (size_t)N}, {&g.iperm, (size_t)N}, {&g.ideg, (size_t)I},
      {&g.usched, (size_t)su}, {&g.isched, (size_t)si}};
  for (auto& w : want) {
    void* p = nullptr;
    cudaError_t ce = cudaMalloc(&p, (w.n ? w.n : 4) * sizeof(int32_t));
    if (ce != cudaSuccess) {
      cache_drop_locked();
      set_error("cudaMalloc of the cached index failed: %s", cudaGetErrorString(ce));
      return (int)ce;
    }
    g_cache.owned.push_back(p);
    *w.p = static_cast<int32_t*>(p);
  }
  size_t wsb = 0;
  TRY(mmsbm_graph_workspace_bytes(N, U, I, R, &wsb));
  char* ws;
  TRY(sc.alloc(&ws, wsb));
  int rc = mmsbm_graph_build(t.u, t.i, t.r, N, U, I, R, g.useg, g.uadj, g.uperm, g.udeg, g.iseg, g.iadj, g.iperm,
                             g.ideg, g.usched, g.isched, ws, wsb, sc.st);
  if (rc == 0) {
    cudaError_t ce = cudaStreamSynchronize(sc.st);
    if (ce != cudaSuccess) { set_error("index build failed: %s", cudaGetErrorString(ce)); rc = (int)ce; }
  }
  if (rc) { cache_drop_locked(); return rc; }
  g_cache.valid = true; g_cache.dev = dev; g_cache.data = data; g_cache.N = N;
  g_cache.U = U; g_cache.I = I; g_cache.R = R; g_cache.sum = sum; g_cache.g = g;
  *out = g;
  return 0;
}

static int check_common(const void* data, int64_t N, const void* th, int U, int K, const void* et,
                        int I, int L, const void* pr, int R, const char* who) {
  MMSBM_REQUIRE(N >= 0 && U > 0 && I > 0 && K > 0 && L > 0 && R > 0, MMSBM_EINVAL, "%s: bad size", who);
  MMSBM_REQUIRE((N == 0 || data) && th && et && pr, MMSBM_EINVAL, "%s: null pointer", who);
  return 0;
}

}  // namespace mmsbm

using namespace mmsbm;

extern "C" int mmsbm_abi_version(void) { return MMSBM_ABI_VERSION; }
extern "C" const char* mmsbm_last_error(void) { return g_err; }
extern "C" int64_t mmsbm_launch_count(void) { return g_launches; }
extern "C" int mmsbm_row_stride(int32_t k) { return row_stride(k); }
extern "C" int mmsbm_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return MMSBM_ENODEV;
  }
  return n;
}

extern "C" int mmsbm_index_cache_stats(int64_t* hits, int64_t* misses) {
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  if (hits) *hits = g_cache.hits;
  if (misses) *misses = g_cache.misses;
  return 0;
}
extern "C" int mmsbm_index_cache_clear(void) {
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  cache_drop_locked();
  return 0;
}

extern "C" int mmsbm_split_triples(const int64_t* data, int64_t N, int32_t U, int32_t I, int32_t R,
                                   int32_t* u, int32_t* i, int32_t* r, int32_t* bad, void* stream) {
  MMSBM_REQUIRE(N >= 0 && U > 0 && I > 0 && bad && (N == 0 || (data && u && i && r)), MMSBM_EINVAL,
                "mmsbm_split_triples: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MMSBM_CUDA(cudaMemsetAsync(bad, 0, sizeof(int32_t), st));
  if (N > 0) {
    split_triples_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(data, N, U, I, R, u, i, r, bad);
    MMSBM_LAUNCH_CHECK("split_triples_kernel");
  }
  return 0;
}

extern "C" int mmsbm_host_compute_omegas(const int64_t* data, int64_t N, const double* theta, int32_t U,
                                         int32_t K, const double* eta, int32_t I, int32_t L,
                                         const double* pr, int32_t R, double* out) {
  TRY(check_common(data, N, theta, U, K, eta, I, L, pr, R, "mmsbm_host_compute_omegas"));
  MMSBM_REQUIRE(N == 0 || out, MMSBM_EINVAL, "mmsbm_host_compute_omegas: null output");
  Scope sc; TRY(sc.open());
  DevTriples t; TRY(upload_triples(sc, data, N, U, I, R, &t));
  double *dth, *det, *dpr, *dom;
  TRY(upload_rows(sc, theta, (size_t)U, K, &dth));
  TRY(upload_rows(sc, eta, (size_t)I, L, &det));
  TRY(sc.alloc(&dpr, (size_t)K * L * R));
  MMSBM_CUDA(cudaMemcpyAsync(dpr, pr, (size_t)K * L * R * 8, cudaMemcpyHostToDevice, sc.st));
  TRY(sc.alloc(&dom, (size_t)N * K * L));
  TRY(mmsbm_compute_omegas(t.u, t.i, t.r, N, K, L, R, dth, det, dpr, dom, sc.st));
  if (N > 0) MMSBM_CUDA(cudaMemcpyAsync(out, dom, (size_t)N * K * L * 8, cudaMemcpyDeviceToHost, sc.st));
  MMSBM_CUDA(cudaStreamSynchronize(sc.st));
  return 0;
}

extern "C" int mmsbm_host_update_coefficients(const int64_t* data, int64_t N, const double* theta,
                                              int32_t U, int32_t K, const double* eta, int32_t I,
                                              int32_t L, const double* pr, int32_t R, double* n_theta,
                                              double* n_eta, double* n_pr) {
  TRY(check_common(data, N, theta, U, K, eta, I, L, pr, R, "mmsbm_host_update_coefficients"));
  MMSBM_REQUIRE(n_theta && n_eta && n_pr, MMSBM_EINVAL, "mmsbm_host_update_coefficients: null output");
  Scope sc; TRY(sc.open());
  DevGraph g; bool hit = false;
  TRY(cached_graph(sc, data, N, U, I, R, &g, &hit));
  const int ldk = row_stride(K), ldl = row_stride(L);
  double *dth, *det, *dpr, *oth, *oet, *opr;
  TRY(upload_rows(sc, theta, (size_t)U, K, &dth));
  TRY(upload_rows(sc, eta, (size_t)I, L, &det));
  TRY(sc.alloc(&dpr, (size_t)K * L * R));
  MMSBM_CUDA(cudaMemcpyAsync(dpr, pr, (size_t)K * L * R * 8, cudaMemcpyHostToDevice, sc.st));
  TRY(sc.alloc(&oth, (size_t)U * ldk)); TRY(sc.alloc(&oet, (size_t)I * ldl));
  TRY(sc.alloc(&opr, (size_t)K * L * R));
  size_t wsb = 0; TRY(mmsbm_em_workspace_bytes(N, U, I, R, K, L, 1, &wsb));
  char* ws; TRY(sc.alloc(&ws, wsb));
  TRY(mmsbm_em_step(g.useg, g.uadj, g.udeg, g.iseg, g.iadj, g.ideg, g.usched, g.isched, N, U, I, R, K, L, 1, dth, det, dpr,
                    oth, oet, opr, MMSBM_RAW_THETA | MMSBM_RAW_ETA_PR, ws, wsb, sc.st));
  TRY(download_rows(sc, oth, (size_t)U, K, n_theta));
  TRY(download_rows(sc, oet, (size_t)I, L, n_eta));
  MMSBM_CUDA(cudaMemcpyAsync(n_pr, opr, (size_t)K * L * R * 8, cudaMemcpyDeviceToHost, sc.st));
  MMSBM_CUDA(cudaStreamSynchronize(sc.st));
  return 0;
}

extern "C" int mmsbm_host_prod_dist(const int64_t* data, int64_t M, const double* theta, int32_t U,
                                    int32_t K, const double* eta, int32_t I, int32_t L,
                                    const double* pr, int32_t R, double* rat) {
  TRY(check_common(data, M, theta, U, K, eta, I, L, pr, R, "mmsbm_host_prod_dist"));
  MMSBM_REQUIRE(M == 0 || rat, MMSBM_EINVAL, "mmsbm_host_prod_dist: null output");
  Scope sc; TRY(sc.open());
  DevTriples t; TRY(upload_triples(sc, data, M, U, I, /*R=*/0, &t));  // rating column is unused
  double *dth, *det, *dpr, *drat;
  TRY(upload_rows(sc, theta, (size_t)U, K, &dth));
  TRY(upload_rows(sc, eta, (size_t)I, L, &det));
  TRY(sc.alloc(&dpr, (size_t)K * L * R));
  MMSBM_CUDA(cudaMemcpyAsync(dpr, pr, (size_t)K * L * R * 8, cudaMemcpyHostToDevice, sc.st));
  TRY(sc.alloc(&drat, (size_t)M * R));
  TRY(mmsbm_prod_dist(t.u, t.i, M, U, I, R, K, L, 1, dth, det, dpr, drat, sc.st));
  if (M > 0) MMSBM_CUDA(cudaMemcpyAsync(rat, drat, (size_t)M * R * 8, cudaMemcpyDeviceToHost, sc.st));
  MMSBM_CUDA(cudaStreamSynchronize(sc.st));
  return 0;
}

extern "C" int mmsbm_host_likelihood(const int64_t* data, int64_t N, const double* theta, int32_t U,
                                     int32_t K, const double* eta, int32_t I, int32_t L,
                                     const double* pr, int32_t R, double* out) {
  TRY(check_common(data, N, theta, U, K, eta, I, L, pr, R, "mmsbm_host_likelihood"));
  MMSBM_REQUIRE(out, MMSBM_EINVAL, "mmsbm_host_likelihood: null output");
  Scope sc; TRY(sc.open());
  DevGraph g; bool hit = false;
  TRY(cached_graph(sc, data, N, U, I, R, &g, &hit));
  double *dth, *det, *dpr, *dout;
  TRY(upload_rows(sc, theta, (size_t)U, K, &dth));
  TRY(upload_rows(sc, eta, (size_t)I, L, &det));
  TRY(sc.alloc(&dpr, (size_t)K * L * R));
  MMSBM_CUDA(cudaMemcpyAsync(dpr, pr, (size_t)K * L * R * 8, cudaMemcpyHostToDevice, sc.st));
  TRY(sc.alloc(&dout, 1));
  size_t wsb = 0; TRY(mmsbm_likelihood_workspace_bytes(N, U, I, R, K, L, 1, &wsb));
  char* ws; TRY(sc.alloc(&ws, wsb));
  TRY(mmsbm_likelihood(g.useg, g.uadj, g.usched, N, U, I, R, K, L, 1, dth, det, dpr, dout, ws, wsb, sc.st));
  MMSBM_CUDA(cudaMemcpyAsync(out, dout, 8, cudaMemcpyDeviceToHost, sc.st));
  MMSBM_CUDA(cudaStreamSynchronize(sc.st));
  return 0;
}

extern "C" int mmsbm_host_fit(const int64_t* data, int64_t N, int32_t U, int32_t I, int32_t R, int32_t K,
                              int32_t L, int32_t S, int32_t iterations, const double* theta0,
                              const double* eta0, const double* pr0, double* theta_out, double* eta_out,
                              double* pr_out, double* lik_out) {
  TRY(check_common(data, N, theta0, U, K, eta0, I, L, pr0, R, "mmsbm_host_fit"));
  MMSBM_REQUIRE(S > 0 && iterations >= 0 && theta_out && eta_out && pr_out && lik_out, MMSBM_EINVAL,
                "mmsbm_host_fit: bad argument");
  Scope sc; TRY(sc.open());
  Trace tr(sc.st);
  DevTriples t; TRY(upload_triples(sc, data, N, U, I, R, &t));
  tr.mark("rows H2D + split");
  DevGraph g; TRY(build_graph(sc, t, N, U, I, R, &g));
  tr.mark("index build");
  const int ldk = row_stride(K), ldl = row_stride(L);
  const size_t prn = (size_t)S * K * L * R;
  double *tha, *eta_a, *pra, *thb, *etb, *prb, *dlik;
  TRY(upload_rows(sc, theta0, (size_t)S * U, K, &tha));
  TRY(upload_rows(sc, eta0, (size_t)S * I, L, &eta_a));
  TRY(sc.alloc(&pra, prn));
  MMSBM_CUDA(cudaMemcpyAsync(pra, pr0, prn * 8, cudaMemcpyHostToDevice, sc.st));
  TRY(sc.alloc(&thb, (size_t)S * U * ldk)); TRY(sc.alloc(&etb, (size_t)S * I * ldl));
  TRY(sc.alloc(&prb, prn)); TRY(sc.alloc(&dlik, (size_t)S));
  size_t wsb = 0, lwsb = 0;
  TRY(mmsbm_em_workspace_bytes(N, U, I, R, K, L, S, &wsb));
  // the likelihood runs after the EM loop and batches its runs to the workspace it is given:
  // one allocation serves both (at least the likelihood's minimum)
  TRY(mmsbm_likelihood_min_workspace_bytes(N, U, I, R, K, L, S, &lwsb));
  if (lwsb > wsb) wsb = lwsb;
  lwsb = wsb;
  char *ws, *lws;
  TRY(sc.alloc(&ws, wsb));
  lws = ws;
  tr.mark("params H2D + alloc");
  TRY(mmsbm_em_run(g.useg, g.uadj, g.udeg, g.iseg, g.iadj, g.ideg, g.usched, g.isched, N, U, I, R, K, L, S, iterations, tha,
                   eta_a, pra, thb, etb, prb, ws, wsb, sc.st));
  tr.mark("EM iterations");
  const bool in_a = (iterations % 2) == 0;
  double* th = in_a ? tha : thb; double* et = in_a ? eta_a : etb; double* pr = in_a ? pra : prb;
  TRY(mmsbm_likelihood(g.useg, g.uadj, g.usched, N, U, I, R, K, L, S, th, et, pr, dlik, lws, lwsb, sc.st));
  tr.mark("likelihood");
  TRY(download_rows(sc, th, (size_t)S * U, K, theta_out));
  TRY(download_rows(sc, et, (size_t)S * I, L, eta_out));
  MMSBM_CUDA(cudaMemcpyAsync(pr_out, pr, prn * 8, cudaMemcpyDeviceToHost, sc.st));
  MMSBM_CUDA(cudaMemcpyAsync(lik_out, dlik, (size_t)S * 8, cudaMemcpyDeviceToHost, sc.st));
  MMSBM_CUDA(cudaStreamSynchronize(sc.st));
  tr.mark("results D2H");
  return 0;
}

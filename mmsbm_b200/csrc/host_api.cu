// Host-pointer entry points of libmmsbm_b200: what a ctypes stub in the reference would bind
// (INTEGRATION.md).  They own their device memory, copy in, run the device-level calls of
// this library on a private stream, copy out and synchronise.  No CPU arithmetic happens
// here: even the int64 -> int32 split and the row padding run on the device.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
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

// everything one host call allocates: stream-ordered allocations from the device's default
// memory pool (kept warm between calls: release threshold = unlimited), freed on scope exit
struct Scope {
  std::vector<void*> ptrs;
  cudaStream_t st = nullptr;
  ~Scope() {
    if (st) {
      for (void* p : ptrs) cudaFreeAsync(p, st);
      cudaStreamSynchronize(st);
      cudaStreamDestroy(st);
    }
  }
  template <typename T>
  int alloc(T** out, size_t count) {
    void* p = nullptr;
    MMSBM_CUDA(cudaMallocAsync(&p, count ? count * sizeof(T) : 16, st));
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
    cudaMemPool_t pool;
    MMSBM_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    uint64_t keep = UINT64_MAX;
    MMSBM_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
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
  DevTriples t; TRY(upload_triples(sc, data, N, U, I, R, &t));
  DevGraph g; TRY(build_graph(sc, t, N, U, I, R, &g));
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
  DevTriples t; TRY(upload_triples(sc, data, N, U, I, R, &t));
  DevGraph g; TRY(build_graph(sc, t, N, U, I, R, &g));
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

// Micro-benchmark (not product code): how fast can one B200 gather 160-byte rows from an
// L2-resident table?  Variants: (A) per-lane cp.async.bulk into shared memory + LDS.128,
// (B) G lanes per row with LDG.128 (the em_step.cu layout), (C) one lane per row with
// 256-bit loads, (D) one lane per row with 128-bit loads.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo gather_bench.cu -o gather_bench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

constexpr int ROWD = 20;            // doubles per row
constexpr int ROWB = ROWD * 8;      // 160 bytes
constexpr int SROW = 176;           // padded smem row stride (11 x 16 B)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// (A) TMA bulk gather, double buffered per warp
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_tma(const double* __restrict__ table, const int* __restrict__ idx,
                                                    long n, double* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* tile = sm + (size_t)warp * 2 * 32 * SROW;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + (size_t)WARPS * 2 * 32 * SROW) + warp * 2;
  if (lane == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const long gw = (long)blockIdx.x * WARPS + warp, nw = (long)gridDim.x * WARPS;
  const long chunks = n / 32;
  double acc = 0.0;
  uint32_t phase[2] = {0, 0};
  long c = gw;
  auto issue = [&](long chunk, int buf) {
    int id = idx[chunk * 32 + lane];
    if (lane == 0) mbar_expect_tx(&bars[buf], 32 * ROWB);
    __syncwarp();
    bulk_g2s(tile + ((size_t)buf * 32 + lane) * SROW, table + (size_t)id * ROWD, ROWB, &bars[buf]);
  };
  int buf = 0;
  if (c < chunks) issue(c, 0);
  for (; c < chunks; c += nw) {
    long nxt = c + nw;
    if (nxt < chunks) issue(nxt, buf ^ 1);
    mbar_wait(&bars[buf], phase[buf]);
    phase[buf] ^= 1;
    const double2* row = reinterpret_cast<const double2*>(tile + ((size_t)buf * 32 + lane) * SROW);
#pragma unroll
    for (int k = 0; k < ROWD / 2; ++k) { double2 v = row[k]; acc = fma(v.x, 1.0000001, acc); acc = fma(v.y, 0.9999999, acc); }
    __syncwarp();
    buf ^= 1;
  }
  if (acc == 123.456) out[0] = acc;
  out[1 + blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// (B) G lanes per row, LDG.128, UN rows-steps in flight
template <int G, int UN>
__global__ void __launch_bounds__(256) k_group(const double* __restrict__ table, const int* __restrict__ idx, long n,
                                               double* out) {
  const int lane = threadIdx.x & 31;
  const int RPS = 32 / G, grp = lane / G, q = lane % G;
  const long gw = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long)gridDim.x * blockDim.x) >> 5;
  const long per = (long)UN * RPS, chunks = n / per;
  double acc = 0.0;
  constexpr int CH = (ROWD / 2 + G - 1) / G;
  for (long c = gw; c < chunks; c += nw) {
    int my = (lane < per) ? idx[c * per + lane] : 0;
    double2 x[UN][CH];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      int id = __shfl_sync(0xffffffffu, my, (u * RPS + grp) & 31);
      const double2* row = reinterpret_cast<const double2*>(table + (size_t)id * ROWD);
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        int chunk = k * G + q;
        x[u][k] = (grp < RPS && chunk < ROWD / 2) ? __ldg(row + chunk) : make_double2(0, 0);
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u)
#pragma unroll
      for (int k = 0; k < CH; ++k) { acc = fma(x[u][k].x, 1.0000001, acc); acc = fma(x[u][k].y, 0.9999999, acc); }
  }
  out[1 + blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__device__ __forceinline__ void ld256(const double* p, double& a, double& b, double& c, double& d) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

// (C)/(D) one lane per row
template <bool WIDE>
__global__ void __launch_bounds__(256) k_lane(const double* __restrict__ table, const int* __restrict__ idx, long n,
                                              double* out) {
  const long t0 = (long)blockIdx.x * blockDim.x + threadIdx.x, nt = (long)gridDim.x * blockDim.x;
  double acc = 0.0;
  for (long i = t0; i < n; i += nt) {
    const double* row = table + (size_t)idx[i] * ROWD;
    if (WIDE) {
      double v[ROWD];
#pragma unroll
      for (int k = 0; k < ROWD; k += 4) ld256(row + k, v[k], v[k + 1], v[k + 2], v[k + 3]);
#pragma unroll
      for (int k = 0; k < ROWD; ++k) acc = fma(v[k], 1.0000001, acc);
    } else {
      double2 v[ROWD / 2];
#pragma unroll
      for (int k = 0; k < ROWD / 2; ++k) v[k] = __ldg(reinterpret_cast<const double2*>(row) + k);
#pragma unroll
      for (int k = 0; k < ROWD / 2; ++k) { acc = fma(v[k].x, 1.0000001, acc); acc = fma(v[k].y, 0.9999999, acc); }
    }
  }
  out[1 + t0] = acc;
}


// (E) G lanes per row, ONE 256-bit load per lane (rows of 4*G doubles): the em kernel's layout
template <int G, int UN>
__global__ void __launch_bounds__(256) k_group256(const double* __restrict__ table, const int* __restrict__ idx, long n,
                                                  double* out) {
  const int lane = threadIdx.x & 31;
  constexpr int RPS = 32 / G, RD = 4 * G;
  const int grp = lane / G, q = lane % G;
  const long gw = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long)gridDim.x * blockDim.x) >> 5;
  const long per = (long)UN * RPS, chunks = n / per;
  double acc = 0.0;
  for (long c = gw; c < chunks; c += nw) {
    int my = (lane < per) ? idx[c * per + lane] : 0;
    double v[UN][4];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      int id = __shfl_sync(0xffffffffu, my, (u * RPS + grp) & 31);
      const double* p = table + (size_t)id * RD + 4 * (grp < RPS ? q : 0);
      ld256(p, v[u][0], v[u][1], v[u][2], v[u][3]);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc = fma(v[u][k], 1.0000001, acc);
  }
  out[1 + blockIdx.x * blockDim.x + threadIdx.x] = acc;
}


// (F) the em kernel's pair layout through TMA: one cp.async.bulk per rating copies its 320-byte
// pair row into a per-warp ring of STAGES chunk buffers (SLOTS rows each, completion on an
// mbarrier per stage); the rows are consumed from shared memory with the kernel's lane mapping
// (10 lanes x 32 bytes per rating, 3 ratings per step).  Global->register traffic through the
// L1 LSU data pipe is replaced by TMA writes + LDS reads, and STAGES-1 chunks stay in flight.
template <int WARPS, int STAGES, int UN>
__global__ void __launch_bounds__(WARPS * 32) k_tma_pair(const double* __restrict__ table, const int* __restrict__ idx,
                                                         long n, double* out) {
  constexpr int RB = 320, RPS = 3, SLOTS = UN * RPS;
  extern __shared__ __align__(128) unsigned char sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* tile = sm + (size_t)warp * STAGES * SLOTS * RB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + (size_t)WARPS * STAGES * SLOTS * RB) + warp * STAGES;
  if (lane == 0)
    for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const long gw = (long)blockIdx.x * WARPS + warp, nw = (long)gridDim.x * WARPS;
  const long chunks = n / SLOTS;
  const int grp = lane / 10, lig = lane % 10;
  const bool on = grp < RPS;
  double acc = 0.0;
  auto issue = [&](long chunk, int st) {
    if (lane == 0) mbar_expect_tx(&bars[st], SLOTS * RB);
    __syncwarp();
    if (lane < SLOTS) {
      const int id = idx[chunk * SLOTS + lane];
      bulk_g2s(tile + ((size_t)st * SLOTS + lane) * RB, table + (size_t)id * (RB / 8), RB, &bars[st]);
    }
  };
  long c = gw, cn = gw;
  int st_issue = 0;
  for (int s = 0; s < STAGES - 1 && cn < chunks; ++s, cn += nw) { issue(cn, st_issue); st_issue = (st_issue + 1) % STAGES; }
  int st = 0;
  uint32_t phase = 0;               // bit s = parity of stage s
  for (; c < chunks; c += nw) {
    if (cn < chunks) { issue(cn, st_issue); st_issue = (st_issue + 1) % STAGES; cn += nw; }
    mbar_wait(&bars[st], (phase >> st) & 1u);
    phase ^= 1u << st;
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      if (on) {
        const double2* p = reinterpret_cast<const double2*>(tile + ((size_t)st * SLOTS + u * RPS + grp) * RB + lig * 32);
        const double2 a = p[0], b = p[1];
        acc = fma(a.x, 1.0000001, acc); acc = fma(a.y, 0.9999999, acc);
        acc = fma(b.x, 1.0000001, acc); acc = fma(b.y, 0.9999999, acc);
      }
    }
    __syncwarp();
    st = (st + 1) % STAGES;
  }
  out[1 + blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <typename F>
static void timeit(const char* name, long n, F launch) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  launch(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int r = 0; r < 5; ++r) launch();
  CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms, a, b)); ms /= 5;
  printf("%-28s %8.3f ms  %7.2f Grows/s  %7.1f GB/s\n", name, ms, n / ms * 1e-6, n * (double)ROWB / ms * 1e-6);
}

int main(int argc, char** argv) {
  const long nrows = argc > 1 ? atol(argv[1]) : 138000, n = 20000000L / 96 * 96 * 8;
  std::vector<int> h(n);
  uint64_t s = 88172645463325252ull;
  for (long i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % nrows); }
  double* table; int* idx; double* out;
  CK(cudaMalloc(&table, nrows * ROWB)); CK(cudaMemset(table, 0, nrows * ROWB));
  CK(cudaMalloc(&idx, n * 4)); CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&out, 64 << 20));
  printf("table %ld rows (%.1f MB), %ld gathers\n", nrows, nrows * ROWB / 1e6, n);
  {
    constexpr int W = 8;
    size_t smem = (size_t)W * 2 * 32 * SROW + W * 2 * 8;
    CK(cudaFuncSetAttribute(k_tma<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int occ : {1, 2, 3})
      timeit(occ == 1 ? "A tma  8w x1/SM" : occ == 2 ? "A tma  8w x2/SM" : "A tma  8w x3/SM", n,
             [&] { k_tma<W><<<148 * occ, W * 32, smem>>>(table, idx, n, out); });
  }
  {
    constexpr int W = 16;
    size_t smem = (size_t)W * 2 * 32 * SROW + W * 2 * 8;
    CK(cudaFuncSetAttribute(k_tma<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    timeit("A tma 16w x1/SM", n, [&] { k_tma<W><<<148, W * 32, smem>>>(table, idx, n, out); });
  }
  timeit("B ldg128 G=10 UN=8", n, [&] { k_group<10, 8><<<148 * 8, 256>>>(table, idx, n, out); });
  timeit("B ldg128 G=5  UN=4", n, [&] { k_group<5, 4><<<148 * 8, 256>>>(table, idx, n, out); });
  timeit("B ldg128 G=2  UN=2", n, [&] { k_group<2, 2><<<148 * 8, 256>>>(table, idx, n, out); });
  {
    // same number of BYTES gathered: 160-byte rows with 5 lanes vs 320-byte rows with 10 lanes
    double* t2; CK(cudaMalloc(&t2, nrows * 320)); CK(cudaMemset(t2, 0, nrows * 320));
    timeit("E ld256 G=5 160B rows UN=2", n, [&] { k_group256<5, 2><<<148 * 12, 256>>>(table, idx, n, out); });
    timeit("E ld256 G=5 160B rows UN=4", n, [&] { k_group256<5, 4><<<148 * 8, 256>>>(table, idx, n, out); });
    timeit("E ld256 G=10 320B rows UN=4 (n/2 rows)", n / 2, [&] { k_group256<10, 4><<<148 * 12, 256>>>(t2, idx, n / 2, out); });
    timeit("E ld256 G=10 320B rows UN=8 (n/2 rows)", n / 2, [&] { k_group256<10, 8><<<148 * 8, 256>>>(t2, idx, n / 2, out); });
  }
  {
    double* t2; CK(cudaMalloc(&t2, nrows * 320)); CK(cudaMemset(t2, 0, nrows * 320));
#define RUN_F(W, ST, UNv, OCC)                                                                          \
    {                                                                                                   \
      size_t smem = (size_t)W * ST * (UNv * 3) * 320 + W * ST * 8;                                      \
      CK(cudaFuncSetAttribute(k_tma_pair<W, ST, UNv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      char nm[96]; snprintf(nm, sizeof nm, "F tma pair %dw st%d un%d x%d/SM (n/2 rows)", W, ST, UNv, OCC); \
      timeit(nm, n / 2, [&] { k_tma_pair<W, ST, UNv><<<148 * OCC, W * 32, smem>>>(t2, idx, n / 2, out); }); \
    }
    RUN_F(8, 2, 3, 3) RUN_F(8, 3, 3, 3) RUN_F(8, 4, 3, 2) RUN_F(8, 3, 3, 4) RUN_F(8, 4, 3, 4)
    RUN_F(8, 3, 4, 3) RUN_F(8, 3, 8, 2) RUN_F(16, 3, 3, 2) RUN_F(8, 6, 3, 2)
  }
  timeit("C ld256 lane/row", n, [&] { k_lane<true><<<148 * 8, 256>>>(table, idx, n, out); });
  timeit("D ld128 lane/row", n, [&] { k_lane<false><<<148 * 8, 256>>>(table, idx, n, out); });
  return 0;
}

// On-device reductions around the EM loop: the reference's "likelihood"
// (src/expectation_maximization.py:157-167), prod_dist (src/kernels_numpy.py:86-97), the
// prediction statistics (src/mmsbm.py:488-539), the mean over runs (src/mmsbm.py:315) and a
// materialising compute_omegas (src/kernels_numpy.py:21-36) for the plugin shim.
// All sums are two-stage with a fixed order (bit-reproducible).
#include <math.h>

#include "common.cuh"

namespace mmsbm {

constexpr int kLikWarps = 8;
constexpr int kLikCtas = 148 * 4;   // persistent CTAs per run

// ---- likelihood ----------------------------------------------------------------------------
// sum_n sum_kl w~ (log w~ - log S~_n), w~ = max(theta_k eta_l pr_klr, eps), S~ = max(sum w, eps).
// The clamp acts on each (k,l), so every rating costs K*L element visits; what can be saved is
// the logarithm: for w >= eps, log w = log theta_k + log eta_l + log pr_klr (three table values,
// exact to rounding), for w < eps it is the constant log eps.  One log per (rating, l) remains.
// Per rating: sum_kl w~ log w~ - log S~ * sum_kl w~, in a single sweep over k.
// One warp walks whole user segments of the (user, level)-grouped index: theta_u and its logs
// are staged once per user, the level is implied by the group; lanes own l, the loop runs over k.
struct LikArgs {
  const int32_t* useg; const int32_t* uadj;
  const double* theta; const double* eta; const double* pr;
  double* partial;      // [S][kLikCtas*kLikWarps]
  double* tables;       // [S][2][R][K][L] in global memory when they do not fit shared memory
  int U, I, R, K, L, ldk, ldl;
};

// pr and log pr in [R][K][L] order (global-memory variant of the tables)
__global__ void lik_tables_kernel(const double* pr, int K, int L, int R, double* tables) {
  const int run = blockIdx.y, n = K * L * R;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int kl = t / R, r = t - kl * R;
  const double v = __ldg(pr + (size_t)run * n + t);
  double* base = tables + (size_t)run * 2 * n;
  base[(size_t)r * K * L + kl] = v;
  base[(size_t)n + (size_t)r * K * L + kl] = log(v);
}

__global__ void __launch_bounds__(kLikWarps * 32) likelihood_kernel(const LikArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int K = A.K, L = A.L, R = A.R;
  const double kLogEps = log(kEps);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int run = blockIdx.y;
  const double* Pq;                                          // [R][K][L]
  const double* Lq;                                          // [R][K][L]  log pr
  double* th_s;                                              // [warps][2][K]  theta, log theta
  if (A.tables) {                                            // tables too large for shared memory
    Pq = A.tables + (size_t)run * 2 * K * L * R;
    Lq = Pq + (size_t)K * L * R;
    th_s = reinterpret_cast<double*>(smem_raw);
  } else {
    double* Ps = reinterpret_cast<double*>(smem_raw);
    double* Ls = Ps + (size_t)R * K * L;
    th_s = Ls + (size_t)R * K * L;
    const double* prs = A.pr + (size_t)run * K * L * R;
    for (int t = threadIdx.x; t < K * L * R; t += blockDim.x) {
      int kl = t / R, r = t - kl * R;
      const double v = __ldg(prs + t);
      Ps[(size_t)r * K * L + kl] = v;
      Ls[(size_t)r * K * L + kl] = log(v);
    }
    __syncthreads();
    Pq = Ps; Lq = Ls;
  }
  double* th = th_s + warp * 2 * K;
  double* lth = th + K;
  const double* theta_run = A.theta + (size_t)run * A.U * A.ldk;
  const double* eta_run = A.eta + (size_t)run * A.I * A.ldl;
  const int gw = blockIdx.x * kLikWarps + warp, nw = gridDim.x * kLikWarps;
  double acc = 0.0;
  for (int u = gw; u < A.U; u += nw) {
    __syncwarp();
    for (int k = lane; k < K; k += 32) {
      const double v = __ldg(theta_run + (size_t)u * A.ldk + k);
      th[k] = v;
      lth[k] = log(v);
    }
    __syncwarp();
    for (int r = 0; r < R; ++r) {
      const int lo = __ldg(A.useg + (size_t)u * R + r), hi = __ldg(A.useg + (size_t)u * R + r + 1);
      const double* Pr = Pq + (size_t)r * K * L;
      const double* Lr = Lq + (size_t)r * K * L;
      for (int j = lo; j < hi; ++j) {
        const int item = __ldg(A.uadj + j);
        MMSBM_DEV_CHECK(item >= 0 && item < A.I);
        const double* erow = eta_run + (size_t)item * A.ldl;
        double tot = 0.0, s1 = 0.0, s2 = 0.0;   // sum w, sum w~ log w~, sum w~
        for (int l0 = 0; l0 < L; l0 += 32) {
          const int l = l0 + lane;
          if (l < L) {
            const double e = __ldg(erow + l);
            const double le = log(e);
            for (int k = 0; k < K; ++k) {
              const double w = __dmul_rn(__dmul_rn(th[k], e), Pr[k * L + l]);   // reference order
              const bool small = w < kEps;
              const double wc = small ? kEps : w;
              const double lw = small ? kLogEps : (lth[k] + le) + Lr[k * L + l];
              tot += w;
              s1 = fma(wc, lw, s1);
              s2 += wc;
            }
          }
        }
        tot = warp_sum(tot);
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        acc += s1 - log(fmax(tot, kEps)) * s2;
      }
    }
  }
  if (lane == 0) A.partial[(size_t)run * nw + gw] = acc;
}

__global__ void __launch_bounds__(256) sum_partials_kernel(const double* partial, int n, double* out) {
  __shared__ double sm[256];
  const double* p = partial + (size_t)blockIdx.x * n;
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) acc += p[i];
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = sm[0];
}

// ---- prod_dist -----------------------------------------------------------------------------
struct ProdArgs {
  const int32_t* user; const int32_t* item;
  const double* theta; const double* eta; const double* pr;
  double* rat;          // [S][M][R]
  int64_t M;
  int U, I, R, K, L, ldk, ldl;
};

__global__ void __launch_bounds__(256) prod_dist_kernel(const ProdArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int K = A.K, L = A.L, R = A.R;
  double* Pq = reinterpret_cast<double*>(smem_raw);          // [R][K][L]
  double* th_s = Pq + (size_t)R * K * L;                     // [warps][K]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int run = blockIdx.y;
  const double* prs = A.pr + (size_t)run * K * L * R;
  for (int t = threadIdx.x; t < K * L * R; t += blockDim.x) {
    int kl = t / R, r = t - kl * R;
    Pq[(size_t)r * K * L + kl] = __ldg(prs + t);
  }
  __syncthreads();
  double* th = th_s + warp * K;
  const double* theta_run = A.theta + (size_t)run * A.U * A.ldk;
  const double* eta_run = A.eta + (size_t)run * A.I * A.ldl;
  for (int64_t m = (int64_t)blockIdx.x * nwarp + warp; m < A.M; m += (int64_t)gridDim.x * nwarp) {
    const int u = __ldg(A.user + m), it = __ldg(A.item + m);
    MMSBM_DEV_CHECK(u >= 0 && u < A.U && it >= 0 && it < A.I);
    __syncwarp();
    for (int k = lane; k < K; k += 32) th[k] = __ldg(theta_run + (size_t)u * A.ldk + k);
    __syncwarp();
    const double* erow = eta_run + (size_t)it * A.ldl;
    for (int r = 0; r < R; ++r) {
      const double* Pr = Pq + (size_t)r * K * L;
      double acc = 0.0;
      for (int l0 = 0; l0 < L; l0 += 32) {
        const int l = l0 + lane;
        if (l < L) {
          const double e = __ldg(erow + l);
          for (int k = 0; k < K; ++k) acc = fma(th[k] * e, Pr[k * L + l], acc);
        }
      }
      acc = warp_sum(acc);
      if (lane == 0) A.rat[((size_t)run * A.M + m) * R + r] = acc;
    }
  }
}

// ---- prediction statistics -----------------------------------------------------------------
struct StatArgs {
  const double* rat; const int32_t* real;
  int64_t M; int R;
  int64_t* cnt_partial;   // [S][blocks][5]
  double* s2_partial;     // [S][blocks]
  int32_t* pred;          // [S][M] or null
};

__global__ void __launch_bounds__(256) stats_kernel(const StatArgs A) {
  __shared__ long long c_s[256][5];
  __shared__ double d_s[256];
  const int run = blockIdx.y;
  long long c[5] = {0, 0, 0, 0, 0};
  double s2p = 0.0;
  for (int64_t m = (int64_t)blockIdx.x * 256 + threadIdx.x; m < A.M; m += (int64_t)gridDim.x * 256) {
    const double* row = A.rat + ((size_t)run * A.M + m) * A.R;
    double best = row[0], tot = 0.0, expect = 0.0;
    int arg = 0;
    for (int r = 0; r < A.R; ++r) {
      const double v = row[r];
      if (v > best) { best = v; arg = r; }   // first maximum, as np.argmax
      tot += v;
      expect = fma(v, (double)r, expect);    // rat @ [0..R-1]  (src/mmsbm.py:518)
    }
    if (A.pred) A.pred[(size_t)run * A.M + m] = arg;
    if (tot != 0.0) {                        // rows with an all-zero distribution are dropped
      const int real = A.real[m];
      const int gap = abs(arg - real);
      c[0] += 1;
      c[1] += (gap == 0);
      c[2] += (gap <= 1);
      c[3] += gap;
      c[4] += ((double)real == rint(expect));  // np.round: half to even
      s2p += fabs(expect - (double)real);
    }
  }
  for (int k = 0; k < 5; ++k) c_s[threadIdx.x][k] = c[k];
  d_s[threadIdx.x] = s2p;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      for (int k = 0; k < 5; ++k) c_s[threadIdx.x][k] += c_s[threadIdx.x + s][k];
      d_s[threadIdx.x] += d_s[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const size_t b = (size_t)run * gridDim.x + blockIdx.x;
    for (int k = 0; k < 5; ++k) A.cnt_partial[b * 5 + k] = c_s[0][k];
    A.s2_partial[b] = d_s[0];
  }
}

__global__ void stats_final_kernel(const int64_t* cnt_partial, const double* s2_partial, int blocks,
                                   int64_t* counts, double* s2pond) {
  const int run = blockIdx.x;
  if (threadIdx.x < 5) {
    long long acc = 0;
    for (int b = 0; b < blocks; ++b) acc += cnt_partial[((size_t)run * blocks + b) * 5 + threadIdx.x];
    counts[run * 5 + threadIdx.x] = acc;
  } else if (threadIdx.x == 5) {
    double acc = 0.0;
    for (int b = 0; b < blocks; ++b) acc += s2_partial[(size_t)run * blocks + b];
    s2pond[run] = acc;
  }
}

__global__ void mean_runs_kernel(const double* rat, int64_t n, int S, double* mean) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  double acc = rat[t];
  for (int s = 1; s < S; ++s) acc += rat[(size_t)s * n + t];   // slab order, as np.mean(axis=0)
  mean[t] = acc / (double)S;
}

__global__ void omegas_kernel(const int32_t* user, const int32_t* item, const int32_t* level,
                              int64_t total, int K, int L, int R, int ldk, int ldl,
                              const double* theta, const double* eta, const double* pr, double* out) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int KL = K * L;
  const int64_t n = t / KL;
  const int kl = (int)(t - n * KL), k = kl / L, l = kl - k * L;
  const int u = user[n], i = item[n], r = level[n];
  // (theta * eta) * pr, the reference's left-to-right product (src/kernels_numpy.py:32-36)
  out[t] = __dmul_rn(__dmul_rn(theta[(size_t)u * ldk + k], eta[(size_t)i * ldl + l]),
                     pr[(size_t)kl * R + r]);
}

constexpr int kStatBlocks = 296;

}  // namespace mmsbm

using namespace mmsbm;

constexpr size_t kLikTableBytes = 512 * 1024;   // room for [2][R][K][L] per run when spilled to global

extern "C" int mmsbm_likelihood_workspace_bytes(int32_t U, int32_t S, size_t* bytes) {
  MMSBM_REQUIRE(bytes && U > 0 && S > 0, MMSBM_EINVAL, "mmsbm_likelihood_workspace_bytes: bad argument");
  *bytes = align_up((size_t)S * kLikCtas * kLikWarps * 8) + align_up((size_t)S * kLikTableBytes) + 256;
  return 0;
}

extern "C" int mmsbm_likelihood(const int32_t* useg, const int32_t* uadj, int64_t N, int32_t U,
                                int32_t I, int32_t R, int32_t K, int32_t L, int32_t S,
                                const double* theta, const double* eta, const double* pr, double* out,
                                void* ws, size_t ws_bytes, void* stream) {
  MMSBM_REQUIRE(useg && uadj && theta && eta && pr && out && ws, MMSBM_EINVAL,
                "mmsbm_likelihood: null pointer");
  MMSBM_REQUIRE(U > 0 && I > 0 && R > 0 && K > 0 && L > 0 && S > 0 && N >= 0, MMSBM_EINVAL,
                "mmsbm_likelihood: bad size");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Arena arena(ws, ws_bytes);
  const int nw = kLikCtas * kLikWarps;
  double* partial = arena.take<double>((size_t)S * nw);
  MMSBM_REQUIRE(partial, MMSBM_ENOMEM, "mmsbm_likelihood: workspace too small");
  LikArgs a{useg, uadj, theta, eta, pr, partial, nullptr, U, I, R, K, L, row_stride(K), row_stride(L)};
  size_t smem = (2 * (size_t)R * K * L + 2 * (size_t)kLikWarps * K) * 8;
  if (smem > 200 * 1024) {                       // keep the two tables in global memory (L2 resident)
    const size_t tb = 2 * (size_t)R * K * L * 8;
    MMSBM_REQUIRE(tb <= kLikTableBytes, MMSBM_ERANGE, "mmsbm_likelihood: K*L*R = %d too large", K * L * R);
    a.tables = arena.take<double>((size_t)S * 2 * R * K * L);
    MMSBM_REQUIRE(a.tables, MMSBM_ENOMEM, "mmsbm_likelihood: workspace too small");
    const int n = K * L * R;
    lik_tables_kernel<<<dim3((n + 255) / 256, S), 256, 0, st>>>(pr, K, L, R, a.tables);
    MMSBM_LAUNCH_CHECK("lik_tables_kernel");
    smem = 2 * (size_t)kLikWarps * K * 8;
  }
  MMSBM_CUDA(cudaFuncSetAttribute(likelihood_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  likelihood_kernel<<<dim3(kLikCtas, S), kLikWarps * 32, smem, st>>>(a);
  MMSBM_LAUNCH_CHECK("likelihood_kernel");
  sum_partials_kernel<<<S, 256, 0, st>>>(partial, nw, out);
  MMSBM_LAUNCH_CHECK("sum_partials_kernel");
  return 0;
}

extern "C" int mmsbm_prod_dist(const int32_t* user, const int32_t* item, int64_t M, int32_t U,
                               int32_t I, int32_t R, int32_t K, int32_t L, int32_t S,
                               const double* theta, const double* eta, const double* pr, double* rat,
                               void* stream) {
  MMSBM_REQUIRE(theta && eta && pr && (M == 0 || (user && item && rat)), MMSBM_EINVAL,
                "mmsbm_prod_dist: null pointer");
  MMSBM_REQUIRE(U > 0 && I > 0 && R > 0 && K > 0 && L > 0 && S > 0 && M >= 0, MMSBM_EINVAL,
                "mmsbm_prod_dist: bad size");
  if (M == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProdArgs a{user, item, theta, eta, pr, rat, M, U, I, R, K, L, row_stride(K), row_stride(L)};
  size_t smem = ((size_t)R * K * L + 8 * (size_t)K) * 8;
  MMSBM_REQUIRE(smem <= 227 * 1024, MMSBM_ERANGE, "mmsbm_prod_dist: K*L*R too large for shared memory");
  MMSBM_CUDA(cudaFuncSetAttribute(prod_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t want = (M + 7) / 8;
  unsigned grid = (unsigned)(want < 148 * 8 ? want : 148 * 8);
  prod_dist_kernel<<<dim3(grid, S), 256, smem, st>>>(a);
  MMSBM_LAUNCH_CHECK("prod_dist_kernel");
  return 0;
}

extern "C" int mmsbm_stats_workspace_bytes(int64_t M, int32_t S, size_t* bytes) {
  MMSBM_REQUIRE(bytes && M >= 0 && S > 0, MMSBM_EINVAL, "mmsbm_stats_workspace_bytes: bad argument");
  *bytes = align_up((size_t)S * kStatBlocks * 5 * 8) + align_up((size_t)S * kStatBlocks * 8) + 256;
  return 0;
}

extern "C" int mmsbm_predict_stats(const double* rat, const int32_t* real, int64_t M, int32_t R,
                                   int32_t S, int64_t* counts, double* s2pond, int32_t* pred, void* ws,
                                   size_t ws_bytes, void* stream) {
  MMSBM_REQUIRE(counts && s2pond && ws && (M == 0 || (rat && real)), MMSBM_EINVAL,
                "mmsbm_predict_stats: null pointer");
  MMSBM_REQUIRE(M >= 0 && R > 0 && S > 0, MMSBM_EINVAL, "mmsbm_predict_stats: bad size");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Arena arena(ws, ws_bytes);
  int64_t* cp = arena.take<int64_t>((size_t)S * kStatBlocks * 5);
  double* sp = arena.take<double>((size_t)S * kStatBlocks);
  MMSBM_REQUIRE(cp && sp, MMSBM_ENOMEM, "mmsbm_predict_stats: workspace too small");
  StatArgs a{rat, real, M, R, cp, sp, pred};
  stats_kernel<<<dim3(kStatBlocks, S), 256, 0, st>>>(a);
  MMSBM_LAUNCH_CHECK("stats_kernel");
  stats_final_kernel<<<S, 32, 0, st>>>(cp, sp, kStatBlocks, counts, s2pond);
  MMSBM_LAUNCH_CHECK("stats_final_kernel");
  return 0;
}

extern "C" int mmsbm_mean_over_runs(const double* rat, int64_t n, int32_t S, double* mean, void* stream) {
  MMSBM_REQUIRE((n == 0 || (rat && mean)) && n >= 0 && S > 0, MMSBM_EINVAL, "mmsbm_mean_over_runs: bad argument");
  if (n == 0) return 0;
  mean_runs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(rat, n, S, mean);
  MMSBM_LAUNCH_CHECK("mean_runs_kernel");
  return 0;
}

extern "C" int mmsbm_compute_omegas(const int32_t* user, const int32_t* item, const int32_t* level,
                                    int64_t N, int32_t K, int32_t L, int32_t R, const double* theta,
                                    const double* eta, const double* pr, double* omegas, void* stream) {
  MMSBM_REQUIRE(theta && eta && pr && (N == 0 || (user && item && level && omegas)), MMSBM_EINVAL,
                "mmsbm_compute_omegas: null pointer");
  MMSBM_REQUIRE(N >= 0 && K > 0 && L > 0 && R > 0, MMSBM_EINVAL, "mmsbm_compute_omegas: bad size");
  if (N == 0) return 0;
  const int64_t total = N * K * L;
  omegas_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      user, item, level, total, K, L, R, row_stride(K), row_stride(L), theta, eta, pr, omegas);
  MMSBM_LAUNCH_CHECK("omegas_kernel");
  return 0;
}

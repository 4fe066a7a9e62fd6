// Fused E-step / M-step of the MMSBM EM iteration for sm_100a.
//
// Replaces update_coefficients (src/kernels_numpy.py:43-79) + normalize_with_d x2 +
// normalize_with_self (src/expectation_maximization.py:118-155), i.e. the loop body of
// src/mmsbm.py:244-250, for S runs at once.  The omega[N,K,L] tensor of the reference is
// never formed.  omega is rank-1 in (k,l) once the rating r is fixed, so for a segment
// that fixes one side ("owner": a user in the by-user pass, an item in the by-item pass)
//
//     w_r[b]   = sum_a own[a] P[a][b][r]                      (once per segment)   [1]
//     S_n      = sum_b w_r[b] nbr_n[b]                        (per rating: one gathered row)
//     g_r[b]  += nbr_n[b] / max(S_n, eps)                     (per rating)         [2]
//     n_own[a] = own[a] * sum_{r,b} P[a][b][r] g_r[b]         (once per segment)   [3]
//     n_pr[a][b][r] = P[a][b][r] * sum_segments own[a] g_r[b] (rank-1 per segment) [4]
//
// which is the reference's sum reassociated (agreement ~1e-15 relative, tests/test_em_gpu.py).
//
// Kernels of one iteration (all runs in grid.y, run-major so one run's tables stay in L2):
//   prep_p_kernel        P in the two operand layouts of each side
//   small_gemm_kernel    [1] for every user and item: W = own x Pw   (register-blocked fp64 FMA)
//   segment_pass_kernel  [2] THE HOT KERNEL: one warp per segment streams its ratings;
//                        per rating it moves one neighbour row (8*NB bytes, one 256-bit load
//                        per lane of a group of G lanes) and 4 bytes of index.  S_n is a
//                        3-level shuffle reduction over the group, g_r stays in registers and
//                        is reduced across groups once per (segment, rating level); rows are
//                        stored grouped by level (include/mmsbm_b200.h).  w comes from W via
//                        cp.async (prefetched one segment ahead), g overwrites W in place.
//   small_gemm_kernel    [3] n_own = (G x Pn) o own / max(deg,1)   (normalisation fused)
//   pr_accumulate_kernel [4] block-private register accumulators over a slab of segments
//   pr_finalize_kernel   fixed-order reduce over slabs, x P, normalise over ratings
// No atomics on data, every sum has a fixed order: results are bit-reproducible.
#include <stdlib.h>

#include "common.cuh"

namespace mmsbm {

constexpr int kWarps = 8;       // warps per CTA of the segment pass
constexpr int kPrSlabs = 64;
constexpr int kPrThreads = 256;
constexpr int kPrAcc = 8;       // accumulators per thread per output tile
constexpr int kPrBatch = 16;    // segments staged per smem batch
constexpr int kGemmThreads = 256;
constexpr int kGemmBK = 32;     // k-slab of the small GEMM (even)

struct alignas(16) double4_t { double x, y, z, w; };

// predicated 256-bit read-only load (LDG.E.ENL2.256 on sm_100a): a lane's 32-byte chunk of a
// row; the registers keep their (finite) previous contents when !pred
__device__ __forceinline__ void ldg256_if(double4_t& v, const double* p, bool pred) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t"
      "@p ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];\n\t}"
      : "+d"(v.x), "+d"(v.y), "+d"(v.z), "+d"(v.w) : "l"(p), "r"((int)pred));
}
__device__ __forceinline__ void stg256(double* p, const double4_t& v) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w)
               : "memory");
}
__device__ __forceinline__ double4_t lds32(const double* p) {   // two 128-bit shared loads
  const double2 a = *reinterpret_cast<const double2*>(p);
  const double2 b = *reinterpret_cast<const double2*>(p + 2);
  return double4_t{a.x, a.y, b.x, b.y};
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 1/x for x in [eps, huge): MUFU.RCP64H seed (~2^-20) + two Newton steps -> <= ~1 ulp
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

// ---- P in operand layout -------------------------------------------------------------------
// For a side with owner dim NA (stride lda), neighbour dim NB (stride NBp), RNB = R*NBp and
// o = r*NBp + b:   Pw[a][o] (lda x RNB)   Pn[o][a] (RNB x lda),  zero in every padded slot.
// user side: a=k, b=l;  item side: a=l, b=k.  One thread per padded (k,l,r).
__global__ void prep_p_kernel(const double* pr, int K, int L, int R, int ldk, int ldl, double* pw_u,
                              double* pn_u, double* pw_i, double* pn_i) {
  const int run = blockIdx.y;
  const int total = ldk * ldl * R;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int r = t % R, kl = t / R, lp = kl % ldl, kp = kl / ldl;
  double v = 0.0;
  if (kp < K && lp < L) v = __ldg(pr + (((size_t)run * K + kp) * L + lp) * R + r);
  const size_t base = (size_t)run * total;
  const int rnb_u = R * ldl, rnb_i = R * ldk;
  pw_u[base + (size_t)kp * rnb_u + r * ldl + lp] = v;
  pn_u[base + (size_t)(r * ldl + lp) * ldk + kp] = v;
  pw_i[base + (size_t)lp * rnb_i + r * ldk + kp] = v;
  pn_i[base + (size_t)(r * ldk + kp) * ldl + lp] = v;
}

// ---- C[M x N] = A[M x Kd] . B[Kd x N] for tall-skinny fp64 problems (N, Kd <= ~1k) ------------
// One 4x4 output tile per thread, rg row groups x N/4 column groups per CTA, k in slabs of
// kGemmBK staged in shared memory.  EPI: C = C o own / max(deg,1).  fp64 FMA pipe, no tensor
// cores (B200 DMMA is no faster than the vector pipe and the path is not GEMM-bound).
struct GemmArgs {
  const double* A;      // [S][M][lda]
  const double* B;      // [S][Kd][N]
  double* C;            // [S][M][N]
  const double* own;    // EPI: [S][M][N]
  const int32_t* deg;   // EPI: [M]
  int M, N, Kd, lda, rg, normalize;
};

template <bool EPI>
__global__ void __launch_bounds__(kGemmThreads) small_gemm_kernel(const GemmArgs g) {
  extern __shared__ __align__(32) unsigned char smem_raw[];
  const int BM = 4 * g.rg, AS = kGemmBK + 2;
  double* As = reinterpret_cast<double*>(smem_raw);          // [BM][AS]
  double* Bs = As + (size_t)BM * AS;                         // [kGemmBK][N]
  const int run = blockIdx.y, row0 = blockIdx.x * BM;
  const int ncg = g.N >> 2;
  const int rgid = threadIdx.x / ncg, cg = threadIdx.x - rgid * ncg;
  const bool active = rgid < g.rg;
  const double* A = g.A + (size_t)run * g.M * g.lda;
  const double* B = g.B + (size_t)run * g.Kd * g.N;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

  for (int k0 = 0; k0 < g.Kd; k0 += kGemmBK) {
    const int bk = min(kGemmBK, g.Kd - k0);            // even: Kd is a multiple of 4
    __syncthreads();
    for (int t = threadIdx.x; t < BM * bk; t += kGemmThreads) {
      const int m = t / bk, kk = t - m * bk;
      As[m * AS + kk] = (row0 + m < g.M) ? __ldg(A + (size_t)(row0 + m) * g.lda + k0 + kk) : 0.0;
    }
    for (int t = threadIdx.x; t < bk * g.N; t += kGemmThreads) Bs[t] = __ldg(B + (size_t)k0 * g.N + t);
    __syncthreads();
    if (active) {
      const double* ap = As + (size_t)(4 * rgid) * AS;
      const double* bp = Bs + 4 * cg;
      for (int kk = 0; kk < bk; kk += 2) {
        double2 a[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const double2*>(ap + i * AS + kk);
        const double4_t b0 = lds32(bp + (size_t)kk * g.N);
        const double4_t b1 = lds32(bp + (size_t)(kk + 1) * g.N);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[i][0] = fma(a[i].x, b0.x, acc[i][0]); acc[i][1] = fma(a[i].x, b0.y, acc[i][1]);
          acc[i][2] = fma(a[i].x, b0.z, acc[i][2]); acc[i][3] = fma(a[i].x, b0.w, acc[i][3]);
          acc[i][0] = fma(a[i].y, b1.x, acc[i][0]); acc[i][1] = fma(a[i].y, b1.y, acc[i][1]);
          acc[i][2] = fma(a[i].y, b1.z, acc[i][2]); acc[i][3] = fma(a[i].y, b1.w, acc[i][3]);
        }
      }
    }
  }
  if (!active) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = row0 + 4 * rgid + i;
    if (m >= g.M) continue;
    const size_t off = ((size_t)run * g.M + m) * g.N + 4 * cg;
    double4_t v{acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
    if (EPI) {
      const double2 o0 = __ldg(reinterpret_cast<const double2*>(g.own + off));
      const double2 o1 = __ldg(reinterpret_cast<const double2*>(g.own + off + 2));
      v.x *= o0.x; v.y *= o0.y; v.z *= o1.x; v.w *= o1.y;
      if (g.normalize) {
        const double d = (double)max(__ldg(g.deg + m), 1);
        v.x = v.x / d; v.y = v.y / d; v.z = v.z / d; v.w = v.w / d;
      }
    }
    stg256(g.C + off, v);
  }
}

// ---- the segment pass ------------------------------------------------------------------------
struct SegArgs {
  const int32_t* seg;   // [nseg*R+1]
  const int32_t* adj;   // [N] neighbour ids, grouped by (segment, level)
  const double* nbr;    // [S][nnbr][NBp]
  double* wg;           // [S][nseg][R*NBp]  in: w   out: g (in place)
  int nseg, nnbr, NBp, R, G, RPS, segs_per_cta;
};

inline size_t seg_smem_bytes(const SegArgs& a) {
  return (size_t)kWarps * 2 * a.R * a.NBp * 8 + 32;
}

template <int CH, int UN>
__global__ void __launch_bounds__(kWarps * 32, 2)
segment_pass_kernel(const SegArgs A) {
  extern __shared__ __align__(32) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int run = blockIdx.y;
  const int R = A.R, NBp = A.NBp, RNB = R * NBp;
  const int NCH = NBp >> 2;                      // 32-byte chunks per neighbour row
  const int G = A.G, RPS = A.RPS;
  const int SLOTS = UN * RPS;                    // ratings per chunk of work (<= 32)

  double* wbuf = reinterpret_cast<double*>(smem_raw) + (size_t)warp * 2 * RNB;   // [2][RNB]
  int* ctr = reinterpret_cast<int*>(smem_raw + (size_t)kWarps * 2 * RNB * 8);
  if (threadIdx.x == 0) *ctr = kWarps;           // warps start on segments 0..kWarps-1
  __syncthreads();

  const int grp = lane / G, q = lane - grp * G;
  const bool lane_on = grp < RPS;
  const int seg_lo = blockIdx.x * A.segs_per_cta;
  const int seg_hi = min(seg_lo + A.segs_per_cta, A.nseg);
  const double* nbr_run = A.nbr + (size_t)run * A.nnbr * NBp;
  double* wg_run = A.wg + (size_t)run * A.nseg * RNB;
  // lane-constant chunk offsets (in doubles), validity, and the add mask of the 3-level
  // shuffle reduction over a group (G <= 8): bit `off` set iff lane q adds lane q+off
  int coff[CH];
  bool con[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int chunk = c * G + q;
    con[c] = lane_on && chunk < NCH;
    coff[c] = con[c] ? 4 * chunk : 0;
  }
  int addm = 0;
#pragma unroll
  for (int off = 4; off > 0; off >>= 1)
    if (off < G && q + off < G) addm |= off;
  const int leader = grp * G;

  // w row of a segment -> shared memory, asynchronously (16-byte pieces)
  auto fetch_w = [&](int s_, int b_) {
    const double* src = wg_run + (size_t)s_ * RNB;
    double* dst = wbuf + (size_t)b_ * RNB;
    for (int p = lane; p < (RNB >> 1); p += 32) cp_async16(dst + 2 * p, src + 2 * p);
  };

  int sg = seg_lo + warp, buf = 0;
  int bend_pref = 0;                             // lane r <= R holds the start of level r
  if (sg < seg_hi) {
    if (lane <= R) bend_pref = __ldg(A.seg + (size_t)sg * R + lane);
    fetch_w(sg, 0);
  }
  cp_async_commit();

  // gathered rows; never-loaded or stale entries are finite and always multiplied by zero
  double4_t x[UN][CH];
#pragma unroll
  for (int un = 0; un < UN; ++un)
#pragma unroll
    for (int c = 0; c < CH; ++c) x[un][c] = double4_t{0.0, 0.0, 0.0, 0.0};

  while (sg < seg_hi) {
    const int bend_reg = bend_pref;
    // claim the next segment, start fetching its boundaries and its w row
    int t = 0;
    if (lane == 0) t = atomicAdd(ctr, 1);
    const int sg_next = seg_lo + __shfl_sync(kFull, t, 0);
    if (sg_next < seg_hi) {
      if (lane <= R) bend_pref = __ldg(A.seg + (size_t)sg_next * R + lane);
      fetch_w(sg_next, buf ^ 1);
    }
    cp_async_commit();
    const int beg = __shfl_sync(kFull, bend_reg, 0), end = __shfl_sync(kFull, bend_reg, R);
    int cur_ids = 0;
    if (lane < SLOTS && beg + lane < end) cur_ids = ld_stream(A.adj + beg + lane);
    cp_async_wait<1>();                          // this segment's w has landed
    __syncwarp();
    const double* wb = wbuf + (size_t)buf * RNB;
    double* gout = wg_run + (size_t)sg * RNB;

    int cur_r = 0, w_lvl = -1;
    double4_t g[CH], wr[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      g[c] = double4_t{0.0, 0.0, 0.0, 0.0};
      wr[c] = double4_t{0.0, 0.0, 0.0, 0.0};
    }

    auto flush = [&](int r) {                    // g_r: sum over the groups, then to global
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        double4_t v = g[c];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          if (off < RPS) {
            const double tx = __shfl_down_sync(kFull, v.x, off * G);
            const double ty = __shfl_down_sync(kFull, v.y, off * G);
            const double tz = __shfl_down_sync(kFull, v.z, off * G);
            const double tw = __shfl_down_sync(kFull, v.w, off * G);
            if (grp + off < RPS) { v.x += tx; v.y += ty; v.z += tz; v.w += tw; }
          }
        }
        if (grp == 0 && con[c]) stg256(gout + r * NBp + coff[c], v);
        g[c] = double4_t{0.0, 0.0, 0.0, 0.0};
      }
    };

    for (int base = beg; base < end; base += SLOTS) {
      // ---- gather: one 256-bit load per (step, chunk); rows past the end are skipped ----
#pragma unroll
      for (int un = 0; un < UN; ++un) {
        const int slot = un * RPS + grp;
        const bool valid = lane_on && (base + slot < end);
        const int id = __shfl_sync(kFull, cur_ids, slot & 31);
        const double* row = nbr_run + (size_t)id * NBp;
#pragma unroll
        for (int c = 0; c < CH; ++c) ldg256_if(x[un][c], row + coff[c], valid && con[c]);
      }
      // next chunk's ids (independent of the row loads above)
      {
        const int nxt = base + SLOTS + lane;
        cur_ids = (lane < SLOTS && nxt < end) ? ld_stream(A.adj + nxt) : 0;
      }
      // ---- rating level of every slot, lanes <-> slots (rows are sorted by level) ----
      int r_slot = 0;
      {
        const int j = base + lane;
        for (int r = 1; r < R; ++r) r_slot += (j >= __shfl_sync(kFull, bend_reg, r));
      }
      const int nvalid = min(SLOTS, end - base);
      const int r_first = __shfl_sync(kFull, r_slot, 0);
      const int r_last = __shfl_sync(kFull, r_slot, nvalid - 1);

      // ---- per step: S = <w_r, row>, 1/max(S, eps) ----
      double inv[UN];
      int r_un[UN];
#pragma unroll
      for (int un = 0; un < UN; ++un) {
        const int slot = un * RPS + grp;
        r_un[un] = __shfl_sync(kFull, r_slot, slot & 31);
        if (r_un[un] != w_lvl) {                 // level change: reload this lane's w chunks
          w_lvl = r_un[un];
#pragma unroll
          for (int c = 0; c < CH; ++c)
            if (con[c]) wr[c] = lds32(wb + w_lvl * NBp + coff[c]);
        }
        double part = 0.0;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          part = fma(x[un][c].x, wr[c].x, part);
          part = fma(x[un][c].y, wr[c].y, part);
          part = fma(x[un][c].z, wr[c].z, part);
          part = fma(x[un][c].w, wr[c].w, part);
        }
        double tp = __shfl_down_sync(kFull, part, 4);
        if (addm & 4) part += tp;
        tp = __shfl_down_sync(kFull, part, 2);
        if (addm & 2) part += tp;
        tp = __shfl_down_sync(kFull, part, 1);
        if (addm & 1) part += tp;
        const double tot = __shfl_sync(kFull, part, leader);
        const bool valid = lane_on && (base + slot < end);
        inv[un] = valid ? fast_rcp(fmax(tot, kEps)) : 0.0;
      }

      // ---- g_r += row / S, level by level (usually one level per chunk) ----
      for (int r = r_first;; ++r) {
        while (cur_r < r) { flush(cur_r); ++cur_r; }
#pragma unroll
        for (int un = 0; un < UN; ++un) {
          const double im = (r_un[un] == r) ? inv[un] : 0.0;
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            g[c].x = fma(x[un][c].x, im, g[c].x); g[c].y = fma(x[un][c].y, im, g[c].y);
            g[c].z = fma(x[un][c].z, im, g[c].z); g[c].w = fma(x[un][c].w, im, g[c].w);
          }
        }
        if (r >= r_last) break;
      }
    }
    while (cur_r < R) { flush(cur_r); ++cur_r; }
    __syncwarp();
    buf ^= 1;
    sg = sg_next;
  }
  cp_async_wait<0>();
}

// ---- n_pr: Acc[a][r][b] = sum_seg own[seg][a] g[seg][r][b], per-CTA private accumulators,
//      one partial per slab, reduced in fixed order by pr_finalize_kernel -------------------
struct PrArgs {
  const double* own;   // [S][nseg][lda]
  const double* g;     // [S][nseg][RNB]
  double* partial;     // [S][kPrSlabs][NA*RNB]
  int nseg, NA, lda, RNB;
};

__global__ void __launch_bounds__(kPrThreads) pr_accumulate_kernel(const PrArgs A) {
  extern __shared__ __align__(32) unsigned char smem_raw[];
  double* own_s = reinterpret_cast<double*>(smem_raw);   // [kPrBatch][NA]
  double* g_s = own_s + kPrBatch * A.NA;                 // [kPrBatch][RNB]
  const int run = blockIdx.y, slab = blockIdx.x;
  const int per = (A.nseg + kPrSlabs - 1) / kPrSlabs;
  const int s0 = slab * per, s1 = min(s0 + per, A.nseg);
  const int nout = A.NA * A.RNB;
  const double* own_run = A.own + (size_t)run * A.nseg * A.lda;
  const double* g_run = A.g + (size_t)run * A.nseg * A.RNB;
  double* dst = A.partial + ((size_t)run * kPrSlabs + slab) * nout;

  for (int tile = 0; tile < nout; tile += kPrThreads * kPrAcc) {
    double acc[kPrAcc];
    int ia[kPrAcc], ib[kPrAcc];
#pragma unroll
    for (int t = 0; t < kPrAcc; ++t) {
      acc[t] = 0.0;
      int o = tile + t * kPrThreads + threadIdx.x;
      int oc = min(o, nout - 1);
      ia[t] = oc / A.RNB;
      ib[t] = oc - ia[t] * A.RNB;
    }
    for (int b0 = s0; b0 < s1; b0 += kPrBatch) {
      const int nb = min(kPrBatch, s1 - b0);
      __syncthreads();
      for (int t = threadIdx.x; t < nb * A.NA; t += kPrThreads) {
        int sg = t / A.NA, a = t - sg * A.NA;
        own_s[t] = __ldg(own_run + (size_t)(b0 + sg) * A.lda + a);
      }
      for (int t = threadIdx.x; t < nb * A.RNB; t += kPrThreads)
        g_s[t] = __ldg(g_run + (size_t)b0 * A.RNB + t);
      __syncthreads();
      for (int sg = 0; sg < nb; ++sg) {
#pragma unroll
        for (int t = 0; t < kPrAcc; ++t)
          acc[t] = fma(own_s[sg * A.NA + ia[t]], g_s[sg * A.RNB + ib[t]], acc[t]);
      }
    }
#pragma unroll
    for (int t = 0; t < kPrAcc; ++t) {
      int o = tile + t * kPrThreads + threadIdx.x;
      if (o < nout) dst[o] = acc[t];
    }
  }
}

struct PrFinArgs {
  const double* partial;  // [S][kPrSlabs][NA*RNB]
  const double* pr;       // [S][K][L][R]
  double* pr_out;         // [S][K][L][R]
  int K, L, R, NA, NBp, transposed, normalize;
};

// one thread per (k,l): n_pr[k][l][r] = pr[k][l][r] * sum_slabs Acc; then the row over r is
// divided by its sum (a sum that is exactly zero divides by one), expectation_maximization.py:154
__global__ void pr_finalize_kernel(const PrFinArgs A) {
  const int run = blockIdx.y;
  const int kl = blockIdx.x * blockDim.x + threadIdx.x;
  if (kl >= A.K * A.L) return;
  const int k = kl / A.L, l = kl - k * A.L;
  const int a = A.transposed ? l : k, b = A.transposed ? k : l;
  const int RNB = A.R * A.NBp, nout = A.NA * RNB;
  const double* part = A.partial + (size_t)run * kPrSlabs * nout;
  const double* prs = A.pr + ((size_t)run * A.K * A.L + kl) * A.R;
  double* dst = A.pr_out + ((size_t)run * A.K * A.L + kl) * A.R;
  double tot = 0.0;
  for (int r = 0; r < A.R; ++r) {
    double acc = 0.0;
    const int o = a * RNB + r * A.NBp + b;
    for (int s = 0; s < kPrSlabs; ++s) acc += part[(size_t)s * nout + o];
    acc *= prs[r];
    dst[r] = acc;
    tot += acc;
  }
  if (A.normalize) {
    const double d = (tot == 0.0) ? 1.0 : tot;
    for (int r = 0; r < A.R; ++r) dst[r] = dst[r] / d;
  }
}

// post-all-reduce epilogue (rating-sharded runs)
__global__ void finalize_eta_kernel(double* eta, const int32_t* deg, int n_items, int ldl, int L,
                                    size_t total) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  int col = (int)(t % ldl);
  size_t row = t / ldl;
  int item = (int)(row % n_items);
  if (col < L) eta[t] = eta[t] / (double)max(__ldg(deg + item), 1);
}

__global__ void finalize_pr_kernel(double* pr, int KL_total, int R) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= KL_total) return;
  double* row = pr + (size_t)t * R;
  double tot = 0.0;
  for (int r = 0; r < R; ++r) tot += row[r];
  const double d = (tot == 0.0) ? 1.0 : tot;
  for (int r = 0; r < R; ++r) row[r] = row[r] / d;
}

// ------------------------------------------------------------------------------------------
struct PassShape { int CH, UN, G, RPS; };

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

static PassShape choose_shape(int NBp) {
  // NCH 32-byte chunks per row are spread over G lanes x CH chunks per lane; G <= 8 keeps the
  // per-rating shuffle reduction at 3 levels.  UN steps (RPS ratings each) are in flight.
  const int NCH = NBp / 4;
  int CH = 8;
  for (int c = 1; c <= 8; c *= 2) {
    if ((NCH + c - 1) / c <= 8) { CH = c; break; }
  }
  int CHenv = env_int("MMSBM_CH", 0);
  if ((CHenv == 1 || CHenv == 2 || CHenv == 4 || CHenv == 8) && (NCH + CHenv - 1) / CHenv <= 8) CH = CHenv;
  int G = (NCH + CH - 1) / CH;
  int Genv = env_int("MMSBM_G", 0);
  if (Genv >= G && Genv <= 8) G = Genv;      // the group reduction has 3 shuffle levels
  int RPS = 32 / G;
  int UN = CH >= 4 ? 1 : 4 / CH;
  while (UN > 1 && UN * RPS > 32) UN >>= 1;
  int UNenv = env_int("MMSBM_UN", 0);
  if ((UNenv == 1 || UNenv == 2 || UNenv == 4) && UNenv * CH <= 4 && UNenv * RPS <= 32) UN = UNenv;
  return PassShape{CH, UN, G, RPS};
}

static int segs_per_cta_for(int nseg) {
  // aim at >= 8 waves of 2 CTAs/SM on 148 SMs, at least one segment per warp
  int spc = nseg / (148 * 2 * 8);
  if (spc < kWarps) spc = kWarps;
  if (spc > 64) spc = 64;
  return env_int("MMSBM_SPC", spc);
}

static int launch_segment_pass(SegArgs a, int n_runs, cudaStream_t st) {
  const PassShape sh = choose_shape(a.NBp);
  a.G = sh.G; a.RPS = sh.RPS;
  a.segs_per_cta = segs_per_cta_for(a.nseg);
  dim3 grid((a.nseg + a.segs_per_cta - 1) / a.segs_per_cta, n_runs);
  dim3 block(kWarps * 32);
  size_t smem = seg_smem_bytes(a);
  MMSBM_REQUIRE(smem <= 227 * 1024, MMSBM_ERANGE,
                "segment pass needs %zu bytes of shared memory (R=%d, row stride %d)", smem, a.R, a.NBp);
#define MMSBM_SEG_CASE(CHv, UNv)                                                              \
  if (sh.CH == CHv && sh.UN == UNv) {                                                         \
    auto kern = segment_pass_kernel<CHv, UNv>;                                                \
    MMSBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<grid, block, smem, st>>>(a);                                                       \
    MMSBM_LAUNCH_CHECK("segment_pass_kernel");                                                \
    return 0;                                                                                 \
  }
  MMSBM_SEG_CASE(1, 1) MMSBM_SEG_CASE(1, 2) MMSBM_SEG_CASE(1, 4)
  MMSBM_SEG_CASE(2, 1) MMSBM_SEG_CASE(2, 2)
  MMSBM_SEG_CASE(4, 1) MMSBM_SEG_CASE(8, 1)
#undef MMSBM_SEG_CASE
  set_error("no segment-pass variant for CH=%d UN=%d", sh.CH, sh.UN);
  return MMSBM_ERANGE;
}

template <bool EPI>
static int launch_gemm(GemmArgs g, int n_runs, cudaStream_t st) {
  const int ncg = g.N / 4;
  MMSBM_REQUIRE(g.N % 4 == 0 && g.Kd % 4 == 0 && ncg >= 1 && ncg <= kGemmThreads, MMSBM_ERANGE,
                "small gemm: N=%d Kd=%d not supported (N <= 1024, multiples of 4)", g.N, g.Kd);
  int rg = kGemmThreads / ncg;
  if (rg > 32) rg = 32;
  g.rg = rg;
  const int BM = 4 * rg;
  size_t smem = ((size_t)BM * (kGemmBK + 2) + (size_t)kGemmBK * g.N) * 8;
  MMSBM_REQUIRE(smem <= 227 * 1024, MMSBM_ERANGE, "small gemm needs %zu bytes of shared memory", smem);
  auto kern = small_gemm_kernel<EPI>;
  MMSBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<dim3((g.M + BM - 1) / BM, n_runs), kGemmThreads, smem, st>>>(g);
  MMSBM_LAUNCH_CHECK("small_gemm_kernel");
  return 0;
}

struct EmDims {
  int U, I, R, K, L, S, ldk, ldl, rnb_u, rnb_i;
  bool emit_items;   // the side with fewer segments carries the pr accumulation
  int nseg_e, NA_e, NBp_e;
  size_t p_elems, wg_u_elems, wg_i_elems, partial_elems;
};

static EmDims em_dims(int U, int I, int R, int K, int L, int S) {
  EmDims d;
  d.U = U; d.I = I; d.R = R; d.K = K; d.L = L; d.S = S;
  d.ldk = row_stride(K); d.ldl = row_stride(L);
  d.rnb_u = R * d.ldl; d.rnb_i = R * d.ldk;
  d.emit_items = (I <= U);
  d.nseg_e = d.emit_items ? I : U;
  d.NA_e = d.emit_items ? L : K;
  d.NBp_e = d.emit_items ? d.ldk : d.ldl;
  d.p_elems = (size_t)S * d.ldk * d.ldl * R;
  d.wg_u_elems = (size_t)S * U * d.rnb_u;
  d.wg_i_elems = (size_t)S * I * d.rnb_i;
  d.partial_elems = (size_t)S * kPrSlabs * d.NA_e * R * d.NBp_e;
  return d;
}

}  // namespace mmsbm

using namespace mmsbm;

extern "C" int mmsbm_em_workspace_bytes(int32_t U, int32_t I, int32_t R, int32_t K, int32_t L,
                                        int32_t S, size_t* bytes) {
  MMSBM_REQUIRE(bytes && U > 0 && I > 0 && R > 0 && K > 0 && L > 0 && S > 0, MMSBM_EINVAL,
                "mmsbm_em_workspace_bytes: bad argument");
  EmDims d = em_dims(U, I, R, K, L, S);
  *bytes = 4 * align_up(d.p_elems * 8) + align_up(d.wg_u_elems * 8) + align_up(d.wg_i_elems * 8) +
           align_up(d.partial_elems * 8) + 256;
  return 0;
}

static int em_step_impl(const int32_t* useg, const int32_t* uadj, const int32_t* udeg,
                        const int32_t* iseg, const int32_t* iadj, const int32_t* ideg,
                        int64_t N, int32_t U, int32_t I, int32_t R, int32_t K, int32_t L,
                        int32_t S, const double* theta, const double* eta, const double* pr,
                        double* theta_out, double* eta_out, double* pr_out, int32_t flags,
                        void* ws, size_t ws_bytes, void* stream, cudaEvent_t* ev) {
  MMSBM_REQUIRE(useg && uadj && udeg && iseg && iadj && ideg && theta && eta && pr && theta_out &&
                    eta_out && pr_out && ws, MMSBM_EINVAL, "mmsbm_em_step: null pointer");
  MMSBM_REQUIRE(N >= 0 && U > 0 && I > 0 && R > 0 && K > 0 && L > 0 && S > 0, MMSBM_EINVAL,
                "mmsbm_em_step: bad size");
  MMSBM_REQUIRE(K <= 256 && L <= 256 && R <= 31, MMSBM_ERANGE,
                "mmsbm_em_step: K, L <= 256 and R <= 31 supported (K=%d L=%d R=%d)", K, L, R);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EmDims d = em_dims(U, I, R, K, L, S);
  Arena arena(ws, ws_bytes);
  double* pw_u = arena.take<double>(d.p_elems);
  double* pn_u = arena.take<double>(d.p_elems);
  double* pw_i = arena.take<double>(d.p_elems);
  double* pn_i = arena.take<double>(d.p_elems);
  double* wg_u = arena.take<double>(d.wg_u_elems);
  double* wg_i = arena.take<double>(d.wg_i_elems);
  double* partial = arena.take<double>(d.partial_elems);
  MMSBM_REQUIRE(pw_u && pn_u && pw_i && pn_i && wg_u && wg_i && partial, MMSBM_ENOMEM,
                "mmsbm_em_step: workspace too small (%zu)", ws_bytes);
  int rc;
#define MMSBM_MARK(k) do { if (ev) MMSBM_CUDA(cudaEventRecord(ev[k], st)); } while (0)
  MMSBM_MARK(0);
  // ---- P tables and w = own x Pw for every user and item ----
  {
    const int total = d.ldk * d.ldl * R;
    prep_p_kernel<<<dim3((total + 255) / 256, S), 256, 0, st>>>(pr, K, L, R, d.ldk, d.ldl, pw_u, pn_u, pw_i, pn_i);
    MMSBM_LAUNCH_CHECK("prep_p_kernel");
    GemmArgs gu{theta, pw_u, wg_u, nullptr, nullptr, U, d.rnb_u, d.ldk, d.ldk, 0, 0};
    if ((rc = launch_gemm<false>(gu, S, st))) return rc;
    GemmArgs gi{eta, pw_i, wg_i, nullptr, nullptr, I, d.rnb_i, d.ldl, d.ldl, 0, 0};
    if ((rc = launch_gemm<false>(gi, S, st))) return rc;
  }
  MMSBM_MARK(1);
  // ---- by-user pass: g of every user (gathers eta rows) ----
  {
    SegArgs a{useg, uadj, eta, wg_u, U, I, d.ldl, R, 0, 0, 0};
    if ((rc = launch_segment_pass(a, S, st))) return rc;
  }
  MMSBM_MARK(2);
  // ---- by-item pass: g of every item (gathers theta rows) ----
  {
    SegArgs a{iseg, iadj, theta, wg_i, I, U, d.ldk, R, 0, 0, 0};
    if ((rc = launch_segment_pass(a, S, st))) return rc;
  }
  MMSBM_MARK(3);
  // ---- theta' and eta' = (g x Pn) o own / max(deg,1) ----
  {
    GemmArgs gu{wg_u, pn_u, theta_out, theta, udeg, U, d.ldk, d.rnb_u, d.rnb_u, 0,
                (flags & MMSBM_RAW_THETA) ? 0 : 1};
    if ((rc = launch_gemm<true>(gu, S, st))) return rc;
    GemmArgs gi{wg_i, pn_i, eta_out, eta, ideg, I, d.ldl, d.rnb_i, d.rnb_i, 0,
                (flags & MMSBM_RAW_ETA_PR) ? 0 : 1};
    if ((rc = launch_gemm<true>(gi, S, st))) return rc;
  }
  MMSBM_MARK(4);
  // ---- pr' ----
  PrArgs pa{};
  pa.own = d.emit_items ? eta : theta;
  pa.g = d.emit_items ? wg_i : wg_u;
  pa.partial = partial;
  pa.nseg = d.nseg_e; pa.NA = d.NA_e; pa.lda = d.emit_items ? d.ldl : d.ldk;
  pa.RNB = R * d.NBp_e;
  size_t smem = (size_t)kPrBatch * (pa.NA + pa.RNB) * 8;
  MMSBM_CUDA(cudaFuncSetAttribute(pr_accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
  pr_accumulate_kernel<<<dim3(kPrSlabs, S), kPrThreads, smem, st>>>(pa);
  MMSBM_LAUNCH_CHECK("pr_accumulate_kernel");
  MMSBM_MARK(5);

  PrFinArgs fa{};
  fa.partial = partial; fa.pr = pr; fa.pr_out = pr_out;
  fa.K = K; fa.L = L; fa.R = R; fa.NA = d.NA_e; fa.NBp = d.NBp_e;
  fa.transposed = d.emit_items ? 1 : 0;
  fa.normalize = (flags & MMSBM_RAW_ETA_PR) ? 0 : 1;
  pr_finalize_kernel<<<dim3((K * L + 127) / 128, S), 128, 0, st>>>(fa);
  MMSBM_LAUNCH_CHECK("pr_finalize_kernel");
  MMSBM_MARK(6);
#undef MMSBM_MARK
  return 0;
}

extern "C" int mmsbm_em_step(const int32_t* useg, const int32_t* uadj, const int32_t* udeg,
                             const int32_t* iseg, const int32_t* iadj, const int32_t* ideg,
                             int64_t N, int32_t U, int32_t I, int32_t R, int32_t K, int32_t L,
                             int32_t S, const double* theta, const double* eta, const double* pr,
                             double* theta_out, double* eta_out, double* pr_out, int32_t flags,
                             void* ws, size_t ws_bytes, void* stream) {
  return em_step_impl(useg, uadj, udeg, iseg, iadj, ideg, N, U, I, R, K, L, S, theta, eta, pr,
                      theta_out, eta_out, pr_out, flags, ws, ws_bytes, stream, nullptr);
}

// Same step with CUDA events between its stages; synchronises the stream and writes the device
// time in ms of {P tables + w GEMMs, by-user pass, by-item pass, n GEMMs, pr accumulate,
// pr finalize}.
extern "C" int mmsbm_em_step_profiled(const int32_t* useg, const int32_t* uadj, const int32_t* udeg,
                                      const int32_t* iseg, const int32_t* iadj, const int32_t* ideg,
                                      int64_t N, int32_t U, int32_t I, int32_t R, int32_t K,
                                      int32_t L, int32_t S, const double* theta, const double* eta,
                                      const double* pr, double* theta_out, double* eta_out,
                                      double* pr_out, int32_t flags, void* ws, size_t ws_bytes,
                                      void* stream, float* ms6) {
  MMSBM_REQUIRE(ms6, MMSBM_EINVAL, "mmsbm_em_step_profiled: null output");
  cudaEvent_t ev[7];
  for (int k = 0; k < 7; ++k) MMSBM_CUDA(cudaEventCreate(&ev[k]));
  int rc = em_step_impl(useg, uadj, udeg, iseg, iadj, ideg, N, U, I, R, K, L, S, theta, eta, pr,
                        theta_out, eta_out, pr_out, flags, ws, ws_bytes, stream, ev);
  if (rc == 0) {
    cudaError_t e = cudaEventSynchronize(ev[6]);
    if (e != cudaSuccess) { set_error("cudaEventSynchronize: %s", cudaGetErrorString(e)); rc = (int)e; }
    for (int k = 0; k < 6 && rc == 0; ++k) cudaEventElapsedTime(&ms6[k], ev[k], ev[k + 1]);
  }
  for (int k = 0; k < 7; ++k) cudaEventDestroy(ev[k]);
  return rc;
}

extern "C" int mmsbm_em_run(const int32_t* useg, const int32_t* uadj, const int32_t* udeg,
                            const int32_t* iseg, const int32_t* iadj, const int32_t* ideg,
                            int64_t N, int32_t U, int32_t I, int32_t R, int32_t K, int32_t L,
                            int32_t S, int32_t iterations, double* theta_a, double* eta_a,
                            double* pr_a, double* theta_b, double* eta_b, double* pr_b, void* ws,
                            size_t ws_bytes, void* stream) {
  MMSBM_REQUIRE(iterations >= 0, MMSBM_EINVAL, "mmsbm_em_run: negative iteration count");
  for (int it = 0; it < iterations; ++it) {
    const bool fwd = (it & 1) == 0;
    int rc = mmsbm_em_step(useg, uadj, udeg, iseg, iadj, ideg, N, U, I, R, K, L, S,
                           fwd ? theta_a : theta_b, fwd ? eta_a : eta_b, fwd ? pr_a : pr_b,
                           fwd ? theta_b : theta_a, fwd ? eta_b : eta_a, fwd ? pr_b : pr_a, 0, ws,
                           ws_bytes, stream);
    if (rc) return rc;
  }
  return 0;
}

extern "C" int mmsbm_em_finalize(double* eta, const int32_t* ideg, int32_t I, int32_t L, double* pr,
                                 int32_t K, int32_t R, int32_t S, void* stream) {
  MMSBM_REQUIRE(eta && ideg && pr && I > 0 && L > 0 && K > 0 && R > 0 && S > 0, MMSBM_EINVAL,
                "mmsbm_em_finalize: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int ldl = row_stride(L);
  size_t total = (size_t)S * I * ldl;
  finalize_eta_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(eta, ideg, I, ldl, L, total);
  MMSBM_LAUNCH_CHECK("finalize_eta_kernel");
  int kl = S * K * L;
  finalize_pr_kernel<<<(kl + 127) / 128, 128, 0, st>>>(pr, kl, R);
  MMSBM_LAUNCH_CHECK("finalize_pr_kernel");
  return 0;
}

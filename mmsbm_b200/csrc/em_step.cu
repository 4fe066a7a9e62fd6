// Fused E-step / M-step of the MMSBM EM iteration for sm_100a.
//
// Replaces update_coefficients (src/kernels_numpy.py:43-79) + normalize_with_d x2 +
// normalize_with_self (src/expectation_maximization.py:118-155), i.e. the loop body of
// src/mmsbm.py:244-250, for S runs at once.  The omega[N,K,L] tensor of the reference is
// never formed.  omega is rank-1 in (k,l) once the rating r is fixed, so for a segment
// that fixes one side ("owner": a user in the by-user pass, an item in the by-item pass)
//
//     w_r[b]   = sum_a own[a] P[a][b][r]                      (once per segment)
//     S_n      = sum_b w_r[b] nbr_n[b]                        (per rating: one gathered row)
//     g_r[b]  += nbr_n[b] / max(S_n, eps)                     (per rating)
//     n_own[a] = own[a] * sum_{r,b} P[a][b][r] g_r[b]         (once per segment)
//     n_pr[a][b][r] = P[a][b][r] * sum_segments own[a] g_r[b] (rank-1 update per segment)
//
// which is the reference's sum reassociated (agreement ~1e-15 relative, tests/test_em_gpu.py).
// Per rating the kernel moves one neighbour row (8*NB bytes, 256-bit coalesced loads by a
// group of lanes) and 4 bytes of index; everything else stays in registers / shared memory.
//
// Mapping: one warp per segment; inside a warp, groups of G lanes own one rating each
// (RPS = 32/G ratings per step), lane q of a group holds CH 32-byte chunks of the row
// (chunk c*G+q, fetched with one 256-bit load).  S_n is a shuffle reduction over the group; g_r lives in registers and is
// reduced across groups once per (segment, rating level) -- rows are stored grouped by
// rating level (include/mmsbm_b200.h) so the level is uniform except at group boundaries.
// All sums have a fixed order: results are bit-reproducible run to run.
#include <stdlib.h>

#include "common.cuh"

namespace mmsbm {

constexpr int kWarps = 8;  // warps per CTA of the segment pass
constexpr int kPrSlabs = 64;
constexpr int kPrThreads = 256;
constexpr int kPrAcc = 8;    // accumulators per thread per output tile
constexpr int kPrBatch = 16; // segments staged per smem batch

struct SegArgs {
  const int32_t* seg;   // [nseg*R+1]
  const int32_t* adj;   // [N] neighbour ids
  const int32_t* deg;   // [nseg]
  const double* own;    // [S][nseg][lda]
  const double* nbr;    // [S][nnbr][ldb]
  const double* pr;     // [S][K][L][R]
  double* own_out;      // [S][nseg][lda]
  double* gout;         // [S][nseg][R*NBp] or null
  int nseg, nnbr, NA, NB, lda, ldb, R, K, L;
  int transposed;       // 0: a=k,b=l (by user)   1: a=l,b=k (by item)
  int G, RPS, normalize, segs_per_cta;
};

// row stride of the staged P[a][r][b] table: R*NBp is a multiple of 4, +2 keeps rows 16-byte
// aligned with APs/2 odd, so 128-bit reads of consecutive rows by consecutive lanes (the
// epilogue) fall in distinct 16-byte banks
__host__ __device__ inline int ps_stride(int R, int NBp) { return R * NBp + 2; }

// shared-memory carve-up of the segment pass (all regions 32-byte aligned)
__host__ __device__ inline size_t ps_bytes(int NA, int R, int NBp) {
  return ((size_t)NA * ps_stride(R, NBp) * 8 + 31) & ~(size_t)31;
}
__host__ __device__ inline size_t warp_bytes(int NAp, int R, int NBp) {
  return (size_t)(NAp + R * NBp) * 8 + (size_t)((R + 1 + 7) / 8 * 8) * 4;
}
inline size_t seg_smem_bytes(const SegArgs& a) {
  return ps_bytes(a.NA, a.R, a.ldb) + kWarps * warp_bytes(a.lda, a.R, a.ldb) + 32;
}

struct alignas(16) double4_t { double x, y, z, w; };

// one 256-bit read-only load (LDG.E.ENL2.256 on sm_100a): a lane's 32-byte chunk of a row
__device__ __forceinline__ double4_t ldg256(const double* p) {
  double4_t v;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}

__device__ __forceinline__ double4_t lds32(const double* p) {   // two 128-bit shared loads
  const double2 a = *reinterpret_cast<const double2*>(p);
  const double2 b = *reinterpret_cast<const double2*>(p + 2);
  return double4_t{a.x, a.y, b.x, b.y};
}
__device__ __forceinline__ void sts32(double* p, const double4_t& v) {
  *reinterpret_cast<double2*>(p) = make_double2(v.x, v.y);
  *reinterpret_cast<double2*>(p + 2) = make_double2(v.z, v.w);
}

// 1/x for x in [eps, huge): MUFU.RCP64H seed (~2^-20) + two Newton steps -> <= ~1 ulp
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

// predicated 256-bit load: registers keep their (finite) previous contents when !pred
__device__ __forceinline__ void ldg256_if(double4_t& v, const double* p, bool pred) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t"
      "@p ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];\n\t}"
      : "+d"(v.x), "+d"(v.y), "+d"(v.z), "+d"(v.w) : "l"(p), "r"((int)pred));
}

template <int CH, int UN, bool EMIT>
__global__ void __launch_bounds__(kWarps * 32, 2)
segment_pass_kernel(const SegArgs A) {
  extern __shared__ __align__(32) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int run = blockIdx.y;
  const int R = A.R, NA = A.NA, NAp = A.lda, NBp = A.ldb;
  const int RNB = R * NBp, APs = ps_stride(R, NBp);
  const int NCH = NBp >> 2;                      // 32-byte chunks per neighbour row
  const int G = A.G, RPS = A.RPS;
  const int SLOTS = UN * RPS;                    // ratings per chunk of work (<= 32)

  double* Ps = reinterpret_cast<double*>(smem_raw);                    // [NA][APs]
  const size_t per_warp = warp_bytes(NAp, R, NBp);
  unsigned char* wbase = smem_raw + ps_bytes(NA, R, NBp) + warp * per_warp;
  double* wg = reinterpret_cast<double*>(wbase);                       // [R][NBp] w then g
  double* own_s = wg + RNB;                                            // [NAp]
  int* bend = reinterpret_cast<int*>(own_s + NAp);                     // [R+1]
  int* ctr = reinterpret_cast<int*>(smem_raw + ps_bytes(NA, R, NBp) + kWarps * per_warp);

  // ---- stage P[a][r][b] (zero padded) ----
  for (int t = threadIdx.x; t < NA * APs; t += blockDim.x) Ps[t] = 0.0;
  if (threadIdx.x == 0) *ctr = kWarps;           // warps start on segments 0..kWarps-1
  __syncthreads();
  {
    const double* prs = A.pr + (size_t)run * A.K * A.L * R;
    const int LR = A.L * R;
    for (int t = threadIdx.x; t < A.K * LR; t += blockDim.x) {
      int k = t / LR, rem = t - k * LR, l = rem / R, r = rem - l * R;
      int a = A.transposed ? l : k, b = A.transposed ? k : l;
      Ps[a * APs + r * NBp + b] = __ldg(prs + t);
    }
  }
  __syncthreads();

  const int grp = lane / G, q = lane - grp * G;
  const bool lane_on = grp < RPS;
  const int seg_lo = blockIdx.x * A.segs_per_cta;
  const int seg_hi = min(seg_lo + A.segs_per_cta, A.nseg);
  const double* own_run = A.own + (size_t)run * A.nseg * NAp;
  const double* nbr_run = A.nbr + (size_t)run * A.nnbr * NBp;
  double* out_run = A.own_out + (size_t)run * A.nseg * NAp;
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

  // prefetched state of the NEXT segment of this warp: owner-row element and boundary
  int sg = seg_lo + warp;
  double own_pref = 0.0;
  int bend_pref = 0;
  if (sg < seg_hi) {
    if (lane < NA) own_pref = __ldg(own_run + (size_t)sg * NAp + lane);
    if (lane <= R) bend_pref = __ldg(A.seg + (size_t)sg * R + lane);
  }

  // gathered rows; never-loaded or stale entries are finite and always multiplied by zero
  double4_t x[UN][CH];
#pragma unroll
  for (int un = 0; un < UN; ++un)
#pragma unroll
    for (int c = 0; c < CH; ++c) x[un][c] = double4_t{0.0, 0.0, 0.0, 0.0};

  while (sg < seg_hi) {
    // ---- owner row, group boundaries (first 32 entries were prefetched) ----
    if (lane < NAp) own_s[lane] = own_pref;
    if (lane <= R) bend[lane] = bend_pref;
    for (int a = lane + 32; a < NAp; a += 32)
      own_s[a] = (a < NA) ? __ldg(own_run + (size_t)sg * NAp + a) : 0.0;
    for (int r = lane + 32; r <= R; r += 32) bend[r] = __ldg(A.seg + (size_t)sg * R + r);
    // claim the next segment and start fetching its row now
    int t = 0;
    if (lane == 0) t = atomicAdd(ctr, 1);
    const int sg_next = seg_lo + __shfl_sync(kFull, t, 0);
    own_pref = 0.0;
    if (sg_next < seg_hi) {
      if (lane < NA) own_pref = __ldg(own_run + (size_t)sg_next * NAp + lane);
      if (lane <= R) bend_pref = __ldg(A.seg + (size_t)sg_next * R + lane);
    }
    __syncwarp();

    const int beg = bend[0], end = bend[R];
    // first chunk's neighbour ids: issue before the w computation to hide their latency
    int cur_ids = 0;
    if (lane < SLOTS && beg + lane < end) cur_ids = ld_stream(A.adj + beg + lane);

    // ---- w[r][b] = sum_a own[a] P[a][r][b]: a lane owns output pairs, 128-bit reads ----
    for (int o2 = lane; o2 < (RNB >> 1); o2 += 32) {
      double2 acc0 = make_double2(0.0, 0.0), acc1 = make_double2(0.0, 0.0);
      const double* pcol = Ps + 2 * o2;
      int a = 0;
#pragma unroll 2
      for (; a + 1 < NA; a += 2) {
        const double2 ow = *reinterpret_cast<const double2*>(own_s + a);
        const double2 p0 = *reinterpret_cast<const double2*>(pcol + a * APs);
        const double2 p1 = *reinterpret_cast<const double2*>(pcol + (a + 1) * APs);
        acc0.x = fma(ow.x, p0.x, acc0.x); acc0.y = fma(ow.x, p0.y, acc0.y);
        acc1.x = fma(ow.y, p1.x, acc1.x); acc1.y = fma(ow.y, p1.y, acc1.y);
      }
      if (a < NA) {
        const double ow = own_s[a];
        const double2 p0 = *reinterpret_cast<const double2*>(pcol + a * APs);
        acc0.x = fma(ow, p0.x, acc0.x); acc0.y = fma(ow, p0.y, acc0.y);
      }
      *reinterpret_cast<double2*>(wg + 2 * o2) = make_double2(acc0.x + acc1.x, acc0.y + acc1.y);
    }
    __syncwarp();

    int cur_r = 0;
    double4_t g[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) g[c] = double4_t{0.0, 0.0, 0.0, 0.0};

    auto flush = [&](int r) {
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
        if (grp == 0 && con[c]) sts32(wg + r * NBp + coff[c], v);
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
        for (int r = 1; r < R; ++r) r_slot += (j >= bend[r]);
      }
      const int nvalid = min(SLOTS, end - base);
      const int r_first = __shfl_sync(kFull, r_slot, 0);
      const int r_last = __shfl_sync(kFull, r_slot, nvalid - 1);

      // ---- per step: S = <w_r, row>, 1/max(S, eps); straight-line so the UN chains overlap ----
      double inv[UN];
      int r_un[UN];
#pragma unroll
      for (int un = 0; un < UN; ++un) {
        const int slot = un * RPS + grp;
        r_un[un] = __shfl_sync(kFull, r_slot, slot & 31);
        double part = 0.0;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          double4_t w = lds32(wg + r_un[un] * NBp + coff[c]);
          if (!con[c]) w = double4_t{0.0, 0.0, 0.0, 0.0};
          part = fma(x[un][c].x, w.x, part);
          part = fma(x[un][c].y, w.y, part);
          part = fma(x[un][c].z, w.z, part);
          part = fma(x[un][c].w, w.w, part);
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

    // ---- n_own[a] = own[a] * sum_o P[a][o] g[o]  (/ max(deg,1) when normalising) ----
    double scale = 1.0;
    if (A.normalize) scale = (double)max(__ldg(A.deg + sg), 1);
    double* orow_out = out_run + (size_t)sg * NAp;
    for (int a = lane; a < NAp; a += 32) {
      double acc = 0.0;
      if (a < NA) {
        const double* prow = Ps + a * APs;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 2
        for (int o = 0; o < RNB; o += 4) {     // RNB is a multiple of 4
          const double2 p0 = *reinterpret_cast<const double2*>(prow + o);
          const double2 p1 = *reinterpret_cast<const double2*>(prow + o + 2);
          const double2 g0 = *reinterpret_cast<const double2*>(wg + o);
          const double2 g1 = *reinterpret_cast<const double2*>(wg + o + 2);
          a0 = fma(p0.x, g0.x, a0); a1 = fma(p0.y, g0.y, a1);
          a2 = fma(p1.x, g1.x, a2); a3 = fma(p1.y, g1.y, a3);
        }
        acc = ((a0 + a1) + (a2 + a3)) * own_s[a];
        if (A.normalize) acc = acc / scale;
      }
      orow_out[a] = acc;
    }
    if (EMIT) {
      double* gdst = A.gout + ((size_t)run * A.nseg + sg) * RNB;
      for (int o = lane; o < RNB; o += 32) gdst[o] = wg[o];
    }
    __syncwarp();
    sg = sg_next;
  }
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
  // per-rating shuffle reduction at <= 3 levels.  UN steps (RPS ratings each) are in flight.
  const int NCH = NBp / 4;
  int CH = 8;
  for (int c = 1; c <= 8; c *= 2) {
    if ((NCH + c - 1) / c <= 8) { CH = c; break; }
  }
  int CHenv = env_int("MMSBM_CH", 0);
  if ((CHenv == 1 || CHenv == 2 || CHenv == 4 || CHenv == 8) && (NCH + CHenv - 1) / CHenv <= 32) CH = CHenv;
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

template <bool EMIT>
static int launch_segment_pass(const SegArgs& a, const PassShape& sh, int n_runs, cudaStream_t st) {
  dim3 grid((a.nseg + a.segs_per_cta - 1) / a.segs_per_cta, n_runs);
  dim3 block(kWarps * 32);
  size_t smem = seg_smem_bytes(a);
  MMSBM_REQUIRE(smem <= 227 * 1024, MMSBM_ERANGE,
                "segment pass needs %zu bytes of shared memory (K=%d L=%d R=%d)", smem, a.K, a.L, a.R);
#define MMSBM_SEG_CASE(CHv, UNv)                                                              \
  if (sh.CH == CHv && sh.UN == UNv) {                                                         \
    auto kern = segment_pass_kernel<CHv, UNv, EMIT>;                                          \
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

static int segs_per_cta_for(int nseg) {
  // aim at >= 8 waves of 2 CTAs/SM on 148 SMs, at least one segment per warp
  int spc = nseg / (148 * 2 * 8);
  if (spc < kWarps) spc = kWarps;
  if (spc > 64) spc = 64;
  return env_int("MMSBM_SPC", spc);
}

struct EmDims {
  int U, I, R, K, L, S, ldk, ldl;
  bool emit_items;   // the side with fewer segments carries the pr accumulation
  int nseg_e, NA_e, NBp_e;
  size_t g_elems, partial_elems;
};

static EmDims em_dims(int U, int I, int R, int K, int L, int S) {
  EmDims d;
  d.U = U; d.I = I; d.R = R; d.K = K; d.L = L; d.S = S;
  d.ldk = row_stride(K); d.ldl = row_stride(L);
  d.emit_items = (I <= U);
  d.nseg_e = d.emit_items ? I : U;
  d.NA_e = d.emit_items ? L : K;
  d.NBp_e = d.emit_items ? d.ldk : d.ldl;
  d.g_elems = (size_t)S * d.nseg_e * R * d.NBp_e;
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
  *bytes = align_up(d.g_elems * 8) + align_up(d.partial_elems * 8) + 256;
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
  MMSBM_REQUIRE(K <= 256 && L <= 256 && R <= 64, MMSBM_ERANGE,
                "mmsbm_em_step: K, L <= 256 and R <= 64 supported (K=%d L=%d R=%d)", K, L, R);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EmDims d = em_dims(U, I, R, K, L, S);
  Arena arena(ws, ws_bytes);
  double* gbuf = arena.take<double>(d.g_elems);
  double* partial = arena.take<double>(d.partial_elems);
  MMSBM_REQUIRE(gbuf && partial, MMSBM_ENOMEM, "mmsbm_em_step: workspace too small (%zu)", ws_bytes);

#define MMSBM_MARK(k) do { if (ev) MMSBM_CUDA(cudaEventRecord(ev[k], st)); } while (0)
  MMSBM_MARK(0);
  // ---- by-user pass: theta' ----
  SegArgs ua{};
  ua.seg = useg; ua.adj = uadj; ua.deg = udeg;
  ua.own = theta; ua.nbr = eta; ua.pr = pr; ua.own_out = theta_out;
  ua.gout = d.emit_items ? nullptr : gbuf;
  ua.nseg = U; ua.nnbr = I; ua.NA = K; ua.NB = L; ua.lda = d.ldk; ua.ldb = d.ldl;
  ua.R = R; ua.K = K; ua.L = L; ua.transposed = 0;
  PassShape ush = choose_shape(d.ldl);
  ua.G = ush.G; ua.RPS = ush.RPS;
  ua.normalize = (flags & MMSBM_RAW_THETA) ? 0 : 1;
  ua.segs_per_cta = segs_per_cta_for(U);
  int rc = d.emit_items ? launch_segment_pass<false>(ua, ush, S, st)
                        : launch_segment_pass<true>(ua, ush, S, st);
  if (rc) return rc;
  MMSBM_MARK(1);

  // ---- by-item pass: eta' ----
  SegArgs ia{};
  ia.seg = iseg; ia.adj = iadj; ia.deg = ideg;
  ia.own = eta; ia.nbr = theta; ia.pr = pr; ia.own_out = eta_out;
  ia.gout = d.emit_items ? gbuf : nullptr;
  ia.nseg = I; ia.nnbr = U; ia.NA = L; ia.NB = K; ia.lda = d.ldl; ia.ldb = d.ldk;
  ia.R = R; ia.K = K; ia.L = L; ia.transposed = 1;
  PassShape ish = choose_shape(d.ldk);
  ia.G = ish.G; ia.RPS = ish.RPS;
  ia.normalize = (flags & MMSBM_RAW_ETA_PR) ? 0 : 1;
  ia.segs_per_cta = segs_per_cta_for(I);
  rc = d.emit_items ? launch_segment_pass<true>(ia, ish, S, st)
                    : launch_segment_pass<false>(ia, ish, S, st);
  if (rc) return rc;
  MMSBM_MARK(2);

  // ---- pr' ----
  PrArgs pa{};
  pa.own = d.emit_items ? eta : theta;
  pa.g = gbuf; pa.partial = partial;
  pa.nseg = d.nseg_e; pa.NA = d.NA_e; pa.lda = d.emit_items ? d.ldl : d.ldk;
  pa.RNB = R * d.NBp_e;
  size_t smem = (size_t)kPrBatch * (pa.NA + pa.RNB) * 8;
  MMSBM_CUDA(cudaFuncSetAttribute(pr_accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
  pr_accumulate_kernel<<<dim3(kPrSlabs, S), kPrThreads, smem, st>>>(pa);
  MMSBM_LAUNCH_CHECK("pr_accumulate_kernel");
  MMSBM_MARK(3);

  PrFinArgs fa{};
  fa.partial = partial; fa.pr = pr; fa.pr_out = pr_out;
  fa.K = K; fa.L = L; fa.R = R; fa.NA = d.NA_e; fa.NBp = d.NBp_e;
  fa.transposed = d.emit_items ? 1 : 0;
  fa.normalize = (flags & MMSBM_RAW_ETA_PR) ? 0 : 1;
  pr_finalize_kernel<<<dim3((K * L + 127) / 128, S), 128, 0, st>>>(fa);
  MMSBM_LAUNCH_CHECK("pr_finalize_kernel");
  MMSBM_MARK(4);
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

// Same step with CUDA events around its four launches; synchronises the stream and writes
// the device time of {by-user pass, by-item pass, pr accumulate, pr finalize} in ms.
extern "C" int mmsbm_em_step_profiled(const int32_t* useg, const int32_t* uadj, const int32_t* udeg,
                                      const int32_t* iseg, const int32_t* iadj, const int32_t* ideg,
                                      int64_t N, int32_t U, int32_t I, int32_t R, int32_t K,
                                      int32_t L, int32_t S, const double* theta, const double* eta,
                                      const double* pr, double* theta_out, double* eta_out,
                                      double* pr_out, int32_t flags, void* ws, size_t ws_bytes,
                                      void* stream, float* ms4) {
  MMSBM_REQUIRE(ms4, MMSBM_EINVAL, "mmsbm_em_step_profiled: null output");
  cudaEvent_t ev[5];
  for (int k = 0; k < 5; ++k) MMSBM_CUDA(cudaEventCreate(&ev[k]));
  int rc = em_step_impl(useg, uadj, udeg, iseg, iadj, ideg, N, U, I, R, K, L, S, theta, eta, pr,
                        theta_out, eta_out, pr_out, flags, ws, ws_bytes, stream, ev);
  if (rc == 0) {
    cudaError_t e = cudaEventSynchronize(ev[4]);
    if (e != cudaSuccess) { set_error("cudaEventSynchronize: %s", cudaGetErrorString(e)); rc = (int)e; }
    for (int k = 0; k < 4 && rc == 0; ++k) cudaEventElapsedTime(&ms4[k], ev[k], ev[k + 1]);
  }
  for (int k = 0; k < 5; ++k) cudaEventDestroy(ev[k]);
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

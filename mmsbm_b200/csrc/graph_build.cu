// On-device index build: rows grouped by (user, rating) and by (item, rating), original
// order kept inside a group (== numpy stable argsort of id*R + rating).
//
// Replaces the O((U+I)*N) python scans of MMSBM._prepare_objects (src/mmsbm.py:100-122).
// Integer work only, bit-exact by construction:
//   1. key[n] = id[n]*R + level[n];  histogram + exclusive scan -> seg[nkeys+1], degrees
//   2. stable LSD radix sort of (key, n) pairs, 8 bits per pass, ceil(bits(nkeys)/8) passes.
//      Each warp owns a contiguous chunk of kChunk rows: pass A counts digits per chunk,
//      an exclusive scan over [digit][chunk] gives every chunk its output offsets, pass B
//      re-reads the chunk in order and places rows with __match_any_sync ranks (stable).
//   3. adj[j] = other_id[perm[j]]
// HBM-bound streaming; no tensor cores.
#include "common.cuh"

namespace mmsbm {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kChunk = 2048;       // rows per warp chunk
constexpr int kSortWarps = 8;      // warps per CTA
constexpr int kScanBlock = 1024;
constexpr int kScanItems = 4;      // elements per thread in the block scan

__global__ void make_keys_kernel(const int32_t* id, const int32_t* level, int64_t n, int R,
                                 uint32_t* key, uint32_t* val, int32_t* hist, int* bad,
                                 int n_ids) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  int a = id[t], r = level[t];
  val[t] = (uint32_t)t;
  if (a < 0 || a >= n_ids || r < 0 || r >= R) {  // caller error: keep every access in bounds
    atomicExch(bad, 1);
    key[t] = 0u;
    return;
  }
  uint32_t k = (uint32_t)a * (uint32_t)R + (uint32_t)r;
  key[t] = k;
  atomicAdd(hist + k, 1);
}

// ---- exclusive scan of int32 arrays: tile sums -> serial-over-tiles scan -> add -------------
__global__ void __launch_bounds__(kScanBlock) scan_tiles_kernel(const int32_t* in, int32_t* out,
                                                                int32_t* tile_sums, int64_t n) {
  __shared__ int32_t warp_tot[kScanBlock / 32];
  const int64_t tile0 = (int64_t)blockIdx.x * kScanBlock * kScanItems;
  const int64_t i0 = tile0 + (int64_t)threadIdx.x * kScanItems;
  int32_t v[kScanItems], run = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = (i0 + k < n) ? in[i0 + k] : 0;
    run += v[k];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int32_t inc = run;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    int32_t t = __shfl_up_sync(kFull, inc, off);
    if (lane >= off) inc += t;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int32_t w = warp_tot[lane];
    int32_t winc = w;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      int32_t t = __shfl_up_sync(kFull, winc, off);
      if (lane >= off) winc += t;
    }
    warp_tot[lane] = winc - w;
    if (lane == 31) tile_sums[blockIdx.x] = winc;
  }
  __syncthreads();
  int32_t excl = warp_tot[warp] + inc - run;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (i0 + k < n) out[i0 + k] = excl;
    excl += v[k];
  }
}

__global__ void __launch_bounds__(kScanBlock) scan_sums_kernel(int32_t* tile_sums, int64_t n_tiles) {
  __shared__ int32_t warp_tot[kScanBlock / 32];
  __shared__ int32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t base = 0; base < n_tiles; base += kScanBlock) {
    int64_t i = base + threadIdx.x;
    int32_t v = (i < n_tiles) ? tile_sums[i] : 0;
    int32_t inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      int32_t t = __shfl_up_sync(kFull, inc, off);
      if (lane >= off) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int32_t w = warp_tot[lane], winc = w;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        int32_t t = __shfl_up_sync(kFull, winc, off);
        if (lane >= off) winc += t;
      }
      warp_tot[lane] = winc - w;
    }
    __syncthreads();
    const int32_t carry = carry_s;
    int32_t excl = carry + warp_tot[warp] + inc - v;
    if (i < n_tiles) tile_sums[i] = excl;
    __syncthreads();
    if (threadIdx.x == kScanBlock - 1) carry_s = excl + v;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kScanBlock) scan_add_kernel(int32_t* out, const int32_t* tile_sums,
                                                              int64_t n) {
  const int64_t i0 = ((int64_t)blockIdx.x * kScanBlock + threadIdx.x) * kScanItems;
  const int32_t add = tile_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
    if (i0 + k < n) out[i0 + k] += add;
}

static int exclusive_scan(const int32_t* in, int32_t* out, int64_t n, int32_t* tile_sums,
                          cudaStream_t st) {
  if (n <= 0) return 0;
  const int64_t per = (int64_t)kScanBlock * kScanItems;
  const int64_t tiles = (n + per - 1) / per;
  scan_tiles_kernel<<<(unsigned)tiles, kScanBlock, 0, st>>>(in, out, tile_sums, n);
  MMSBM_LAUNCH_CHECK("scan_tiles_kernel");
  scan_sums_kernel<<<1, kScanBlock, 0, st>>>(tile_sums, tiles);
  MMSBM_LAUNCH_CHECK("scan_sums_kernel");
  scan_add_kernel<<<(unsigned)tiles, kScanBlock, 0, st>>>(out, tile_sums, n);
  MMSBM_LAUNCH_CHECK("scan_add_kernel");
  return 0;
}
static int64_t scan_tiles_for(int64_t n) {
  const int64_t per = (int64_t)kScanBlock * kScanItems;
  return (n + per - 1) / per;
}

// ---- radix pass A: per-chunk digit counts, written [digit][chunk] --------------------------
__global__ void __launch_bounds__(kSortWarps * 32)
radix_count_kernel(const uint32_t* key, int64_t n, int shift, int64_t n_chunks, int32_t* counts) {
  __shared__ int32_t hist[kSortWarps][kRadix];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t chunk = (int64_t)blockIdx.x * kSortWarps + warp;
  for (int d = lane; d < kRadix; d += 32) hist[warp][d] = 0;
  __syncwarp();
  if (chunk < n_chunks) {
    const int64_t lo = chunk * kChunk, hi = min(lo + (int64_t)kChunk, n);
    for (int64_t i = lo + lane; i < hi; i += 32)
      atomicAdd(&hist[warp][(key[i] >> shift) & (kRadix - 1)], 1);
    __syncwarp();
    for (int d = lane; d < kRadix; d += 32) counts[(int64_t)d * n_chunks + chunk] = hist[warp][d];
  }
}

// ---- radix pass B: stable placement -------------------------------------------------------
__global__ void __launch_bounds__(kSortWarps * 32)
radix_scatter_kernel(const uint32_t* key, const uint32_t* val, int64_t n, int shift,
                     int64_t n_chunks, const int32_t* offsets, uint32_t* key_out,
                     uint32_t* val_out) {
  __shared__ int32_t off[kSortWarps][kRadix];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t chunk = (int64_t)blockIdx.x * kSortWarps + warp;
  if (chunk >= n_chunks) return;
  for (int d = lane; d < kRadix; d += 32) off[warp][d] = offsets[(int64_t)d * n_chunks + chunk];
  __syncwarp();
  const int64_t lo = chunk * kChunk, hi = min(lo + (int64_t)kChunk, n);
  const unsigned lt = (1u << lane) - 1u;
  for (int64_t base = lo; base < hi; base += 32) {
    const int64_t i = base + lane;
    const bool on = i < hi;
    uint32_t k = on ? key[i] : 0u, v = on ? val[i] : 0u;
    // rows past the end get a digit no live row can share a ballot with
    const int d = on ? (int)((k >> shift) & (kRadix - 1)) : kRadix + lane;
    const unsigned peers = __match_any_sync(kFull, d);
    const int rank = __popc(peers & lt);
    int32_t dst = 0;
    if (on) dst = off[warp][d] + rank;
    __syncwarp();
    if (on && rank == 0) off[warp][d] += __popc(peers);
    __syncwarp();
    MMSBM_DEV_CHECK(!on || (dst >= 0 && dst < n));
    if (on) { key_out[dst] = k; val_out[dst] = v; }
  }
}

__global__ void gather_adj_kernel(const uint32_t* perm, const int32_t* other, int64_t n,
                                  int32_t* adj, int32_t* perm_out) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  uint32_t p = perm[t];
  MMSBM_DEV_CHECK(p < n);
  adj[t] = other[p];
  perm_out[t] = (int32_t)p;
}

__global__ void degrees_kernel(const int32_t* seg, int n_ids, int R, int32_t* deg) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_ids) return;
  deg[t] = seg[(int64_t)(t + 1) * R] - seg[(int64_t)t * R];
}

__global__ void set_last_kernel(int32_t* seg, int64_t nkeys, int32_t n) { seg[nkeys] = n; }

// ---- work schedule of the segment pass ------------------------------------------------------
// A segment (all ratings of a user, or of an item) is cut into pieces of at most
// MMSBM_PIECE_LEN ratings; one warp processes one piece.  Single-piece segments write their g row
// in place; the pieces of a long segment write partial rows into numbered slots that a fix-up
// kernel adds up in piece order (deterministic).  This bounds the work of one warp whatever the
// degree distribution (real rating graphs are heavy-tailed: a handful of items hold 1e5+ ratings).
// sched layout (int32): [0] pieces P, [1] slots, [2] long segments, [3] piece length,
//   piece_seg[Pmax] piece_idx[Pmax] piece_slot[Pmax] long_seg[Lmax] long_slot0[Lmax+1]
__host__ __device__ inline int64_t sched_pmax(int64_t N, int64_t nseg) { return nseg + N / MMSBM_PIECE_LEN + 1; }
// capacities: a long segment has > PIECE_LEN ratings, so there are at most N/PIECE_LEN of them;
// each contributes ceil(deg/PIECE_LEN) <= deg/PIECE_LEN + 1 slots, i.e. at most 2N/PIECE_LEN slots
__host__ __device__ inline int64_t sched_lmax(int64_t N) { return N / MMSBM_PIECE_LEN + 1; }

__global__ void sched_count_kernel(const int32_t* deg, int nseg, int32_t* pps, int32_t* nslot, int32_t* nlong) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nseg) return;
  int p = (deg[t] + MMSBM_PIECE_LEN - 1) / MMSBM_PIECE_LEN;
  if (p < 1) p = 1;
  pps[t] = p;
  nslot[t] = p > 1 ? p : 0;
  nlong[t] = p > 1 ? 1 : 0;
}

__global__ void sched_fill_kernel(const int32_t* deg, int nseg, const int32_t* piece_base,
                                  const int32_t* slot_base, const int32_t* long_base, int64_t pmax,
                                  int64_t lmax, int32_t* sched) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nseg) return;
  int32_t* piece_seg = sched + 4;
  int32_t* piece_idx = piece_seg + pmax;
  int32_t* piece_slot = piece_idx + pmax;
  int32_t* long_seg = piece_slot + pmax;
  int32_t* long_slot0 = long_seg + lmax;
  int p = (deg[t] + MMSBM_PIECE_LEN - 1) / MMSBM_PIECE_LEN;
  if (p < 1) p = 1;
  const int pb = piece_base[t], sb = slot_base[t];
  MMSBM_DEV_CHECK(pb >= 0 && pb + p <= pmax &&
                  (p == 1 || (sb + p <= 2 * (lmax - 1) + 1 && long_base[t] < lmax)));
  for (int k = 0; k < p; ++k) {
    piece_seg[pb + k] = t;
    piece_idx[pb + k] = k;
    piece_slot[pb + k] = p > 1 ? sb + k : -1;
  }
  if (p > 1) {
    long_seg[long_base[t]] = t;
    long_slot0[long_base[t]] = sb;
  }
  if (t == nseg - 1) {                       // totals and the closing slot boundary
    const int n_slots = sb + (p > 1 ? p : 0);
    const int n_long = long_base[t] + (p > 1 ? 1 : 0);
    sched[0] = pb + p;
    sched[1] = n_slots;
    sched[2] = n_long;
    sched[3] = MMSBM_PIECE_LEN;
    long_slot0[n_long] = n_slots;
  }
}

struct GraphWs {
  uint32_t *key_a, *key_b, *val_a, *val_b;
  int32_t *hist, *counts, *tile_sums;
  int32_t *sa, *sb, *sc;   // schedule scratch: per-segment counts, scanned in place
  int* bad;
};

static size_t graph_ws_layout(int64_t N, int64_t max_keys, int64_t max_segs, void* base, size_t cap,
                              GraphWs* out) {
  const int64_t n_chunks = (N + kChunk - 1) / kChunk;
  const int64_t n_counts = n_chunks * kRadix;
  const int64_t big = n_counts > max_keys + 1 ? n_counts : max_keys + 1;
  Arena a(base ? base : (void*)nullptr, base ? cap : (size_t)-1);
  size_t need = 0;
  auto bump = [&](size_t bytes) { need += align_up(bytes); };
  bump(N * 4); bump(N * 4); bump(N * 4); bump(N * 4);
  bump((max_keys + 1) * 4); bump(n_counts * 4); bump((scan_tiles_for(big) + 1) * 4); bump(256);
  bump(max_segs * 4); bump(max_segs * 4); bump(max_segs * 4);
  if (base && out) {
    out->key_a = a.take<uint32_t>(N); out->key_b = a.take<uint32_t>(N);
    out->val_a = a.take<uint32_t>(N); out->val_b = a.take<uint32_t>(N);
    out->hist = a.take<int32_t>(max_keys + 1);
    out->counts = a.take<int32_t>(n_counts);
    out->tile_sums = a.take<int32_t>(scan_tiles_for(big) + 1);
    out->bad = a.take<int>(64);
    out->sa = a.take<int32_t>(max_segs); out->sb = a.take<int32_t>(max_segs); out->sc = a.take<int32_t>(max_segs);
    if (!out->bad || !out->sc) return 0;
  }
  return need;
}

static int build_one(const int32_t* id, const int32_t* other, const int32_t* level, int64_t N,
                     int n_ids, int R, int32_t* seg, int32_t* adj, int32_t* perm, int32_t* deg,
                     GraphWs& w, cudaStream_t st) {
  const int64_t nkeys = (int64_t)n_ids * R;
  MMSBM_CUDA(cudaMemsetAsync(w.hist, 0, (nkeys + 1) * 4, st));
  MMSBM_CUDA(cudaMemsetAsync(w.bad, 0, sizeof(int), st));
  if (N > 0) {
    make_keys_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(id, level, N, R, w.key_a, w.val_a,
                                                                  w.hist, w.bad, n_ids);
    MMSBM_LAUNCH_CHECK("make_keys_kernel");
  }
  int rc = exclusive_scan(w.hist, seg, nkeys, w.tile_sums, st);
  if (rc) return rc;
  set_last_kernel<<<1, 1, 0, st>>>(seg, nkeys, (int32_t)N);
  MMSBM_LAUNCH_CHECK("set_last_kernel");
  degrees_kernel<<<(n_ids + 255) / 256, 256, 0, st>>>(seg, n_ids, R, deg);
  MMSBM_LAUNCH_CHECK("degrees_kernel");
  if (N == 0) return 0;

  int bits = 0;
  while (((int64_t)1 << bits) < nkeys) ++bits;
  const int passes = bits == 0 ? 0 : (bits + kRadixBits - 1) / kRadixBits;
  const int64_t n_chunks = (N + kChunk - 1) / kChunk;
  const unsigned grid = (unsigned)((n_chunks + kSortWarps - 1) / kSortWarps);
  uint32_t *ka = w.key_a, *kb = w.key_b, *va = w.val_a, *vb = w.val_b;
  for (int p = 0; p < passes; ++p) {
    const int shift = p * kRadixBits;
    radix_count_kernel<<<grid, kSortWarps * 32, 0, st>>>(ka, N, shift, n_chunks, w.counts);
    MMSBM_LAUNCH_CHECK("radix_count_kernel");
    rc = exclusive_scan(w.counts, w.counts, n_chunks * kRadix, w.tile_sums, st);
    if (rc) return rc;
    radix_scatter_kernel<<<grid, kSortWarps * 32, 0, st>>>(ka, va, N, shift, n_chunks, w.counts, kb, vb);
    MMSBM_LAUNCH_CHECK("radix_scatter_kernel");
    uint32_t* t = ka; ka = kb; kb = t;
    t = va; va = vb; vb = t;
  }
  gather_adj_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(va, other, N, adj, perm);
  MMSBM_LAUNCH_CHECK("gather_adj_kernel");
  return 0;
}

static int build_schedule(const int32_t* deg, int nseg, int64_t N, int32_t* sched, GraphWs& w,
                          cudaStream_t st) {
  const unsigned grid = (unsigned)((nseg + 255) / 256);
  sched_count_kernel<<<grid, 256, 0, st>>>(deg, nseg, w.sa, w.sb, w.sc);
  MMSBM_LAUNCH_CHECK("sched_count_kernel");
  int rc;
  if ((rc = exclusive_scan(w.sa, w.sa, nseg, w.tile_sums, st))) return rc;
  if ((rc = exclusive_scan(w.sb, w.sb, nseg, w.tile_sums, st))) return rc;
  if ((rc = exclusive_scan(w.sc, w.sc, nseg, w.tile_sums, st))) return rc;
  sched_fill_kernel<<<grid, 256, 0, st>>>(deg, nseg, w.sa, w.sb, w.sc, sched_pmax(N, nseg), sched_lmax(N), sched);
  MMSBM_LAUNCH_CHECK("sched_fill_kernel");
  return 0;
}

}  // namespace mmsbm

using namespace mmsbm;

extern "C" int mmsbm_sched_elems(int64_t N, int32_t nseg, int64_t* elems) {
  MMSBM_REQUIRE(elems && N >= 0 && nseg > 0, MMSBM_EINVAL, "mmsbm_sched_elems: bad argument");
  *elems = 4 + 3 * sched_pmax(N, nseg) + 2 * sched_lmax(N) + 1;
  return 0;
}

extern "C" int mmsbm_graph_workspace_bytes(int64_t N, int32_t U, int32_t I, int32_t R, size_t* bytes) {
  MMSBM_REQUIRE(bytes && N >= 0 && U > 0 && I > 0 && R > 0, MMSBM_EINVAL,
                "mmsbm_graph_workspace_bytes: bad argument");
  MMSBM_REQUIRE(N < ((int64_t)1 << 31) && (int64_t)U * R < ((int64_t)1 << 31) &&
                    (int64_t)I * R < ((int64_t)1 << 31), MMSBM_ERANGE,
                "index build: N, U*R and I*R must be below 2^31");
  const int64_t mk = (int64_t)(U > I ? U : I) * R;
  *bytes = graph_ws_layout(N, mk, U > I ? U : I, nullptr, 0, nullptr);
  return 0;
}

extern "C" int mmsbm_graph_build(const int32_t* user, const int32_t* item, const int32_t* level,
                                 int64_t N, int32_t U, int32_t I, int32_t R, int32_t* useg,
                                 int32_t* uadj, int32_t* uperm, int32_t* udeg, int32_t* iseg,
                                 int32_t* iadj, int32_t* iperm, int32_t* ideg, int32_t* usched,
                                 int32_t* isched, void* ws, size_t ws_bytes, void* stream) {
  MMSBM_REQUIRE(useg && udeg && iseg && ideg && usched && isched && ws, MMSBM_EINVAL,
                "mmsbm_graph_build: null pointer");
  MMSBM_REQUIRE(N == 0 || (user && item && level && uadj && uperm && iadj && iperm), MMSBM_EINVAL,
                "mmsbm_graph_build: null pointer");
  size_t need = 0;
  int rc = mmsbm_graph_workspace_bytes(N, U, I, R, &need);
  if (rc) return rc;
  MMSBM_REQUIRE(ws_bytes >= need, MMSBM_ENOMEM, "mmsbm_graph_build: workspace %zu < %zu", ws_bytes, need);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GraphWs w{};
  const int64_t mk = (int64_t)(U > I ? U : I) * R;
  graph_ws_layout(N, mk, U > I ? U : I, ws, ws_bytes, &w);
  MMSBM_REQUIRE(w.bad && w.sc, MMSBM_ENOMEM, "mmsbm_graph_build: workspace carve-up failed");
  if ((rc = build_one(user, item, level, N, U, R, useg, uadj, uperm, udeg, w, st))) return rc;
  if ((rc = build_schedule(udeg, U, N, usched, w, st))) return rc;
  if ((rc = build_one(item, user, level, N, I, R, iseg, iadj, iperm, ideg, w, st))) return rc;
  return build_schedule(ideg, I, N, isched, w, st);
}

// One side only (a rank of a sharded run holds a CSR over its own users built from their ratings
// and a CSC over its own items built from theirs -- two different row sets): `id` selects the
// segment (already shifted to start at 0), `other` is the neighbour id kept in adj (global).
extern "C" int mmsbm_graph_build_side(const int32_t* id, const int32_t* other, const int32_t* level,
                                      int64_t N, int32_t n_ids, int32_t R, int32_t* seg, int32_t* adj,
                                      int32_t* perm, int32_t* deg, int32_t* sched, void* ws,
                                      size_t ws_bytes, void* stream) {
  MMSBM_REQUIRE(seg && deg && sched && ws, MMSBM_EINVAL, "mmsbm_graph_build_side: null pointer");
  MMSBM_REQUIRE(N == 0 || (id && other && level && adj && perm), MMSBM_EINVAL,
                "mmsbm_graph_build_side: null pointer");
  size_t need = 0;
  int rc = mmsbm_graph_workspace_bytes(N, n_ids, n_ids, R, &need);
  if (rc) return rc;
  MMSBM_REQUIRE(ws_bytes >= need, MMSBM_ENOMEM, "mmsbm_graph_build_side: workspace %zu < %zu", ws_bytes, need);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GraphWs w{};
  graph_ws_layout(N, (int64_t)n_ids * R, n_ids, ws, ws_bytes, &w);
  MMSBM_REQUIRE(w.bad && w.sc, MMSBM_ENOMEM, "mmsbm_graph_build_side: workspace carve-up failed");
  if ((rc = build_one(id, other, level, N, n_ids, R, seg, adj, perm, deg, w, st))) return rc;
  return build_schedule(deg, n_ids, N, sched, w, st);
}

// The work schedule alone, for a contiguous RANGE of segments of an index that was built over all
// ids (a rank of a sharded run takes slices of the full index: seg + lo*R, deg + lo, the whole adj
// array -- positions in seg are absolute -- and only needs its own schedule).  n_ratings = ratings of
// the range (it sizes the schedule: mmsbm_sched_elems(n_ratings, n_segments)).
extern "C" int mmsbm_sched_workspace_bytes(int32_t nseg, size_t* bytes) {
  MMSBM_REQUIRE(bytes && nseg > 0, MMSBM_EINVAL, "mmsbm_sched_workspace_bytes: bad argument");
  *bytes = 3 * align_up((size_t)nseg * 4) + align_up((size_t)(scan_tiles_for(nseg) + 1) * 4) + 256;
  return 0;
}
extern "C" int mmsbm_sched_build(const int32_t* deg, int32_t nseg, int64_t n_ratings, int32_t* sched, void* ws,
                                 size_t ws_bytes, void* stream) {
  MMSBM_REQUIRE(deg && sched && ws && nseg > 0 && n_ratings >= 0, MMSBM_EINVAL, "mmsbm_sched_build: bad argument");
  size_t need = 0;
  mmsbm_sched_workspace_bytes(nseg, &need);
  MMSBM_REQUIRE(ws_bytes >= need, MMSBM_ENOMEM, "mmsbm_sched_build: workspace %zu < %zu", ws_bytes, need);
  Arena a(ws, ws_bytes);
  GraphWs w{};
  w.sa = a.take<int32_t>(nseg); w.sb = a.take<int32_t>(nseg); w.sc = a.take<int32_t>(nseg);
  w.tile_sums = a.take<int32_t>(scan_tiles_for(nseg) + 1);
  MMSBM_REQUIRE(w.tile_sums, MMSBM_ENOMEM, "mmsbm_sched_build: workspace carve-up failed");
  return build_schedule(deg, nseg, n_ratings, sched, w, static_cast<cudaStream_t>(stream));
}

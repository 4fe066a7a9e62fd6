// ONE set of S runs sharded over the GPUs of a box (one process per GPU): the multi-GPU form of
// the loop src/mmsbm.py:243-250 for a fit that is too slow (or too large) for one GPU.
//
// Sharding ("owner computes" on both sides, SURVEY.md section 8e.3 taken one step further):
//   rank g owns a contiguous USER range and a contiguous ITEM range, both balanced by rating
//   count.  It holds a CSR over its users built from ALL their ratings and a CSC over its items
//   built from ALL theirs, so n_theta of its users and n_eta of its items are complete local sums:
//   no partial theta / eta sums ever cross the NVLink, and no item-side work is replicated.
//   What a pass gathers are rows of the OTHER side for arbitrary ids, so every rank keeps full
//   gather tables (eta for the by-user pass, theta for the by-item pass, in the run-interleaved
//   layouts of segment_pass.cuh) in an EXCHANGE BUFFER that its peers map through CUDA IPC.
//
// One iteration on a rank (main stream `st`, copy stream `cs`):
//   prep_p, W_u = theta_own x Pw, W_i = eta_own x Pw                      local rows only
//   by-user pass over own users (gathers tables[cur].eta)      -> g_u -> theta_own'
//   interleave theta_own' into own tables[nxt].theta; cs: push that slice into every peer's
//     buffer with the COPY ENGINES (cudaMemcpy2DAsync over NVLink) -- overlaps the by-item pass
//   by-item pass over own items (gathers tables[cur].theta)    -> g_i -> eta_own'
//   interleave eta_own' into tables[nxt].eta; cs: push
//   n_pr partial over own items (or users), then ONE small ncclAllReduce of n_pr [S][K][L][R]
//     after the rank's pushes have completed: it is the only collective of the iteration and
//     doubles as the barrier that tells every rank all slices have landed (and that nobody
//     still reads the tables that the next iteration overwrites); normalise pr.
// The SMs never move parameter rows: slices travel by DMA while the segment pass computes.
// Sums have a fixed order for a fixed world size; vs one GPU only the n_pr order differs.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "em_internal.cuh"

namespace mmsbm {

// ---- NCCL through dlopen (the library torch.distributed already loaded) --------------------
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

#define MMSBM_NCCL(expr)                                                                   \
  do {                                                                                     \
    ncclResult_t r__ = (expr);                                                             \
    if (r__ != ncclSuccess) {                                                              \
      set_error("%s failed: %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "?"); \
      return 1000 + (int)r__;                                                              \
    }                                                                                      \
  } while (0)

static int nccl_ready() {
  MMSBM_REQUIRE(g_nccl.handle, MMSBM_EINVAL, "NCCL is not loaded: call mmsbm_nccl_load first");
  return 0;
}

// ---- layout of the exchange buffer ------------------------------------------------------------
// Two halves (the tables an iteration reads, the tables it fills for the next one); each half:
//   eta  tables of the by-user pass:  hexa [S/6][I][6][ldl] | pairs [S/2][I][2][ldl] | plain [S][I][ldl]
//   theta tables of the by-item pass:                         pairs [S/2][U][2][ldk] | plain [S][U][ldk]
// Every table is addressed as on one GPU (segment_pass.cuh); only the run groups the pass
// really uses (plan_runs) are filled and exchanged.  256 bytes of scratch close the buffer.
struct Table { size_t off; int gs, group0, groups, ld, n_all; };     // off: bytes inside a half
struct GridLayout {
  Table eta[3], theta[2];
  size_t half_bytes, total_bytes;
};

static GridLayout grid_layout(int U, int I, int K, int L, int S) {
  const int ldk = row_stride(K), ldl = row_stride(L);
  const RunPlan pu = plan_runs(ldl, S, true), pi = plan_runs(ldk, S, false);
  GridLayout g{};
  size_t off = 0;
  auto put = [&](Table& t, int gs, int group0, int groups, int ld, int n_all, int addr_groups) {
    t = Table{off, gs, group0, groups, ld, n_all};
    off += align_up((size_t)addr_groups * gs * n_all * ld * 8);
  };
  put(g.eta[0], 6, 0, pu.hexas, ldl, I, pu.hexas);
  put(g.eta[1], 2, pu.pair_group0, pu.pairs, ldl, I, pu.pairs ? S / 2 : 0);
  put(g.eta[2], 1, pu.single_from, S - pu.single_from, ldl, I, S > pu.single_from ? S : 0);
  put(g.theta[0], 2, pi.pair_group0, pi.pairs, ldk, U, pi.pairs ? S / 2 : 0);
  put(g.theta[1], 1, pi.single_from, S - pi.single_from, ldk, U, S > pi.single_from ? S : 0);
  g.half_bytes = off;
  g.total_bytes = 2 * off + 256;
  return g;
}

static const double* table_ptr(const void* base, const GridLayout& g, int half, const Table& t) {
  if (t.groups == 0) return nullptr;
  return reinterpret_cast<const double*>(static_cast<const char*>(base) + (size_t)half * g.half_bytes + t.off);
}

struct PushStreams {
  cudaStream_t cs = nullptr;
  cudaEvent_t ready = nullptr, pushed = nullptr;
};

// own rows [lo, lo + n_own) of one parameter (plain [S][n_own][ld]) -> the tables of half `half`:
// interleave into the own buffer on `st`, then DMA the slice into every peer's buffer on `cs`
static int publish_side(const mmsbm_shard_t& sh, const GridLayout& g, const Table* tabs, int n_tabs,
                        const double* own_rows, int n_own, int lo, int half, cudaStream_t st,
                        const PushStreams& ps) {
  char* mine = static_cast<char*>(sh.exchange_dev[sh.rank]);
  for (int k = 0; k < n_tabs; ++k) {
    const Table& t = tabs[k];
    if (t.groups == 0) continue;
    double* dst = reinterpret_cast<double*>(mine + (size_t)half * g.half_bytes + t.off);
    int rc = launch_interleave(own_rows, dst, n_own, t.n_all, lo, t.ld, t.group0, t.groups, t.gs, st);
    if (rc) return rc;
  }
  if (sh.world == 1) return 0;
  MMSBM_CUDA(cudaEventRecord(ps.ready, st));
  MMSBM_CUDA(cudaStreamWaitEvent(ps.cs, ps.ready, 0));
  for (int d = 1; d < sh.world; ++d) {
    const int peer = (sh.rank + d) % sh.world;               // staggered: no two ranks start on one peer
    char* theirs = static_cast<char*>(sh.exchange_dev[peer]);
    for (int k = 0; k < n_tabs; ++k) {
      const Table& t = tabs[k];
      if (t.groups == 0) continue;
      const size_t row_bytes = (size_t)t.gs * t.ld * 8;
      const size_t pitch = (size_t)t.n_all * row_bytes;
      const size_t o = (size_t)half * g.half_bytes + t.off + ((size_t)t.group0 * t.n_all + lo) * row_bytes;
      MMSBM_CUDA(cudaMemcpy2DAsync(theirs + o, pitch, mine + o, pitch, (size_t)n_own * row_bytes,
                                   (size_t)t.groups, cudaMemcpyDeviceToDevice, ps.cs));
    }
  }
  return 0;
}

static int check_shard(const mmsbm_shard_t* sh, const char* who) {
  MMSBM_REQUIRE(sh, MMSBM_EINVAL, "%s: null shard", who);
  MMSBM_REQUIRE(sh->n_users > 0 && sh->n_items > 0 && sh->n_levels > 0 && sh->K > 0 && sh->L > 0 &&
                    sh->n_runs > 0 && sh->n_users_own > 0 && sh->n_items_own > 0 && sh->n_ratings_u >= 0 &&
                    sh->n_ratings_i >= 0, MMSBM_EINVAL, "%s: bad size", who);
  MMSBM_REQUIRE(sh->user_lo >= 0 && sh->user_lo + sh->n_users_own <= sh->n_users && sh->item_lo >= 0 &&
                    sh->item_lo + sh->n_items_own <= sh->n_items, MMSBM_EINVAL, "%s: own range outside the ids", who);
  MMSBM_REQUIRE(sh->world >= 1 && sh->rank >= 0 && sh->rank < sh->world && sh->exchange_dev, MMSBM_EINVAL,
                "%s: bad rank / world / exchange", who);
  for (int r = 0; r < sh->world; ++r)
    MMSBM_REQUIRE(sh->exchange_dev[r], MMSBM_EINVAL, "%s: exchange pointer of rank %d is null", who, r);
  MMSBM_REQUIRE(sh->world == 1 || sh->nccl_comm, MMSBM_EINVAL, "%s: world > 1 needs an NCCL communicator", who);
  MMSBM_REQUIRE(sh->K <= 32 && sh->L <= 32 && sh->n_levels <= 31, MMSBM_ERANGE,
                "%s: sharded runs support K, L <= 32 and R <= 31", who);
  return 0;
}

struct ShardDims {
  int ldk, ldl, rnb_u, rnb_i;
  bool emit_items;
  int64_t pmax_u, pmax_i, lmax_u, lmax_i, smax_u, smax_i;
  size_t p_elems, wg_u, wg_i, partial, slots_u, slots_i, ctr;
};

static ShardDims shard_dims(const mmsbm_shard_t& s) {
  ShardDims d;
  const int S = s.n_runs, R = s.n_levels;
  d.ldk = row_stride(s.K); d.ldl = row_stride(s.L);
  d.rnb_u = R * d.ldl; d.rnb_i = R * d.ldk;
  d.emit_items = s.n_items <= s.n_users;
  d.pmax_u = (int64_t)s.n_users_own + s.n_ratings_u / MMSBM_PIECE_LEN + 1;      // = graph_build.cu
  d.pmax_i = (int64_t)s.n_items_own + s.n_ratings_i / MMSBM_PIECE_LEN + 1;
  d.lmax_u = s.n_ratings_u / MMSBM_PIECE_LEN + 1; d.lmax_i = s.n_ratings_i / MMSBM_PIECE_LEN + 1;
  d.smax_u = 2 * (s.n_ratings_u / MMSBM_PIECE_LEN) + 1; d.smax_i = 2 * (s.n_ratings_i / MMSBM_PIECE_LEN) + 1;
  d.p_elems = (size_t)S * d.ldk * d.ldl * R;
  d.wg_u = (size_t)S * s.n_users_own * d.rnb_u;
  d.wg_i = (size_t)S * s.n_items_own * d.rnb_i;
  d.partial = (size_t)S * kPrSlabs * (d.emit_items ? s.L * R * d.ldk : s.K * R * d.ldl);
  d.slots_u = (size_t)S * d.smax_u * d.rnb_u;
  d.slots_i = (size_t)S * d.smax_i * d.rnb_i;
  d.ctr = 2 * ((size_t)S + 8);
  return d;
}

}  // namespace mmsbm

using namespace mmsbm;

// ---- NCCL plumbing -----------------------------------------------------------------------------
extern "C" int mmsbm_nccl_load(const char* path) {
  if (g_nccl.handle) return 0;
  void* h = dlopen((path && *path) ? path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  MMSBM_REQUIRE(h, MMSBM_ENODEV, "cannot load NCCL (%s): %s", (path && *path) ? path : "libnccl.so.2", dlerror());
  NcclApi api;
  api.handle = h;
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
  api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(h, "ncclAllReduce"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
  MMSBM_REQUIRE(api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GetErrorString,
                MMSBM_ENODEV, "NCCL library lacks a required symbol");
  g_nccl = api;
  return 0;
}

extern "C" int mmsbm_nccl_unique_id(unsigned char* id128) {
  int rc = nccl_ready();
  if (rc) return rc;
  MMSBM_REQUIRE(id128, MMSBM_EINVAL, "mmsbm_nccl_unique_id: null output");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  MMSBM_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, 128);
  return 0;
}

extern "C" int mmsbm_nccl_comm_init(const unsigned char* id128, int32_t rank, int32_t world, void** comm) {
  int rc = nccl_ready();
  if (rc) return rc;
  MMSBM_REQUIRE(id128 && comm && world >= 1 && rank >= 0 && rank < world, MMSBM_EINVAL,
                "mmsbm_nccl_comm_init: bad argument");
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t c = nullptr;
  MMSBM_NCCL(g_nccl.CommInitRank(&c, world, id, rank));
  *comm = c;
  return 0;
}

extern "C" int mmsbm_nccl_comm_destroy(void* comm) {
  if (!comm || !g_nccl.handle) return 0;
  MMSBM_NCCL(g_nccl.CommDestroy(static_cast<ncclComm_t>(comm)));
  return 0;
}

// ---- peer-mappable device memory (CUDA IPC) -----------------------------------------------------
extern "C" int mmsbm_ipc_alloc(size_t bytes, void** dev_ptr, unsigned char* handle64) {
  MMSBM_REQUIRE(dev_ptr && handle64 && bytes > 0, MMSBM_EINVAL, "mmsbm_ipc_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  void* p = nullptr;
  MMSBM_CUDA(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return (int)e;
  }
  memcpy(handle64, &h, 64);
  *dev_ptr = p;
  return 0;
}
extern "C" int mmsbm_ipc_open(const unsigned char* handle64, void** dev_ptr) {
  MMSBM_REQUIRE(dev_ptr && handle64, MMSBM_EINVAL, "mmsbm_ipc_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  MMSBM_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
extern "C" int mmsbm_ipc_close(void* dev_ptr) {
  if (dev_ptr) MMSBM_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return 0;
}
extern "C" int mmsbm_ipc_free(void* dev_ptr) {
  if (dev_ptr) MMSBM_CUDA(cudaFree(dev_ptr));
  return 0;
}

// ---- sizes ------------------------------------------------------------------------------------------
extern "C" int mmsbm_shard_exchange_bytes(int32_t U, int32_t I, int32_t K, int32_t L, int32_t S, size_t* bytes) {
  MMSBM_REQUIRE(bytes && U > 0 && I > 0 && K > 0 && L > 0 && S > 0, MMSBM_EINVAL,
                "mmsbm_shard_exchange_bytes: bad argument");
  *bytes = grid_layout(U, I, K, L, S).total_bytes;
  return 0;
}

extern "C" int mmsbm_shard_workspace_bytes(const mmsbm_shard_t* sh, size_t* bytes) {
  int rc = check_shard(sh, "mmsbm_shard_workspace_bytes");
  if (rc) return rc;
  MMSBM_REQUIRE(bytes, MMSBM_EINVAL, "mmsbm_shard_workspace_bytes: null output");
  const ShardDims d = shard_dims(*sh);
  *bytes = 4 * align_up(d.p_elems * 8) + align_up(d.wg_u * 8) + align_up(d.wg_i * 8) + align_up(d.partial * 8) +
           align_up(d.slots_u * 8) + align_up(d.slots_i * 8) + align_up(d.ctr * 4) + 256;
  return 0;
}

// The tables of half `half` <- this rank's rows of theta and eta, on every rank; returns after the
// NCCL barrier has been enqueued on `stream` (every rank must call it).
extern "C" int mmsbm_shard_publish(const mmsbm_shard_t* sh, const double* theta_own, const double* eta_own,
                                   int32_t half, void* stream) {
  int rc = check_shard(sh, "mmsbm_shard_publish");
  if (rc) return rc;
  MMSBM_REQUIRE(theta_own && eta_own && (half == 0 || half == 1), MMSBM_EINVAL, "mmsbm_shard_publish: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const GridLayout g = grid_layout(sh->n_users, sh->n_items, sh->K, sh->L, sh->n_runs);
  PushStreams ps;
  struct Cleanup {
    PushStreams& p;
    ~Cleanup() {
      if (p.ready) cudaEventDestroy(p.ready);
      if (p.pushed) cudaEventDestroy(p.pushed);
      if (p.cs) cudaStreamDestroy(p.cs);
    }
  } cleanup{ps};
  if (sh->world > 1) {
    MMSBM_CUDA(cudaStreamCreateWithFlags(&ps.cs, cudaStreamNonBlocking));
    MMSBM_CUDA(cudaEventCreateWithFlags(&ps.ready, cudaEventDisableTiming));
    MMSBM_CUDA(cudaEventCreateWithFlags(&ps.pushed, cudaEventDisableTiming));
  }
  if ((rc = publish_side(*sh, g, g.theta, 2, theta_own, sh->n_users_own, sh->user_lo, half, st, ps))) return rc;
  if ((rc = publish_side(*sh, g, g.eta, 3, eta_own, sh->n_items_own, sh->item_lo, half, st, ps))) return rc;
  if (sh->world > 1) {
    if ((rc = nccl_ready())) return rc;
    MMSBM_CUDA(cudaEventRecord(ps.pushed, ps.cs));
    MMSBM_CUDA(cudaStreamWaitEvent(st, ps.pushed, 0));
    double* scratch = reinterpret_cast<double*>(static_cast<char*>(sh->exchange_dev[sh->rank]) + 2 * g.half_bytes);
    MMSBM_CUDA(cudaMemsetAsync(scratch, 0, 8, st));
    MMSBM_NCCL(g_nccl.AllReduce(scratch, scratch, 1, ncclFloat64, ncclSum, static_cast<ncclComm_t>(sh->nccl_comm), st));
  }
  return 0;
}

// `iterations` EM steps of the sharded loop.  (theta, eta)_a/_b: this rank's OWN rows, plain layout
// [S][n_users_own][ldk] / [S][n_items_own][ldl]; pr_a/_b: [S][K][L][R], replicated.  The tables of half
// `half` must hold the parameters of the _a buffers on every rank (mmsbm_shard_publish).  The result is
// in _a (tables in `half`) when iterations is even, else in _b (tables in half ^ 1).
// prof (optional, 2 floats): mean device ms per iteration, and of it the mean ms between the end of the
// rank's compute and the end of the n_pr all-reduce (exposed exchange + barrier); synchronises.
extern "C" int mmsbm_em_run_sharded(const mmsbm_shard_t* shp, int32_t iterations, double* theta_a, double* eta_a,
                                    double* pr_a, double* theta_b, double* eta_b, double* pr_b, int32_t half,
                                    void* ws, size_t ws_bytes, void* stream, float* prof) {
  int rc = check_shard(shp, "mmsbm_em_run_sharded");
  if (rc) return rc;
  const mmsbm_shard_t& sh = *shp;
  MMSBM_REQUIRE(iterations >= 0 && theta_a && eta_a && pr_a && theta_b && eta_b && pr_b && ws &&
                    (half == 0 || half == 1), MMSBM_EINVAL, "mmsbm_em_run_sharded: bad argument");
  MMSBM_REQUIRE(sh.useg_dev && sh.udeg_dev && sh.usched_dev && sh.iseg_dev && sh.ideg_dev && sh.isched_dev &&
                    (sh.n_ratings_u == 0 || sh.uadj_dev) && (sh.n_ratings_i == 0 || sh.iadj_dev), MMSBM_EINVAL,
                "mmsbm_em_run_sharded: null index pointer");
  if (sh.world > 1 && (rc = nccl_ready())) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int S = sh.n_runs, R = sh.n_levels, K = sh.K, L = sh.L;
  const int Uo = sh.n_users_own, Io = sh.n_items_own;
  const ShardDims d = shard_dims(sh);
  const GridLayout g = grid_layout(sh.n_users, sh.n_items, K, L, S);
  Arena arena(ws, ws_bytes);
  double* pw_u = arena.take<double>(d.p_elems);
  double* pn_u = arena.take<double>(d.p_elems);
  double* pw_i = arena.take<double>(d.p_elems);
  double* pn_i = arena.take<double>(d.p_elems);
  double* wg_u = arena.take<double>(d.wg_u);
  double* wg_i = arena.take<double>(d.wg_i);
  double* partial = arena.take<double>(d.partial);
  double* slots_u = arena.take<double>(d.slots_u);
  double* slots_i = arena.take<double>(d.slots_i);
  int32_t* counters = arena.take<int32_t>(d.ctr);
  MMSBM_REQUIRE(pw_u && pn_u && pw_i && pn_i && wg_u && wg_i && partial && slots_u && slots_i && counters,
                MMSBM_ENOMEM, "mmsbm_em_run_sharded: workspace too small (%zu)", ws_bytes);

  PushStreams ps;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // profiling: start, compute done, exchange done, end
  struct Cleanup {
    PushStreams& p;
    cudaEvent_t* ev;
    ~Cleanup() {
      if (p.ready) cudaEventDestroy(p.ready);
      if (p.pushed) cudaEventDestroy(p.pushed);
      if (p.cs) cudaStreamDestroy(p.cs);
      for (int k = 0; k < 4; ++k) if (ev[k]) cudaEventDestroy(ev[k]);
    }
  } cleanup{ps, ev};
  if (sh.world > 1) {
    MMSBM_CUDA(cudaStreamCreateWithFlags(&ps.cs, cudaStreamNonBlocking));
    MMSBM_CUDA(cudaEventCreateWithFlags(&ps.ready, cudaEventDisableTiming));
    MMSBM_CUDA(cudaEventCreateWithFlags(&ps.pushed, cudaEventDisableTiming));
  }
  if (prof) for (int k = 0; k < 4; ++k) MMSBM_CUDA(cudaEventCreate(&ev[k]));
  double prof_iter = 0.0, prof_exch = 0.0;

  const int64_t n_big = sh.n_ratings_u > sh.n_ratings_i ? sh.n_ratings_u : sh.n_ratings_i;
  const bool dyn = env_int("MMSBM_DYN", n_big >= ((int64_t)1 << 22) ? 1 : 0) != 0;
  const void* mine = sh.exchange_dev[sh.rank];

  for (int it = 0; it < iterations; ++it) {
    const bool fwd = (it & 1) == 0;
    const int cur = half ^ (it & 1), nxt = cur ^ 1;
    const double* th = fwd ? theta_a : theta_b; const double* et = fwd ? eta_a : eta_b;
    const double* pr = fwd ? pr_a : pr_b;
    double* th_n = fwd ? theta_b : theta_a; double* et_n = fwd ? eta_b : eta_a; double* pr_n = fwd ? pr_b : pr_a;
    if (prof) MMSBM_CUDA(cudaEventRecord(ev[0], st));
    if (dyn) MMSBM_CUDA(cudaMemsetAsync(counters, 0, d.ctr * 4, st));
    if ((rc = launch_prep_p(pr, K, L, R, d.ldk, d.ldl, S, pw_u, pn_u, pw_i, pn_i, st))) return rc;
    if ((rc = launch_w(th, pw_u, wg_u, Uo, d.ldk, d.rnb_u, S, st))) return rc;
    if ((rc = launch_w(et, pw_i, wg_i, Io, d.ldl, d.rnb_i, S, st))) return rc;
    // ---- by-user pass over the own users: gathers eta rows of ALL items ----
    {
      SegArgs a{sh.useg_dev, sh.uadj_dev, sh.usched_dev, table_ptr(mine, g, cur, g.eta[2]), wg_u, slots_u, d.pmax_u,
                d.lmax_u, d.smax_u, Uo, sh.n_items, d.ldl, R, 0, 0, 0, dyn ? counters : nullptr};
      if ((rc = launch_segment_pass_and_fixup(a, table_ptr(mine, g, cur, g.eta[1]),
                                              table_ptr(mine, g, cur, g.eta[0]), sh.n_ratings_u, S, st))) return rc;
    }
    if ((rc = launch_n(wg_u, pn_u, th, sh.udeg_dev, th_n, Uo, d.ldk, d.rnb_u, 1, S, st))) return rc;
    // theta' of the own users -> every rank's next tables; the DMA overlaps the by-item pass
    if ((rc = publish_side(sh, g, g.theta, 2, th_n, Uo, sh.user_lo, nxt, st, ps))) return rc;
    // ---- by-item pass over the own items: gathers theta rows of ALL users ----
    {
      SegArgs a{sh.iseg_dev, sh.iadj_dev, sh.isched_dev, table_ptr(mine, g, cur, g.theta[1]), wg_i, slots_i, d.pmax_i,
                d.lmax_i, d.smax_i, Io, sh.n_users, d.ldk, R, 0, 0, 0, dyn ? counters + d.ctr / 2 : nullptr};
      if ((rc = launch_segment_pass_and_fixup(a, table_ptr(mine, g, cur, g.theta[0]), nullptr,
                                              sh.n_ratings_i, S, st))) return rc;
    }
    if ((rc = launch_n(wg_i, pn_i, et, sh.ideg_dev, et_n, Io, d.ldl, d.rnb_i, 1, S, st))) return rc;
    if ((rc = publish_side(sh, g, g.eta, 3, et_n, Io, sh.item_lo, nxt, st, ps))) return rc;
    // ---- n_pr: partial over the own segments of the emitting side, summed over the ranks ----
    if ((rc = launch_pr(d.emit_items ? et : th, d.emit_items ? wg_i : wg_u, partial, pr, pr_n,
                        d.emit_items ? Io : Uo, d.emit_items ? L : K, d.emit_items ? d.ldl : d.ldk,
                        d.emit_items ? d.ldk : d.ldl, K, L, R, S, d.emit_items, sh.world == 1, st))) return rc;
    if (prof) MMSBM_CUDA(cudaEventRecord(ev[1], st));
    if (sh.world > 1) {
      MMSBM_CUDA(cudaEventRecord(ps.pushed, ps.cs));
      MMSBM_CUDA(cudaStreamWaitEvent(st, ps.pushed, 0));      // this rank's slices have landed everywhere
      MMSBM_NCCL(g_nccl.AllReduce(pr_n, pr_n, (size_t)S * K * L * R, ncclFloat64, ncclSum,
                                  static_cast<ncclComm_t>(sh.nccl_comm), st));
      if ((rc = launch_finalize_pr(pr_n, S * K * L, R, st))) return rc;
    }
    if (prof) {
      MMSBM_CUDA(cudaEventRecord(ev[2], st));
      MMSBM_CUDA(cudaEventSynchronize(ev[2]));
      float a = 0.f, b = 0.f;
      cudaEventElapsedTime(&a, ev[0], ev[2]);
      cudaEventElapsedTime(&b, ev[1], ev[2]);
      prof_iter += a; prof_exch += b;
    }
  }
  if (prof) {
    prof[0] = iterations > 0 ? (float)(prof_iter / iterations) : 0.f;
    prof[1] = iterations > 0 ? (float)(prof_exch / iterations) : 0.f;
  }
  // the copy stream is destroyed on return: its work is ordered before the last all-reduce on `st`
  return 0;
}

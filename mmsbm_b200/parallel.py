"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on the box, gloo in
the CPU tests) for rendezvous and the small host-side exchanges.  Three ways the path shards
(SURVEY.md section 8e):

  runs        independent ``sampling`` runs round-robin over ranks; no data-path
              collective, one gather of the results at the end (src/mmsbm.py:182-185
              fans the same runs out to a process pool);
  folds       cv_fit folds x runs are the same thing one level up;
  ratings     ONE set of runs split over all ranks (``ShardedEngine``): rank g owns a
              contiguous USER range and a contiguous ITEM range, both balanced by rating count,
              with all the ratings of those users (CSR) and all the ratings of those items (CSC).
              n_theta / n_eta of owned ids are complete local sums; the rows a pass gathers are
              kept in an exchange buffer that every rank fills on every peer by copy-engine DMA
              over NVLink (CUDA IPC) while the next pass computes, and the only collective per
              iteration is one NCCL all-reduce of n_pr.  The loop itself runs inside the library
              (mmsbm_em_run_sharded, csrc/sharded_run.cu); this module only sets it up.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist


def dist_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_runs(n_runs, rank, world):
    """Run indices of ``rank``: round-robin, so every rank gets ceil or floor(S/world)."""
    return list(range(rank, n_runs, world))


def shard_jobs(n_folds, n_runs, rank, world):
    """(fold, run) jobs of ``rank`` for cv_fit: the folds x runs grid in fold-major order cut into
    ``world`` contiguous blocks whose sizes differ by at most one (src/mmsbm.py:420-457 runs the
    jobs serially).  Contiguous blocks keep the runs of a fold together: a rank encodes and
    indexes as few folds as possible and batches their runs in one launch."""
    jobs = [(f, s) for f in range(n_folds) for s in range(n_runs)]
    base, extra = divmod(len(jobs), world)
    lo = rank * base + min(rank, extra)
    return jobs[lo:lo + base + (1 if rank < extra else 0)]


def gather_runs(local, n_runs):
    """``local``: {run index: result dict} of this rank -> list of all S results, in run
    order, on every rank."""
    rank, world = dist_info()
    if world == 1:
        return [local[s] for s in range(n_runs)]
    boxes = [None] * world
    dist.all_gather_object(boxes, local)
    merged = {}
    for b in boxes:
        merged.update(b)
    return [merged[s] for s in range(n_runs)]


def broadcast_seed(seed):
    """One seed for every rank.  With ``seed=None`` each process would draw its own OS entropy:
    ranks would then build different child seeds (different theta0/eta0/pr0 for the 'same' run)
    and different cv folds.  Rank 0 draws the entropy and everybody uses it."""
    rank, world = dist_info()
    if world == 1 or seed is not None:
        return seed
    box = [int(np.random.SeedSequence().entropy) if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def balanced_partition(counts, world):
    """Contiguous id ranges [lo, hi) per rank with the sums of ``counts`` as equal as a prefix
    split allows and AT LEAST ONE id per rank (a rank without ids would sit out of the
    collectives and hang the others).  Returns world+1 int64 boundaries."""
    counts = np.asarray(counts, dtype=np.int64)
    n = len(counts)
    if world > n:
        raise ValueError(f"{world} ranks for {n} ids: every rank needs at least one")
    csum = np.concatenate([[0], np.cumsum(counts)])
    total = csum[-1]
    bounds = [0]
    for r in range(1, world):
        cut = int(np.searchsorted(csum, total * r // world, side="left"))
        cut = max(cut, bounds[-1] + 1)              # at least one id for rank r-1 ...
        cut = min(cut, n - (world - r))             # ... and for each of the ranks still to come
        bounds.append(cut)
    bounds.append(n)
    return np.asarray(bounds, dtype=np.int64)


def user_partition(user_degree, world):
    """Contiguous user ranges balanced by rating count (see ``balanced_partition``)."""
    return balanced_partition(user_degree, world)


def shard_rows(data, n_users, n_items, rank, world):
    """The two row sets of ``rank``: rows of its own users (user ids shifted to start at 0, item
    ids global) and rows of its own items (item ids shifted, user ids global), plus the two
    partitions.  Every rating appears in exactly one rank's user rows and one rank's item rows."""
    data = np.asarray(data)
    ub = balanced_partition(np.bincount(data[:, 0], minlength=n_users), world)
    ib = balanced_partition(np.bincount(data[:, 1], minlength=n_items), world)
    ulo, uhi, ilo, ihi = int(ub[rank]), int(ub[rank + 1]), int(ib[rank]), int(ib[rank + 1])
    rows_u = data[(data[:, 0] >= ulo) & (data[:, 0] < uhi)].copy()
    rows_u[:, 0] -= ulo
    rows_i = data[(data[:, 1] >= ilo) & (data[:, 1] < ihi)].copy()
    rows_i[:, 1] -= ilo
    return rows_u, rows_i, ub, ib


def destroy_communicators():
    """Destroy the cached NCCL communicators of this process (before the process group goes)."""
    from . import _lib
    lib = _lib.load()
    for comm in _COMM_CACHE.values():
        lib.mmsbm_nccl_comm_destroy(comm)
    _COMM_CACHE.clear()


def _nccl_library_path():
    """The libnccl this process already has mapped (torch's bundled one), so that the library
    binds the same NCCL the process group uses; None lets dlopen search for libnccl.so.2."""
    try:
        with open("/proc/self/maps") as fh:
            for line in fh:
                if "libnccl" in line and ".so" in line:
                    return line.split()[-1]
    except OSError:
        pass
    try:
        import nvidia
        import os
        for base in nvidia.__path__:
            cand = os.path.join(base, "nccl", "lib", "libnccl.so.2")
            if os.path.exists(cand):
                return cand
    except ImportError:
        pass
    return None


_COMM_CACHE = {}          # (device, rank, world) -> ncclComm_t of this process (created once, kept)
SEGMENT_COST = 48         # ratings an id is worth in the partition balance (see ShardedEngine)


class ShardedEngine:
    """S runs sharded over the ranks of the default process group by user range x item range
    (module docstring).  Same surface as ``Engine``: ``set_params`` / ``run`` / ``likelihood`` /
    ``get_params``; full-size arrays go in and come out on every rank.

    Set-up: every rank uploads the rows once and builds the FULL index on its GPU (one sort per
    side, a few ms even at 1e8 ratings), derives the two partitions from the device degree
    vectors, and then works on SLICES of that index: ``seg + lo*R`` and ``deg + lo`` of its own
    range (positions in ``seg`` are absolute, so ``adj`` is shared) plus a schedule built for the
    range alone (``mmsbm_sched_build``).  No host-side row shuffling."""

    def __init__(self, data, n_users, n_items, n_levels, K, L, device=None, pretend=None):
        """``pretend=(rank, world)`` (measurement only, one process): take the ranges rank ``rank`` of
        ``world`` would own, with no peers -- the gather tables are filled once from all rows and
        then only the own slices are refreshed, so the NUMBERS are meaningless after the first
        iteration but the kernels do exactly one rank's work (profiles/scripts/rank_compute.py)."""
        from . import _lib
        from .engine import Engine
        self._lib = _lib
        self.lib = _lib.load(require_device=True)
        self.rank, self.world = dist_info()
        self._pretend = pretend
        if pretend is not None:
            self.rank, self.world = int(pretend[0]), int(pretend[1])
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None \
            else torch.device(device)
        self.U, self.I, self.R, self.K, self.L = int(n_users), int(n_items), int(n_levels), int(K), int(L)
        if self.world > min(self.U, self.I):
            raise ValueError(f"{self.world} ranks for {self.U} users x {self.I} items: every rank needs one of each")
        self.ldk, self.ldl = self.lib.mmsbm_row_stride(self.K), self.lib.mmsbm_row_stride(self.L)
        self.S = 0
        self._exchange = None
        self._peers = []
        self._comm = None
        self._ptrs = None
        base = Engine(data, self.U, self.I, self.R, self.K, self.L, device=self.device)   # the full index
        self._base = base
        self.N = base.N
        R = self.R
        with torch.cuda.device(self.device):
            # balance = ratings + SEGMENT_COST per id: what an id costs besides its ratings (its W / G
            # rows, the two contractions, the per-segment part of the pass) is worth about that many
            # ratings (ML-20M shape: 4.5 ns per user-run against 0.125 ns per rating-update, plus the
            # in-pass overhead); without it the rank that owns the long tail of a heavy-tailed id
            # distribution does most of the per-id work (Zipf ids, 8 GPUs: 3.5 instead of 1.1 ms)
            self.ub = balanced_partition(base.udeg[:self.U].cpu().numpy().astype(np.int64) + SEGMENT_COST, self.world)
            self.ib = balanced_partition(base.ideg[:self.I].cpu().numpy().astype(np.int64) + SEGMENT_COST, self.world)
            self.ulo, self.uhi = int(self.ub[self.rank]), int(self.ub[self.rank + 1])
            self.ilo, self.ihi = int(self.ib[self.rank]), int(self.ib[self.rank + 1])
            self.Uo, self.Io = self.uhi - self.ulo, self.ihi - self.ilo
            ends = torch.stack([base.useg[self.ulo * R], base.useg[self.uhi * R],
                                base.iseg[self.ilo * R], base.iseg[self.ihi * R]]).cpu().numpy()
            self.Nu, self.Ni = int(ends[1] - ends[0]), int(ends[3] - ends[2])
            self.useg, self.uadj, self.udeg = base.useg[self.ulo * R:], base.uadj, base.udeg[self.ulo:]
            self.iseg, self.iadj, self.ideg = base.iseg[self.ilo * R:], base.iadj, base.ideg[self.ilo:]
            self.usched = self._build_sched(self.udeg, self.Uo, self.Nu)
            self.isched = self._build_sched(self.ideg, self.Io, self.Ni)
            if self.world > 1 and pretend is None:
                self._open_nccl()
        if pretend is not None:                 # the library sees a single rank that owns a sub-range
            self._true_world, self.world = self.world, 1
            self.rank = 0

    # ------------------------------------------------------------------ set-up
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _build_sched(self, deg, n_seg, n_ratings):
        lib, _lib = self.lib, self._lib
        ne, need = C.c_int64(0), C.c_size_t(0)
        _lib.check(lib.mmsbm_sched_elems(n_ratings, n_seg, C.byref(ne)), "sched_elems")
        _lib.check(lib.mmsbm_sched_workspace_bytes(n_seg, C.byref(need)), "sched_workspace_bytes")
        sched = torch.empty(ne.value, dtype=torch.int32, device=self.device)
        ws = torch.empty(max(need.value, 16), dtype=torch.uint8, device=self.device)
        _lib.check(lib.mmsbm_sched_build(deg.data_ptr(), n_seg, n_ratings, sched.data_ptr(), ws.data_ptr(),
                                         need.value, self._stream()), "sched_build")
        torch.cuda.current_stream(self.device).synchronize()
        return sched

    def _open_nccl(self):
        """One communicator per process and group shape, created on first use and kept (creating
        one costs about a second; engines come and go)."""
        lib, _lib = self.lib, self._lib
        key = (self.device.index, self.rank, self.world)
        if key in _COMM_CACHE:
            self._comm = _COMM_CACHE[key]
            return
        path = _nccl_library_path()
        _lib.check(lib.mmsbm_nccl_load(path.encode() if path else None), "nccl_load")
        uid = (C.c_ubyte * 128)()
        if self.rank == 0:
            _lib.check(lib.mmsbm_nccl_unique_id(uid), "nccl_unique_id")
        box = [bytes(uid)]
        dist.broadcast_object_list(box, src=0)
        uid = (C.c_ubyte * 128).from_buffer_copy(box[0])
        comm = C.c_void_p()
        _lib.check(lib.mmsbm_nccl_comm_init(uid, self.rank, self.world, C.byref(comm)), "nccl_comm_init")
        self._comm = _COMM_CACHE[key] = comm

    def _open_exchange(self, S):
        lib, _lib = self.lib, self._lib
        self.close_exchange()
        need = C.c_size_t(0)
        _lib.check(lib.mmsbm_shard_exchange_bytes(self.U, self.I, self.K, self.L, S, C.byref(need)),
                   "shard_exchange_bytes")
        self.exchange_bytes = need.value
        ptrs = (C.c_void_p * self.world)()
        if self.world == 1:
            self._exchange_t = torch.empty(need.value, dtype=torch.uint8, device=self.device)
            ptrs[0] = self._exchange_t.data_ptr()
        else:
            mine, handle = C.c_void_p(), (C.c_ubyte * 64)()
            _lib.check(lib.mmsbm_ipc_alloc(need.value, C.byref(mine), handle), "ipc_alloc")
            self._exchange = mine
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle))
            for r in range(self.world):
                if r == self.rank:
                    ptrs[r] = mine.value
                    continue
                p = C.c_void_p()
                _lib.check(lib.mmsbm_ipc_open((C.c_ubyte * 64).from_buffer_copy(handles[r]), C.byref(p)),
                           f"ipc_open (rank {r})")
                self._peers.append(p)
                ptrs[r] = p.value
        self._ptrs = ptrs

    def close_exchange(self):
        """Unmap the peers' buffers and free the own one (collective: peers must be done with it)."""
        if self._exchange is None and not self._peers:
            return
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier()
        for p in self._peers:
            self.lib.mmsbm_ipc_close(p)
        self._peers = []
        if self.world > 1:
            dist.barrier()
        if self._exchange is not None:
            self.lib.mmsbm_ipc_free(self._exchange)
            self._exchange = None

    def close(self):
        """Release the exchange buffer (collective).  The communicator stays cached for the next
        engine of this process; ``destroy_communicators`` releases it."""
        self.close_exchange()
        self._comm = None

    def _shard(self):
        s = self._lib.Shard()
        s.useg, s.uadj, s.udeg, s.usched = (t.data_ptr() for t in (self.useg, self.uadj, self.udeg, self.usched))
        s.iseg, s.iadj, s.ideg, s.isched = (t.data_ptr() for t in (self.iseg, self.iadj, self.ideg, self.isched))
        s.n_ratings_u, s.n_ratings_i = self.Nu, self.Ni
        s.n_users_own, s.user_lo, s.n_items_own, s.item_lo = self.Uo, self.ulo, self.Io, self.ilo
        s.n_users, s.n_items, s.n_levels, s.K, s.L, s.n_runs = self.U, self.I, self.R, self.K, self.L, self.S
        s.rank, s.world = self.rank, self.world
        s.exchange = C.cast(self._ptrs, C.POINTER(C.c_void_p))
        s.nccl_comm = self._comm
        return s

    # ---------------------------------------------------------------- parameters
    @staticmethod
    def _pad(x, ld):
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.shape[-1] == ld:
            return x
        out = np.zeros(x.shape[:-1] + (ld,), dtype=np.float64)
        out[..., :x.shape[-1]] = x
        return out

    def set_params(self, theta, eta, pr):
        """theta [S,U,K], eta [S,I,L], pr [S,K,L,R]: the FULL arrays on every rank (each keeps
        its own rows), then the gather tables are filled on all ranks (collective)."""
        theta, eta, pr = np.asarray(theta), np.asarray(eta), np.asarray(pr)
        if theta.ndim == 2:
            theta, eta, pr = theta[None], eta[None], pr[None]
        S = theta.shape[0]
        if theta.shape != (S, self.U, self.K) or eta.shape != (S, self.I, self.L) \
                or pr.shape != (S, self.K, self.L, self.R):
            raise ValueError("parameter shapes do not match the engine")
        dev = self.device
        with torch.cuda.device(dev):
            if S != self.S or getattr(self, "_ptrs", None) is None:
                self.S = S
                self._open_exchange(S)
                self._shard_c = self._shard()
                need = C.c_size_t(0)
                self._lib.check(self.lib.mmsbm_shard_workspace_bytes(C.byref(self._shard_c), C.byref(need)),
                                "shard_workspace_bytes")
                self._ws = torch.empty(max(need.value, 16), dtype=torch.uint8, device=dev)
                self._ws_bytes = need.value
            if self._pretend is not None:       # fill the tables from ALL rows once (a shard that owns everything)
                full = self._shard()
                full.n_users_own, full.user_lo, full.n_items_own, full.item_lo = self.U, 0, self.I, 0
                th_all = torch.from_numpy(self._pad(theta, self.ldk)).to(dev)
                et_all = torch.from_numpy(self._pad(eta, self.ldl)).to(dev)
                self._lib.check(self.lib.mmsbm_shard_publish(C.byref(full), th_all.data_ptr(), et_all.data_ptr(),
                                                             0, self._stream()), "shard_publish (all rows)")
                torch.cuda.synchronize(dev)
            self.theta = torch.from_numpy(self._pad(theta[:, self.ulo:self.uhi], self.ldk)).to(dev)
            self.eta = torch.from_numpy(self._pad(eta[:, self.ilo:self.ihi], self.ldl)).to(dev)
            self.pr = torch.from_numpy(np.ascontiguousarray(pr, dtype=np.float64)).to(dev)
            self._alt = (torch.empty_like(self.theta), torch.empty_like(self.eta), torch.empty_like(self.pr))
            self._half = 0
            self._lib.check(self.lib.mmsbm_shard_publish(C.byref(self._shard_c), self.theta.data_ptr(),
                                                         self.eta.data_ptr(), self._half, self._stream()),
                            "shard_publish")

    def run(self, iterations, prof=False):
        """``iterations`` EM steps of all S runs (the loop lives in the library).  With ``prof``
        returns the mean device ms per iteration of [whole iteration, wait for n_pr at its end,
        P tables + W, pass 1, n + publish + pr partial, pass 2, n + publish, 0] (measuring
        synchronises every iteration)."""
        iterations = int(iterations)
        if iterations <= 0:
            return None
        a, b = (self.theta, self.eta, self.pr), self._alt
        out = (C.c_float * 8)() if prof else None
        self._lib.check(self.lib.mmsbm_em_run_sharded(
            C.byref(self._shard_c), iterations, a[0].data_ptr(), a[1].data_ptr(), a[2].data_ptr(),
            b[0].data_ptr(), b[1].data_ptr(), b[2].data_ptr(), self._half, self._ws.data_ptr(),
            self._ws_bytes, self._stream(), C.addressof(out) if prof else None), "em_run_sharded")
        if iterations % 2:
            (self.theta, self.eta, self.pr), self._alt = self._alt, (self.theta, self.eta, self.pr)
            self._half ^= 1
        return [float(x) for x in out] if prof else None

    # ---------------------------------------------------------------- results
    def _full(self, own, lo, hi, n_all):
        """[S][n_all][ld] on every rank from the own rows: zeros elsewhere + one sum over the
        ranks (x + 0 is exact, so this is a gather)."""
        full = torch.zeros((self.S, n_all, own.shape[-1]), dtype=torch.float64, device=self.device)
        full[:, lo:hi] = own
        if self.world > 1:
            dist.all_reduce(full, op=dist.ReduceOp.SUM)
        return full

    def likelihood_device(self):
        """Per-run likelihood: each rank sums over the ratings of its own users, then the ranks'
        partial sums are added."""
        _lib = self._lib
        eta_full = self._full(self.eta, self.ilo, self.ihi, self.I)
        out = torch.empty(max(self.S, 1), dtype=torch.float64, device=self.device)
        dims = (self.Nu, self.Uo, self.I, self.R, self.K, self.L, self.S)
        need = C.c_size_t(0)
        _lib.check(self.lib.mmsbm_likelihood_min_workspace_bytes(*dims, C.byref(need)),
                   "likelihood_min_workspace_bytes")
        ws = torch.empty(max(need.value, 16), dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.mmsbm_likelihood(
            self.useg.data_ptr(), self.uadj.data_ptr(), self.usched.data_ptr(), *dims,
            self.theta.data_ptr(), eta_full.data_ptr(), self.pr.data_ptr(), out.data_ptr(),
            ws.data_ptr(), need.value, self._stream()), "likelihood")
        lik = out[:self.S].clone()
        if self.world > 1:
            dist.all_reduce(lik, op=dist.ReduceOp.SUM)
        return lik

    def likelihood(self):
        return self.likelihood_device().cpu().numpy()

    def get_params(self):
        """Full (theta [S,U,K], eta [S,I,L], pr [S,K,L,R]) as numpy on every rank."""
        th = self._full(self.theta, self.ulo, self.uhi, self.U).cpu().numpy()[..., :self.K]
        et = self._full(self.eta, self.ilo, self.ihi, self.I).cpu().numpy()[..., :self.L]
        return np.ascontiguousarray(th), np.ascontiguousarray(et), self.pr.cpu().numpy()

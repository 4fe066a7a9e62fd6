#!/usr/bin/env python
"""Benchmark of the EM hot path on synthetic ratings of a BASELINE.json shape.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload ml20m|ml1m|ml100k|netflix] [--iters-per-step T]

metric = rating-updates/s = ratings x EM iterations x runs / time.
A *step* is one fit-sized pass of the hot path: T EM iterations (default 400, the
reference's default ``iterations``; BASELINE.json's ML-20M config names none) of all S
runs (``sampling``) over the same synthetic ratings.

  value     K steps on data and parameters already resident in HBM (Engine.run ->
            mmsbm_em_run), CUDA events on the launching stream, max over ranks.
  e2e       the same T iterations through the host-pointer C ABI (mmsbm_host_fit, what
            the reference-side ctypes stub calls): H2D of the int64 [N,3] rows and of
            theta0/eta0/pr0 from pinned memory, index build, EM loop, likelihood, D2H of
            the fitted parameters -- all inside the timed region.
  roofline  dominant kernel = segment_pass_kernel (by-user + by-item launch of one
            iteration): algorithmic bytes B_alg*N*S / their CUDA-event
            time (mmsbm_em_step_profiled), against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the oracle port (numpy restatement of the reference) on the host, one
            thread, on a bounded row sample of the same shape; cpu_baseline_numba = the numba
            port (oracle/mmsbm_oracle_numba.py, the backend the reference picks by itself on a
            CPU host), all host threads, same sample.

N > 1 (torchrun): independent runs shard over ranks with no data-path collective; every
rank processes S runs of its own seeds over a replica of the ratings (weak scaling).
``--impl reference`` times the reference's CPU algorithm (oracle port; the reference is
pure Python and cannot travel) with one process per run, as its spawn pool does; the numba
port by default (``--cpu-backend``), host cores split evenly between the processes.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (U, I, N, K, L, S)   R = 5 everywhere (BASELINE.json configs)
    "ml100k": (943, 1682, 100_000, 10, 10, 1),
    "ml1m": (6040, 3706, 1_000_000, 10, 10, 8),
    "ml20m": (138_000, 27_000, 20_000_000, 20, 20, 8),
    "netflix": (480_000, 17_700, 100_000_000, 32, 32, 1),
}
R = 5
FALLBACK_HBM_GBS = 6650.0   # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def _zipf_ids(g, n_ids, size):
    """Heavy-tailed ids: P(id = k) ~ 1/(k+1) (Zipf exponent 1), by inverse-CDF sampling."""
    cdf = np.cumsum(1.0 / np.arange(1, n_ids + 1))
    cdf /= cdf[-1]
    return np.searchsorted(cdf, g.random(size), side="left").astype(np.int64)


def synth_triples(U, I, N, seed=0, ids="uniform"):
    """Uniform ids like the reference's benchmark_mmsbm.py:23-31, or Zipf-like heavy-tailed ids
    (SURVEY.md 8d); every id forced to appear."""
    g = np.random.default_rng(seed)
    data = np.empty((N, 3), dtype=np.int64)
    if ids == "zipf":
        data[:, 0] = _zipf_ids(g, U, N)
        data[:, 1] = _zipf_ids(g, I, N)
    else:
        data[:, 0] = g.integers(0, U, N)
        data[:, 1] = g.integers(0, I, N)
    data[:, 2] = g.integers(0, R, N)
    data[:U, 0] = np.arange(U)
    data[:I, 1] = np.arange(I)
    data[:R, 2] = np.arange(R)
    return data


def seeded_inits(data, U, I, K, L, seeds):
    """theta0/eta0/pr0 exactly as src/mmsbm.py:224-233 for each SeedSequence child."""
    du = np.maximum(np.bincount(data[:, 0], minlength=U), 1)[:, None]
    di = np.maximum(np.bincount(data[:, 1], minlength=I), 1)[:, None]
    th, et, pr = [], [], []
    for s in seeds:
        g = np.random.default_rng(s)
        th.append(g.random((U, K)) / du)
        et.append(g.random((I, L)) / di)
        p = g.random((K, L, R))
        pr.append(p / p.sum(axis=2, keepdims=True))
    return np.stack(th), np.stack(et), np.stack(pr)


def b_alg(U, I, N, K, L, S):
    """Algorithmic bytes per rating-update, SURVEY.md section 8(d)."""
    return 8.0 * (K + L) + 16.0 / S + 16.0 * (K * U + L * I) / N


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU arms
def _cpu_sample(U, I, K, L, n_rows, seed=0):
    data = synth_triples(U, I, max(n_rows, max(U, I)), seed=seed)
    return data


def cpu_baseline_port(U, I, K, L, n_rows, repeats=3):
    """Oracle port, one thread, one EM iteration of one run on an n_rows sample."""
    from oracle import mmsbm_oracle as orc
    data = _cpu_sample(U, I, K, L, n_rows)
    fu, fi = orc.degree_factors(data, K, L)
    th, et, pr = (a[0] for a in seeded_inits(data, U, I, K, L, [1]))
    orc.em_iteration(data[:2000], th, et, pr, fu, fi)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        orc.em_iteration(data, th, et, pr, fu, fi, chunk=50_000)
        best = min(best, time.perf_counter() - t0)
    return {"value": data.shape[0] / best, "unit": "rating-updates/s", "cores": 1, "kind": "port",
            "sample": f"{data.shape[0]} rows of the same U/I/K/L, 1 run, 1 EM iteration incl. the three "
                      f"normalisations, best of {repeats}; oracle/mmsbm_oracle.py (numpy restatement of "
                      f"kernels_numpy.update_coefficients)"}


def cpu_baseline_numba(U, I, K, L, n_rows, repeats=3):
    """Numba port (the backend the reference's load_backend("auto") picks on a host without
    CuPy), all host threads in its parallel omega phase, one EM iteration of one run."""
    try:
        import numba
        from oracle import mmsbm_oracle as orc, mmsbm_oracle_numba as onb
    except ImportError as e:        # numba missing on this host: say so instead of guessing
        return {"unavailable": str(e)}
    data = _cpu_sample(U, I, K, L, n_rows)
    fu, fi = orc.degree_factors(data, K, L)
    th, et, pr = (a[0] for a in seeded_inits(data, U, I, K, L, [1]))
    onb.em_iteration(data[:2000], th, et, pr, fu, fi)             # JIT compile outside the timing
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        onb.em_iteration(data, th, et, pr, fu, fi)
        best = min(best, time.perf_counter() - t0)
    return {"value": data.shape[0] / best, "unit": "rating-updates/s", "cores": int(numba.get_num_threads()),
            "kind": "port",
            "sample": f"{data.shape[0]} rows of the same U/I/K/L, 1 run, 1 EM iteration incl. the three "
                      f"normalisations, best of {repeats}, JIT excluded; oracle/mmsbm_oracle_numba.py "
                      f"(restatement of kernels_numba.update_coefficients: parallel omega phase, serial scatter)"}


_W = {}


def _ref_worker_init(U, I, K, L, n_rows, backend, threads):
    from oracle import mmsbm_oracle as orc
    data = _cpu_sample(U, I, K, L, n_rows)
    _W["orc"], _W["data"] = orc, data
    _W["f"] = orc.degree_factors(data, K, L)
    _W["shape"] = (U, I, K, L)
    _W["step"] = lambda th, et, pr: orc.em_iteration(data, th, et, pr, *_W["f"], chunk=50_000)
    if backend == "numba":
        import numba
        from oracle import mmsbm_oracle_numba as onb
        numba.set_num_threads(max(1, min(threads, numba.config.NUMBA_NUM_THREADS)))
        _W["step"] = lambda th, et, pr: onb.em_iteration(data, th, et, pr, *_W["f"])
        th, et, pr = (a[0] for a in seeded_inits(data, U, I, K, L, [0]))
        onb.em_iteration(data[:2000], th, et, pr, *_W["f"])       # JIT compile in the initializer


def _ref_worker_step(args):
    seed, iters = args
    orc, data = _W["orc"], _W["data"]
    U, I, K, L = _W["shape"]
    th, et, pr = (a[0] for a in seeded_inits(data, U, I, K, L, [seed]))
    for _ in range(iters):
        th, et, pr = _W["step"](th, et, pr)
    return float(th.sum())


def run_reference_arm(args, shape):
    """The reference's CPU algorithm (oracle port), one process per run like its spawn pool
    (src/mmsbm.py:182-185), on a bounded row sample."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    U, I, N, K, L, S = shape
    n_rows = max(args.cpu_rows, max(U, I))
    cores = os.cpu_count() or 1
    procs = min(S, cores)
    iters = 1
    backend = args.cpu_backend
    if backend == "auto":           # the reference's own order on a CPU host: numba, then numpy
        try:
            import numba  # noqa: F401
            backend = "numba"
        except ImportError:
            backend = "numpy"
    threads = max(1, cores // procs) if backend == "numba" else 1
    ctx = mp.get_context("spawn" if backend == "numba" else "fork")   # numba's thread pool does not survive fork
    with ctx.Pool(processes=procs, initializer=_ref_worker_init,
                  initargs=(U, I, K, L, n_rows, backend, threads)) as pool:
        for _ in range(args.warmup):
            pool.map(_ref_worker_step, [(s, iters) for s in range(S)])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_ref_worker_step, [(s, iters) for s in range(S)])
        dt = time.perf_counter() - t0
    value = n_rows * iters * S * args.steps / dt
    sample = (f"{n_rows} rows of the {args.workload} shape (same U/I/K/L/R), {S} runs x {iters} EM iteration per "
              f"step, one process per run ({procs} processes x {threads} thread(s), {cores} host cores), "
              f"{backend} port of the reference kernels, init included")
    line = {
        "impl": "reference", "metric": "rating-updates/sec", "value": value, "unit": "rating-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "users": U, "items": I, "ratings": N, "K": K, "L": L, "R": R,
                   "sampling": S, "cpu_sample_rows": n_rows},
        "cpu_baseline": {"value": value, "unit": "rating-updates/s", "cores": procs * threads, "kind": "port",
                         "backend": backend, "sample": sample},
        "e2e": {"value": value, "unit": "rating-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def run_b200_arm(args, shape):
    import torch
    import torch.distributed as dist
    from mmsbm_b200 import _lib
    from mmsbm_b200.engine import Engine

    U, I, N, K, L, S = shape
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    T = args.iters_per_step
    lib = _lib.load(require_device=True)

    t0 = time.perf_counter()
    data = synth_triples(U, I, N, seed=0, ids=args.ids)
    seeds = np.random.default_rng(1).bit_generator._seed_seq.spawn(S * world)[rank * S:(rank + 1) * S]
    th0, et0, pr0 = seeded_inits(data, U, I, K, L, seeds)
    gen_s = time.perf_counter() - t0

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng = Engine(data, U, I, R, K, L)
    torch.cuda.synchronize()
    build_ms = (time.perf_counter() - t0) * 1e3
    eng.set_params(th0, et0, pr0)

    if args.shard == "ratings" and world > 1:
        # ONE set of S runs, ratings split by user range over the ranks, n_eta / n_pr all-reduced
        # over NCCL every iteration (BASELINE.json configs[4]); strong scaling, no e2e / roofline
        from mmsbm_b200.parallel import RatingShardedEngine
        del eng
        seeds0 = np.random.default_rng(1).bit_generator._seed_seq.spawn(S)
        th0, et0, pr0 = seeded_inits(data, U, I, K, L, seeds0)
        sh = RatingShardedEngine(data, U, I, R, K, L)
        sh.set_params(th0, et0, pr0)
        for _ in range(args.warmup):
            sh.run(T)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        clocks = ClockSampler(local_rank)
        launches0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            sh.run(T)
        e1.record()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        clk = clocks.stop()
        lik = sh.likelihood()
        if rank == 0:
            print(json.dumps({
                "metric": "rating-updates/sec", "value": float(N) * T * S * args.steps / (ms * 1e-3),
                "unit": "rating-updates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": args.workload, "users": U, "items": I, "ratings": N, "K": K, "L": L,
                           "R": R, "sampling": S, "iterations_per_step": T, "ids": f"{args.ids}, seed 0",
                           "parallelism": f"ratings sharded by user range over {world} GPUs, NCCL all-reduce "
                                          "of n_eta and n_pr every iteration",
                           "local_ratings": int(sh.engine.N)},
                "roofline": None, "cpu_baseline": None, "e2e": None,
                "gpu_launches": int(_lib.launch_count() - launches0), "clocks": clk,
                "ms_per_iteration": ms / args.steps / T, "likelihood_run0": float(lik[0])}), flush=True)
        dist.destroy_process_group()
        return

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        eng.run(T)
    barrier()
    clocks = ClockSampler(local_rank)
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        eng.run(T)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - launches0
    clk = clocks.stop()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    updates_per_step = float(N) * T * S * world
    value = updates_per_step * args.steps / (ms * 1e-3)

    # ---- per-kernel device time of one iteration (CUDA events inside the library) ----
    ms4 = (ctypes.c_float * 7)()
    acc = np.zeros(7)
    reps = 5
    for _ in range(reps):
        b = eng._alt
        _lib.check(lib.mmsbm_em_step_profiled(
            *eng._graph_args(), eng.N, U, I, R, K, L, S, eng.theta.data_ptr(), eng.eta.data_ptr(),
            eng.pr.data_ptr(), b[0].data_ptr(), b[1].data_ptr(), b[2].data_ptr(), 0,
            eng._ws.data_ptr(), eng._ws_bytes, eng._stream(), ctypes.addressof(ms4)), "em_step_profiled")
        acc += np.array(list(ms4))
        eng.swap()
    k_ms = acc / reps
    peak, peak_src = hbm_peak()
    alg_bytes = b_alg(U, I, N, K, L, S) * N * S           # both launches of segment_pass_kernel
    seg_ms = float(k_ms[1] + k_ms[3])
    achieved = alg_bytes / (seg_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": None, "peak_source": peak_src, "kernel": "segment_pass_kernel (by-user + by-item launch)",
        "alg_bytes_per_update": b_alg(U, I, N, K, L, S),
        "kernel_ms": {"p_tables_w": float(k_ms[0]), "by_user": float(k_ms[1]), "n_users": float(k_ms[2]),
                      "by_item": float(k_ms[3]), "n_items": float(k_ms[4]), "pr_accumulate": float(k_ms[5]),
                      "pr_finalize": float(k_ms[6])},
        "share_of_iteration": seg_ms / float(k_ms.sum()),
    }
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):   # dram bytes per launch from the committed ncu --set full capture
        try:
            roofline["traffic"] = json.load(open(prof)).get(args.workload)
        except Exception:
            pass

    # ---- end to end through the host-pointer C ABI ----
    e2e = None
    if not args.no_e2e:
        def pinned(a):
            t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            return t, t.numpy()
        keep = [pinned(a) for a in (data, th0, et0, pr0)]
        h_data, h_th, h_et, h_pr = (k[1] for k in keep)
        outs = [torch.empty(a.shape, dtype=torch.float64).pin_memory() for a in (th0, et0, pr0)]
        lik = torch.empty(S, dtype=torch.float64).pin_memory()

        def fit_once():
            _lib.check(lib.mmsbm_host_fit(
                h_data.ctypes.data, N, U, I, R, K, L, S, T, h_th.ctypes.data, h_et.ctypes.data, h_pr.ctypes.data,
                outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), lik.data_ptr()), "host_fit")
        fit_once()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fit_once()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        h2d = int(h_data.nbytes + h_th.nbytes + h_et.nbytes + h_pr.nbytes)
        d2h = int(sum(o.numel() * 8 for o in outs) + S * 8)
        e2e = {"value": updates_per_step * args.steps / dt, "unit": "rating-updates/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": dt / args.steps * 1e3,
               "api": "mmsbm_host_fit (C ABI, host pointers): H2D rows+theta0/eta0/pr0, index build, "
                      f"{T} EM iterations, likelihood, D2H theta/eta/pr/likelihood",
               "likelihood_run0": float(lik[0])}

    cpu = cpu_nb = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline_port(U, I, K, L, args.cpu_rows)
        cpu_nb = cpu_baseline_numba(U, I, K, L, args.cpu_rows)

    if rank == 0:
        line = {
            "metric": "rating-updates/sec", "value": value, "unit": "rating-updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": args.workload, "users": U, "items": I, "ratings": N, "K": K, "L": L, "R": R,
                       "sampling_per_gpu": S, "iterations_per_step": T, "ids": f"{args.ids}, seed 0",
                       "init": "reference seeded init, model seed 1",
                       "parallelism": f"runs sharded over {world} GPU(s), no data-path collective",
                       "l2": "no explicit flush: one iteration touches the parameters of all runs and both "
                             "index arrays (> 126 MB L2 at ml20m); see DESIGN.md"},
            "roofline": roofline, "cpu_baseline": cpu, "cpu_baseline_numba": cpu_nb, "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clk, "index_build_ms": build_ms, "host_datagen_s": gen_s,
            "ms_per_iteration": ms / args.steps / T,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ml20m", choices=sorted(WORKLOADS))
    ap.add_argument("--iters-per-step", type=int, default=400)
    ap.add_argument("--cpu-rows", type=int, default=200_000)
    ap.add_argument("--cpu-backend", default="auto", choices=["auto", "numba", "numpy"],
                    help="--impl reference: which port of the reference kernels to time")
    ap.add_argument("--ids", default="uniform", choices=["uniform", "zipf"])
    ap.add_argument("--shard", default="runs", choices=["runs", "ratings"],
                    help="N > 1 only: shard independent runs (default, weak scaling) or the ratings of "
                         "every run by user range with a per-iteration all-reduce (strong scaling)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    shape = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, shape)
    else:
        run_b200_arm(args, shape)


if __name__ == "__main__":
    main()

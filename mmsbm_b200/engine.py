"""Device-resident EM engine: the host-side owner of the buffers the sm_100a kernels
work on.  PyTorch is used for plumbing only (device memory, streams, and -- in
parallel.py -- torch.distributed); every computation is a call into
libmmsbm_b200.so through the device-pointer C ABI (include/mmsbm_b200.h).

One ``Engine`` = one encoded training set on one GPU + the parameters of S runs
(``sampling``) in the padded device layout theta [S][U][ldk], eta [S][I][ldl],
pr [S][K][L][R].  It stands where the reference keeps ``train`` plus the per-run
numpy arrays of ``run_one_sampling`` (src/mmsbm.py:187-269).
"""
import numpy as np
import torch

from . import _lib



class Engine:
    def __init__(self, data, n_users, n_items, n_levels, K, L, device=None):
        """``data``: int [N,3] encoded (user, item, rating) rows (numpy, host)."""
        self.lib = _lib.load(require_device=True)
        if not torch.cuda.is_available():
            raise ImportError("mmsbm_b200: torch sees no CUDA device; there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None \
            else torch.device(device)
        data = np.asarray(data)
        if data.ndim != 2 or data.shape[1] != 3:
            raise ValueError("data must have shape [N,3]")
        self.N = int(data.shape[0])
        self.U, self.I, self.R, self.K, self.L = int(n_users), int(n_items), int(n_levels), int(K), int(L)
        self.ldk, self.ldl = self.lib.mmsbm_row_stride(self.K), self.lib.mmsbm_row_stride(self.L)
        self.S = 0
        self.theta = self.eta = self.pr = None
        self._alt = None
        self._ws = None
        with torch.cuda.device(self.device):
            # the int64 [N,3] rows go up as they are; the int32 split and the id range check
            # run on the device (mmsbm_split_triples)
            raw = torch.from_numpy(np.ascontiguousarray(data, dtype=np.int64)).to(self.device)
            self.cols = torch.empty((3, max(self.N, 1)), dtype=torch.int32, device=self.device)
            bad = torch.zeros(1, dtype=torch.int32, device=self.device)
            _lib.check(self.lib.mmsbm_split_triples(
                raw.data_ptr(), self.N, self.U, self.I, self.R, self.cols[0].data_ptr(),
                self.cols[1].data_ptr(), self.cols[2].data_ptr(), bad.data_ptr(), self._stream()),
                "split_triples")
            if int(bad.item()):
                raise ValueError("data holds an id outside [0,U) x [0,I) x [0,R)")
            del raw
            self._build_graph()

    @classmethod
    def for_prediction(cls, n_users, n_items, n_levels, K, L, device=None):
        """An engine without a training set: holds parameters (``set_params``) and serves
        ``prod_dist_device`` / ``mean_over_runs`` -- what a model restored by ``MMSBM.load``
        needs for ``predict``.  ``run`` / ``likelihood`` need the ratings and refuse."""
        self = cls.__new__(cls)
        self.lib = _lib.load(require_device=True)
        if not torch.cuda.is_available():
            raise ImportError("mmsbm_b200: torch sees no CUDA device; there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None \
            else torch.device(device)
        self.N = 0
        self.U, self.I, self.R, self.K, self.L = int(n_users), int(n_items), int(n_levels), int(K), int(L)
        self.ldk, self.ldl = self.lib.mmsbm_row_stride(self.K), self.lib.mmsbm_row_stride(self.L)
        self.S = 0
        self.theta = self.eta = self.pr = None
        self._alt = None
        self._ws = None
        self.cols = None
        self.useg = None                          # no index structure
        return self

    def _need_ratings(self):
        if getattr(self, "useg", None) is None:
            raise RuntimeError("this engine holds parameters only (restored model): "
                               "fit again to iterate or to compute a likelihood")

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _empty(self, n, dtype):
        return torch.empty(max(int(n), 1), dtype=dtype, device=self.device)

    def _bytes(self, n):
        return torch.empty(max(int(n), 16), dtype=torch.uint8, device=self.device)

    def _build_graph(self):
        i32 = torch.int32
        N, U, I, R = self.N, self.U, self.I, self.R
        self.useg, self.uadj, self.uperm, self.udeg = (
            self._empty(U * R + 1, i32), self._empty(N, i32), self._empty(N, i32), self._empty(U, i32))
        self.iseg, self.iadj, self.iperm, self.ideg = (
            self._empty(I * R + 1, i32), self._empty(N, i32), self._empty(N, i32), self._empty(I, i32))
        n_u, n_i = _lib.C.c_int64(0), _lib.C.c_int64(0)
        _lib.check(self.lib.mmsbm_sched_elems(N, U, _lib.C.byref(n_u)), "sched_elems")
        _lib.check(self.lib.mmsbm_sched_elems(N, I, _lib.C.byref(n_i)), "sched_elems")
        self.usched, self.isched = self._empty(n_u.value, i32), self._empty(n_i.value, i32)
        need = _lib.C.c_size_t(0)
        _lib.check(self.lib.mmsbm_graph_workspace_bytes(N, U, I, R, _lib.C.byref(need)), "graph_workspace_bytes")
        ws = self._bytes(need.value)
        _lib.check(self.lib.mmsbm_graph_build(
            self.cols[0].data_ptr(), self.cols[1].data_ptr(), self.cols[2].data_ptr(), N, U, I, R,
            self.useg.data_ptr(), self.uadj.data_ptr(), self.uperm.data_ptr(), self.udeg.data_ptr(),
            self.iseg.data_ptr(), self.iadj.data_ptr(), self.iperm.data_ptr(), self.ideg.data_ptr(),
            self.usched.data_ptr(), self.isched.data_ptr(),
            ws.data_ptr(), need.value, self._stream()), "graph_build")
        torch.cuda.current_stream(self.device).synchronize()   # ws is released on return
        del ws
        self.cols = None            # the int32 columns are only an input of the build

    def _graph_args(self):
        return (self.useg.data_ptr(), self.uadj.data_ptr(), self.udeg.data_ptr(),
                self.iseg.data_ptr(), self.iadj.data_ptr(), self.ideg.data_ptr(),
                self.usched.data_ptr(), self.isched.data_ptr())

    # ---------------------------------------------------------------- parameters
    def _pad(self, x, ld):
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.shape[-1] == ld:
            return x
        out = np.zeros(x.shape[:-1] + (ld,), dtype=np.float64)
        out[..., :x.shape[-1]] = x
        return out

    def set_params(self, theta, eta, pr):
        """theta [S,U,K], eta [S,I,L], pr [S,K,L,R] (numpy, host layout of the reference)."""
        theta, eta, pr = np.asarray(theta), np.asarray(eta), np.asarray(pr)
        if theta.ndim == 2:
            theta, eta, pr = theta[None], eta[None], pr[None]
        S = theta.shape[0]
        if theta.shape != (S, self.U, self.K) or eta.shape != (S, self.I, self.L) \
                or pr.shape != (S, self.K, self.L, self.R):
            raise ValueError("parameter shapes do not match the engine")
        dev = self.device
        self.S = S
        self.theta = torch.from_numpy(self._pad(theta, self.ldk)).to(dev)
        self.eta = torch.from_numpy(self._pad(eta, self.ldl)).to(dev)
        self.pr = torch.from_numpy(np.ascontiguousarray(pr, dtype=np.float64)).to(dev)
        self._alt = (torch.empty_like(self.theta), torch.empty_like(self.eta), torch.empty_like(self.pr))
        if getattr(self, "useg", None) is None:   # prediction-only engine: no EM workspace
            self._ws, self._ws_bytes = None, 0
            return
        need = _lib.C.c_size_t(0)
        _lib.check(self.lib.mmsbm_em_workspace_bytes(self.N, self.U, self.I, self.R, self.K, self.L, S,
                                                     _lib.C.byref(need)), "em_workspace_bytes")
        self._ws = self._bytes(need.value)
        self._ws_bytes = need.value

    def get_params(self):
        """numpy (theta [S,U,K], eta [S,I,L], pr [S,K,L,R])."""
        th = self.theta.cpu().numpy()[..., :self.K]
        et = self.eta.cpu().numpy()[..., :self.L]
        return np.ascontiguousarray(th), np.ascontiguousarray(et), self.pr.cpu().numpy()

    # ------------------------------------------------------------------- EM loop
    def run(self, iterations):
        """``iterations`` EM steps for all S runs, asynchronous on the current stream."""
        if iterations <= 0:
            return
        self._need_ratings()
        a = (self.theta, self.eta, self.pr)
        b = self._alt
        _lib.check(self.lib.mmsbm_em_run(
            *self._graph_args(), self.N, self.U, self.I, self.R, self.K, self.L, self.S, int(iterations),
            a[0].data_ptr(), a[1].data_ptr(), a[2].data_ptr(),
            b[0].data_ptr(), b[1].data_ptr(), b[2].data_ptr(),
            self._ws.data_ptr(), self._ws_bytes, self._stream()), "em_run")
        if iterations % 2:
            self.swap()

    def step_raw(self, flags):
        """One step into the alternate buffers with ``flags`` (RAW_THETA / RAW_ETA_PR);
        returns the three output tensors without swapping."""
        self._need_ratings()
        b = self._alt
        _lib.check(self.lib.mmsbm_em_step(
            *self._graph_args(), self.N, self.U, self.I, self.R, self.K, self.L, self.S,
            self.theta.data_ptr(), self.eta.data_ptr(), self.pr.data_ptr(),
            b[0].data_ptr(), b[1].data_ptr(), b[2].data_ptr(), int(flags),
            self._ws.data_ptr(), self._ws_bytes, self._stream()), "em_step")
        return b

    def finalize(self, eta, pr, ideg=None):
        """Post-all-reduce epilogue, in place."""
        ideg = self.ideg if ideg is None else ideg
        _lib.check(self.lib.mmsbm_em_finalize(eta.data_ptr(), ideg.data_ptr(), self.I, self.L,
                                              pr.data_ptr(), self.K, self.R, self.S, self._stream()),
                   "em_finalize")

    def swap(self):
        (self.theta, self.eta, self.pr), self._alt = self._alt, (self.theta, self.eta, self.pr)

    # ---------------------------------------------------------------- reductions
    def likelihood_device(self):
        self._need_ratings()
        out = self._empty(self.S, torch.float64)
        dims = (self.N, self.U, self.I, self.R, self.K, self.L, self.S)
        need = _lib.C.c_size_t(0)
        _lib.check(self.lib.mmsbm_likelihood_min_workspace_bytes(*dims, _lib.C.byref(need)),
                   "likelihood_min_workspace_bytes")
        if self._ws is not None and self._ws_bytes >= need.value:
            # the EM workspace is idle between iterations (same stream): the likelihood batches its
            # runs to whatever it is given
            ws, ws_bytes = self._ws, self._ws_bytes
        else:
            _lib.check(self.lib.mmsbm_likelihood_workspace_bytes(*dims, _lib.C.byref(need)),
                       "likelihood_workspace_bytes")
            ws, ws_bytes = self._bytes(need.value), need.value
        _lib.check(self.lib.mmsbm_likelihood(
            self.useg.data_ptr(), self.uadj.data_ptr(), self.usched.data_ptr(), self.N, self.U, self.I,
            self.R, self.K, self.L, self.S, self.theta.data_ptr(), self.eta.data_ptr(), self.pr.data_ptr(),
            out.data_ptr(), ws.data_ptr(), ws_bytes, self._stream()), "likelihood")
        ws.record_stream(torch.cuda.current_stream(self.device))
        return out[:self.S]

    def likelihood(self):
        return self.likelihood_device().cpu().numpy()

    def prod_dist_device(self, test):
        """rat [S,M,R] on the device for int [M,3] (or [M,2]) test rows."""
        test = np.asarray(test)
        M = int(test.shape[0])
        if M and (test[:, 0].min() < 0 or test[:, 0].max() >= self.U
                  or test[:, 1].min() < 0 or test[:, 1].max() >= self.I):
            raise ValueError("test rows hold an id unseen in training")
        tu = torch.from_numpy(np.ascontiguousarray(test[:, 0], dtype=np.int32)).to(self.device)
        ti = torch.from_numpy(np.ascontiguousarray(test[:, 1], dtype=np.int32)).to(self.device)
        rat = torch.empty((self.S, M, self.R), dtype=torch.float64, device=self.device)
        _lib.check(self.lib.mmsbm_prod_dist(
            tu.data_ptr(), ti.data_ptr(), M, self.U, self.I, self.R, self.K, self.L, self.S,
            self.theta.data_ptr(), self.eta.data_ptr(), self.pr.data_ptr(), rat.data_ptr(),
            self._stream()), "prod_dist")
        for t in (tu, ti):
            t.record_stream(torch.cuda.current_stream(self.device))
        return rat

    def mean_over_runs(self, rat):
        S, M, R = rat.shape
        out = torch.empty((M, R), dtype=torch.float64, device=self.device)
        _lib.check(self.lib.mmsbm_mean_over_runs(rat.data_ptr(), M * R, S, out.data_ptr(), self._stream()),
                   "mean_over_runs")
        return out


def predict_stats(rat, real, device=None, want_pred=False):
    """Prediction statistics of src/mmsbm.py:488-539 on the GPU.

    ``rat``: torch cuda tensor or numpy, [S,M,R] or [M,R]; ``real``: encoded ratings [M].
    Returns a list of S dicts (numpy scalars as in the reference) and, optionally, the
    argmax predictions [S,M] (numpy int32)."""
    lib = _lib.load(require_device=True)
    if not isinstance(rat, torch.Tensor):
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        rat = torch.from_numpy(np.ascontiguousarray(rat, dtype=np.float64)).to(dev)
    if rat.dim() == 2:
        rat = rat[None]
    rat = rat.contiguous()
    dev = rat.device
    S, M, R = rat.shape
    real_d = torch.from_numpy(np.ascontiguousarray(real, dtype=np.int32)).to(dev)
    counts = torch.empty((S, 5), dtype=torch.int64, device=dev)
    s2p = torch.empty((S,), dtype=torch.float64, device=dev)
    pred = torch.empty((S, max(M, 1)), dtype=torch.int32, device=dev) if want_pred else None
    need = _lib.C.c_size_t(0)
    _lib.check(lib.mmsbm_stats_workspace_bytes(M, S, _lib.C.byref(need)), "stats_workspace_bytes")
    ws = torch.empty(max(need.value, 16), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(lib.mmsbm_predict_stats(rat.data_ptr(), real_d.data_ptr(), M, R, S, counts.data_ptr(),
                                       s2p.data_ptr(), pred.data_ptr() if want_pred else None,
                                       ws.data_ptr(), need.value, stream), "predict_stats")
    c = counts.cpu().numpy()
    s = s2p.cpu().numpy()
    out = []
    for k in range(S):
        n = c[k, 0]
        with np.errstate(divide="ignore", invalid="ignore"):
            out.append({
                "accuracy": np.float64(c[k, 1]) / n,
                "one_off_accuracy": np.float64(c[k, 2]) / n,
                "mae": 1 - np.float64(c[k, 4]) / n,
                "s2": np.int64(c[k, 3]),
                "s2pond": np.float64(s[k]),
            })
    if want_pred:
        return out, pred.cpu().numpy()[:, :M]
    return out

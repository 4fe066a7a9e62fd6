"""CPU oracle, numba flavour -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

Restates the algorithm of the reference's numba backend (src/kernels_numba.py:19-84), the
backend its ``load_backend("auto")`` picks on a host without CuPy (src/backend.py:16-22), so
that ``bench.py`` can time "the reference's numba path" on the GPU box's host cores next to the
numpy port (oracle/mmsbm_oracle.py).  Only ``tests/`` and the CPU-baseline / ``--impl reference``
legs of ``bench.py`` import it.

What the reference's numba backend does, and this file does likewise:
  * phase 1 (parallel over ratings, src/kernels_numba.py:19-37): materialise
    omega[n,k,l] = theta[u,k] * eta[i,l] * pr[k,l,r];
  * phase 2 (serial, :50-84): S_n = sum_kl omega, inc = omega / (S_n + eps) -- eps is ADDED here,
    the numpy backend clamps with max() instead (src/kernels_numpy.py:51-52) -- then one serial
    sweep over the ratings scatters row sums into n_theta, column sums into n_eta and the whole
    [K,L] slab into n_pr[:, :, r];
  * fastmath is on in both phases, so sums may be reassociated: agreement with the numpy oracle
    is to ~1e-12 relative, not bitwise.

Parity status: PINNED through tests/test_oracle_golden.py (omega against the reference numba
kernel's output on the toy of tests/test_backends.py, stored in tests/golden/toy_backends.npz;
the M-step sums against the numpy oracle, itself bit-pinned to the reference).
"""
import numpy as np
from numba import njit, prange

EPS = float(np.finfo(np.float64).eps)


@njit(parallel=True, fastmath=True, cache=False)
def omegas(data, theta, eta, pr):
    """[N,K,L] unnormalised responsibilities (src/kernels_numba.py:19-37)."""
    n_rows, K, L = data.shape[0], theta.shape[1], eta.shape[1]
    out = np.empty((n_rows, K, L))
    for n in prange(n_rows):
        u, i, r = data[n, 0], data[n, 1], data[n, 2]
        for k in range(K):
            t = theta[u, k]
            for l in range(L):
                out[n, k, l] = t * eta[i, l] * pr[k, l, r]
    return out


@njit(fastmath=True, cache=False)
def _scatter(data, om, n_theta, n_eta, n_pr):
    """Serial normalise-and-scatter sweep (src/kernels_numba.py:50-84)."""
    n_rows, K, L = om.shape
    for n in range(n_rows):
        u, i, r = data[n, 0], data[n, 1], data[n, 2]
        s = 0.0
        for k in range(K):
            for l in range(L):
                s += om[n, k, l]
        inv = 1.0 / (s + EPS)
        for k in range(K):
            row = 0.0
            for l in range(L):
                inc = om[n, k, l] * inv
                row += inc
                n_eta[i, l] += inc
                n_pr[k, l, r] += inc
            n_theta[u, k] += row


def em_sums(data, theta, eta, pr):
    """(n_theta[U,K], n_eta[I,L], n_pr[K,L,R]), unnormalised, like
    kernels_numba.update_coefficients."""
    data = np.ascontiguousarray(data, dtype=np.int64)
    om = omegas(data, theta, eta, pr)
    n_theta, n_eta, n_pr = np.zeros_like(theta), np.zeros_like(eta), np.zeros_like(pr)
    _scatter(data, om, n_theta, n_eta, n_pr)
    return n_theta, n_eta, n_pr


def em_iteration(data, theta, eta, pr, fu, fi, chunk=None):
    """One EM iteration: sums + the three normalisations of
    src/expectation_maximization.py:118-155 (shared with the numpy oracle).  ``chunk`` bounds the
    [N,K,L] temporary; the reference materialises it whole."""
    from . import mmsbm_oracle as orc
    if chunk is None or chunk >= data.shape[0]:
        nt, ne, npr = em_sums(data, theta, eta, pr)
    else:
        nt, ne, npr = np.zeros_like(theta), np.zeros_like(eta), np.zeros_like(pr)
        for lo in range(0, data.shape[0], chunk):
            a, b, c = em_sums(data[lo:lo + chunk], theta, eta, pr)
            nt += a; ne += b; npr += c
    return orc.scale_by_degree(nt, fu), orc.scale_by_degree(ne, fi), orc.normalize_pr(npr)

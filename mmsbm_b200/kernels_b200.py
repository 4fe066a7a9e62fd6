"""B200 kernel plugin with the reference's backend contract (src/backend.py:16-22).

Three module-level callables, all ``(data, theta, eta, pr)``, numpy float64 in and
out, ``data`` an int64 ``[N,3]`` array of encoded ``(user, item, rating)`` rows:

    compute_omegas      -> ndarray [N,K,L]          (src/kernels_numpy.py:21-36)
    update_coefficients -> (n_theta, n_eta, n_pr)   (src/kernels_numpy.py:43-79), unnormalised
    prod_dist           -> ndarray [M,R]            (src/kernels_numpy.py:86-97)

Each call goes through the host-pointer C ABI of libmmsbm_b200.so, which copies to
the GPU, runs the sm_100a kernels and copies back.  Importing this module raises
``ImportError`` when the library or a CUDA device is missing -- the reference's own
signal for an unavailable backend; nothing here computes on the CPU.
"""
import numpy as np

from . import _lib

__all__ = ["compute_omegas", "update_coefficients", "prod_dist", "likelihood"]

_L = _lib.load(require_device=True)


def _prep(data, theta, eta, pr):
    data = np.ascontiguousarray(data, dtype=np.int64)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    eta = np.ascontiguousarray(eta, dtype=np.float64)
    pr = np.ascontiguousarray(pr, dtype=np.float64)
    if data.ndim != 2 or data.shape[1] != 3:
        raise ValueError("data must have shape [N,3]")
    if theta.ndim != 2 or eta.ndim != 2 or pr.ndim != 3 or pr.shape[:2] != (theta.shape[1], eta.shape[1]):
        raise ValueError("theta [U,K], eta [I,L], pr [K,L,R] expected")
    return data, theta, eta, pr


def _p(a):
    return a.ctypes.data


def _args(data, theta, eta, pr):
    return (_p(data), data.shape[0], _p(theta), theta.shape[0], theta.shape[1],
            _p(eta), eta.shape[0], eta.shape[1], _p(pr), pr.shape[2])


def compute_omegas(data, theta, eta, pr):
    data, theta, eta, pr = _prep(data, theta, eta, pr)
    out = np.empty((data.shape[0], theta.shape[1], eta.shape[1]))
    _lib.check(_L.mmsbm_host_compute_omegas(*_args(data, theta, eta, pr), _p(out)), "compute_omegas")
    return out


def update_coefficients(data, theta, eta, pr):
    data, theta, eta, pr = _prep(data, theta, eta, pr)
    n_theta, n_eta, n_pr = np.empty_like(theta), np.empty_like(eta), np.empty_like(pr)
    _lib.check(_L.mmsbm_host_update_coefficients(*_args(data, theta, eta, pr),
                                                 _p(n_theta), _p(n_eta), _p(n_pr)),
               "update_coefficients")
    return n_theta, n_eta, n_pr


def prod_dist(data, theta, eta, pr):
    data, theta, eta, pr = _prep(data, theta, eta, pr)
    out = np.empty((data.shape[0], pr.shape[2]))
    _lib.check(_L.mmsbm_host_prod_dist(*_args(data, theta, eta, pr), _p(out)), "prod_dist")
    return out


def likelihood(data, theta, eta, pr):
    """The reference's compute_likelihood (src/expectation_maximization.py:157-167)."""
    data, theta, eta, pr = _prep(data, theta, eta, pr)
    out = np.empty(1)
    _lib.check(_L.mmsbm_host_likelihood(*_args(data, theta, eta, pr), _p(out)), "likelihood")
    return np.float64(out[0])

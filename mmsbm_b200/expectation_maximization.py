"""EM driver with the reference's contract (src/expectation_maximization.py:7-189).

Same constructor, same method names, numpy arrays at the surface.  The three kernel
calls dispatch to the B200 backend (backend.load_backend); ``compute_likelihood`` is an
on-device reduction; the two normalisations are the reference's one-line array
divisions -- inside the fitted loop (MMSBM.fit) they never run on the host: the CUDA
EM step applies them as its epilogue.
"""
import numpy as np

from .backend import load_backend


class ExpectationMaximization:
    def __init__(self, dims, user_indices, item_indices, rating_indices,
                 norm_factors, backend: str = "auto", debug: bool = False):
        self._dims = dims
        self._user_indices = user_indices
        self._item_indices = item_indices
        self._rating_indices = rating_indices
        self._normalization_factors = norm_factors
        self._debug = debug
        (self._compute_omegas,
         self._update_coeffs,
         self._prod_dist,
         self._backend) = load_backend(backend)
        if self._debug:
            print(f"Using {self._backend} backend")

    def compute_omegas(self, data, theta, eta, pr):
        """omega[n,k,l] = theta[u,k] eta[i,l] pr[k,l,r], shape [N,K,L]."""
        return self._compute_omegas(data, theta, eta, pr)

    def update_coefficients(self, data, theta, eta, pr):
        """Unnormalised (n_theta [U,K], n_eta [I,L], n_pr [K,L,R])."""
        return self._update_coeffs(data, theta, eta, pr)

    def normalize_with_d(self, df, type_):
        return df / self._normalization_factors[type_]

    @staticmethod
    def normalize_with_self(df):
        """pr[k,l,:] / sum_r pr[k,l,r]; a sum that is exactly zero divides by one."""
        tot = df.sum(axis=2)
        return df / np.where(tot == 0, 1, tot)[:, :, None]

    def compute_likelihood(self, data, theta, eta, pr):
        from . import kernels_b200
        return kernels_b200.likelihood(data, theta, eta, pr)

    @staticmethod
    def prod_dist(x, theta, eta, pr):
        """Rating distribution of one (user, item) row."""
        from . import kernels_b200
        row = np.array([[int(x[0]), int(x[1]), 0]], dtype=np.int64)
        return kernels_b200.prod_dist(row, theta, eta, pr)[0]

    def compute_prod_dist(self, data, theta, eta, pr):
        """rat[n,r] = sum_kl theta[u,k] eta[i,l] pr[k,l,r] for every row of ``data``."""
        return self._prod_dist(data, theta, eta, pr)

"""Drop-in plugin for the reference's loader: with this file's directory on
``sys.path``, ``backend.load_backend("b200")`` of eudald-seeslab/mmsbm
(src/backend.py:16-22) imports it and gets the three kernels.  See INTEGRATION.md."""
from mmsbm_b200.kernels_b200 import compute_omegas, update_coefficients, prod_dist  # noqa: F401

__all__ = ["compute_omegas", "update_coefficients", "prod_dist"]

"""Backend loader with the reference's signature (src/backend.py:4-28).

``load_backend(name)`` returns ``(compute_omegas, update_coefficients, prod_dist,
name)``.  The only backend is the CUDA one: ``"b200"`` (``"auto"`` resolves to it).
The reference's CPU backends (numpy / numba) and CuPy are what this package
replaces; asking for them, or running without the built library or a GPU, raises
``ImportError`` exactly as the reference does for an unavailable backend -- there
is no CPU fallback.
"""
from importlib import import_module

BACKENDS = ("b200",)


def load_backend(name: str = "auto"):
    order = list(BACKENDS) if name == "auto" else [name]
    last_error = None
    for backend in order:
        if backend not in BACKENDS:
            last_error = ModuleNotFoundError(
                f"backend '{backend}' is not part of mmsbm_b200 (available: {BACKENDS})")
            continue
        try:
            mod = import_module(f"mmsbm_b200.kernels_{backend}")
            return mod.compute_omegas, mod.update_coefficients, mod.prod_dist, backend
        except ImportError as e:  # library not built, or no CUDA device
            last_error = e
    raise ImportError(f"Could not load any backend. Last error: {last_error}")

"""TEST / BENCH INFRASTRUCTURE -- installs the UNMODIFIED reference into ``baseline/_ref``.

    python oracle/install_reference.py [--force]

The reference (eudald-seeslab/mmsbm v1.0.7, /root/reference) is a flat set of pure-Python
modules (setup.py:10-12 ``py_modules``), so "building" it is one offline pip install from a
writable copy of the tree (the build writes egg-info next to setup.py and /root/reference is
read-only):

    pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
        --target baseline/_ref /tmp/<copy of /root/reference>

``--no-deps``: numpy / pandas / tqdm (setup.py:28-32, unpinned) are already in the image and
there is no index to resolve them from.  ``baseline/_ref`` is git-ignored (no reference source
enters the history) but NOT gpurun-ignored, so the installed modules travel to the GPU box
where /root/reference does not exist.  Consumers (the only ones allowed to touch it):

  * bench.py ``--impl reference`` and the ``cpu_baseline`` leg: the reference's own
    ``kernels_numba`` / ``kernels_numpy`` ``update_coefficients`` + its
    ``ExpectationMaximization`` normalisations, timed on the box's host cores
    (``cpu_baseline.kind = "reference"``; the oracle port is the fallback when the
    directory is absent);
  * tests/test_reference_plugin.py: the reference's ``MMSBM(backend="b200")`` through its
    own spawn pool, loading this repo's ``kernels_b200`` plugin (INTEGRATION.md level 1).

Nothing under ``mmsbm_b200/`` imports it.
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"
TARGET = os.path.join(ROOT, "baseline", "_ref")
MODULES = ("mmsbm", "expectation_maximization", "data_handler", "helpers", "backend", "logger",
           "kernels_numpy", "kernels_numba", "kernels_cupy")


def installed():
    return all(os.path.exists(os.path.join(TARGET, m + ".py")) for m in MODULES)


def install(force=False):
    """Returns the target directory, or None when /root/reference is not there (GPU box)."""
    if installed() and not force:
        return TARGET
    if not os.path.isdir(REFERENCE):
        return None
    tmp = tempfile.mkdtemp(prefix="mmsbm_ref_")
    try:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REFERENCE, src)
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        os.makedirs(TARGET)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", TARGET, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode:
            sys.stderr.write(res.stdout[-2000:] + res.stderr[-2000:])
            raise RuntimeError("pip install of the reference failed")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    assert installed(), "reference modules missing after the install"
    return TARGET


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))

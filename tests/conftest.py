import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def rel_err(a, b):
    """max|a-b| / max|b|  (max-norm relative to the field magnitude; SURVEY.md section 7, hard part 4)"""
    a = np.asarray(a)
    b = np.asarray(b)
    scale = np.abs(b).max()
    if scale == 0.0:
        return float(np.abs(a - b).max())
    return float(np.abs(a - b).max() / scale)


# groups of hydrovs components: (components compared, components that set the scale).  Single velocity
# components (ufbar_x ...) are measured against the magnitude of the whole velocity vector: for a flat
# interface the x components are pure rounding noise (1e-18) in both codes.
HYDRO_GROUPS = {
    "rho": ([0], [0]), "phi": ([1], [1]), "u_f": ([2, 3, 4], [2, 3, 4]), "rho_tot": ([5], [5]), "u_g": ([6, 7, 8], [6, 7, 8]),
    "a_f": ([9, 10, 11], [9, 10, 11]), "a_g": ([12, 13, 14], [12, 13, 14]), "u_b": ([15, 16, 17], [15, 16, 17]),
    "nfbar": ([18], [2, 3, 4]), "ngbar": ([19], [6, 7, 8]), "ufbar": ([20], [2, 3, 4]), "ugbar": ([21], [6, 7, 8]),
}


# velocities / accelerations that are identically zero by symmetry carry rounding noise of order 1e-18 in both
# codes; they are compared against this floor (lattice units, sound speed 0.577) instead of their own magnitude
VEL_FLOOR = 1e-4
# accelerations are -cs2*alpha0*grad(density): a gradient that vanishes by symmetry is a difference of densities that
# agree to ~1 ulp, i.e. O(alpha0 * 1e-16) absolute
ACC_FLOOR = 1e-3


def assert_hydro_close(got, want, tol, what=""):
    for name, (idx, sidx) in HYDRO_GROUPS.items():
        scale = np.abs(want[sidx]).max()
        if name in ("a_f", "a_g"):
            scale = max(scale, ACC_FLOOR)
        elif name not in ("rho", "phi", "rho_tot"):
            scale = max(scale, VEL_FLOOR)
        err = np.abs(got[idx] - want[idx]).max()
        e = float(err / scale) if scale > 0 else float(err)
        assert e <= tol, f"{what} hydrovs[{name}] rel err {e:.3e} > {tol:.1e}"


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as om
    om.build()
    return om


@pytest.fixture(scope="session")
def bflbm():
    import bflbm_b200
    return bflbm_b200

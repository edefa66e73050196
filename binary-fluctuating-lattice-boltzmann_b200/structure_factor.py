"""Host mirror of FHDeX's StructFact as the reference driver uses it (main_run_job.cpp:299-310, 342-349, 50-54), over
include/bflbm_sf.h: the hydrodynamic fields are transformed and accumulated on the GPU, only the averaged spectra come back."""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import _build
from .lattice import BflbmError, Lattice

#: the reference's pair lists, main_run_job.cpp:301-306 (indices into hydrovs, VariableNames order)
REFERENCE_PAIRS = list(zip([0, 1, 0, 2, 3, 4, 6, 7, 8, 2, 9, 15, 16, 17, 15, 18, 19, 20, 21, 20, 20, 21],
                           [0, 1, 1, 2, 3, 4, 6, 7, 8, 6, 9, 15, 16, 17, 16, 18, 19, 20, 21, 21, 18, 18]))

_sf_lib = None


def _load():
    global _sf_lib
    if _sf_lib is None:
        if not os.path.exists(_build.SF_LIB):
            raise BflbmError(f"{_build.SF_LIB} not built: run bflbm_b200.build_sf() (there is no CPU fallback)")
        ctypes.CDLL(_build.LIB, mode=ctypes.RTLD_GLOBAL)
        lib = ctypes.CDLL(_build.SF_LIB)
        vp, ip, dpp = ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p
        lib.bflbm_sf_create.restype, lib.bflbm_sf_create.argtypes = ip, [vp, ip, dpp, dpp, dpp, ctypes.POINTER(vp)]
        lib.bflbm_sf_create_multi.restype, lib.bflbm_sf_create_multi.argtypes = ip, [vp, ip, dpp, dpp, dpp, ctypes.POINTER(vp)]
        lib.bflbm_sf_destroy.restype, lib.bflbm_sf_destroy.argtypes = ip, [vp]
        lib.bflbm_sf_accumulate.restype, lib.bflbm_sf_accumulate.argtypes = ip, [vp]
        lib.bflbm_sf_reset.restype, lib.bflbm_sf_reset.argtypes = ip, [vp]
        lib.bflbm_sf_samples.restype, lib.bflbm_sf_samples.argtypes = ctypes.c_longlong, [vp]
        lib.bflbm_sf_get.restype, lib.bflbm_sf_get.argtypes = ip, [vp, ip, dpp, dpp]
        _sf_lib = lib
    return _sf_lib


class StructureFactor:
    """StructFact(ba, dm, var_names, var_scaling, pairA, pairB) for a whole-box Lattice or a MultiLattice (slabs on several GPUs)."""

    def __init__(self, lat: Lattice, pairs=REFERENCE_PAIRS, var_scaling=None):
        self.lib = _load()
        self.lat = lat
        self.pairs = [(int(a), int(b)) for a, b in pairs]
        a = np.ascontiguousarray([p[0] for p in self.pairs], dtype=np.int32)
        b = np.ascontiguousarray([p[1] for p in self.pairs], dtype=np.int32)
        sc = None if var_scaling is None else np.ascontiguousarray(var_scaling, dtype=np.float64)
        self.h = ctypes.c_void_p()
        create = self.lib.bflbm_sf_create_multi if hasattr(lat, "ngpus") else self.lib.bflbm_sf_create
        rc = create(lat.h, len(self.pairs), a.ctypes.data, b.ctypes.data, None if sc is None else sc.ctypes.data, ctypes.byref(self.h))
        if rc:
            raise BflbmError(f"bflbm_sf_create failed ({rc})")

    def close(self):
        if self.h:
            self.lib.bflbm_sf_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def fort_structure(self):
        """StructFact::FortStructure(hydrovs, 0): add the current state's A_k conj(B_k) of every pair."""
        rc = self.lib.bflbm_sf_accumulate(self.h)
        if rc:
            raise BflbmError(f"bflbm_sf_accumulate failed ({rc})")

    def reset(self):
        self.lib.bflbm_sf_reset(self.h)

    @property
    def samples(self) -> int:
        return int(self.lib.bflbm_sf_samples(self.h))

    def result(self, zero_avg: bool = True, imag: bool = False):
        """What StructFact::WritePlotFile writes: (npairs, nz, ny, nx) sample means on the shifted k grid."""
        shape = (len(self.pairs), self.lat.nz, self.lat.ny, self.lat.nx)
        re = np.empty(shape)
        im = np.empty(shape) if imag else None
        rc = self.lib.bflbm_sf_get(self.h, 1 if zero_avg else 0, re.ctypes.data, None if im is None else im.ctypes.data)
        if rc:
            raise BflbmError(f"bflbm_sf_get failed ({rc})")
        return (re, im) if imag else re

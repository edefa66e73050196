"""CPU tests of the boundary: the C-ABI library builds, loads and exports every symbol include/bflbm.h
declares; it refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "bflbm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bflbm_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(bflbm):
    bflbm.build()
    lib = ctypes.CDLL(bflbm.LIB)
    names = declared_symbols()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/bflbm.h but not exported: {missing}"


def test_structure_factor_library_exports_its_header(bflbm):
    bflbm.build_sf()
    ctypes.CDLL(bflbm.LIB, mode=ctypes.RTLD_GLOBAL)
    lib = ctypes.CDLL(bflbm.SF_LIB)
    src = open(os.path.join(ROOT, "include", "bflbm_sf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b(bflbm_sf_[a-z0-9_]+)\s*\(", src)))
    assert len(names) == 7
    assert not [n for n in names if not hasattr(lib, n)]


def test_python_binding_covers_header(bflbm):
    lib = bflbm.load_library()
    for n in declared_symbols():
        assert getattr(lib, n).argtypes is not None or n in ("bflbm_last_error", "bflbm_version"), n


def test_no_cpu_fallback(bflbm):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(bflbm.BflbmError, match="no CUDA device|CUDA"):
        bflbm.Lattice(8)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the package or include/ may reference it."""
    pkg = os.path.join(ROOT, "binary-fluctuating-lattice-boltzmann_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".H")):
                txt = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "oracle" not in txt.lower() or fn == "README.md", f"{fn} mentions the oracle"


def test_params_defaults(bflbm):
    p = bflbm.Params()
    assert (p.kBT, p.tau_f, p.tau_g, p.alpha0, p.alpha1, p.kappa, p.rho_lo, p.rho_hi, p.seed) == (0.0, 0.5, 0.5, 4.0, 0.0, 4.0, 0.0, 1.0, 12345)
    assert len(bflbm.VARIABLE_NAMES) == 22 and bflbm.VARIABLE_NAMES[5] == "p_bulk"


def philox4x32_10_numpy(ctr, key):
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c = [int(v) for v in ctr]
    k = [int(v) for v in key]
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c[3] ^ k[1]) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
    return tuple(c)


# Random123 known-answer vectors for philox4x32-10 (kat_vectors of the Random123 distribution)
PHILOX_KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


@pytest.mark.parametrize("ctr,key,want", PHILOX_KAT)
def test_philox_numpy_restatement_kat(ctr, key, want):
    assert philox4x32_10_numpy(ctr, key) == want

"""On-GPU structure-factor accumulator (include/bflbm_sf.h) against the numpy restatement of the
reference's convention (tests/stats.py: AMReX_DFT.H:19-183; pairs of main_run_job.cpp:301-306)."""
import numpy as np
import pytest

import stats

pytestmark = pytest.mark.gpu


def test_structure_factor_accumulator_matches_numpy(bflbm):
    bflbm.build_sf()
    shape = (12, 10, 16)  # nx, ny, nz : even and odd half-spectrum sizes, non-cubic
    prm = bflbm.Params(kBT=1e-5, alpha0=1.5, kappa=0.1, rho_lo=0.1, rho_hi=3.0, seed=5)
    pairs = bflbm.REFERENCE_PAIRS
    with bflbm.Lattice(*shape, params=prm) as lat:
        lat.init_droplet(0.3)
        ref = stats.StructureFactor(pairs)
        with bflbm.StructureFactor(lat, pairs) as sf:
            for _ in range(4):
                lat.step(3)
                sf.fort_structure()
                ref.add(lat.hydrovars())
            assert sf.samples == 4
            got, got_im = sf.result(zero_avg=True, imag=True)
            want = ref.result()
            scale = np.abs(want).max(axis=(1, 2, 3), keepdims=True) + 1e-300
            assert (np.abs(got - want) / scale).max() < 1e-11
            # Hermitian symmetry of the completed spectrum: Im S(-k) = -Im S(k); auto-correlations are real
            for p, (a, b) in enumerate(pairs):
                if a == b:
                    assert np.abs(got_im[p]).max() <= 1e-12 * scale[p].max()
            sf.reset()
            assert sf.samples == 0


def test_structure_factor_needs_a_whole_box(bflbm):
    bflbm.build_sf()
    with bflbm.Lattice(16, 8, 16, params=bflbm.Params(), slab=(0, 8)) as lat:
        with pytest.raises(bflbm.BflbmError):
            bflbm.StructureFactor(lat, [(0, 0)])

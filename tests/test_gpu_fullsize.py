"""BASELINE.json's full-size configurations through size-independent properties (the CPU oracle cannot run them):

  * exact conservation laws of the step (LBM_binary.H:506-511: meq_0 = rho_s, force and noise moments 0 vanish;
    xi^g_{1..3} = -xi^f_{1..3}): both species' masses to round-off, total momentum unchanged by collision and noise
    in a force-free mixture;
  * lattice symmetries of the deterministic step: translating the initial state by a lattice vector translates the
    result (bitwise when the shift is a multiple of the tile, to round-off otherwise);
  * the fluctuating step is a pure function of (seed, global cell, step): same seed -> same bits, run to run;
  * nothing non-finite after a burst of steps.

Sizes: 256^3 droplet with noise (configs[2]), 512^3 mixture with noise (configs[3] / configs[4], one GPU's share).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _sum_fields(lat):
    hb = lat.hydrovars_bar()
    rho, phi = hb[0], hb[1]
    mom = np.array([(rho * hb[2 + k] + phi * hb[6 + k]).sum() for k in range(3)])
    return rho.sum(), phi.sum(), mom, hb


def test_config3_droplet_256cubed_conservation_and_reproducibility(bflbm):
    prm = bflbm.Params(kBT=1e-5, alpha0=1.5, kappa=0.1, rho_lo=0.1, rho_hi=3.0, seed=7)
    n = 256
    with bflbm.Lattice(n, n, n, params=prm) as lat:
        lat.init_droplet(0.2)
        m0r, m0p, _, _ = _sum_fields(lat)
        lat.step(50)
        assert lat.check_nan() == 0
        m1r, m1p, _, hb1 = _sum_fields(lat)
        assert abs(m1r - m0r) <= 1e-12 * m0r and abs(m1p - m0p) <= 1e-12 * m0p, "species masses are conserved exactly"
        mr, mp = lat.total_mass()  # the device-side reduction sees the same masses
        assert abs(mr - m1r) <= 1e-11 * m1r and abs(mp - m1p) <= 1e-11 * m1p
    with bflbm.Lattice(n, n, n, params=prm) as lat:
        lat.init_droplet(0.2)
        lat.step(50)
        assert np.array_equal(lat.hydrovars_bar(), hb1), "same (seed, cell, step) keys => same bits, run to run"


def test_deterministic_step_commutes_with_lattice_translations(bflbm):
    """kBT = 0 droplet at 128^3: shift the initial populations by (32, 8, 0) cells (multiples of the tile in x and y, so
    every brick sees the data of another brick: bitwise) and by (3, 5, 7) (round-off only: the density sums group
    differently; in z also because the restart builds the densities of the two slab-face planes as local + halo part)."""
    n = 128
    prm = bflbm.Params(kBT=0.0, alpha0=1.5, kappa=0.1, rho_lo=0.1, rho_hi=3.0)
    with bflbm.Lattice(n, n, n, params=prm) as lat:
        lat.init_droplet(0.25)
        f0, g0 = lat.populations()
        lat.set_tiling(4)
        lat.init_from_populations(f0, g0)  # same entry point for the reference run and the shifted runs
        lat.step(10)
        ref = lat.hydrovars()
        for shift, exact in (((0, 8, 32), True), ((7, 5, 3), False)):  # numpy axes are (z, y, x)
            fs = np.ascontiguousarray(np.roll(f0, shift, axis=(1, 2, 3)))
            gs = np.ascontiguousarray(np.roll(g0, shift, axis=(1, 2, 3)))
            lat.init_from_populations(fs, gs)
            lat.step(10)
            got = np.roll(lat.hydrovars(), tuple(-s for s in shift), axis=(1, 2, 3))
            if exact:
                assert np.array_equal(got, ref)
            else:
                scale = np.abs(ref).max(axis=(1, 2, 3), keepdims=True) + 1e-300
                assert (np.abs(got - ref) / scale).max() < 1e-12


def test_config5_mixture_512cubed_conservation(bflbm):
    """One GPU's share of the weak-scaling job (the bench workload): force-free mixture (alpha0 = 0), noise on.
    Masses exact; the total momentum receives neither force nor noise (xi^g = -xi^f) and stays at round-off."""
    prm = bflbm.Params(kBT=1e-5, alpha0=0.0, seed=11)
    n = 512
    with bflbm.Lattice(n, n, n, params=prm) as lat:
        lat.init_mixture()
        lat.step(20)
        assert lat.check_nan() == 0
        mr, mp = lat.total_mass()
        assert abs(mr - n ** 3) <= 1e-12 * n ** 3 and abs(mp - n ** 3) <= 1e-12 * n ** 3
        hb = lat.hydrovars_bar()
        rho, phi = hb[0], hb[1]
        for k in range(3):
            p = float((rho * hb[2 + k]).sum() + (phi * hb[6 + k]).sum())
            # per-cell momentum noise is O(sqrt(kBT)) ~ 3e-3; an unbalanced sum over 1.3e8 cells would be ~ 40
            assert abs(p) < 1e-6, f"total momentum component {k} = {p:.3e}"
        var = float(((rho - rho.mean()) ** 2).mean() / (prm.kBT * 3.0))
        assert 0.5 < var < 1.05, f"density fluctuations are building up towards kBT/cs2: {var:.3f} after 20 steps"

"""Depth test of the in-kernel Gaussian generator (Philox4x32-10 + fp32 Box-Muller on the MUFU fast paths, 22-bit radius, 1024
angles per word: csrc/philox.cuh) at the sample count a production run draws in a second: 2^20 cells x 300 steps x 33 draws =
1.04e10 normals, binned on the device (bflbm_debug_normal_statistics).

  * moments: mean, variance, skewness, kurtosis within 5 sigma of N(0, 1) at that N (5e-5 on the mean);
  * chi-square of the pooled histogram, 220 bins of 0.05 over [-5.5, 5.5], against the exact normal bin probabilities;
  * tails: the generator is bounded, |n| <= sqrt(2 ln 2^23) = 5.6467 (U >= 2^-23); the mass in 5 <= |n| < 5.65 must be the normal one,
    nothing lies beyond 5.65 (a normal stream would put ~160 of 1e10 there: the documented truncation, 1.6e-8 of the mass);
  * pair-joint chi-square: the two normals made from ONE Philox word (cos / sin branch), 32 x 32 bins over [-4, 4)^2, against the
    product of the marginals -- where a dependence between a pair's members would show.
The reference draws from amrex::RandomNormal (third party, unpinned): there is no stream to compare with, only N(0, 1)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _phi_cdf(x):
    return 0.5 * (1.0 + math.erf(x / math.sqrt(2.0)))


def test_ten_billion_normals(bflbm):
    lib = bflbm.load_library()
    ncells, nsteps, nbins, lo, hi = 1 << 20, 300, 240, -6.0, 6.0
    hist = np.zeros(nbins + 2, dtype=np.uint64)
    joint = np.zeros(32 * 32, dtype=np.uint64)
    mom = np.zeros(4)
    rc = lib.bflbm_debug_normal_statistics(20261018, ncells, 0, nsteps, nbins, lo, hi, hist.ctypes.data, joint.ctypes.data, mom.ctypes.data)
    assert rc == 0
    N = ncells * nsteps * 33
    assert int(hist.sum()) == N
    m = mom / N
    assert abs(m[0]) < 5 / math.sqrt(N), f"mean {m[0]:.3e}"
    assert abs(m[1] - 1) < 5 * math.sqrt(2 / N), f"variance {m[1]:.8f}"
    assert abs(m[2]) < 5 * math.sqrt(15 / N), f"third moment {m[2]:.3e}"
    assert abs(m[3] - 3) < 5 * math.sqrt(96 / N), f"fourth moment {m[3]:.6f}"
    # pooled histogram
    edges = lo + (hi - lo) / nbins * np.arange(nbins + 1)
    p = np.array([_phi_cdf(edges[i + 1]) - _phi_cdf(edges[i]) for i in range(nbins)])
    inner = (edges[:-1] >= -5.5 - 1e-9) & (edges[1:] <= 5.5 + 1e-9)
    O, E = hist[:nbins][inner].astype(float), N * p[inner]
    chi2, dof = float(((O - E) ** 2 / E).sum()), int(inner.sum()) - 1
    assert chi2 < dof + 8 * math.sqrt(2 * dof), f"chi2 = {chi2:.1f} for {dof} dof"
    # tails
    assert hist[nbins] == 0 and hist[nbins + 1] == 0, "nothing outside [-6, 6)"
    beyond = edges[:-1] >= 5.65 - 1e-9
    assert hist[:nbins][beyond].sum() == 0 and hist[:nbins][edges[1:] <= -5.65 + 1e-9].sum() == 0, "bounded at 5.6467"
    shell = ((edges[:-1] >= 5.0 - 1e-9) & (edges[1:] <= 5.65 + 1e-9)) | ((edges[1:] <= -5.0 + 1e-9) & (edges[:-1] >= -5.65 - 1e-9))
    got = float(hist[:nbins][shell].sum())
    want = N * 2 * (_phi_cdf(5.6467) - _phi_cdf(5.0))
    assert abs(got - want) < 6 * math.sqrt(want), f"mass in 5 <= |n| < 5.65: {got:.0f}, normal {want:.0f}"
    # the two members of a Box-Muller pair are independent
    J = joint.reshape(32, 32).astype(float)
    npairs = ncells * nsteps * 9
    e1 = np.array([_phi_cdf(-4 + 0.25 * (i + 1)) - _phi_cdf(-4 + 0.25 * i) for i in range(32)])
    Ej = npairs * np.outer(e1, e1)
    chi2j, dofj = float(((J - Ej) ** 2 / Ej).sum()), 32 * 32 - 1
    assert chi2j < dofj + 8 * math.sqrt(2 * dofj), f"joint chi2 = {chi2j:.1f} for {dofj} dof"
    print(f"N = {N:.3e}: mean {m[0]:.2e} var-1 {m[1] - 1:.2e} m3 {m[2]:.2e} m4-3 {m[3] - 3:.2e}; chi2 {chi2:.1f}/{dof}; "
          f"tail shell {got:.0f} vs {want:.0f}; joint chi2 {chi2j:.1f}/{dofj}")

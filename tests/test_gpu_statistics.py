"""Statistical parity of the fluctuating step (north_star: "equipartition of velocity variance, the noise-covariance
matrix, and the static structure factor S(k) and capillary-wave spectrum must agree within stated tolerances").

The reference's random stream (amrex::RandomNormal) is third-party and unpinned, so with noise on the CUDA path and
the reference can only agree statistically.  Each case below is run through the CUDA library with the SAME driver
function, analysis code (tests/stats.py) and parameters that produced the committed reference statistics
tests/golden/stats_*.json (tests/golden/make_stats_golden.py, oracle/_ref = the reference's own headers), and is held
against (i) those reference statistics and (ii) the closed-form expectation the reference's notebooks use.

Tolerances (stated per assertion) are ~5 sigma of the sampling error of the estimator at the case's sample count.
"""
import importlib.util
import json
import os

import numpy as np
import pytest

import stats

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _gold_module():
    spec = importlib.util.spec_from_file_location("make_stats_golden", os.path.join(HERE, "golden", "make_stats_golden.py"))
    # the generator imports the oracle loader at module scope; it is only USED by its golden_* functions
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _record(name, got):
    """BFLBM_STATS_OUT=<dir>: keep the measured statistics next to the reference's (profiles/ evidence)."""
    out = os.environ.get("BFLBM_STATS_OUT")
    if out:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, f"gpu_stats_{name}.json"), "w") as fh:
            json.dump(got, fh, indent=1)


def _gold(name):
    with open(os.path.join(HERE, "golden", f"stats_{name}.json")) as fh:
        return json.load(fh)


def test_mixture_equipartition_and_structure_factor(bflbm):
    """Config: 16^3 homogeneous mixture rho = phi = 1, alpha0 = 0, tau = 1/2, kBT = 1e-5 (Parameters:1-19 scaled down);
    600 snapshots every 10 steps after 1000.  Mixture.ipynb cells 1-2 normalisations."""
    G = _gold_module()
    ref = _gold("mixture")
    C = G.MIXTURE
    with bflbm.Lattice(*C["shape"], params=bflbm.Params(**C["params"], seed=987654321)) as lat:
        lat.init_mixture()
        got = G.run_mixture(lat.step, lat.hydrovars, C)
    _record("mixture", got)
    N = np.prod(C["shape"])
    eq, req = got["equipartition"], ref["equipartition"]
    # closed-form expectations: conserved densities carry the finite-size factor 1 - 1/N; the real species velocities
    # follow v = 3/4 vbar + 1/4 vbar' + 1/2 xibar at tau = 1/2 (Mixture.ipynb cell 1, LBM_binary.H:266-272)
    theory = {"rho": 1 - 1 / N, "phi": 1 - 1 / N, "rho_phi": 0.0, "ub": 1 - 1 / N, "uf_real": 0.75, "ug_real": 0.75, "ufug_real": 0.25,
              "ufbar": 1.0, "ugbar": 1.0, "xibar_f": 0.5}
    for k, want in theory.items():
        assert abs(eq[k] - want) < 0.02, f"equipartition[{k}] = {eq[k]:.4f}, theory {want:.4f} (tolerance 0.02 absolute)"
        assert abs(eq[k] - req[k]) < 0.02, f"equipartition[{k}] = {eq[k]:.4f}, reference code {req[k]:.4f} (tolerance 0.02 absolute)"
    # static structure factors, shell averaged: flat in k (= 1 normalised), cross term 0; 3 % per shell
    s, rs = np.array(got["sf_shells_normalised"]), np.array(ref["sf_shells_normalised"])
    want = np.array([1, 1, 0, 1, 1, 1.0])[:, None]
    assert np.abs(s - want).max() < 0.03, f"S(k) shells deviate from flat: {s}"
    assert np.abs(s - rs).max() < 0.03, f"S(k) shells deviate from the reference code's: {np.abs(s - rs).max():.4f}"


def test_noise_covariance_matrix(bflbm, oracle_mod):
    """NoiseCovariance.ipynb cell 3 on a droplet state (densities from ~0.004 to 3.3): the variance of every noise
    component is the amplitude^2 of LBM_binary.H:113-127, components are uncorrelated, f/g momentum noise is exactly
    opposite.  200 frames of 8^3 = 102 400 samples per component: sigma(variance ratio) = 0.44 %."""
    G = _gold_module()
    ref = _gold("noise")
    C = G.NOISE
    f, g = oracle_mod.droplet_populations(*C["shape"], C["radius"], C["params"]["kappa"], C["rho_lo"], C["rho_hi"])
    with bflbm.Lattice(*C["shape"], params=bflbm.Params(**C["params"], rho_lo=C["rho_lo"], rho_hi=C["rho_hi"], seed=31337)) as lat:
        lat.init_from_populations(f, g)

        def frames():
            for _ in range(C["frames"]):
                lat.step(1)
                fn, gn = lat.noise()
                hb = lat.hydrovars_bar()
                yield fn, gn, hb[0], hb[1]
        got = G.run_noise(frames, C)
    _record("noise", got)
    vf, vg = np.array(got["variance_ratio_f"]), np.array(got["variance_ratio_g"])
    assert np.abs(vf - 1).max() < 0.025 and np.abs(vg - 1).max() < 0.025, f"noise variance / theory: f {vf}, g {vg}"
    assert np.abs(vf - np.array(ref["variance_ratio_f"])).max() < 0.035  # two independent estimates, 0.44 % sigma each
    assert np.abs(vg - np.array(ref["variance_ratio_g"])).max() < 0.035
    assert got["max_dev_from_expected_corr"] < 0.025, "noise correlation matrix is not identity (+ the -1 f/g momentum block)"
    assert np.allclose(got["momentum_fg_corr"], -vf[:3], rtol=0, atol=1e-12), "g momentum noise must be exactly -f momentum noise"


def test_capillary_wave_spectrum(bflbm, oracle_mod):
    """Flat interface (Parameters:22-37 recipe, alpha0 = 1.5, kappa = 0.1, rho in [0.1, 3], scaled to 2 x 32 x 40):
    3000 deterministic steps, then kBT = 1e-5; interface height per column every 100 steps for 150 000 steps;
    Flat_Interface.ipynb cells 4, 7, 9 analysis.  Compared mode by mode with the reference code's spectrum.
    (The reference's own spectrum sits ~40x below kBT ny / (gamma nx k^2) with the Laplace-law gamma ~ 0.012 of
    Surface_Tension.ipynb -- the same deficit the authors found for the droplet shape modes, SURVEY.md section 4 --
    so the acceptance is GPU-vs-reference on the same estimator, not a theoretical gamma.)"""
    path = os.path.join(HERE, "golden", "stats_capillary.json")
    if not os.path.exists(path):
        pytest.skip("tests/golden/stats_capillary.json not generated")
    G = _gold_module()
    ref = _gold("capillary")
    C = ref["case"]
    C["shape"] = tuple(C["shape"])
    f, g = oracle_mod.stripe_populations(*C["shape"], C["frac"], C["params"]["kappa"], C["rho_lo"], C["rho_hi"])
    # the same ensemble as the reference side: `replicas` independent runs (own seed each), spectra averaged
    nrep = int(ref.get("replicas", 1))
    runs = []
    for r in range(nrep):
        prm = dict(C["params"], kBT=0.0, rho_lo=C["rho_lo"], rho_hi=C["rho_hi"], seed=271828 + 7919 * r)
        with bflbm.Lattice(*C["shape"], params=bflbm.Params(**prm)) as lat:
            lat.init_from_populations(f, g)
            runs.append(G.run_capillary(lat.step, lambda: lat.hydrovars_bar()[0], lambda kbt: lat.set_params(kBT=kbt), C))
    got = G.combine_capillary(runs, C)
    _record("capillary", got)
    # deterministic relaxation: same interface position as the reference code (free-running 3000 steps)
    assert abs(got["h_det"] - ref["h_det"]) < 1e-8 * ref["h_det"], (got["h_det"], ref["h_det"])
    assert abs(got["h_mean"] - ref["h_mean"]) < 0.02, (got["h_mean"], ref["h_mean"])
    k, p, rp = np.array(got["k"]), np.array(got["hk2"]), np.array(ref["hk2"])
    low = k <= 1.2  # the capillary regime (k * interface width < 1)
    ratio = p[low] / rp[low]
    # 16 x 1500 frames on each side: ~3 % per mode and side (the slowest modes decorrelate over tens of frames)
    tol_mode, tol_mean = (0.10, 0.04) if nrep >= 8 else (0.25, 0.10)
    assert np.abs(ratio - 1).max() < tol_mode, f"<|h_k|^2> GPU / reference per mode: {ratio}"
    assert abs(np.mean(ratio) - 1) < tol_mean, f"mean spectrum ratio {np.mean(ratio):.3f}"
    assert abs(got["gamma_lowk"] / ref["gamma_lowk"] - 1) < tol_mean + 0.01, (got["gamma_lowk"], ref["gamma_lowk"])
    # k^-2 law of the capillary regime: k^2 <|h_k|^2> flat within 35 % over the three lowest modes
    flat = (k ** 2 * p)[:3]
    assert flat.max() / flat.min() < 1.35, flat


def test_droplet_shape_mode_variance(bflbm, oracle_mod):
    """configs[2] (Droplet_Fluctuation) scaled to 24^3: droplet r = 0.3, alpha0 = 1.5, kappa = 0.1, rho in [0.1, 3];
    2000 deterministic steps, then kBT = 2e-5; every 40 steps the relative principal semi-axes of the mass-weighted
    covariance of rho (LBM_hydrovs.H:258-335) for 24 000 steps; the two shape-mode sums of Droplet_Fluctuation.ipynb cells
    22-25 against the reference code's (600 frames, shape modes decorrelate within ~10 frames: ~20 % sampling error each)."""
    path = os.path.join(HERE, "golden", "stats_droplet.json")
    if not os.path.exists(path):
        pytest.skip("tests/golden/stats_droplet.json not generated")
    G = _gold_module()
    ref = _gold("droplet")
    C = ref["case"]
    C["shape"] = tuple(C["shape"])
    f, g = oracle_mod.droplet_populations(*C["shape"], C["radius"], C["params"]["kappa"], C["rho_lo"], C["rho_hi"])
    nrep = int(ref.get("replicas", 1))
    runs = []
    for r in range(nrep):
        prm = dict(C["params"], kBT=0.0, rho_lo=C["rho_lo"], rho_hi=C["rho_hi"], seed=161803 + 7919 * r)
        with bflbm.Lattice(*C["shape"], params=bflbm.Params(**prm)) as lat:
            lat.init_from_populations(f, g)
            runs.append(G.run_droplet(lat.step, lambda: lat.hydrovars_bar()[0], lambda kbt: lat.set_params(kBT=kbt), C))
    got = G.combine_droplet(runs)
    _record("droplet", got)
    assert np.allclose(got["axes_det"], ref["axes_det"], rtol=1e-9, atol=0), "deterministic relaxation: same droplet shape"
    assert np.abs(np.array(got["axes_mean"]) - np.array(ref["axes_mean"])).max() < (1e-3 if nrep >= 8 else 2e-3)
    lo, hi = (0.75, 1.25) if nrep >= 8 else (0.6, 1.6)  # 16 x 600 frames per side: ~6 % each
    for key in ("sum_plus", "sum_minus"):
        ratio = got[key] / ref[key]
        assert lo < ratio < hi, f"{key}: GPU {got[key]:.3e} vs reference {ref[key]:.3e} (ratio {ratio:.2f})"

#!/usr/bin/env python
"""Generates tests/golden/stats_*.json: fluctuation statistics of the REFERENCE code itself (oracle/_ref = the
reference's own headers compiled over the test shim, with its CPU random stream: std::mt19937 +
std::normal_distribution behind amrex::RandomNormal), analysed with tests/stats.py.

The GPU tests (tests/test_gpu_statistics.py) run the same cases through the CUDA library and apply the same analysis.
amrex::RandomNormal's stream is third-party and unpinned, so agreement is statistical by construction.

Run here (the container that has /root/reference):   python tests/golden/make_stats_golden.py [mixture|noise|capillary ...]
Cases and cost: mixture ~1 min, noise ~10 s on 8 host threads; capillary and droplet are ensembles of REPLICAS independent
runs (different seeds, one host thread each, BFLBM_GOLDEN_WORKERS processes at a time): ~25 core-minutes per replica.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import stats  # noqa: E402
from oracle import oracle as om  # noqa: E402

# ---- case definitions shared with the GPU tests ---------------------------------------------------------------
MIXTURE = dict(shape=(16, 16, 16), params=dict(kBT=1e-5, tau_f=0.5, tau_g=0.5, alpha0=0.0, alpha1=0.0, kappa=4.0),
               equil=1000, steps=6000, every=10)
NOISE = dict(shape=(8, 8, 8), params=dict(kBT=1e-5, tau_f=0.5, tau_g=0.5, alpha0=1.5, alpha1=0.0, kappa=0.1), radius=0.3,
             rho_lo=0.1, rho_hi=3.0, frames=200)
CAPILLARY = dict(shape=(2, 32, 40), params=dict(kBT=1e-5, tau_f=0.5, tau_g=0.5, alpha0=1.5, alpha1=0.0, kappa=0.1),
                 rho_lo=0.1, rho_hi=3.0, frac=0.5, det_steps=3000, equil=10000, steps=150000, every=100)
DROPLET = dict(shape=(24, 24, 24), params=dict(kBT=2e-5, tau_f=0.5, tau_g=0.5, alpha0=1.5, alpha1=0.0, kappa=0.1), radius=0.3,
               rho_lo=0.1, rho_hi=3.0, det_steps=2000, equil=2000, steps=24000, every=40)
REPLICAS = 16  # independent runs per ensemble case (round 1 had one run: +-25 % per capillary mode; 16 runs: +-10 %)
SF_PAIRS = [(0, 0), (1, 1), (0, 1), (15, 15), (16, 16), (17, 17)]  # rho-rho, phi-phi, rho-phi, ub_a-ub_a


def ref(shape, params, seed):
    O = om.RefOracle(*shape, fast=True)
    O.set_params(**params)
    O.set_rng(2, seed)
    return O


def run_mixture(stepper, hydro, C=MIXTURE):
    """stepper(n): advance n steps; hydro(): (22, nz, ny, nx).  Returns the statistics dict."""
    eq = stats.Equipartition(C["params"]["kBT"])
    sf = stats.StructureFactor(SF_PAIRS)
    stepper(C["equil"])
    for _ in range(C["steps"] // C["every"]):
        stepper(C["every"])
        h = hydro()
        eq.add(h)
        sf.add(h)
    kT = C["params"]["kBT"]
    k, shells = sf.shell_means(5)
    # Mixture.ipynb cell 2 normalisations: densities by kBT/cs2 * rho, barycentric velocity by kBT/(rho+phi)
    norm = np.array([kT / stats.CS2, kT / stats.CS2, kT / stats.CS2, kT / 2.0, kT / 2.0, kT / 2.0])
    return {"equipartition": eq.result(), "sf_k": k.tolist(), "sf_shells_normalised": (shells / norm[:, None]).tolist(), "samples": eq.n}


def run_noise(frames_fn, C=NOISE):
    """frames_fn(): iterator of (fn, gn, rho, phi) snapshots."""
    kT, tau = C["params"]["kBT"], C["params"]["tau_f"]
    zs, rf, rg = [], [], []
    for fn, gn, rho, phi in frames_fn():
        vf, vg = stats.noise_amplitudes2(rho, phi, kT, tau)
        zs.append(np.concatenate([fn[1:] / np.sqrt(vf[1:]), gn[1:] / np.sqrt(vg[1:])]).reshape(36, -1))
        assert not fn[0].any() and not gn[0].any()
    z = np.concatenate(zs, axis=1)
    corr = z @ z.T / z.shape[1]
    expected = np.eye(36)
    for i in range(3):  # species g gets MINUS the momentum noise of species f (LBM_binary.H:117-119)
        expected[i, 18 + i] = expected[18 + i, i] = -1.0
    return {"variance_ratio_f": np.diag(corr)[:18].tolist(), "variance_ratio_g": np.diag(corr)[18:].tolist(),
            "max_dev_from_expected_corr": float(np.abs(corr - expected).max()),
            "momentum_fg_corr": [float(corr[i, 18 + i]) for i in range(3)], "samples": int(z.shape[1])}


def run_capillary(stepper, rho_field, set_kbt, C=CAPILLARY):
    level = 0.5 * (C["rho_lo"] + C["rho_hi"])
    set_kbt(0.0)
    stepper(C["det_steps"])
    h_det = float(stats.interface_height(rho_field(), level).mean())
    set_kbt(C["params"]["kBT"])
    stepper(C["equil"])
    hs = []
    for _ in range(C["steps"] // C["every"]):
        stepper(C["every"])
        hs.append(stats.interface_height(rho_field(), level))
    hs = np.array(hs)
    k, p = stats.capillary_spectrum(hs)
    nx, ny = C["shape"][0], C["shape"][1]
    return {"h_det": h_det, "h_mean": float(hs.mean()), "k": k.tolist(), "hk2": p.tolist(), "frames": int(hs.shape[0]),
            "gamma_lowk": stats.surface_tension_from_spectrum(k, p, C["params"]["kBT"], ny, nx, kmax=0.8)}


def run_droplet(stepper, rho_field, set_kbt, C=DROPLET):
    """configs[2] scaled down: deterministic relaxation of the droplet, then noise; relative principal axes of the
    mass-weighted covariance every `every` steps; shape-mode sums of Droplet_Fluctuation.ipynb cells 22-25."""
    set_kbt(0.0)
    stepper(C["det_steps"])
    ax_det = stats.droplet_axes(rho_field())
    set_kbt(C["params"]["kBT"])
    stepper(C["equil"])
    ax = []
    for _ in range(C["steps"] // C["every"]):
        stepper(C["every"])
        ax.append(stats.droplet_axes(rho_field()))
    ax = np.array(ax)
    plus, minus = stats.shape_mode_variances(ax)
    return {"axes_det": ax_det.tolist(), "axes_mean": ax.mean(axis=0).tolist(), "axes_var": ax.var(axis=0).tolist(),
            "sum_plus": plus, "sum_minus": minus, "frames": int(ax.shape[0])}


def combine_capillary(runs, C=CAPILLARY):
    """Ensemble of independent runs: every run removes its own mean height per column (Flat_Interface.ipynb cell 9), the spectra
    are averaged with equal weights (equal frame counts)."""
    k = np.array(runs[0]["k"])
    p = np.mean([r["hk2"] for r in runs], axis=0)
    nx, ny = C["shape"][0], C["shape"][1]
    return {"h_det": runs[0]["h_det"], "h_mean": float(np.mean([r["h_mean"] for r in runs])), "k": k.tolist(), "hk2": p.tolist(),
            "frames": int(sum(r["frames"] for r in runs)), "replicas": len(runs),
            "gamma_lowk": stats.surface_tension_from_spectrum(k, p, C["params"]["kBT"], ny, nx, kmax=0.8)}


def combine_droplet(runs):
    return {"axes_det": runs[0]["axes_det"], "axes_mean": np.mean([r["axes_mean"] for r in runs], axis=0).tolist(),
            "axes_var": np.mean([r["axes_var"] for r in runs], axis=0).tolist(), "sum_plus": float(np.mean([r["sum_plus"] for r in runs])),
            "sum_minus": float(np.mean([r["sum_minus"] for r in runs])), "frames": int(sum(r["frames"] for r in runs)), "replicas": len(runs)}


def _one_droplet(seed):
    C = DROPLET
    O = ref(C["shape"], dict(C["params"], kBT=0.0), seed)
    O.set_num_threads(1)
    f, g = om.droplet_populations(*C["shape"], C["radius"], C["params"]["kappa"], C["rho_lo"], C["rho_hi"])
    O.init_from_populations(f, g)
    return run_droplet(O.step, lambda: O.hydrovars_bar()[0], lambda kbt: O.set_params(kBT=kbt))


def _one_capillary(seed):
    C = CAPILLARY
    O = ref(C["shape"], dict(C["params"], kBT=0.0), seed)
    O.set_num_threads(1)
    f, g = om.stripe_populations(*C["shape"], C["frac"], C["params"]["kappa"], C["rho_lo"], C["rho_hi"])
    O.init_from_populations(f, g)
    return run_capillary(O.step, lambda: O.hydrovars_bar()[0], lambda kbt: O.set_params(kBT=kbt))


def _ensemble(fn, seed0):
    import multiprocessing as mp
    workers = int(os.environ.get("BFLBM_GOLDEN_WORKERS", "6"))
    with mp.get_context("fork").Pool(workers) as pool:
        return pool.map(fn, [seed0 + 1000 * i for i in range(REPLICAS)], chunksize=1)


def golden_droplet():
    runs = _ensemble(_one_droplet, 1234)
    out = combine_droplet(runs)
    out["per_replica"] = [{"sum_plus": r["sum_plus"], "sum_minus": r["sum_minus"]} for r in runs]
    return out


def golden_mixture():
    C = MIXTURE
    O = ref(C["shape"], C["params"], 20261018)
    O.init_mixture()
    return run_mixture(O.step, O.hydrovars)


def golden_noise():
    C = NOISE
    O = ref(C["shape"], C["params"], 4242)
    f, g = om.droplet_populations(*C["shape"], C["radius"], C["params"]["kappa"], C["rho_lo"], C["rho_hi"])
    O.init_from_populations(f, g)

    def frames():
        for _ in range(C["frames"]):
            O.step(1)
            fn, gn = O.noise()
            hb = O.hydrovars_bar()
            yield fn, gn, hb[0], hb[1]
    return run_noise(frames)


def golden_capillary():
    runs = _ensemble(_one_capillary, 99)
    out = combine_capillary(runs)
    out["per_replica_hk2"] = [r["hk2"] for r in runs]
    return out


if __name__ == "__main__":
    om.build()
    which = sys.argv[1:] or ["mixture", "noise", "capillary", "droplet"]
    for name in which:
        t0 = time.time()
        res = {"mixture": golden_mixture, "noise": golden_noise, "capillary": golden_capillary, "droplet": golden_droplet}[name]()
        res["case"] = {"mixture": MIXTURE, "noise": NOISE, "capillary": CAPILLARY, "droplet": DROPLET}[name]
        res["generator"] = "oracle/_ref (reference headers over oracle/shim, fast build, per-thread mt19937 streams)"
        res["seconds"] = time.time() - t0
        with open(os.path.join(HERE, f"stats_{name}.json"), "w") as fh:
            json.dump(res, fh, indent=1)
        print(name, "done in %.0f s" % res["seconds"])

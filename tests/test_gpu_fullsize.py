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


def _gold(name):
    import json
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    return json.load(open(os.path.join(here, "golden", f"stats_{name}.json")))


def _gold_module():
    import importlib.util
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_stats_golden", os.path.join(here, "golden", "make_stats_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _record(name, got):
    import json
    import os
    out = os.environ.get("BFLBM_STATS_OUT")
    if out:
        os.makedirs(out, exist_ok=True)
        json.dump(got, open(os.path.join(out, f"gpu_stats_{name}.json"), "w"), indent=1)


def test_config2_capillary_waves_at_128cubed(bflbm):
    """BASELINE configs[1]: flat interface at 128^3 with noise, capillary-wave spectrum and surface tension against the reference.
    The reference code can only be run on a thin box (2 x 32 x 40, tests/golden/stats_capillary.json), where the notebook's
    one-slice spectrum (Flat_Interface.ipynb cells 7, 9) is the k_x = 0 row of the sheet's 2-D spectrum (its only other row,
    k_x = pi, carries 1/250 of the weight).  In a 128-wide sheet a slice mixes 128 k_x rows, most of them beyond the capillary
    regime, so the comparison is made on what both boxes resolve: the k_x = 0 modes, <|h(0, k_y)|^2> = kBT nx ny / (gamma k_y^2)
    (numpy 'backward' norm), i.e. gamma(k_y) = kBT ny / (nx k_y^2 <|fft_y(mean_x h)|^2>), over the common band 0.19 <= k_y <= 0.8
    (the thin box's four lowest modes; longer waves of the big box would need more than the recipe's 10 000 equilibration steps).
    600 frames; tolerance 15 % on the band mean (13 modes x ~600/5 independent frames on this side, 16 x 1500 on the other)."""
    import stats
    ref = _gold("capillary")
    C = dict(ref["case"])
    shape, nframes, every = (128, 128, 128), 600, 100
    nx, ny = shape[0], shape[1]
    kT = C["params"]["kBT"]
    level = 0.5 * (C["rho_lo"] + C["rho_hi"])
    prm = dict(C["params"], kBT=0.0, rho_lo=C["rho_lo"], rho_hi=C["rho_hi"], seed=8128)
    with bflbm.Lattice(*shape, params=bflbm.Params(**prm)) as lat:
        lat.init_stripe(C["frac"])
        lat.step(C["det_steps"])
        lat.set_params(kBT=kT)
        lat.step(C["equil"])
        hs = []
        for _ in range(nframes):
            lat.step(every)
            hs.append(stats.interface_height(lat.hydrovars_bar()[0], level))
    hs = np.array(hs)  # (frames, ny, nx)
    hbar = hs.mean(axis=2)
    hk = np.fft.fft(hbar - hbar.mean(axis=0, keepdims=True), axis=1, norm="backward")
    p0 = (np.abs(hk) ** 2).mean(axis=0)[1:ny // 2]
    k = (2 * np.pi * np.fft.fftfreq(ny))[1:ny // 2]
    band = (k >= 0.19) & (k <= 0.8)
    g_gpu = float(np.mean(kT * ny / (nx * k[band] ** 2 * p0[band])))
    rk, rp = np.array(ref["k"]), np.array(ref["hk2"])
    rnx, rny = ref["case"]["shape"][0], ref["case"]["shape"][1]
    rband = (rk >= 0.19) & (rk <= 0.8)
    kx_r = 2 * np.pi * np.fft.fftfreq(rnx)
    # the thin box's slice spectrum with its (tiny) k_x != 0 rows taken out: <|h(0,k)|^2> / nx^2 = p - sum_{kx != 0} theory
    g_ref = float(np.mean([kT * rny / rnx * np.sum(1.0 / (kx_r ** 2 + kk ** 2)) / pp for kk, pp in zip(rk[rband], rp[rband])]))
    ks, ps = stats.capillary_spectrum(hs)
    got = {"gamma_kx0_band": g_gpu, "gamma_reference_band": g_ref, "k": k.tolist(), "h_kx0_k2": p0.tolist(), "frames": nframes,
           "gamma_slice_estimator_lowk": stats.surface_tension_from_spectrum(ks, ps, kT, ny, nx, kmax=0.8), "h_mean": float(hs.mean())}
    _record("capillary_128cubed", got)
    assert abs(g_gpu / g_ref - 1) < 0.15, f"gamma from the k_x = 0 modes, 0.19 <= k_y <= 0.8: GPU 128^3 {g_gpu:.4f} vs reference code {g_ref:.4f}"
    flat = k[band] ** 2 * p0[band]  # the k^-2 law of the k_x = 0 row
    assert flat.max() / flat.min() < 2.0, flat


def test_config3_droplet_shape_modes_scale_with_radius(bflbm, oracle_mod):
    """BASELINE configs[2] (droplet with noise): shape-mode variance against the reference, scaled.  The relative semi-axes
    a_i / R of a droplet fluctuate with variance ~ kBT / (gamma R^2) (Droplet_Fluctuation.ipynb cells 22-25), so the two shape sums
    times R^2 are size independent.  The reference code ran 24^3 (R = 7.2); here 48^3 (R = 14.4, 8 runs of 120 000 steps -- the
    l = 2 modes of the bigger droplet relax ~8 x slower) must reproduce sum * R^2 within a factor 0.6 .. 1.7, and the 256^3 droplet
    of configs[2] itself is checked for what converges fast at that size: velocity equipartition in the bulk."""
    G = _gold_module()
    ref = _gold("droplet")
    C = dict(ref["case"])
    C.update(shape=(48, 48, 48), det_steps=4000, equil=8000, steps=120000, every=80)
    f, g = oracle_mod.droplet_populations(*C["shape"], C["radius"], C["params"]["kappa"], C["rho_lo"], C["rho_hi"])
    runs = []
    for r in range(8):
        prm = dict(C["params"], kBT=0.0, rho_lo=C["rho_lo"], rho_hi=C["rho_hi"], seed=4800 + r)
        with bflbm.Lattice(*C["shape"], params=bflbm.Params(**prm)) as lat:
            lat.init_from_populations(f, g)
            runs.append(G.run_droplet(lat.step, lambda: lat.hydrovars_bar()[0], lambda kbt: lat.set_params(kBT=kbt), C))
    got = G.combine_droplet(runs)
    scale = (48 / ref["case"]["shape"][0]) ** 2
    got.update(scaled_ratio_plus=got["sum_plus"] * scale / ref["sum_plus"], scaled_ratio_minus=got["sum_minus"] * scale / ref["sum_minus"])
    _record("droplet_48cubed", got)
    for key in ("scaled_ratio_plus", "scaled_ratio_minus"):
        assert 0.6 < got[key] < 1.7, f"{key} = {got[key]:.2f} (sum * R^2 against the reference code's 24^3 droplet)"
    # configs[2] at its own size: kBT-equipartition of the barycentric velocity in the bulk of the 256^3 box.  The analytic droplet
    # launches sound waves that take ~1e4 steps to damp at this size, so the thermal part is isolated as the difference of two
    # runs that differ only in their seed (same deterministic flow): <(u_A - u_B)^2> (rho + phi) / 2 = 3 kBT.
    n = 256
    far = np.zeros((n, n, n), dtype=bool)
    far[:40] = far[-40:] = True  # slabs far from the droplet (radius 51 around the centre)
    fields = []
    for seed in (256, 257):
        prm = bflbm.Params(kBT=1e-5, alpha0=1.5, kappa=0.1, rho_lo=0.1, rho_hi=3.0, seed=seed)
        with bflbm.Lattice(n, n, n, params=prm) as lat:
            lat.init_droplet(0.2)
            lat.step(400)
            h = lat.hydrovars()
            fields.append((h[5][far].copy(), np.stack([h[15][far], h[16][far], h[17][far]])))
    (rt, ua), (_, ub) = fields
    eq = float((((ua - ub) ** 2).sum(axis=0) * rt).mean() / (2 * 3 * 1e-5))
    assert 0.85 < eq < 1.1, f"<(u_A - u_B)^2> (rho + phi) / (6 kBT) = {eq:.3f} in the bulk of the 256^3 droplet box"

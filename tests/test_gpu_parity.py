"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Tolerance (BASELINE.json north_star): with kT = 0 the rho, phi and u fields agree within 1e-12 relative
per step, measured as max|delta| / max|field| over a group of components that share a scale, from an
IDENTICAL state (one-step restart).  With noise on, the GPU's own normals are injected into the oracle
(bflbm_get_normals), which turns the fluctuating step into the same deterministic check.
"""
import glob
import os

import numpy as np
import pytest

from conftest import assert_hydro_close, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-12
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
ALGOS = ["fused", "twopass"]


def params_of(d):
    return {k[2:]: float(d[k]) for k in d.files if k.startswith("p_")}


def make_lattice(bflbm, shape, params, algo="fused", lz=None, **kw):
    nx, ny, nz = shape
    P = bflbm.Params(**{**params, **kw})
    lat = bflbm.Lattice(nx, ny, nz, params=P)
    lat.set_algorithm(algo)
    if lz:
        lat.set_tiling(lz)
    return lat


def test_philox_kat_on_device(bflbm):
    from test_abi import PHILOX_KAT
    for ctr, key, want in PHILOX_KAT:
        assert bflbm.philox4x32_10(ctr, key) == want


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("path", [p for p in GOLDEN if "noise" not in p], ids=lambda p: os.path.basename(p)[:-4])
def test_one_step_restart_vs_golden(bflbm, path, algo):
    """From the fixture's initial populations: every step, restarted from the REFERENCE state of the step before."""
    d = np.load(path)
    shape = [int(v) for v in d["shape"]]
    steps = [0] + [int(s) for s in d["steps"]]
    prm = params_of(d)
    with make_lattice(bflbm, shape, prm, algo) as lat:
        # state 0: observers reproduce hydrovs / hydrovsbar of the reference's LBM_init
        lat.init_from_populations(d["f0"], d["g0"])
        f, g = lat.populations()
        assert np.array_equal(f, d["f0"]) and np.array_equal(g, d["g0"]), "get_populations must return what was put in"
        assert_hydro_close(lat.hydrovars(), d["h0"], TOL, "init")
        assert rel_err(lat.hydrovars_bar()[[0, 1, 5]], d["hb0"][[0, 1, 5]]) <= TOL
        # consecutive fixture states that are exactly one step apart
        for a, b in zip(steps[:-1], steps[1:]):
            if b - a != 1:
                continue
            fa, ga = (d["f0"], d["g0"]) if a == 0 else (d[f"f_{a}"], d[f"g_{a}"])
            lat.init_from_populations(fa, ga)
            lat.step(1)
            f, g = lat.populations()
            assert rel_err(f, d[f"f_{b}"]) <= TOL and rel_err(g, d[f"g_{b}"]) <= TOL
            assert_hydro_close(lat.hydrovars(), d[f"h_{b}"], TOL, f"step {b}")


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("path", [p for p in GOLDEN if "noise" not in p], ids=lambda p: os.path.basename(p)[:-4])
def test_free_running_vs_golden(bflbm, path, algo):
    d = np.load(path)
    shape = [int(v) for v in d["shape"]]
    with make_lattice(bflbm, shape, params_of(d), algo) as lat:
        lat.init_from_populations(d["f0"], d["g0"])
        done = 0
        for s in [int(v) for v in d["steps"]]:
            lat.step(s - done)
            done = s
            f, g = lat.populations()
            assert rel_err(f, d[f"f_{s}"]) <= 10 * TOL and rel_err(g, d[f"g_{s}"]) <= 10 * TOL
            assert_hydro_close(lat.hydrovars(), d[f"h_{s}"], 10 * TOL, f"step {s}")
            hb = lat.hydrovars_bar()
            assert rel_err(hb[[0, 1, 5]], d[f"hb_{s}"][[0, 1, 5]]) <= 10 * TOL
            vscale = max(np.abs(d[f"hb_{s}"][[2, 3, 4, 6, 7, 8]]).max(), 1e-4)
            assert np.abs(hb[[2, 3, 4, 6, 7, 8]] - d[f"hb_{s}"][[2, 3, 4, 6, 7, 8]]).max() <= 10 * TOL * vscale
        assert lat.step_count == done


@pytest.mark.parametrize("init,arg", [("stripe", 0.5), ("droplet", 0.2), ("droplet", 0.3), ("mixture", None)])
@pytest.mark.parametrize("shape", [(32, 32, 32), (8, 24, 20), (20, 12, 33)])
def test_analytic_inits_vs_oracle(bflbm, oracle_mod, init, arg, shape):
    """LBM_init_mixture / _stripe / _droplet incl. the integer divisions and box[0]-for-z quirk on odd / non-cubic boxes."""
    prm = dict(kBT=0.0, tau_f=0.5, tau_g=0.5, alpha0=1.5, alpha1=0.0, kappa=0.6, rho_lo=0.1, rho_hi=3.0)
    O = oracle_mod.PortOracle(*shape)
    O.set_params(**prm)
    args = [] if arg is None else [arg]
    getattr(O, "init_" + init)(*args)
    with make_lattice(bflbm, shape, prm) as lat:
        getattr(lat, "init_" + init)(*args)
        f, g = lat.populations()
        fo, go = O.populations()
        # device tanh vs glibc tanh: a few ulp
        assert rel_err(f, fo) <= 1e-14 and rel_err(g, go) <= 1e-14
        assert_hydro_close(lat.hydrovars(), O.hydrovars(), TOL, "init")
        lat.step(3)
        O.step(3)
        assert_hydro_close(lat.hydrovars(), O.hydrovars(), 10 * TOL, "3 steps")


@pytest.mark.parametrize("algo", ALGOS)
def test_config1_flat_interface_32cubed(bflbm, oracle_mod, algo):
    """BASELINE config 1: flat interface 32^3, kT = 0.  Per-step (one-step restart) parity at steps
    0, 1, 10, 50 of the oracle trajectory, and the free-running state after 100 steps."""
    shape = (32, 32, 32)
    for prm in (dict(kBT=0.0, tau_f=0.5, tau_g=0.5, alpha0=4.0, alpha1=0.0, kappa=4.0, rho_lo=0.0, rho_hi=1.0),
                dict(kBT=0.0, tau_f=0.5, tau_g=0.5, alpha0=1.5, alpha1=0.0, kappa=0.1, rho_lo=0.1, rho_hi=3.0)):
        O = oracle_mod.PortOracle(*shape)
        O.set_params(**prm)
        O.init_stripe(0.5)
        f0, g0 = O.populations()
        with make_lattice(bflbm, shape, prm, algo) as lat, make_lattice(bflbm, shape, prm, algo) as free:
            free.init_from_populations(f0, g0)
            done = 0
            for s in (0, 1, 10, 50):
                O.step(s - done)
                free.step(s - done)
                done = s
                fo, go = O.populations()
                lat.init_from_populations(fo, go)
                lat.step(1)
                O.step(1)
                free.step(1)
                done += 1
                assert_hydro_close(lat.hydrovars(), O.hydrovars(), TOL, f"restart at {s}")
                fl, gl = lat.populations()
                fo, go = O.populations()
                assert rel_err(fl, fo) <= TOL and rel_err(gl, go) <= TOL
            O.step(100 - done)
            free.step(100 - done)
            # free-running: rounding differences are not amplified for this relaxing interface
            assert_hydro_close(free.hydrovars(), O.hydrovars(), 1e-10, "free-running 100 steps")
            m0 = f0.sum(), g0.sum()
            m = free.total_mass()
            assert abs(m[0] - m0[0]) <= 1e-12 * m0[0] and abs(m[1] - m0[1]) <= 1e-12 * m0[1]


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("path", [p for p in GOLDEN if "noise" in p], ids=lambda p: os.path.basename(p)[:-4])
def test_noise_amplitudes_vs_golden(bflbm, path, algo):
    """The reference's noise fields for given (rho, phi) and given normals: amplitude law of LBM_binary.H:113-127.
    GPU noise / GPU normals must reproduce fn / normals of the fixture (same state => same amplitudes)."""
    d = np.load(path)
    shape = [int(v) for v in d["shape"]]
    with make_lattice(bflbm, shape, params_of(d), algo) as lat:
        lat.init_from_populations(d["f_1"], d["g_1"])
        fn, gn = lat.noise()
        n = lat.normals()  # (nz, ny, nx, 33)
        ref_n = d["normals"][1]  # normals fed before step 1 = the draws of the noise generated at the END of step 1
        with np.errstate(divide="ignore", invalid="ignore"):
            amp_f = d["fn_1"] / np.moveaxis(ref_n, -1, 0)[[0, 0, 1, 2] + [3 + 2 * (a - 4) for a in range(4, 19)]]
            amp_g = d["gn_1"] / np.moveaxis(ref_n, -1, 0)[[0, 0, 1, 2] + [4 + 2 * (a - 4) for a in range(4, 19)]]
        nn = np.moveaxis(n, -1, 0)
        for a in range(1, 19):
            df = 3 + 2 * (a - 4) if a >= 4 else a - 1
            dg = 4 + 2 * (a - 4) if a >= 4 else a - 1
            assert rel_err(fn[a], amp_f[a] * nn[df]) <= 1e-12, f"fn[{a}]"
            assert rel_err(gn[a], amp_g[a] * nn[dg]) <= 1e-12, f"gn[{a}]"
        assert not fn[0].any() and not gn[0].any()
        assert np.array_equal(gn[1:4], -fn[1:4])


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("tau", [(0.5, 0.5), (0.8, 0.6)])
def test_fluctuating_step_with_injected_normals(bflbm, oracle_mod, algo, tau):
    """Noise on: drive the oracle with the GPU's normals; populations, noise and all 22 hydro fields (the real
    velocities contain xi/2, LBM_binary.H:266-272) must then agree to 1e-12 for several consecutive steps."""
    shape = (12, 10, 14)
    prm = dict(kBT=1e-5, tau_f=tau[0], tau_g=tau[1], alpha0=1.5, alpha1=0.0, kappa=0.1, rho_lo=0.1, rho_hi=3.0)
    with make_lattice(bflbm, shape, prm, algo, seed=2024, step0=17) as lat:
        lat.init_droplet(0.3)
        f0, g0 = lat.populations()
        O = oracle_mod.PortOracle(*shape)
        O.set_params(**prm)
        O.set_normals(lat.normals())
        O.init_from_populations(f0, g0)
        assert_hydro_close(lat.hydrovars(), O.hydrovars(), TOL, "init")
        for s in range(4):
            fn, gn = lat.noise()
            fo, go = O.noise()
            assert rel_err(fn, fo) <= TOL and rel_err(gn, go) <= TOL
            lat.step(1)
            O.set_normals(lat.normals())  # normals of the noise generated at the end of this step
            O.step(1)
            fl, gl = lat.populations()
            fo, go = O.populations()
            assert rel_err(fl, fo) <= 10 * TOL and rel_err(gl, go) <= 10 * TOL, f"step {s}"
            assert_hydro_close(lat.hydrovars(), O.hydrovars(), 10 * TOL, f"step {s}")
        assert lat.step_count == 17 + 4


@pytest.mark.parametrize("tau", [(0.5, 0.5), (0.7, 0.6)])
def test_reference_state_noise_with_injected_normals(bflbm, oracle_mod, tau):
    """USE_REF_STATE (LBM_binary.H:12, 92-107): noise amplitudes from the equilibrium profiles at the COM-shifted cell.  The CPU port
    of that branch is pinned bit for bit to the reference headers built with -DUSE_REF_STATE (tests/test_oracle.py); here the
    CUDA path is driven next to it with the GPU's normals.  The equilibrium state is displaced by (2, -1, 3) cells, so the integer
    shift the device recomputes after every step (update_com) is non-zero in every direction."""
    shape = (10, 12, 14)
    prm = dict(kBT=2e-5, tau_f=tau[0], tau_g=tau[1], alpha0=1.5, alpha1=0.0, kappa=0.1, rho_lo=0.1, rho_hi=3.0)
    f, g = oracle_mod.droplet_populations(*shape, 0.3, 0.1, 0.1, 3.0)
    P0 = oracle_mod.PortOracle(*shape)
    P0.set_params(**dict(prm, kBT=0.0))
    P0.init_from_populations(f, g)
    P0.step(5)
    hb = P0.hydrovars_bar()
    rho_eq, phi_eq = (np.roll(hb[k], (3, -1, 2), axis=(0, 1, 2)).copy() for k in (0, 1))
    # a smooth asymmetric modulation moves the centre of mass by a NON-integer amount: the shift is an integer truncation of
    # (com - com_ref), which must not sit on a rounding knife edge
    zz, yy, xx = np.meshgrid(*(np.arange(n) for n in shape[::-1]), indexing="ij")
    rho_eq *= 1.0 + 0.3 * np.cos(2 * np.pi * (xx + 0.3) / shape[0]) * np.cos(2 * np.pi * (yy - 1.1) / shape[1]) * np.sin(2 * np.pi * (zz + 0.7) / shape[2])
    O = oracle_mod.PortOracle(*shape)
    O.set_params(**prm)
    O.set_equilibrium(rho_eq, phi_eq, rho_eq + phi_eq)
    with bflbm.Lattice(*shape[:3], params=bflbm.Params(**prm, seed=77)) as lat:
        lat.set_reference_state(rho_eq, phi_eq, rho_eq + phi_eq)
        com = np.array([(rho_eq * c).sum() / rho_eq.sum() for c in np.meshgrid(*(np.arange(n) for n in shape[::-1]), indexing="ij")])[::-1]
        assert np.allclose(lat.reference_com(), com, rtol=1e-12)
        lat.init_from_populations(f, g)
        O.set_normals(lat.normals())
        O.init_from_populations(f, g)
        fn, gn = lat.noise()
        fo, go = O.noise()
        assert rel_err(fn, fo) < TOL and rel_err(gn, go) < TOL, "reference-state noise amplitudes after the restart"
        assert_hydro_close(lat.hydrovars(), O.hydrovars(), TOL, "ref-state restart")
        for s in range(4):
            lat.step(1)
            O.set_normals(lat.normals())
            O.step(1)
            fn, gn = lat.noise()
            fo, go = O.noise()
            assert rel_err(fn, fo) < 10 * TOL and rel_err(gn, go) < 10 * TOL, f"noise after step {s + 1}"
            assert_hydro_close(lat.hydrovars(), O.hydrovars(), 10 * TOL, f"ref-state step {s + 1}")
        lat.step(20)  # graph-replayed chunk with the device-side centre of mass in the loop
        assert lat.check_nan() == 0
        # switching the reference state off gives the shipped noise (current densities) again
        lat.set_reference_state(None)
        O.set_equilibrium(None, None, None)
        f1, g1 = lat.populations()
        lat.init_from_populations(f1, g1)
        O.set_normals(lat.normals())
        O.init_from_populations(f1, g1)
        assert rel_err(lat.noise()[0], O.noise()[0]) < TOL


def test_normals_are_standard_and_keyed(bflbm):
    """Counter-based noise: N(0,1) moments, independence across draws/cells/steps, reproducibility, seed/step keys."""
    prm = dict(kBT=1e-5)
    with make_lattice(bflbm, (32, 32, 32), prm, seed=1) as a, make_lattice(bflbm, (32, 32, 32), prm, seed=1) as b, \
            make_lattice(bflbm, (32, 32, 32), prm, seed=2) as c:
        for lat in (a, b, c):
            lat.init_mixture()
        na, nb, nc = a.normals(), b.normals(), c.normals()
        assert np.array_equal(na, nb), "same (seed, cell, step) => same numbers"
        assert not np.array_equal(na, nc)
        x = na.reshape(-1, 33)
        n = x.shape[0]
        assert abs(x.mean()) < 5 / np.sqrt(x.size)
        assert abs(x.var() - 1) < 5 * np.sqrt(2 / x.size)
        assert abs((x ** 4).mean() - 3) < 5 * np.sqrt(96 / x.size)
        assert np.abs(x).max() < 6.8
        cov = (x.T @ x) / n
        assert np.abs(cov - np.eye(33)).max() < 6 / np.sqrt(n), "draws of one cell are uncorrelated"
        # neighbouring cells and consecutive steps are uncorrelated
        assert abs((na[:, :, 1:, 0] * na[:, :, :-1, 0]).mean()) < 5 / np.sqrt(n)
        a.step(1)
        n1 = a.normals()
        assert abs((n1 * na).mean()) < 5 / np.sqrt(na.size)
        # per-draw KS-like check on the CDF at a few points
        from math import erf, sqrt
        for q in (-2.0, -1.0, 0.0, 0.5, 1.5):
            want = 0.5 * (1 + erf(q / sqrt(2)))
            assert abs((x < q).mean() - want) < 5 * np.sqrt(want * (1 - want) / x.size)


def test_fused_and_twopass_agree(bflbm):
    shape = (40, 24, 36)
    prm = dict(kBT=1e-5, alpha0=1.5, kappa=0.1, rho_lo=0.1, rho_hi=3.0, tau_f=0.7, tau_g=0.55)
    with make_lattice(bflbm, shape, prm, "fused") as A, make_lattice(bflbm, shape, prm, "twopass") as B:
        A.init_droplet(0.25)
        B.init_droplet(0.25)
        A.step(10)  # free running with noise: rounding differences grow step by step (1e-12 per step is checked elsewhere)
        B.step(10)
        assert_hydro_close(A.hydrovars(), B.hydrovars(), 1e-11, "fused vs two-pass, 10 fluctuating steps")


@pytest.mark.parametrize("lz", [2, 3, 5, 16])
def test_fused_result_independent_of_brick_height_to_rounding(bflbm, lz):
    shape = (33, 17, 21)  # partial bricks in every direction
    prm = dict(kBT=0.0, alpha0=1.5, kappa=0.1, rho_lo=0.1, rho_hi=3.0)
    with make_lattice(bflbm, shape, prm, "fused", lz=lz) as A, make_lattice(bflbm, shape, prm, "twopass") as B:
        A.init_droplet(0.3)
        B.init_droplet(0.3)
        A.step(7)
        B.step(7)
        assert_hydro_close(A.hydrovars(), B.hydrovars(), 1e-12, f"lz={lz}")


def test_error_behaviour(bflbm):
    with pytest.raises(bflbm.BflbmError):
        bflbm.Lattice(8, params=bflbm.Params(alpha1=0.1))
    with pytest.raises(bflbm.BflbmError):
        bflbm.Lattice(8, params=bflbm.Params(tau_f=0.0))
    with bflbm.Lattice(8) as lat:
        with pytest.raises(bflbm.BflbmError, match="before init"):
            lat.step(1)
        lat.init_mixture()
        assert lat.check_nan() == 0
        f, g = lat.populations()
        f[:] = np.nan
        lat.init_from_populations(f, g)
        with pytest.raises(bflbm.BflbmError, match="non-finite"):
            lat.check_nan()


def test_center_of_mass_and_mass(bflbm, oracle_mod):
    shape = (16, 12, 20)
    prm = dict(kBT=0.0, alpha0=1.5, kappa=0.5, rho_lo=0.1, rho_hi=2.0)
    O = oracle_mod.PortOracle(*shape)
    O.set_params(**prm)
    O.init_droplet(0.3)
    rho = O.hydrovars_bar()[0]
    z, y, x = np.meshgrid(np.arange(shape[2]), np.arange(shape[1]), np.arange(shape[0]), indexing="ij")
    com = np.array([(rho * x).sum(), (rho * y).sum(), (rho * z).sum()]) / rho.sum()
    with make_lattice(bflbm, shape, prm) as lat:
        lat.init_droplet(0.3)
        c, sums = lat.center_of_mass()
        assert np.allclose(c, com, rtol=1e-12)
        assert abs(sums[0] - rho.sum()) < 1e-12 * rho.sum()


# ---------------------------------------------------------------------------------------------------------------------
# slab decomposition (the multi-GPU path), emulated on one GPU: P slabs, device-to-device copies instead of NCCL
@pytest.mark.parametrize("peer", [False, True], ids=["copies", "peer"])
@pytest.mark.parametrize("lz", [2, 4])
@pytest.mark.parametrize("nslabs", [2, 3, 4])
@pytest.mark.parametrize("kbt", [0.0, 1e-5])
def test_slabs_bitwise_equal_to_whole_box(bflbm, nslabs, kbt, lz, peer):
    """results must not depend on the number of slabs.  Same brick height => bit-identical.
    Slabs have >= 3 brick rows, so the overlapped step (end rows + exchange on one stream, interior rows on a second)
    is the path under test; lz = 4 adds planes that the step kernel folds itself from the extended boxes.
    peer: the pack kernel writes each message straight into the neighbour lattice's mailbox and raises its arrival flag,
    the unpack kernel waits for the flag on the device (the multi-GPU protocol, here between lattices of one GPU)."""
    from bflbm_b200.distributed import EmulatedSlabs
    shape = (20, 12, 48)
    prm = bflbm.Params(kBT=kbt, alpha0=1.5, kappa=0.1, rho_lo=0.1, rho_hi=3.0, tau_f=0.7, tau_g=0.55, seed=99)
    with bflbm.Lattice(*shape, params=prm) as whole:
        whole.set_tiling(lz)
        whole.init_droplet(0.3)
        S = EmulatedSlabs(*shape, nslabs, params=prm, brick_lz=lz, peer=peer)
        try:
            S.init_droplet(0.3)
            for _ in range(3):
                whole.step(2)
                S.step(2)
                fw, gw = whole.populations()
                fs, gs = S.gather("populations")
                assert np.array_equal(fw, fs) and np.array_equal(gw, gs)
                assert np.array_equal(whole.hydrovars(), S.gather("hydrovars"))
            if kbt > 0:
                assert np.array_equal(whole.normals(), S.gather("normals")), "noise is keyed by the GLOBAL cell index"
            if peer:
                assert all(lat.halo_error() == 0 for lat in S.lats)
        finally:
            S.close()


def test_slabs_with_one_plane_last_brick_row(bflbm):
    """nz_local % brick height == 1: the last brick row of every slab is a single plane.  The overlapped schedule would
    race there (slab-face fold on the main stream against the interior rows on the second stream), so the library must
    fall back to the single-stream step.  Slab boundaries are not brick boundaries of the whole box here, so agreement is at
    rounding level rather than bitwise; a stale plane would show up at O(1e-3).  Run twice: same bits."""
    from bflbm_b200.distributed import EmulatedSlabs
    shape = (20, 12, 51)
    prm = bflbm.Params(kBT=1e-5, alpha0=1.5, kappa=0.1, rho_lo=0.1, rho_hi=3.0, tau_f=0.7, tau_g=0.55, seed=5)
    with bflbm.Lattice(*shape, params=prm) as whole:
        whole.set_tiling(4)
        whole.init_droplet(0.3)
        whole.step(6)
        want = whole.hydrovars()
    runs = []
    for _ in range(2):
        S = EmulatedSlabs(*shape, 3, params=prm, brick_lz=4)
        try:
            assert all(nzl % 4 == 1 for _, nzl in S.bounds)
            S.init_droplet(0.3)
            S.step(6)
            runs.append(S.gather("hydrovars"))
        finally:
            S.close()
    assert np.array_equal(runs[0], runs[1]), "slab step is not reproducible run to run"
    assert_hydro_close(runs[0], want, 1e-11, "slabs with a one-plane last brick row")


def test_global_array_entry_points_on_slabs(bflbm, oracle_mod):
    """bflbm_init_from_global_populations / bflbm_get_*_into_global: every slab is pointed at the WHOLE box's host arrays and
    reads / fills only its planes (the wrap-around ghost planes of the first and last slab included)."""
    from bflbm_b200.distributed import EmulatedSlabs
    shape = (12, 10, 17)
    prm = dict(kBT=0.0, tau_f=0.5, tau_g=0.5, alpha0=1.5, alpha1=0.0, kappa=0.1, rho_lo=0.1, rho_hi=3.0)
    O = oracle_mod.PortOracle(*shape)
    O.set_params(**prm)
    O.init_droplet(0.3)
    O.step(2)
    f, g = O.populations()
    S = EmulatedSlabs(*shape, 3, params=bflbm.Params(**prm), peer=True)
    try:
        for lat in S.lats:
            lat.init_from_global_populations(f, g)  # peer mode: uploads and packs
        for lat in S.lats:
            bflbm.lattice._check(lat.lib.bflbm_halo_refresh_end(lat.h))
        S.step(3)
        O.step(3)
        h = np.full((22,) + shape[::-1], np.nan)
        fo, go = np.full_like(f, np.nan), np.full_like(g, np.nan)
        for lat in S.lats:
            bflbm.lattice._check(lat.lib.bflbm_get_hydrovars_into_global(lat.h, h.ctypes.data))
            bflbm.lattice._check(lat.lib.bflbm_get_populations_into_global(lat.h, fo.ctypes.data, go.ctypes.data))
        assert np.array_equal(h, S.gather("hydrovars")) and np.array_equal(fo, S.gather("populations")[0])
        assert_hydro_close(h, O.hydrovars(), 10 * TOL, "global-array restart + 3 steps")
    finally:
        S.close()
    with bflbm.Lattice(*shape, params=bflbm.Params(**prm)) as W:  # the same entry point on a whole box
        W.init_from_global_populations(f, g)
        W.step(3)
        assert_hydro_close(W.hydrovars(), O.hydrovars(), 10 * TOL, "whole box through the global-array entry")


def test_multi_with_one_gpu_is_the_whole_box(bflbm):
    """bflbm_multi with ngpus = 1 is a plain whole-box lattice (same bits); N > 1 is covered by tests/test_gpu_multiprocess.py."""
    shape = (16, 12, 20)
    prm = bflbm.Params(kBT=1e-5, alpha0=1.5, kappa=0.1, rho_lo=0.1, rho_hi=3.0, seed=3)
    with bflbm.Lattice(*shape, params=prm) as A, bflbm.MultiLattice(*shape, params=prm, ngpus=1) as M:
        A.init_droplet(0.3)
        M.init_droplet(0.3)
        A.step(7)
        M.step(7)
        assert np.array_equal(A.hydrovars(), M.hydrovars()) and M.step_count == 7
        fa, ga = A.populations()
        M.init_from_populations(fa, ga)
        A.init_from_populations(fa, ga)
        A.step(2)
        M.step(2)
        assert np.array_equal(A.hydrovars_bar(), M.hydrovars_bar())
        assert np.allclose(A.total_mass(), M.total_mass(), rtol=1e-14)
        assert M.check_nan() == 0
        ca, cb = A.droplet_covariance(), M.droplet_covariance()
        assert all(np.array_equal(x, y) for x, y in zip(ca, cb))


def test_slabs_restart_from_populations(bflbm, oracle_mod):
    """LBM_init (restart) on slabs: ghosted upload + halo refresh, then parity with the oracle."""
    from bflbm_b200.distributed import EmulatedSlabs
    shape = (16, 8, 18)
    prm = dict(kBT=0.0, tau_f=0.5, tau_g=0.5, alpha0=1.5, alpha1=0.0, kappa=0.1, rho_lo=0.1, rho_hi=3.0)
    O = oracle_mod.PortOracle(*shape)
    O.set_params(**prm)
    O.init_stripe(0.5)
    O.step(2)
    f, g = O.populations()
    S = EmulatedSlabs(*shape, 3, params=bflbm.Params(**prm))
    try:
        S.init_from_global_populations(f, g)
        assert_hydro_close(S.gather("hydrovars"), O.hydrovars(), TOL, "slab restart")
        S.step(3)
        O.step(3)
        assert_hydro_close(S.gather("hydrovars"), O.hydrovars(), 10 * TOL, "slab restart + 3 steps")
    finally:
        S.close()


@pytest.mark.parametrize("kbt", [0.0, 1e-5])
def test_rate1_fast_path_matches_general_path(bflbm, monkeypatch, kbt):
    """tau_f = tau_g = 1/2 selects the rate == 1 specialisation (non-conserved input moments never read);
    BFLBM_RATE1=0 forces the general relaxation code on the same input: same result to rounding."""
    shape = (24, 20, 16)
    prm = dict(kBT=kbt, alpha0=1.5, kappa=0.1, rho_lo=0.1, rho_hi=3.0, tau_f=0.5, tau_g=0.5)
    with make_lattice(bflbm, shape, prm, "fused") as A:
        monkeypatch.setenv("BFLBM_RATE1", "0")
        with make_lattice(bflbm, shape, prm, "fused") as B:
            monkeypatch.delenv("BFLBM_RATE1")
            A.init_droplet(0.3)
            B.init_droplet(0.3)
            A.step(5)
            B.step(5)
            fa, ga = A.populations()
            fb, gb = B.populations()
            assert np.abs(fa - fb).max() <= 1e-14 * np.abs(fb).max()
            assert np.abs(ga - gb).max() <= 1e-14 * np.abs(gb).max()


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("kbt", [0.0, 1e-5])
@pytest.mark.parametrize("shape", [(32, 32, 32), (8, 40, 24)])
def test_graph_replayed_steps_bitwise_equal_to_single_launches(bflbm, monkeypatch, kbt, shape, algo):
    """bflbm_step(n) replays CUDA graphs of 64 / 16 / 4 / 2 steps once the lattice is in steady state (small boxes are
    launch bound: Parameters:1-37).  The noise key's step counter then comes from device memory.  Same bits as stepping
    with plain launches (BFLBM_GRAPH=0), whatever the split into calls."""
    prm = dict(kBT=kbt, alpha0=1.5, kappa=0.1, rho_lo=0.1, rho_hi=3.0, tau_f=0.5, tau_g=0.5, seed=2024)
    monkeypatch.setenv("BFLBM_GRAPH", "0")
    with make_lattice(bflbm, shape, prm, algo) as P:
        monkeypatch.delenv("BFLBM_GRAPH")
        with make_lattice(bflbm, shape, prm, algo) as Gr:
            P.init_droplet(0.3)
            Gr.init_droplet(0.3)
            l0 = Gr.kernel_launches
            for n in (1, 87, 2, 5, 64):  # 87 = 64 + 16 + 4 + 2 + 1
                P.step(n)
                Gr.step(n)
                fp, gp = P.populations()
                fg, gg = Gr.populations()
                assert np.array_equal(fp, fg) and np.array_equal(gp, gg), f"graph replay differs after a call of {n} steps"
                assert np.array_equal(P.hydrovars(), Gr.hydrovars())
            assert Gr.step_count == P.step_count == 159
            # changing kBT re-captures (the parameters are by-value arguments of the captured launches)
            P.set_params(kBT=2e-5)
            Gr.set_params(kBT=2e-5)
            P.step(20)
            Gr.step(20)
            assert np.array_equal(P.hydrovars(), Gr.hydrovars())
            assert Gr.kernel_launches > l0


def test_droplet_covariance_matches_numpy(bflbm):
    """fittingDropletCovariance (LBM_hydrovs.H:258-335) on the device: second moments by reduction, eigenvalues in
    closed form, against numpy on the downloaded density (positions are cell indices; the covariance is shift invariant)."""
    import stats
    shape = (28, 20, 24)
    prm = dict(kBT=0.0, alpha0=1.5, kappa=0.1, rho_lo=0.1, rho_hi=3.0)
    with make_lattice(bflbm, shape, prm) as lat:
        lat.init_droplet(0.25)
        lat.step(7)
        rho = lat.hydrovars_bar()[0]
        com, cov, eig = lat.droplet_covariance()
    nz, ny, nx = rho.shape
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    pos = np.stack([x, y, z]).astype(float)
    m = rho.sum()
    c = (pos * rho).sum(axis=(1, 2, 3)) / m
    d = pos - c[:, None, None, None]
    want = np.einsum("aijk,bijk,ijk->ab", d, d, rho) / m
    assert np.allclose(com, c, rtol=1e-12)
    assert np.abs(cov - want).max() <= 1e-10 * np.abs(want).max()
    assert np.allclose(eig, np.linalg.eigvalsh(want), rtol=1e-9)
    lam = eig
    axes = np.array([(lam[i] ** 2 / (lam[(i + 1) % 3] * lam[(i + 2) % 3])) ** (1 / 6) for i in range(3)])
    assert np.allclose(axes, stats.droplet_axes(rho), rtol=1e-9)


@pytest.mark.parametrize("shape", [(1, 1, 2), (1, 7, 5), (3, 1, 4), (2, 2, 2), (33, 9, 3), (9, 17, 34)])
@pytest.mark.parametrize("kbt", [0.0, 1e-5])
def test_degenerate_and_ragged_boxes_vs_oracle(bflbm, oracle_mod, shape, kbt):
    """Boxes thinner than the stencil, of width 1, odd, and not multiples of the 32 x 8 tile (partial bricks, bricks that
    are their own periodic neighbours, no fold-in-staging): three steps against the CPU oracle, with injected normals."""
    prm = dict(kBT=kbt, tau_f=0.6, tau_g=0.5, alpha0=1.5, alpha1=0.0, kappa=0.1, rho_lo=0.1, rho_hi=3.0)
    with make_lattice(bflbm, shape, prm, seed=3) as lat:
        lat.init_stripe(0.5)
        f0, g0 = lat.populations()
        O = oracle_mod.PortOracle(*shape)
        O.set_params(**prm)
        if kbt > 0:
            O.set_normals(lat.normals())
        O.init_from_populations(f0, g0)
        for s in range(3):
            lat.step(1)
            if kbt > 0:
                O.set_normals(lat.normals())
            O.step(1)
            fl, gl = lat.populations()
            fo, go = O.populations()
            assert rel_err(fl, fo) <= 10 * TOL and rel_err(gl, go) <= 10 * TOL, f"step {s}"
            assert_hydro_close(lat.hydrovars(), O.hydrovars(), 10 * TOL, f"step {s}")


# The checker on the GPU box is the C port; its pin to the reference's own headers (oracle/_ref, prebuilt, travels with the
# snapshot) is re-run there too, so that the `-m gpu` log carries it next to the parity results.
@pytest.mark.gpu
def test_checker_is_pinned_to_reference_headers_on_this_box(oracle_mod):
    import test_oracle
    test_oracle.test_port_matches_reference_headers_bitwise(oracle_mod)
    test_oracle.test_port_reference_state_noise_matches_ref_build_bitwise(oracle_mod)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [1, 2])
def test_profiled_steps_equal_plain_steps(bflbm, mode):
    """bflbm_set_profiling: CUDA events around every launch (mode 2: read back on demand, 130 steps wrap the event pool once).
    Timing must not change a bit of the result, and every step must be accounted for."""
    shape = (24, 20, 16)
    prm = dict(kBT=1e-5, alpha0=1.5, kappa=0.1, rho_lo=0.1, rho_hi=3.0, seed=11)
    nsteps = 130
    with make_lattice(bflbm, shape, prm, "fused") as A, make_lattice(bflbm, shape, prm, "fused") as B:
        A.init_droplet(0.3)
        B.init_droplet(0.3)
        A.step(nsteps)
        B.set_profiling(mode)
        B.step(nsteps)
        ms, n = B.profile()
        B.set_profiling(0)
        assert n == nsteps
        assert ms[0] > 0 and ms[1] >= 0
        assert np.array_equal(A.hydrovars(), B.hydrovars())
        B.step(3)  # back to graph replay
        A.step(3)
        assert np.array_equal(A.hydrovars(), B.hydrovars())

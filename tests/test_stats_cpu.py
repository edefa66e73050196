"""CPU checks of the analysis recipes (tests/stats.py) and of the committed reference statistics (tests/golden/stats_*.json,
produced from oracle/_ref by tests/golden/make_stats_golden.py): the REFERENCE code's own fluctuations satisfy the
closed-form expectations its notebooks use, which pins the estimators before they are applied to the GPU output."""
import json
import os

import numpy as np

import stats

HERE = os.path.dirname(os.path.abspath(__file__))


def _gold(name):
    with open(os.path.join(HERE, "golden", f"stats_{name}.json")) as fh:
        return json.load(fh)


def test_structure_factor_of_white_noise_is_flat_and_normalised():
    rng = np.random.default_rng(1)
    sf = stats.StructureFactor([(0, 0), (1, 1), (0, 1)])
    for _ in range(40):
        sf.add(rng.standard_normal((2, 8, 10, 12)))
    s = sf.result()
    assert s.shape == (3, 8, 10, 12) and s[0, 4, 5, 6] == 0.0  # k = 0 bin zeroed after the shift (AMReX_DFT.H:170-183)
    _, shells = sf.shell_means(4)
    assert np.abs(shells[:2] - 1).max() < 0.1 and np.abs(shells[2]).max() < 0.1


def test_interface_height_of_a_tanh_stripe():
    nz, ny, nx = 48, 6, 4
    z = np.arange(nz)[:, None, None]
    top = 33.3 + 0.25 * np.sin(2 * np.pi * np.arange(ny) / ny)[None, :, None] + np.zeros((1, 1, nx))
    rho = 0.1 + 2.9 * 0.5 * (np.tanh((z - 12.2) / 1.5) + np.tanh((top - z) / 1.5))
    h = stats.interface_height(rho, 1.55)
    assert h.shape == (ny, nx)
    assert np.abs(h - top[0]).max() < 0.02  # linear interpolation of a tanh of width 1.5: a few 1e-3 cells


def test_capillary_spectrum_recovers_planted_modes():
    rng = np.random.default_rng(2)
    ny, nx, frames, kT, gamma = 32, 2, 4000, 1e-5, 0.0123
    k = 2 * np.pi * np.fft.fftfreq(ny)
    amp = np.zeros(ny)
    amp[1:] = np.sqrt(kT * ny / (gamma * nx * k[1:] ** 2))  # thin-sheet equipartition, 'backward' norm
    hk = amp[None, :] * (rng.standard_normal((frames, ny)) + 1j * rng.standard_normal((frames, ny))) / np.sqrt(2)
    h = np.fft.ifft(hk, axis=1).real * np.sqrt(2)  # real field with the same <|h_k|^2>
    hs = np.repeat(h[:, :, None], nx, axis=2) + 30.0
    kk, p = stats.capillary_spectrum(hs)
    g = stats.surface_tension_from_spectrum(kk, p, kT, ny, nx, kmax=1.0)
    assert abs(g / gamma - 1) < 0.1


def test_droplet_axes_of_an_ellipsoid():
    n = 40
    z, y, x = np.meshgrid(*(np.arange(n) + 0.5 - n / 2,) * 3, indexing="ij")
    sphere = ((x ** 2 + y ** 2 + z ** 2) < 9.0 ** 2).astype(float)
    assert np.abs(stats.droplet_axes(sphere) - 1).max() < 2e-3
    ell = (((x / 7) ** 2 + (y / 9) ** 2 + (z / 11) ** 2) < 1).astype(float)
    want = np.array([7.0, 9.0, 11.0]) / (7.0 * 9.0 * 11.0) ** (1 / 3)
    assert np.abs(stats.droplet_axes(ell) - want).max() < 0.01  # voxelised ellipsoid
    frames = 1.0 + 1e-3 * np.array([[1, -1, 0], [-1, 1, 0], [0, 1, -1], [0, -1, 1.0]])
    plus, minus = stats.shape_mode_variances(frames)
    assert abs(plus - 2e-6) < 1e-12 and abs(minus - 6e-6) < 1e-12  # per frame: (0, 1, -1)^2 and (2, 1, -1)^2 x 1e-6


def test_reference_statistics_satisfy_the_notebooks_expectations():
    m = _gold("mixture")
    n = np.prod(m["case"]["shape"])
    eq = m["equipartition"]
    for key, want in {"rho": 1 - 1 / n, "phi": 1 - 1 / n, "ub": 1 - 1 / n, "uf_real": 0.75, "ug_real": 0.75, "ufug_real": 0.25,
                      "ufbar": 1.0, "ugbar": 1.0, "xibar_f": 0.5, "rho_phi": 0.0}.items():
        assert abs(eq[key] - want) < 0.01, (key, eq[key], want)
    s = np.array(m["sf_shells_normalised"])
    assert np.abs(s - np.array([1, 1, 0, 1, 1, 1.0])[:, None]).max() < 0.02
    nz = _gold("noise")
    assert np.abs(np.array(nz["variance_ratio_f"]) - 1).max() < 0.02 and np.abs(np.array(nz["variance_ratio_g"]) - 1).max() < 0.02
    assert nz["max_dev_from_expected_corr"] < 0.02

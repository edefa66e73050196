"""TEST INFRASTRUCTURE -- the statistical analysis recipes of the reference, restated in numpy.

north_star: "With noise on, the check is statistical: equipartition of velocity variance, the noise-covariance
matrix, and the static structure factor S(k) and capillary-wave spectrum must agree within stated tolerances,
using the reference's AMReX_Analysis/AMReX_DFT analysis on both outputs."

What is followed (all `file:line` relative to /root/reference; yt / skimage / FHDeX are not available here):
  * structure factor     AMReX_DFT.H:19-132 (r2c DFT, Hermitian completion, 1/sqrt(N)), :138-183 (fftshift, k=0 zeroed);
                         variable pairs as main_run_job.cpp:301-310; normalisations of Mixture.ipynb cell 2
  * noise covariance     NoiseCovariance.ipynb cell 3: time/ensemble variance of every dumped noise component divided by
                         the amplitude^2 of LBM_binary.H:113-127
  * equipartition        Mixture.ipynb cells 1-2: <d rho^2> cs2/kBT, <u_b^2>(rho+phi)/kBT, LB ("bar") velocities <u^2> rho/kBT
  * droplet shape modes  LBM_hydrovs.H:258-335 (`fittingDropletCovariance`: mass-weighted covariance of rho about the centre of
                         mass, eigenvalues) and Droplet_Fluctuation.ipynb cells 3, 22-25 (principal semi-axes
                         a_i = lambda_i^(1/3) R / (lambda_j lambda_k)^(1/6); sums over i<j of <(da_i +- da_j)^2>)
  * interface height     Flat_Interface.ipynb cell 4 (`ih_direct`): iso-level (rho_lo+rho_hi)/2 of rho along z, UPPER
                         interface, linear interpolation between the two cells that bracket the level
  * capillary spectrum   Flat_Interface.ipynb cells 7, 9: h(y) of one x-slice, ensemble mean per y removed,
                         numpy fft "backward" norm, k_y = 2 pi fftfreq(ny); <|h_k|^2> = kBT / (gamma k^2) L-normalised

Every function works on plain arrays shaped like the library's outputs: (ncomp, nz, ny, nx).
The same functions are applied to the GPU output and to the CPU oracle output (tests/golden/make_stats_golden.py).
"""
from __future__ import annotations

import numpy as np

CS2 = 1.0 / 3.0
# b_k = sum_i w_i e_ki^2, LBM_d3q19.H:56-76
BNORM = np.array([1.0, 1 / 3, 1 / 3, 1 / 3, 2 / 3, 4 / 3, 4 / 9, 1 / 9, 1 / 9, 1 / 9, 2 / 3, 2 / 3, 2 / 3, 2 / 9, 2 / 9, 2 / 9, 2.0, 4 / 3, 4 / 9])


# ------------------------------------------------------------------------------------------ structure factor
def dft_normalised(a: np.ndarray) -> np.ndarray:
    """AMReX_DFT.H:19-132: full complex spectrum of a real 3-D field, scaled by 1/sqrt(N)."""
    return np.fft.fftn(a) / np.sqrt(a.size)


class StructureFactor:
    """Running <A_k B_k^*> for a list of variable pairs (FHDeX StructFact as used at main_run_job.cpp:299-310,
    342-349), convention of AMReX_DFT.H: 1/sqrt(N) per transform, fftshift on output, the k = 0 bin zeroed."""

    def __init__(self, pairs):
        self.pairs = list(pairs)
        self.acc = None
        self.n = 0

    def add(self, fields: np.ndarray):
        """fields: (ncomp, nz, ny, nx) snapshot; the per-snapshot mean is removed like StructFact's zero_avg."""
        comps = sorted({c for p in self.pairs for c in p})
        ft = {c: dft_normalised(fields[c] - fields[c].mean()) for c in comps}
        cur = np.stack([(ft[a] * np.conj(ft[b])).real for a, b in self.pairs])
        self.acc = cur if self.acc is None else self.acc + cur
        self.n += 1

    def result(self) -> np.ndarray:
        s = self.acc / self.n
        s = np.fft.fftshift(s, axes=(1, 2, 3))
        nz, ny, nx = s.shape[1:]
        s[:, nz // 2, ny // 2, nx // 2] = 0.0  # AMReX_DFT.H:170-183
        return s

    def shell_means(self, nbins=6):
        """Mean of S over |k| shells (k = 0 excluded): flatness check of Mixture.ipynb cell 2."""
        s = self.result()
        nz, ny, nx = s.shape[1:]
        kz, ky, kx = np.meshgrid(np.fft.fftshift(np.fft.fftfreq(nz)), np.fft.fftshift(np.fft.fftfreq(ny)),
                                 np.fft.fftshift(np.fft.fftfreq(nx)), indexing="ij")
        k = 2 * np.pi * np.sqrt(kx ** 2 + ky ** 2 + kz ** 2)
        edges = np.linspace(0, k.max() * (1 + 1e-9), nbins + 1)
        out = np.zeros((s.shape[0], nbins))
        for b in range(nbins):
            m = (k > edges[b]) & (k <= edges[b + 1]) & (k > 0)
            out[:, b] = s[:, m].mean(axis=1) if m.any() else np.nan
        return 0.5 * (edges[1:] + edges[:-1]), out


# ------------------------------------------------------------------------------------------ noise covariance
def noise_amplitudes2(rho, phi, kBT, tau_f):
    """Variance the reference gives each noise component (LBM_binary.H:79-82, 113-127), arrays (19, ...) for f and g."""
    lam = 1.0 / (tau_f + 0.5)
    A = 2.0 * (lam - 0.5 * lam * lam)
    vf = np.zeros((19,) + np.shape(rho))
    vg = np.zeros_like(vf)
    vf[1:4] = vg[1:4] = A * kBT * np.abs(rho * phi / (rho + phi))
    for a in range(4, 19):
        vf[a] = A * kBT / CS2 * BNORM[a] * np.abs(rho)
        vg[a] = A * kBT / CS2 * BNORM[a] * np.abs(phi)
    return vf, vg


# ------------------------------------------------------------------------------------------ equipartition
class Equipartition:
    """Real-space fluctuation variances of a homogeneous mixture, normalised like Mixture.ipynb cells 1-2.
    Feed hydrovs snapshots (22, nz, ny, nx); result() returns a dict of dimensionless ratios (1 = equipartition,
    up to the finite-size factor 1 - 1/N for the conserved densities)."""
    KEYS = ["rho", "phi", "rho_phi", "ub", "uf_real", "ug_real", "ufug_real", "ufbar", "ugbar", "xibar_f"]

    def __init__(self, kBT):
        self.kBT = kBT
        self.s = {k: 0.0 for k in self.KEYS}
        self.n = 0

    def add(self, h: np.ndarray):
        kT = self.kBT
        rho, phi = h[0], h[1]
        r0, p0 = rho.mean(), phi.mean()
        self.s["rho"] += ((rho - r0) ** 2).mean() / (kT / CS2 * r0)
        self.s["phi"] += ((phi - p0) ** 2).mean() / (kT / CS2 * p0)
        self.s["rho_phi"] += ((rho - r0) * (phi - p0)).mean() / (kT / CS2 * np.sqrt(r0 * p0))
        self.s["ub"] += (h[15:18] ** 2).mean() * (r0 + p0) / kT
        self.s["uf_real"] += (h[2:5] ** 2).mean() * r0 / kT
        self.s["ug_real"] += (h[6:9] ** 2).mean() * p0 / kT
        self.s["ufug_real"] += (h[2:5] * h[6:9]).mean() * np.sqrt(r0 * p0) / kT
        self.s["ufbar"] += (h[20] ** 2).mean() * r0 / kT
        self.s["ugbar"] += (h[21] ** 2).mean() * p0 / kT
        self.s["xibar_f"] += (h[18] ** 2).mean() * r0 / kT
        self.n += 1

    def result(self):
        return {k: v / self.n for k, v in self.s.items()}


# ------------------------------------------------------------------------------------------ flat interface
def interface_height(rho: np.ndarray, level: float) -> np.ndarray:
    """Flat_Interface.ipynb cell 4: z position of the UPPER interface (rho falling through `level` with increasing z)
    for every (y, x) column, by linear interpolation between the bracketing cells.  rho: (nz, ny, nx) -> (ny, nx)."""
    nz = rho.shape[0]
    above = rho >= level
    fall = above[:-1] & ~above[1:]  # rho(z) >= level > rho(z+1)
    # the upper interface is the LAST falling crossing along z (the stripe sits in the middle of the box)
    idx = (nz - 2) - np.argmax(fall[::-1], axis=0)
    if not fall.any(axis=0).all():
        raise ValueError("no interface found in some column")
    r0 = np.take_along_axis(rho, idx[None], axis=0)[0]
    r1 = np.take_along_axis(rho, idx[None] + 1, axis=0)[0]
    return idx + (r0 - level) / (r0 - r1)


def capillary_spectrum(h_frames: np.ndarray):
    """Flat_Interface.ipynb cells 7, 9.  h_frames: (frames, ny, nx) interface heights.  Every x-slice is treated like
    the notebook's single slice (sliceIdx_x) and the spectra are averaged over the slices: ensemble mean per y removed,
    numpy fft 'backward' norm along y.  Returns (k_y[1:ny//2], <|h_k|^2>[1:ny//2])."""
    h = np.asarray(h_frames)
    ny = h.shape[1]
    dh = h - h.mean(axis=0, keepdims=True)
    hk = np.fft.fft(dh, axis=1, norm="backward")
    p = (np.abs(hk) ** 2).mean(axis=(0, 2))
    k = 2 * np.pi * np.fft.fftfreq(ny)
    return k[1:ny // 2], p[1:ny // 2]


def surface_tension_from_spectrum(k, p, kBT, ny, nx=1, kmax=None):
    """Equipartition of the capillary modes of an nx x ny sheet, (gamma/2) k^2 |h(kx,ky)|^2 / (nx ny) = kBT/2 in numpy's
    'backward' convention, seen through ONE x-slice transformed along y (the notebook's estimator):
        <|h_slice(ky)|^2> = kBT * ny / (gamma * nx) * sum_kx 1 / (kx^2 + ky^2)
    (for a thin sheet the kx = 0 term dominates: kBT ny / (gamma nx ky^2)).  Returns gamma averaged over modes ky <= kmax."""
    m = np.ones_like(k, dtype=bool) if kmax is None else (k <= kmax)
    kx = 2 * np.pi * np.fft.fftfreq(nx)
    s = np.array([np.sum(1.0 / (kx ** 2 + ky ** 2)) for ky in k[m]])
    return float(np.mean(kBT * ny / nx * s / p[m]))


# ------------------------------------------------------------------------------------------ droplet shape modes
def droplet_axes(rho: np.ndarray) -> np.ndarray:
    """LBM_hydrovs.H:258-335: eigenvalues of the mass-weighted covariance matrix of rho about its centre of mass (cell
    centres at i + 1/2, unit cells), turned into RELATIVE principal semi-axes a_i / R = (lambda_i^2 / (lambda_j lambda_k))^(1/6)
    (Droplet_Fluctuation.ipynb cell 3), ascending.  rho: (nz, ny, nx); the droplet must not straddle the periodic boundary."""
    nz, ny, nx = rho.shape
    z, y, x = np.meshgrid(np.arange(nz) + 0.5, np.arange(ny) + 0.5, np.arange(nx) + 0.5, indexing="ij")
    m = rho.sum()
    pos = np.stack([x, y, z])
    com = (pos * rho).sum(axis=(1, 2, 3)) / m
    d = pos - com[:, None, None, None]
    cov = np.einsum("aijk,bijk,ijk->ab", d, d, rho) / m
    lam = np.linalg.eigvalsh(cov)
    return np.array([(lam[i] ** 2 / (lam[(i + 1) % 3] * lam[(i + 2) % 3])) ** (1.0 / 6.0) for i in range(3)])


def shape_mode_variances(axes_frames: np.ndarray):
    """Droplet_Fluctuation.ipynb cells 22-25 on relative axes s_i = a_i / R - 1 (frames, 3): the two sums that enter
    gamma_(2,0) = 15 kBT / (16 pi R^2 sum_{i<j} <(s_i + s_j)^2>) and gamma_(2,+-2) = 45 kBT / (16 pi R^2 sum_{i<j} <(s_i - s_j)^2>)."""
    s = np.asarray(axes_frames) - 1.0
    plus = sum(((s[:, i] + s[:, j]) ** 2).mean() for i in range(3) for j in range(i + 1, 3))
    minus = sum(((s[:, i] - s[:, j]) ** 2).mean() for i in range(3) for j in range(i + 1, 3))
    return float(plus), float(minus)

"""Generates tests/golden/droplet_fit.json from the reference's own externlib.H (oracle/_ref/libbflbm_ref_fit.so):
coefficients of the (W, R) gradient flow at a few (W, R, dt, C0) and the fit of two density fields -- a synthetic tanh droplet and
the droplet of the authors' recipe relaxed for 400 steps by the reference headers (32^3, alpha0 = 1.5, kappa = 0.1, rho in [0.1, 3],
r = 0.3).  Run where /root/reference exists:  python tests/golden/make_fit_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as om  # noqa: E402

COEF_POINTS = [(0.02, 0.3, 0.02, 1.0), (0.1, 0.3, 0.02, 2.9), (3e-4, 0.32, 0.004, 2.9), (0.005, 0.2, 0.02, 1.5), (0.05, 0.45, 0.0008, 3.0)]


def synthetic(n=32, R=0.27, W=0.0011, c=(0.52, 0.49, 0.5)):
    x = (np.arange(n) + 0.5) / n
    Z, Y, X = np.meshgrid(x, x, x, indexing="ij")
    r = np.sqrt((X - c[0]) ** 2 + (Y - c[1]) ** 2 + (Z - c[2]) ** 2)
    return 0.1 + 2.9 * 0.5 * (1 + np.tanh((R - r) / np.sqrt(2 * W)))


def relaxed(n=32):
    f, g = om.droplet_populations(n, n, n, 0.3, 0.1, 0.1, 3.0)
    O = om.RefOracle(n, n, n)
    O.set_params(kBT=0.0, tau_f=0.5, tau_g=0.5, alpha0=1.5, alpha1=0.0, kappa=0.1)
    O.init_from_populations(f, g)
    O.step(400)
    return O.hydrovars()[0]


if __name__ == "__main__":
    om.build()
    F = om.RefFit()
    out = {"generator": "tests/golden/make_fit_golden.py (oracle/_ref/libbflbm_ref_fit.so: the reference's externlib.H)",
           "coefficients": [{"W": W, "R": R, "dt": dt, "C0": C0, "values": F.coefficients(W, R, 0.2, 0.2, dt, C0).tolist()} for W, R, dt, C0 in COEF_POINTS]}
    for name, rho, W0, R0 in (("synthetic", synthetic(), 0.1, 0.3), ("relaxed", relaxed(), 0.1, 0.3)):
        W, R, u, ok, trace = F.fit(rho, W0, R0)
        mw, mr, com = F.field_terms(rho, W0, R0)
        out[name] = {"W0": W0, "R0": R0, "W": W, "R": R, "undulation": u, "converged": ok, "MfW0": mw, "MfR0": mr, "com": com.tolist(),
                     "trace_head": trace[:5].tolist(), "rho_sum": float(rho.sum()), "rho_min": float(rho.min()), "rho_max": float(rho.max())}
        print(name, W, R, u, ok)
    json.dump(out, open(os.path.join(HERE, "droplet_fit.json"), "w"), indent=1)

"""Regenerates tests/golden/*.npz from oracle/_ref: the REFERENCE's own headers (LBM_d3q19.H, LBM_binary.H
read from /root/reference, compiled unchanged over oracle/shim).  Run in the container that has
/root/reference:   python tests/golden/make_golden.py

Each fixture holds the inputs (sizes, parameters, initial populations or analytic init, injected normals)
and the reference's outputs (populations, hydrovs[22], hydrovsbar[9], noise) after the stated steps.
The reference has no golden vectors of its own; these are outputs of the reference itself.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as om  # noqa: E402


def run_case(name, shape, params, init, steps, noise_seed=None, pops=None):
    nx, ny, nz = shape
    R = om.RefOracle(nx, ny, nz)
    R.set_params(**params)
    rng = np.random.default_rng(noise_seed) if noise_seed is not None else None
    normals = []

    def next_normals():
        if rng is None:
            return
        n = rng.standard_normal((nz, ny, nx, 33))
        normals.append(n)
        R.set_normals(n)

    next_normals()
    if init[0] == "stripe":
        R.init_stripe(init[1])
    elif init[0] == "droplet":
        R.init_droplet(init[1])
    elif init[0] == "mixture":
        R.init_mixture()
    else:
        R.init_from_populations(*pops)
    f0, g0 = R.populations()
    out = dict(shape=np.array(shape), steps=np.array(steps), init=np.array(init[0]),
               init_arg=np.array(init[1] if len(init) > 1 else 0.0),
               f0=f0, g0=g0, h0=R.hydrovars(), hb0=R.hydrovars_bar())
    for k, v in params.items():
        out["p_" + k] = np.array(v)
    done = 0
    for s in steps:
        while done < s:
            next_normals()
            R.step(1)
            done += 1
        f, g = R.populations()
        out[f"f_{s}"], out[f"g_{s}"] = f, g
        out[f"h_{s}"], out[f"hb_{s}"] = R.hydrovars(), R.hydrovars_bar()
        fn, gn = R.noise()
        out[f"fn_{s}"], out[f"gn_{s}"] = fn, gn
    if normals:
        out["normals"] = np.stack(normals)  # normals[k] feeds the noise generated after step k-1 (k=0: init)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim > 1})


if __name__ == "__main__":
    assert om.RefOracle.available(), "oracle/_ref not built (needs /root/reference): make -C oracle ref"
    dflt = dict(kBT=0.0, tau_f=0.5, tau_g=0.5, alpha0=4.0, alpha1=0.0, kappa=4.0)
    # shipped defaults, the three inits, non-cubic boxes
    run_case("stripe_default_8x6x12", (8, 6, 12), dflt, ("stripe", 0.5), [1, 5])
    run_case("droplet_default_10x10x10", (10, 10, 10), dflt, ("droplet", 0.3), [1, 5])
    run_case("mixture_default_6x6x6", (6, 6, 6), dflt, ("mixture",), [1, 3])
    # the authors' flat-interface recipe (Parameters:22-30) entered through the restart path, general tau
    f, g = om.stripe_populations(8, 8, 16, 0.5, 0.1, 0.1, 3.0)
    run_case("stripe_recipe_tau_8x8x16", (8, 8, 16), dict(kBT=0.0, tau_f=0.7, tau_g=0.9, alpha0=1.5, alpha1=0.0, kappa=0.1),
             ("restart",), [1, 4], pops=(f, g))
    # noise on, normals injected (33 per cell in the reference's draw order)
    f, g = om.droplet_populations(8, 8, 8, 0.3, 0.1, 0.0, 3.0)
    run_case("droplet_noise_8x8x8", (8, 8, 8), dict(kBT=1e-5, tau_f=0.5, tau_g=0.5, alpha0=1.5, alpha1=0.0, kappa=0.1),
             ("restart",), [1, 3], noise_seed=7, pops=(f, g))
    run_case("mixture_noise_tau_6x8x6", (6, 8, 6), dict(kBT=2e-5, tau_f=0.8, tau_g=0.6, alpha0=0.5, alpha1=0.0, kappa=4.0),
             ("mixture",), [1, 3], noise_seed=11)

#!/usr/bin/env python
"""Ensemble of fluctuating runs from ONE kBT = 0 checkpoint -- the reference's two-stage workflow (relax at kBT = 0, write
f_checkpoint / g_checkpoint, restart with noise: main_run_job.cpp:244-268, 399-409) run as an ensemble with the
asynchronous transfer calls: while run k steps, the checkpoint of run k + 1 travels to the GPU and the last frame of run
k - 1 travels back (include/bflbm.h, "asynchronous host transfers").

    python examples/ensemble_restarts.py [--n 32] [--runs 4] [--steps 400] [--relax 2000]

Prints, per run, the interface-weighted centre of mass and the mean density of the droplet phase; the runs differ only by
their noise seed.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bflbm_b200 as b  # noqa: E402


def pinned_like(a):
    """A pinned host copy (the copies overlap the steps only from pinned memory); plain numpy if torch cannot pin."""
    try:
        import torch
        t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True).numpy()
    except Exception:
        t = np.empty_like(a)
    t[...] = a
    return t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=32)
    ap.add_argument("--runs", type=int, default=4)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--relax", type=int, default=2000)
    a = ap.parse_args()
    n = a.n
    prm = dict(alpha0=1.5, kappa=0.1, rho_lo=0.1, rho_hi=3.0, tau_f=0.5, tau_g=0.5)  # the authors' recipe, Parameters:22-30

    # stage 1: deterministic relaxation, checkpoint to the host
    with b.Lattice(n, n, n, params=b.Params(kBT=0.0, **prm)) as det:
        det.init_droplet(0.3)
        det.step(a.relax)
        f, g = (pinned_like(x) for x in det.populations())

    # stage 2: the ensemble
    frame = pinned_like(np.zeros((9, n, n, n)))
    z, y, x = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")

    def report(run, fr):
        rho = fr[0]
        m = rho.sum()
        print(f"run {run}: mass {m:.9f}  centre of mass ({(rho * x).sum() / m:.4f}, {(rho * y).sum() / m:.4f}, {(rho * z).sum() / m:.4f})"
              f"  max rho {rho.max():.6f}")

    with b.Lattice(n, n, n, params=b.Params(kBT=1e-5, seed=1000, **prm)) as lat:
        lat.stage_populations(f, g)                   # run 0: nothing to hide behind
        for run in range(a.runs):
            lat.set_params(seed=1000 + run)
            lat.init_from_staged()                    # = init_from_populations(f, g), queued behind the copy
            if run + 1 < a.runs:
                lat.stage_populations(f, g)           # the next run's checkpoint travels during this run's steps
            lat.step(a.steps)
            if run > 0:
                lat.download_wait()                   # frame of run - 1: arrived while this run was stepping
                report(run - 1, frame)
            lat.hydrovars_bar_async(frame)            # travels during the next run
        lat.download_wait()
        report(a.runs - 1, frame)
        assert lat.check_nan() == 0


if __name__ == "__main__":
    main()

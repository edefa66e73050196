"""Host driver (C++): Parameters parser + plotfile IO on the CPU; the full job (init, steps, plotfiles,
checkpoint, restart, equilibrium extraction) on the GPU against the oracle."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_cpp_parser_and_plotfile(tmp_path):
    exe = tmp_path / "host_cpp_test"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-std=c++17", "-O1", "-pthread", os.path.join(ROOT, "tests", "host_cpp_test.cpp"), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe), str(tmp_path)], check=True, capture_output=True, text=True).stdout
    assert "host_cpp_test ok" in out


def test_driver_builds_and_reports_usage(bflbm):
    from bflbm_b200 import host_driver
    exe = host_driver.build()
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr


def read_plotfile(path):
    hdr = open(os.path.join(path, "Header")).read().split("\n")
    ncomp = int(hdr[1])
    names = hdr[2:2 + ncomp]
    raw = open(os.path.join(path, "Level_0", "Cell_D_00000"), "rb").read()
    nl = raw.index(b"\n")
    fab = raw[:nl].decode()
    hi = fab.split("((")[2].split(")")[1].strip(" (").split(",")
    nx, ny, nz = [int(v) + 1 for v in hi]
    data = np.frombuffer(raw[nl + 1:], dtype="<f8").reshape(ncomp, nz, ny, nx)
    return names, data


@pytest.mark.gpu
def test_driver_job_matches_oracle(tmp_path, bflbm, oracle_mod):
    from bflbm_b200 import host_driver
    exe = host_driver.build()
    prm = tmp_path / "Parameters"
    prm.write_text(f"""
system = flat_interface
nx = 8
ny = 12
nz = 16
kBT = 0.
alpha0 = 1.5
kappa = 0.1
rho_lo = 0.1
rho_hi = 3.
nsteps = 20
plot_int = 5
print_int = 10
t_window = 10
root_path = {tmp_path}
plot_fields = hydrovars
""")
    r = subprocess.run([exe, str(prm)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    assert "Run time =" in r.stdout
    base = tmp_path / "data_interface_alpha0_1.50"
    run = base / "lbm_data_shshan_alpha0_1.50_xi_0.0e+00_size8-12-16"
    O = oracle_mod.PortOracle(8, 12, 16)
    O.set_params(kBT=0.0, alpha0=1.5, kappa=0.1, rho_lo=0.1, rho_hi=3.0)
    O.init_stripe(0.5)
    from conftest import assert_hydro_close
    for step in (0, 5, 10, 15, 20):
        O.step(step - (0 if step == 0 else step - 5))
        names, h = read_plotfile(str(run / f"plt{step:07d}"))
        assert names[:6] == ["rho", "phi", "ufx", "ufy", "ufz", "p_bulk"] and len(names) == 22
        assert_hydro_close(h, O.hydrovars(), 1e-11, f"plotfile step {step}")
    # checkpoint = fold/gold of the last step, 19 components each
    _, f = read_plotfile(str(base / "f_checkpoint0000020_alpha0_1.50_xi_0.0e+00_size8-12-16"))
    fo, go = O.populations()
    assert np.abs(f - fo).max() <= 1e-12 * np.abs(fo).max()
    # equilibrium extraction: mean over frames 10..20
    _, rho_eq = read_plotfile(str(base / "equilibrium_rho_alpha0_1.50_size8-12-16"))
    assert rho_eq.shape == (1, 16, 12, 8)
    assert "convergence L1 (rho):" in r.stdout and "equilibrium state: mean of 3 frames" in r.stdout  # PrintConvergence, Debug.H:275-358
    # restart from the checkpoint with noise on (two-stage workflow of ReadMe.ipynb cells 1-3)
    prm2 = tmp_path / "Parameters2"
    prm2.write_text(prm.read_text() + "\nkBT = 1e-6\nstep_continue = 20\nif_continue_from_last_frame = true\nnsteps = 10\nplot_int = 5\nt_window = 5\n"
                    .replace("kBT = 0.\n", ""))
    txt = prm.read_text().replace("kBT = 0.", "kBT = 1e-6").replace("nsteps = 20", "nsteps = 10") + \
        "step_continue = 20\nif_continue_from_last_frame = true\nplot_SF_window = 10\nout_SF_step = 5\n"
    prm2.write_text(txt)
    r2 = subprocess.run([exe, str(prm2)], capture_output=True, text=True)
    assert r2.returncode == 0, r2.stderr + r2.stdout
    run2 = base / "lbm_data_shshan_alpha0_1.50_xi_1.0e-06_size8-12-16_continue"
    names, h30 = read_plotfile(str(run2 / "plt0000030"))
    assert np.isfinite(h30).all() and abs(h30[0].sum() - O.hydrovars()[0].sum()) < 1e-9 * h30[0].sum()
    # structure factors of the window [20, 30], every 5 steps (StructFact::WritePlotFile, main_run_job.cpp:50-54): 22 pairs on
    # the shifted k grid, k = 0 zeroed, auto-correlations non-negative and inversion symmetric
    assert "structure factor: 2 samples written" in r2.stdout
    names, sf = read_plotfile(str(run2) + "/plt_SF0000030")
    assert sf.shape == (22, 16, 12, 8) and names[0] == "struct_fact_real_rho_rho" and np.isfinite(sf).all()
    assert sf[0, 8, 6, 4] == 0.0 and sf[0].min() >= 0.0 and sf[0].max() > 0.0
    inv = np.roll(sf[0][::-1, ::-1, ::-1], 1, axis=(0, 1, 2))  # k -> -k on the shifted grid (even sizes)
    assert np.allclose(inv, sf[0], rtol=1e-9, atol=1e-30)

"""The real multi-GPU paths (skipped with fewer GPUs than ranks; the single-GPU emulation of the same library path is
tests/test_gpu_parity.py::test_slabs_bitwise_equal_to_whole_box):
  * one process per GPU: NCCL send/recv of the halo messages, and peer mode (CUDA-IPC-mapped mailboxes, no collective per step);
  * one process, several devices: bflbm_multi, and the C++ driver with ngpus = N."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("mode", ["nccl", "peer"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_slabs_over_processes_bitwise_equal_to_whole_box(world, mode):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + world + (20 if mode == "peer" else 0)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "mp_slab_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, MP_PEER="1" if mode == "peer" else "0"))
    if os.environ.get("BFLBM_MP_LOG"):
        with open(os.environ["BFLBM_MP_LOG"] + ".full", "a") as fh:
            fh.write(f"==== world {world} mode {mode} rc {r.returncode}\n" + r.stdout[-3000:] + r.stderr[-3000:])
    assert r.returncode == 0 and f"MP_SLAB_OK {world} {mode}" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    log = os.environ.get("BFLBM_MP_LOG")  # keep the evidence (profiles/)
    if log:
        with open(log, "a") as fh:
            fh.write([ln for ln in r.stdout.splitlines() if "MP_SLAB" in ln][-1] + "\n")


@pytest.mark.parametrize("ngpus", [2, 4, 8])
def test_multi_lattice_bitwise_equal_to_whole_box(bflbm, ngpus):
    """bflbm_multi: one process, N devices, peer-to-peer ghost exchange.  Same brick height => same bits as one GPU, with noise,
    general relaxation times, through restart and getters of the whole-box host arrays."""
    if _ngpus() < ngpus:
        pytest.skip(f"needs {ngpus} GPUs")
    nx, ny, nz, lz = 40, 24, 24 * ngpus, 4
    prm = bflbm.Params(kBT=1e-5, alpha0=1.5, kappa=0.1, rho_lo=0.1, rho_hi=3.0, tau_f=0.7, tau_g=0.55, seed=4711)
    with bflbm.Lattice(nx, ny, nz, params=prm) as W, bflbm.MultiLattice(nx, ny, nz, params=prm, ngpus=ngpus, brick_lz=lz) as M:
        W.set_tiling(lz)
        W.init_droplet(0.3)
        M.init_droplet(0.3)
        for _ in range(3):
            W.step(4)
            M.step(4)
            assert np.array_equal(W.hydrovars(), M.hydrovars())
        fw, gw = W.populations()
        fm, gm = M.populations()
        assert np.array_equal(fw, fm) and np.array_equal(gw, gm)
        W.init_from_populations(fw, gw)
        M.init_from_populations(fw, gw)
        W.step(3)
        M.step(3)
        # a restart rebuilds the densities of the slab-face planes as (local + remote): rounding-level dependence on the cut
        hw, hm = W.hydrovars(), M.hydrovars()
        assert np.abs(hw - hm).max() <= 1e-12 * np.abs(hw).max()
        assert np.abs(W.noise()[0] - M.noise()[0]).max() <= 1e-12 * np.abs(W.noise()[0]).max()
        assert M.check_nan() == 0
        assert np.allclose(W.total_mass(), M.total_mass(), rtol=1e-13)
        assert np.allclose(W.droplet_covariance()[2], M.droplet_covariance()[2], rtol=1e-10)


@pytest.mark.parametrize("ngpus", [2, 4])
def test_structure_factor_on_slabs_equals_whole_box(bflbm, ngpus):
    """BASELINE configs[3] analyses S(k) on the decomposition: the accumulator assembles the slabs' hydro fields on one GPU by peer
    copies and transforms there (main_run_job.cpp:342-349 runs FortStructure on the distributed MultiFab).  Same fields => same
    spectra as the one-GPU accumulator, bit for bit."""
    if _ngpus() < ngpus:
        pytest.skip(f"needs {ngpus} GPUs")
    nx, ny, nz, lz = 16, 24, 8 * ngpus, 4
    prm = bflbm.Params(kBT=1e-5, alpha0=0.0, seed=12)
    pairs = [(0, 0), (1, 1), (0, 1), (15, 16), (2, 6)]
    with bflbm.Lattice(nx, ny, nz, params=prm) as W, bflbm.MultiLattice(nx, ny, nz, params=prm, ngpus=ngpus, brick_lz=lz) as M:
        W.set_tiling(lz)
        W.init_mixture()
        M.init_mixture()
        with bflbm.StructureFactor(W, pairs) as SW, bflbm.StructureFactor(M, pairs) as SM:
            for _ in range(4):
                W.step(5)
                M.step(5)
                SW.fort_structure()
                SM.fort_structure()
            a, b = SW.result(imag=True), SM.result(imag=True)
            assert SM.samples == 4 and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def _read_plotfile(path):
    hdr = open(os.path.join(path, "Header")).read().split("\n")
    ncomp = int(hdr[1])
    raw = open(os.path.join(path, "Level_0", "Cell_D_00000"), "rb").read()
    return ncomp, raw[raw.index(b"\n") + 1:]


@pytest.mark.parametrize("ngpus", [2, 8])
def test_driver_on_n_gpus_writes_the_same_plotfiles(bflbm, tmp_path, ngpus):
    """The C++ driver (csrc/host/main_run_job.cpp) with ngpus = N against ngpus = 1: plotfiles and checkpoints bit-identical
    (fluctuating droplet job, same brick height)."""
    if _ngpus() < ngpus:
        pytest.skip(f"needs {ngpus} GPUs")
    from bflbm_b200 import host_driver
    exe = host_driver.build()
    outs = []
    for n in (1, ngpus):
        root = tmp_path / f"run{n}"
        root.mkdir()
        prm = root / "Parameters"
        prm.write_text(f"""
system = droplet
nx = 32
ny = 24
nz = {16 * ngpus}
kBT = 1e-5
alpha0 = 1.5
kappa = 0.1
rho_lo = 0.1
rho_hi = 3.
radius = 0.3
nsteps = 30
plot_int = 10
print_int = 10
out_step = 0
root_path = {root}
plot_fields = hydrovars
ngpus = {n}
brick_lz = 4
""")
        r = subprocess.run([exe, str(prm)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr + r.stdout
        outs.append(root)
    rel = []
    for dirpath, _, files in os.walk(outs[0]):
        for fn in files:
            if fn.startswith("Cell_D"):
                rel.append(os.path.relpath(os.path.join(dirpath, fn), outs[0]))
    assert len(rel) >= 6  # plt 0/10/20/30 + f/g checkpoints
    for p in rel:
        a = open(os.path.join(outs[0], p), "rb").read()
        b = open(os.path.join(outs[1], p), "rb").read()
        assert a == b, f"{p} differs between 1 and {ngpus} GPUs"

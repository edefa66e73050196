"""Worker of tests/test_gpu_multiprocess.py: one rank per GPU (torchrun), NCCL ring exchange between real processes.
Every rank also steps the WHOLE box on its own GPU and compares its slab bit for bit (results must not
depend on the number of GPUs; the noise is keyed by the global cell index)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bflbm_b200 as b  # noqa: E402
from bflbm_b200.distributed import SlabLattice  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    nx, ny, nz, lz = 40, 24, 24 * world, 4
    prm = b.Params(kBT=1e-5, alpha0=1.5, kappa=0.1, rho_lo=0.1, rho_hi=3.0, tau_f=0.7, tau_g=0.55, seed=4711)
    peer = os.environ.get("MP_PEER", "0") == "1"  # halo messages as peer-to-peer stores through CUDA-IPC-mapped mailboxes
    S = SlabLattice(nx, ny, nz, params=prm, device=local, peer=peer)
    S.lat.set_tiling(lz)
    S.init_droplet(0.3)
    with b.Lattice(nx, ny, nz, params=prm, device=local) as whole:
        whole.set_tiling(lz)
        whole.init_droplet(0.3)
        ok = True
        for _ in range(3):
            S.step(4)
            whole.step(4)
            fw, gw = whole.populations()
            fs, gs = S.lat.populations()
            sl = slice(S.z0, S.z0 + S.nzl)
            ok &= np.array_equal(fw[:, sl], fs) and np.array_equal(gw[:, sl], gs)
            ok &= np.array_equal(whole.hydrovars()[:, sl], S.lat.hydrovars())
        # restart through the whole-box host arrays (every rank points at the same global array), then step again
        S.lat.init_from_global_populations(fw, gw)
        if peer:
            b.lattice._check(S.lat.lib.bflbm_halo_refresh_end(S.lat.h))
        else:
            b.lattice._check(S.lat.lib.bflbm_halo_refresh_begin(S.lat.h))
            S._exchange()
            b.lattice._check(S.lat.lib.bflbm_halo_refresh_end(S.lat.h))
        whole.init_from_populations(fw, gw)
        S.step(3)
        whole.step(3)
        # a restart rebuilds the densities of the slab-face planes as (local + remote): rounding-level dependence on the cut
        hw, hs = whole.hydrovars()[:, sl], S.lat.hydrovars()
        ok &= bool(np.abs(hw - hs).max() <= 1e-12 * np.abs(hw).max())
        if peer:
            ok &= S.lat.halo_error() == 0
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    S.lat.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MP_SLAB_OK" if int(t.item()) == 1 else "MP_SLAB_MISMATCH", world, "peer" if peer else "nccl", flush=True)
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()

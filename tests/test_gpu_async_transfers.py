"""Asynchronous host transfers (include/bflbm.h, "asynchronous host transfers"): a checkpoint staged next to a stepping
lattice and a frame downloaded next to the following steps must give exactly what the synchronous calls give --
bflbm_init_from_populations[_slab] (LBM_init, LBM_binary.H:631-661) and bflbm_get_hydrovars[_bar]."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PRM = dict(kBT=1e-5, alpha0=1.5, kappa=0.1, rho_lo=0.1, rho_hi=3.0, tau_f=0.5, tau_g=0.5, seed=77)


def pinned(shape):
    import torch
    return torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()


def checkpoint(bflbm, shape, steps, algo):
    nx, ny, nz = shape
    with bflbm.Lattice(nx, ny, nz, params=bflbm.Params(**PRM)) as lat:
        lat.set_algorithm(algo)
        lat.init_droplet(0.3)
        lat.step(steps)
        return lat.populations()


@pytest.mark.parametrize("algo", ["fused", "twopass"])
@pytest.mark.parametrize("shape", [(24, 20, 16), (40, 16, 33)])
def test_staged_restart_and_async_frames_equal_the_synchronous_calls(bflbm, shape, algo):
    nx, ny, nz = shape
    ck = [checkpoint(bflbm, shape, s, algo) for s in (5, 9)]
    P = bflbm.Params(**PRM)
    # synchronous route: upload, 7 steps, frame, 3 steps, populations -- for both checkpoints
    want = []
    with bflbm.Lattice(nx, ny, nz, params=P) as A:
        A.set_algorithm(algo)
        for f, g in ck:
            A.init_from_populations(f, g)
            A.step(7)
            hb, h = A.hydrovars_bar(), A.hydrovars()
            A.step(3)
            want.append((hb, h, A.populations()))
    # asynchronous route: every transfer has steps of the same lattice queued next to it
    host = [(pinned(f.shape), pinned(g.shape)) for f, g in ck]
    for (pf, pg), (f, g) in zip(host, ck):
        pf[...] = f
        pg[...] = g
    out_hb, out_h = pinned((9, nz, ny, nx)), pinned((22, nz, ny, nx))
    with bflbm.Lattice(nx, ny, nz, params=P) as B:
        B.set_algorithm(algo)
        B.init_mixture()
        B.stage_populations(*host[0])
        B.step(6)                          # runs while checkpoint 0 travels
        for k in range(2):
            B.init_from_staged()
            if k == 0:
                B.stage_populations(*host[1])  # the next checkpoint travels during this interval
            B.step(7)
            B.hydrovars_bar_async(out_hb)
            B.step(3)                      # overlaps the download
            B.download_wait()
            assert np.array_equal(out_hb, want[k][0])
            fb, gb = B.populations()
            assert np.array_equal(fb, want[k][2][0]) and np.array_equal(gb, want[k][2][1])
        # 22-component frame, and a second download requested while the first is in flight
        B.init_from_populations(*ck[1])
        B.step(7)
        B.hydrovars_async(out_h)
        B.hydrovars_bar_async(out_hb)      # waits for the first on the device, then reuses the buffer
        B.download_wait()
        assert np.array_equal(out_hb, want[1][0])
        B.sync()
        assert np.array_equal(out_h, want[1][1])
        B.stage_wait()
        B.release_staging()
        B.step(3)
        fb, gb = B.populations()
        assert np.array_equal(fb, want[1][2][0]) and np.array_equal(gb, want[1][2][1])


def test_staged_ghosted_checkpoint_equals_the_slab_restart(bflbm):
    """ghosted = 1 on slabs: the staged route against bflbm_init_from_populations_slab, two slabs on one GPU."""
    from bflbm_b200.distributed import EmulatedSlabs
    from bflbm_b200.lattice import _check
    shape = (24, 20, 16)
    nx, ny, nz = shape
    f, g = checkpoint(bflbm, shape, 6, "fused")
    P = bflbm.Params(**PRM)
    res = []
    for staged in (False, True):
        S = EmulatedSlabs(nx, ny, nz, 2, params=P, brick_lz=4)
        try:
            if not staged:
                S.init_from_global_populations(f, g)
            else:
                keep = []
                for lat, (z0, nzl) in zip(S.lats, S.bounds):
                    idx = [(z % nz) for z in range(z0 - 1, z0 + nzl + 1)]
                    fg = (np.ascontiguousarray(f[:, idx]), np.ascontiguousarray(g[:, idx]))
                    keep.append(fg)
                    lat.stage_populations(*fg, ghosted=True)
                for lat in S.lats:
                    lat.init_from_staged()
                for lat in S.lats:
                    _check(lat.lib.bflbm_halo_refresh_begin(lat.h))
                S._exchange()
                for lat in S.lats:
                    _check(lat.lib.bflbm_halo_refresh_end(lat.h))
            S.step(5)
            res.append(S.gather("hydrovars"))
        finally:
            S.close()
    assert np.array_equal(res[0], res[1])


def test_async_transfer_errors(bflbm):
    f, g = checkpoint(bflbm, (8, 8, 8), 1, "fused")
    with bflbm.Lattice(8, 8, 8, params=bflbm.Params(**PRM)) as lat:
        with pytest.raises(bflbm.BflbmError):
            lat.init_from_staged()                       # nothing staged
        lat.stage_populations(f, g)
        with pytest.raises(bflbm.BflbmError):
            lat.stage_populations(f, g)                  # one checkpoint at a time
        with pytest.raises(ValueError):
            lat.stage_populations(f[:, :4], g[:, :4])    # wrong shape
        with pytest.raises(bflbm.BflbmError):
            lat.hydrovars_bar_async()                    # not initialised yet
        lat.init_from_staged()
        lat.stage_wait()
        lat.step(2)
        assert lat.check_nan() == 0
        lat.download_wait()                              # nothing pending: a no-op
    with bflbm.Lattice(8, 8, 8, params=bflbm.Params(**PRM), slab=(0, 4)) as slab:
        with pytest.raises(bflbm.BflbmError):
            slab.stage_populations(f[:, :4].copy(), g[:, :4].copy())  # a slab stages ghosted arrays

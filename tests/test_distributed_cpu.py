"""CPU tests (gloo, world_size 2 and 3) of the host-side slab logic: decomposition, ring neighbours and the
message routing used for the per-step ghost exchange (send[lo] -> lower neighbour's recv[hi], ...)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_slab_bounds_cover_the_box():
    from bflbm_b200.distributed import slab_bounds, neighbours
    for nz, world in [(512, 8), (64, 3), (33, 4), (4096, 8), (10, 5)]:
        z = 0
        for r in range(world):
            z0, n = slab_bounds(nz, world, r)
            assert z0 == z and n >= 2
            z += n
        assert z == nz
        sizes = [slab_bounds(nz, world, r)[1] for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        slab_bounds(6, 4, 0)
    assert neighbours(4, 0) == (3, 1) and neighbours(4, 3) == (2, 0) and neighbours(1, 0) == (0, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nz, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bflbm_b200.distributed import slab_bounds, exchange_ring
        ny, nx = 3, 4
        # a global periodic field; every rank owns a slab and must end up with the right ghost planes
        glob = np.arange(nz * ny * nx, dtype=np.float64).reshape(nz, ny, nx)
        z0, nzl = slab_bounds(nz, world, rank)
        own = glob[z0:z0 + nzl]
        for _ in range(3):  # repeated exchanges must not cross-talk (tags / ordering with 2 ranks)
            send_lo, send_hi = torch.from_numpy(own[0].copy()), torch.from_numpy(own[-1].copy())
            recv_lo, recv_hi = torch.empty_like(send_lo), torch.empty_like(send_hi)
            exchange_ring(send_lo, send_hi, recv_lo, recv_hi, rank, world)
            assert np.array_equal(recv_lo.numpy(), glob[(z0 - 1) % nz]), "ghost plane below = neighbour's top plane"
            assert np.array_equal(recv_hi.numpy(), glob[(z0 + nzl) % nz]), "ghost plane above = neighbour's bottom plane"
        # peer-mode plumbing: every rank publishes its mailbox handle once and picks its two ring neighbours' (64 opaque bytes)
        from bflbm_b200.distributed import gather_ring_handles, neighbours
        lo, hi = gather_ring_handles(bytes([rank]) * 64, rank, world)
        nlo, nhi = neighbours(world, rank)
        assert lo == bytes([nlo]) * 64 and hi == bytes([nhi]) * 64
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nz", [(2, 8), (3, 11)])
def test_ring_exchange_gloo(world, nz):
    import bflbm_b200  # noqa: F401  (registers the package in the parent so spawn children can import it)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nz, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res


def test_single_rank_exchange_is_self_wrap():
    from bflbm_b200.distributed import exchange_ring, gather_ring_handles
    assert gather_ring_handles(b"x" * 64, 0, 1) == (b"x" * 64, b"x" * 64)
    a, b = torch.arange(4.0), torch.arange(4.0) + 10
    ra, rb = torch.empty(4), torch.empty(4)
    exchange_ring(a, b, ra, rb, 0, 1)
    assert torch.equal(rb, a) and torch.equal(ra, b)

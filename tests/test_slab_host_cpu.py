"""CPU tests (gloo, world_size 2) of the host side of the slab decomposition: slab bounds, particle
partition, and the neighbour exchange protocol (halo swap, migration counts, variable-length rows)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nairn_mpm_fea_b200 import problem
from nairn_mpm_fea_b200.slab import NeighbourExchange, partition_particles, slab_bounds


def test_slab_bounds_cover_grid():
    for world in (1, 2, 4, 8):
        b = slab_bounds(816, 8, 808, world)
        assert b[0][0] == 0 and b[-1][1] == 816
        for (lo, hi), (lo2, hi2) in zip(b[:-1], b[1:]):
            assert hi == lo2 and hi - lo >= 3
        assert len(b) == world


def test_partition_is_a_partition():
    pr = problem.block3d(ncell=6, margin=2)
    k = (pr.particles["in_elem"] - 1) // (pr.horiz * pr.vert)
    bounds = slab_bounds(pr.depth, int(k.min()), int(k.max()) + 1, 2)
    parts = [partition_particles(pr.particles, pr.horiz, pr.vert, lo, hi) for lo, hi in bounds]
    ids = np.concatenate([p["ids"] for p in parts])
    assert sorted(ids.tolist()) == list(range(pr.nparticles))
    for p, (lo, hi) in zip(parts, bounds):
        kk = (p["in_elem"] - 1) // (pr.horiz * pr.vert)
        assert np.all((kk >= lo) & (kk < hi))
        assert p["pos"].shape == (3, len(p["ids"])) and p["n_nonrigid"] == len(p["ids"])
        assert np.array_equal(p["pos"], pr.particles["pos"][:, p["ids"]])


def test_partition_replicates_rigid_particles():
    """Rigid-BC particles go to every slab; assemble_by_id takes them once."""
    from nairn_mpm_fea_b200.slab import assemble_by_id
    from tests.parity import load_golden
    pr = problem.from_reference_dump(load_golden("block3d_rigid_wall"))
    n, nnr = pr.nparticles, int(pr.particles["n_nonrigid"])
    assert 0 < nnr < n
    k = (pr.particles["in_elem"][:nnr] - 1) // (pr.horiz * pr.vert)
    bounds = slab_bounds(pr.depth, int(k.min()), int(k.max()) + 1, 3)
    parts = [partition_particles(pr.particles, pr.horiz, pr.vert, lo, hi) for lo, hi in bounds]
    assert sum(p["n_nonrigid"] for p in parts) == nnr
    for p in parts:
        assert np.array_equal(p["ids"][p["n_nonrigid"]:], np.arange(nnr, n))
    down = [dict(ids=p["ids"], pos=p["pos"], in_elem=p["in_elem"]) for p in parts]
    whole = assemble_by_id(down, n, n - nnr)
    assert np.array_equal(whole["pos"], pr.particles["pos"]) and np.array_equal(whole["in_elem"], pr.particles["in_elem"])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ex = NeighbourExchange(rank, world)
    ok = True
    # halo swap: every rank fills its send buffers with a rank-tagged pattern
    n = 1000
    send_lo = torch.full((n,), 10.0 * rank + 1, dtype=torch.float64)
    send_hi = torch.full((n,), 10.0 * rank + 2, dtype=torch.float64)
    recv_lo = torch.zeros(n, dtype=torch.float64)
    recv_hi = torch.zeros(n, dtype=torch.float64)
    ex.swap(send_lo, send_hi, recv_lo, recv_hi)
    if rank > 0:
        ok &= bool(torch.all(recv_lo == 10.0 * (rank - 1) + 2))
    else:
        ok &= bool(torch.all(recv_lo == 0))
    if rank < world - 1:
        ok &= bool(torch.all(recv_hi == 10.0 * (rank + 1) + 1))
    else:
        ok &= bool(torch.all(recv_hi == 0))
    # migration: rank r sends r+1 rows down and r+3 rows up
    row = 53
    n_lo, n_hi = (rank + 1 if rank > 0 else 0), (rank + 3 if rank < world - 1 else 0)
    f_lo, f_hi = ex.swap_counts(n_lo, n_hi, torch.device("cpu"))
    ok &= f_lo == ((rank - 1) + 3 if rank > 0 else 0)
    ok &= f_hi == ((rank + 1) + 1 if rank < world - 1 else 0)
    cap = 16
    s_lo = torch.arange(cap * row, dtype=torch.float64) + 1000 * rank
    s_hi = torch.arange(cap * row, dtype=torch.float64) + 1000 * rank + 500
    r_lo = torch.zeros(cap * row, dtype=torch.float64)
    r_hi = torch.zeros(cap * row, dtype=torch.float64)
    ex.swap_rows(s_lo, s_hi, r_lo, r_hi, n_lo, n_hi, f_lo, f_hi, row)
    if f_lo:
        ok &= bool(torch.equal(r_lo[: f_lo * row], torch.arange(f_lo * row, dtype=torch.float64) + 1000 * (rank - 1) + 500))
    if f_hi:
        ok &= bool(torch.equal(r_hi[: f_hi * row], torch.arange(f_hi * row, dtype=torch.float64) + 1000 * (rank + 1)))
    ok &= bool(torch.all(r_lo[f_lo * row:] == 0)) and bool(torch.all(r_hi[f_hi * row:] == 0))
    # the two-integer handshake over a separate CPU group (what SlabSim uses next to NCCL so that it stays off the GPU stream)
    ex2 = NeighbourExchange(rank, world, cpu_group=dist.new_group(backend="gloo"))
    for step in range(3):
        g_lo, g_hi = ex2.swap_counts(100 * step + rank, 200 * step + rank, torch.device("cpu"))
        ok &= g_lo == (200 * step + rank - 1 if rank > 0 else 0)
        ok &= g_hi == (100 * step + rank + 1 if rank < world - 1 else 0)
    q.put((rank, ok))
    dist.destroy_process_group()


def test_neighbour_exchange_gloo_world2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res

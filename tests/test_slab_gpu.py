"""Slab decomposition on the GPU: the multi-slab run must reproduce the reference (golden dumps) to the
same tolerances as the single-GPU run, with particles migrating between slabs.

 * LockstepCluster: every slab in one process on one GPU (same library calls and buffers as the
   multi-process run, device copies instead of NCCL) -- runs wherever one GPU is available;
 * torchrun with 2 ranks over NCCL when the box has >= 2 GPUs."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.parity import TOL_100STEP, compare_particles, load_golden

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def occupied_planes(prob):
    nnr = int(prob.particles.get("n_nonrigid", prob.nparticles))
    k = (np.asarray(prob.particles["in_elem"])[:nnr] - 1) // (prob.horiz * prob.vert)
    return int(k.min()), int(k.max()) + 1


@pytest.mark.parametrize("case,world,sort_interval", [("block3d_fast_crossings", 2, 0), ("block3d_jitter", 2, 5),
                                                       ("block3d_ugimp_usavg", 1, 0), ("block3d_rigid_wall", 2, 4),
                                                       ("block3d_rigid_piston", 2, 0)])
def test_lockstep_slabs_match_reference(case, world, sort_interval):
    from nairn_mpm_fea_b200.problem import from_reference_dump
    from nairn_mpm_fea_b200.slab import LockstepCluster, slab_bounds
    z = load_golden(case)
    prob = from_reference_dump(z)
    first, last = occupied_planes(prob)
    # cut one plane off-centre so the slab face is not a symmetry plane of the block
    bounds = slab_bounds(prob.depth, first, last, world)
    cl = LockstepCluster(prob, bounds, device=0, sort_interval=sort_interval)
    snaps = sorted(int(k[1:].split("/")[0]) for k in z if k.startswith("p") and k.endswith("/pos") and k[1] != "0")
    done = 0
    for s in snaps:
        cl.step(s - done)
        done = s
        got = cl.download()
        errs, bad = compare_particles(got, z, "p%d" % s, TOL_100STEP)
        assert not bad, "%s, %d slabs, after %d steps: %s" % (case, world, s, bad)
        assert np.array_equal(got["in_elem"], z["p%d/inElem" % s])
    if world > 1 and case == "block3d_fast_crossings":
        moved = sum(s.migrated_out for s in cl.sims)
        assert moved > 0, "test problem should push particles across the slab face"
        assert sum(s.migrated_in for s in cl.sims) == moved
    assert sum(s.num_particles() for s in cl.sims) == prob.nparticles + (world - 1) * cl.n_rigid
    cl.close()


def test_lockstep_slab_that_starts_empty():
    """A body moving along z enters a slab that held no particles at upload: the free-flying block of
    block3d_free_ugimp starts in cell planes 4-6 and has 10 particles in plane 3 after 30 steps; slab 0 = planes [0, 4)."""
    from nairn_mpm_fea_b200.problem import from_reference_dump
    from nairn_mpm_fea_b200.slab import LockstepCluster
    z = load_golden("block3d_free_ugimp")
    prob = from_reference_dump(z)
    first, last = occupied_planes(prob)
    assert first == 4
    cl = LockstepCluster(prob, [(0, first), (first, prob.depth)], device=0, sort_interval=3)
    assert cl.sims[0].num_particles() == 0
    cl.step(30)
    got = cl.download()
    errs, bad = compare_particles(got, z, "p30", TOL_100STEP)
    assert not bad, bad
    assert np.array_equal(got["in_elem"], z["p30/inElem"])
    assert cl.sims[0].num_particles() == int(np.count_nonzero((z["p30/inElem"] - 1) // (prob.horiz * prob.vert) < first)) > 0
    cl.close()


@pytest.mark.parametrize("exchange", ["library", "torch"])
@pytest.mark.parametrize("case,extra", [("block3d_fast_crossings", ()), ("block3d_xpic3", ("--no-migration-needed",)),
                                        ("block3d_fmpm2", ("--no-migration-needed",)), ("block3d_rigid_wall", ("--no-migration-needed",))])
def test_two_ranks_over_nccl(case, extra, exchange):
    """2 processes, 2 GPUs, NCCL: halo exchanges (incl. the ones inside XPIC/FMPM iterations), migration, replicated
    rigid particles; compared with the reference dump.  exchange = library: ncclSend/ncclRecv issued by libmpmgpu itself
    (mpmgpu_slab_connect / mpmgpu_slab_step, the production path); torch: the same buffers moved by torch.distributed P2P ops."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "slab_worker.py"),
           case, *extra]
    env = dict(os.environ, MPMGPU_SLAB_TORCH_EXCHANGE="1" if exchange == "torch" else "0")
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "SLAB_OK" in p.stdout

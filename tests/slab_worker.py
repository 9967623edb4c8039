"""torchrun worker: slab-decomposed run over NCCL checked against the golden reference dump."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nairn_mpm_fea_b200.problem import from_reference_dump  # noqa: E402
from nairn_mpm_fea_b200.slab import SlabSim, gather_by_id, partition_particles, slab_bounds  # noqa: E402
from tests.parity import TOL_100STEP, compare_particles, load_golden, xpic_for_step  # noqa: E402


def main(case):
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    z = load_golden(case)
    prob = from_reference_dump(z)
    nnr = int(prob.particles.get("n_nonrigid", prob.nparticles))
    k = (np.asarray(prob.particles["in_elem"])[:nnr] - 1) // (prob.horiz * prob.vert)
    bounds = slab_bounds(prob.depth, int(k.min()), int(k.max()) + 1, world)
    lo, hi = bounds[rank]
    part = partition_particles(prob.particles, prob.horiz, prob.vert, lo, hi)
    sim = SlabSim(prob, part, lo, hi, rank, world, device=local)
    snaps = sorted(int(q[1:].split("/")[0]) for q in z if q.startswith("p") and q.endswith("/pos") and q[1] != "0")
    done = 0
    ok = True
    for s in snaps:
        while done < s:                 # the PeriodicXPIC schedule of the golden run, if it has one
            x = xpic_for_step(z, done + 1)
            if x:
                sim.sim.set_xpic(*x)
            sim.step(1)
            done += 1
        got = gather_by_id(sim.download(), prob.nparticles, n_rigid=prob.nparticles - nnr)
        errs, bad = compare_particles(got, z, "p%d" % s, TOL_100STEP)
        if bad or not np.array_equal(got["in_elem"], z["p%d/inElem" % s]):
            ok = False
            print("rank", rank, "after", s, "steps:", bad)
    moved = torch.tensor([sim.migrated_out], device="cuda")
    dist.all_reduce(moved)
    if rank == 0:
        print("migrated rows:", int(moved.item()))
        if ok and (int(moved.item()) > 0 or "--no-migration-needed" in sys.argv):
            print("SLAB_OK")
    sim.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main(sys.argv[1])

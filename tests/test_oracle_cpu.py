"""Pins the C restatement (oracle/mpm_oracle.c) against the golden dumps of the UNMODIFIED reference:
every task of step 1 (nodes + particles) and whole-step snapshots, same tolerances as the GPU tests.
Runs on CPU."""
import numpy as np
import pytest

from tests.parity import TASK_MAP, TOL_1STEP, TOL_100STEP, compare_nodes, compare_particles, load_golden

CASES = ["block3d_neohookean", "block3d_neohookean_uj1", "block3d_isoplastic", "disks2d_neohookean", "disks2d_isoplastic",
         "disks2d_ugimp_planestrain", "disks2d_linear_planestress", "block3d_jitter", "block3d_ugimp_usavg", "block3d_fast_crossings", "block3d_gravity_damping", "block3d_linear_usl",
         "block3d_ugimp_usf"]
TASK_INDEX = {"initialization": 0, "mass_and_momentum": 1, "post_extrapolation": 2, "update_strains_first": 3, "grid_forces": 4,
              "post_forces": 5, "update_momenta": 6, "update_particles": 7, "update_strains_last": 8, "reset_elements": 9}


def make(z):
    from nairn_mpm_fea_b200.problem import from_reference_dump
    from oracle.port import PortOracle
    return PortOracle(from_reference_dump(z))


@pytest.mark.parametrize("case", CASES)
def test_port_tasks_match_reference(case):
    z = load_golden(case)
    o = make(z)
    for i, nm in enumerate(str(s) for s in z["task_names"]):
        o.run_task(TASK_INDEX[TASK_MAP[nm]])
        pre = "s1/t%d" % i
        errs, bad = compare_nodes(o.download_nodes(), z, pre + "/nodes", TOL_1STEP)
        assert not bad, "%s task %d (%s): nodes %s" % (case, i, nm, bad)
        got = o.download()
        errs, bad = compare_particles(got, z, pre + "/p", TOL_1STEP)
        assert not bad, "%s task %d (%s): particles %s" % (case, i, nm, bad)
        assert np.array_equal(got["in_elem"], z[pre + "/p/inElem"])
    o.close()


@pytest.mark.parametrize("case", CASES)
def test_port_whole_steps_match_reference(case):
    z = load_golden(case)
    o = make(z)
    snaps = sorted(int(k[1:].split("/")[0]) for k in z if k.startswith("p") and k.endswith("/pos") and k[1] != "0")
    done = 0
    for s in snaps:
        o.step(s - done)
        done = s
        got = o.download()
        errs, bad = compare_particles(got, z, "p%d" % s, TOL_1STEP if s == 1 else TOL_100STEP)
        assert not bad, "%s after %d steps: %s" % (case, s, bad)
        assert np.array_equal(got["in_elem"], z["p%d/inElem" % s])
        assert np.array_equal(got["crossings"], z["p%d/crossings" % s])
    o.close()

"""Output side on the GPU, through the C ABI: archive records packed on the device must be
byte-identical to nairn_mpm_fea_b200/archive.py applied to a full download of the same state (archive.py is byte-identical to
the reference CLI's files, tests/test_archive_cpu.py), whatever the internal particle order (fused path: physically sorted);
the global sums must equal numpy sums over the download and be repeatable."""
import numpy as np
import pytest

from nairn_mpm_fea_b200 import archive
from tests.parity import load_golden
from tests.test_device_archive_cpu import reference_sums, sums_close

pytestmark = pytest.mark.gpu

CASES = [("block3d_ugimp_usavg", 1, "iYYYYNNNNNNNNNNNNN"), ("block3d_jitter", 2, "iYYYYYNYYYNNYYYYYY"), ("block3d_neohookean", 2, "iYYYYYNYYYNNYCYYYY"),
         ("block3d_isoplastic", 1, "iYYYYYNYYYNNYYYYYY"), ("disks2d_isoplastic", 1, "iYYYYYNYYYNNYYYYYY"), ("block3d_rigid_wall", 2, "iYYYYNNYNNNNNYNNYN"),
         ("disks2d_rigid_plate", 1, "iYYNNNNNNNNNNNNNNN")]


def make_sim(z, kernel_path):
    from nairn_mpm_fea_b200 import MpmGpu
    from nairn_mpm_fea_b200.problem import from_reference_dump
    prob = from_reference_dump(z)
    return MpmGpu(prob, device=0, kernel_path=kernel_path, sort_interval=3 if kernel_path == 2 else 0), prob


@pytest.mark.parametrize("case,kernel_path,order", CASES)
def test_device_records_equal_host_writer_on_download(case, kernel_path, order):
    z = load_golden(case)
    sim, prob = make_sim(z, kernel_path)
    n = prob.nparticles
    rng = np.random.default_rng(11)
    angles0 = 0.2 * rng.standard_normal((3, n))
    sim.set_archive_origin(angles0=angles0, thickness=1.25)
    for nsteps in (0, 7):                # before any step (upload order) and after steps (sorted order on the fused path)
        if nsteps:
            sim.step(nsteps)
        state = sim.download()
        want, recsize = archive.records(prob, state, order, thickness=np.full(n, 1.25), angles0=angles0)
        assert sim.archive_record_size(order) == recsize
        got = sim.pack_archive(order)
        if got != want:
            a, b = np.frombuffer(got, np.uint8).reshape(n, recsize), np.frombuffer(want, np.uint8).reshape(n, recsize)
            cols = np.nonzero(np.any(a != b, axis=0))[0]
            rows = np.nonzero(np.any(a != b, axis=1))[0]
            raise AssertionError("%s after %d steps: record bytes differ at offsets %s in %d records (record size %d)"
                                 % (case, nsteps, cols[:24].tolist(), rows.size, recsize))
    sim.close()


@pytest.mark.parametrize("case,kernel_path", [(c, k) for c, k, _ in CASES])
def test_global_sums_equal_numpy_over_download(case, kernel_path):
    z = load_golden(case)
    sim, prob = make_sim(z, kernel_path)
    sim.step(5)
    state = sim.download()
    got = sim.global_sums()
    want = reference_sums(prob, state, 3 if prob.is3d else 2)
    assert sums_close(got, want, 1e-11), np.abs(got - want[0]).max(axis=0)
    again = sim.global_sums()
    assert np.array_equal(got, again, equal_nan=True), "fixed summation order: repeated calls must agree bit for bit"
    sim.close()


def test_unsupported_archive_items_are_refused():
    from nairn_mpm_fea_b200.capi import MpmGpuError
    z = load_golden("block3d_ugimp_usavg")
    sim, prob = make_sim(z, 1)
    assert sim.archive_record_size("iYYYYNNNNNNYNNNNNN") == -1          # shear components
    with pytest.raises(MpmGpuError):
        sim.pack_archive("iYYYYNNNNNNYNNNNNN")
    small = np.zeros(16, np.uint8)
    with pytest.raises(MpmGpuError):
        sim.pack_archive("iYYYY", out=small)
    sim.close()

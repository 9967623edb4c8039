"""nairn_mpm_fea_b200/archive.py writes the reference's binary particle archives: checked byte for byte against the files the
reference's own CLI (oracle/_ref/NairnMPM, one thread) writes for the golden cases, with the state taken from the golden
dump of the same step."""
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from nairn_mpm_fea_b200 import archive, problem
from tests.parity import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "NairnMPM")

CASES = ["block3d_ugimp_usavg", "disks2d_ugimp_planestrain", "block3d_neohookean_uj1", "disks2d_isoplastic", "block3d_rigid_wall_lattice"]


FULL_ORDER = "iYYYYYNYYYNNYYYYYY"        # every item this path produces (bytes 6 and 10 are obsolete, 11 = shear components), history 1


@pytest.mark.parametrize("case,order_in", [(c, None) for c in CASES] +
                         [("block3d_neohookean_uj1", FULL_ORDER), ("disks2d_isoplastic", FULL_ORDER), ("disks2d_neohookean", "iYYYYYNYYYNNYCYYYY"),
                          ("disks2d_mooney_planestress", "iYYYYYNYYYNNYCYYYY")])
def test_archives_match_the_reference_cli_byte_for_byte(case, order_in):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/NairnMPM not built")
    z = load_golden(case)
    prob = problem.from_reference_dump(z)
    snaps = sorted(int(k[1:].split("/")[0]) for k in z if k.startswith("p") and k.endswith("/pos") and k[1] != "0")
    step = [s for s in snaps if s <= 40][-1]
    dt_ms = prob.dt * 1.0e3
    xml = str(z["xml"])
    xml = re.sub(r'<ArchiveTime units="ms">[^<]*</ArchiveTime>', '<ArchiveTime units="ms">%r</ArchiveTime>' % (0.999 * dt_ms), xml)
    if order_in is not None:
        xml = re.sub(r"<MPMArchiveOrder>[^<]*</MPMArchiveOrder>", "<MPMArchiveOrder>%s</MPMArchiveOrder>" % order_in, xml)
    tmax = (step + 0.5) * dt_ms
    xml = re.sub(r'(<Timing [^>]*max=")[^"]*(")', r"\g<1>%r\g<2>" % tmax, xml)
    xml = re.sub(r'<MaxTime units="ms">[^<]*</MaxTime>', '<MaxTime units="ms">%r</MaxTime>' % tmax, xml)
    d = tempfile.mkdtemp(prefix="arch_")
    open(os.path.join(d, "in.fmcmd"), "w").write(xml)
    p = subprocess.run([REF, "-np", "1", "in.fmcmd"], cwd=d, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-1000:] + p.stderr[-1000:]
    order = re.search(r"Archive format: (\S+)", p.stdout).group(1)
    crack = re.search(r"Crack archive format: (\S+)", p.stdout).group(1)
    root = re.search(r"<ArchiveRoot>([^<]*)</ArchiveRoot>", xml).group(1)
    want = open(os.path.join(d, root + str(step)), "rb").read()
    pre = "p%d/" % step
    state = dict(pos=z[pre + "pos"], vel=z[pre + "vel"], sp=z[pre + "sp"], pressure=z[pre + "pressure"], ep=z[pre + "ep"],
                 wrot=z[pre + "wrot"], eplast=z[pre + "eplast"], energies=z[pre + "energies"], history=z[pre + "hist"],
                 in_elem=z[pre + "inElem"], crossings=z[pre + "crossings"])
    body, recsize = archive.records(prob, state, order)
    got = archive.header(archive.normalise_order(order), crack, prob.is3d, step * prob.dt) + body
    n = prob.nparticles
    assert len(want) == archive.HEADER_LENGTH + n * recsize, (len(want), recsize, n)
    assert got[:archive.HEADER_LENGTH] == want[:archive.HEADER_LENGTH]
    if got != want:         # say which record field differs before failing
        a = np.frombuffer(got[archive.HEADER_LENGTH:], np.uint8).reshape(n, recsize)
        b = np.frombuffer(want[archive.HEADER_LENGTH:], np.uint8).reshape(n, recsize)
        cols = np.nonzero(np.any(a != b, axis=0))[0]
        raise AssertionError("record bytes differ at offsets %s (record size %d)" % (cols[:24].tolist(), recsize))


def test_unsupported_items_are_refused():
    z = load_golden("block3d_ugimp_usavg")
    prob = problem.from_reference_dump(z)
    with pytest.raises(NotImplementedError):
        archive.records(prob, {}, "iYYYYNNNNNNYNNNNNN")          # shear components

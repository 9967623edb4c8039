"""Reader for the reference's binary particle archives (ver6; System/ArchiveData.cpp:464-489 header,
:806-966 records) -- enough to compare two runs field by field."""
import glob
import os
import re

import numpy as np


def read_archive(path, nparticles=None):
    raw = open(path, "rb").read()
    assert raw[:4] == b"ver6", raw[:8]
    body = raw[64:]
    olen = raw[4]
    order = raw[5:5 + olen].decode()
    if nparticles is None:
        raise ValueError("need particle count")
    rec = len(body) // nparticles
    assert rec * nparticles == len(body), (len(body), nparticles)
    a = np.frombuffer(body, dtype=np.uint8).reshape(nparticles, rec)
    elem = a[:, 0:4].copy().view(np.int32)[:, 0]
    mp = a[:, 4:12].copy().view(np.float64)[:, 0]
    mat = a[:, 12:14].copy().view(np.int16)[:, 0]
    # after the 16-byte fixed head every optional field is a double except an int-sized tail
    # (element crossings) when the order string asks for it: split generically
    nd = (rec - 16) // 8
    tail = (rec - 16) % 8
    dbl = a[:, 16:16 + 8 * nd].copy().view(np.float64)
    tail_bytes = a[:, 16 + 8 * nd:16 + 8 * nd + tail].copy()
    return dict(order=order, elem=elem, mp=mp, mat=mat, doubles=dbl, tail=tail_bytes)


def list_archives(root_prefix):
    """All numbered archive files <root_prefix><step>, sorted by step."""
    out = []
    for f in glob.glob(root_prefix + "*"):
        m = re.match(re.escape(root_prefix) + r"(\d+)$", f)
        if m:
            out.append((int(m.group(1)), f))
    return sorted(out)

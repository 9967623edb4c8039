"""The shipped library really contains the hand-written sm_100a kernels the design describes (checked on the SASS, no GPU
needed): FP64 reductions and warp match in the scatter kernels, cp.async (LDGSTS) node tiles in the gathering kernels,
bulk-copy (TMA) instructions in the pipelined variant, and no kernel built for another architecture."""
import collections
import re
import shutil
import subprocess

import pytest


def _sass():
    from nairn_mpm_fea_b200 import build
    lib = build.build()
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([tool, "-sass", lib], capture_output=True, text=True, check=True).stdout
    feats = collections.defaultdict(collections.Counter)
    archs = set(re.findall(r"arch = (sm_\w+)", out))
    fn = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        if fn is None:
            continue
        for key, pat in (("red_f64", "REDG.E.ADD.F64"), ("match", "MATCH.ANY"), ("ldgsts", "LDGSTS"), ("tma", "UBLKCP"),
                         ("dfma", "DFMA"), ("local", "STL")):
            if pat in line:
                feats[fn][key] += 1
    return archs, feats


def test_fused_kernels_have_the_designed_instructions():
    archs, feats = _sass()
    assert archs == {"sm_100a"}, archs

    def find(prefix):
        hits = [f for f in feats if prefix in f]
        assert hits, prefix
        return hits

    for f in find("k_f1_mass_momentum"):
        assert feats[f]["match"] >= 1 and feats[f]["red_f64"] >= 4 and feats[f]["local"] == 0, (f, feats[f])
    for prefix in ("k_f2_strain_forces", "k_f3_update_momentum", "k_fx_iterate"):
        for f in find(prefix):
            assert feats[f]["match"] >= 1 and feats[f]["red_f64"] >= 3 and feats[f]["ldgsts"] >= 1, (f, feats[f])
    for f in find("k_f4_strain_reset"):
        assert feats[f]["ldgsts"] >= 1 and feats[f]["red_f64"] == 0 and feats[f]["dfma"] > 100, (f, feats[f])
    for f in find("k_f4_pipe"):
        assert feats[f]["tma"] >= 1, (f, feats[f])
    for f in find("k_p2g_mass_momentum"):            # the general path: one FP64 reduction per particle-node pair
        assert feats[f]["red_f64"] >= 1, (f, feats[f])

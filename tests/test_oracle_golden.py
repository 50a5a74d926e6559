"""The oracle (oracle/oracle.c) against the committed golden vectors, which were produced by the
COMPILED UNMODIFIED REFERENCE (tests/golden/make_golden.py).  Bit-exact: this is what pins the oracle."""
import glob
import os

import numpy as np
import pytest

from solaris_b200 import synth
from oraclelib import Oracle, default_nebula

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
INTEGRATORS = {"rkf78": 3, "rk4": 1, "dp": 0}


def load_case(path):
    g = np.load(path)
    s = synth.System({k: g[k] for k in ("counts", "y0", "mass", "radius", "density", "cD", "gammaStokes", "gammaEpstein",
                                        "migStopAt", "type", "migType", "id")})
    s["n"] = int(s["counts"].sum())
    return g, s, bool(g["barycentric"]), (default_nebula() if int(g["with_nebula"]) else None)


def test_golden_files_present():
    assert len(GOLDEN) >= 6


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_compute_matches_reference_golden(path):
    g, s, bary, neb = load_case(path)
    o = Oracle(s, bary, neb)
    for key in g.files:
        if not key.startswith("compute_f"):
            continue
        fl = int(key[len("compute_f"):])
        a = o.compute(float(g["t_compute"]), s.y0, fl)
        assert np.array_equal(a, g[key]), f"{key}: oracle differs from the reference"
        rm3, idx, dist, mig = o.side()
        assert np.array_equal(rm3, g[f"rm3_f{fl}"])
        assert np.array_equal(idx, g[f"nnidx_f{fl}"])
        assert np.array_equal(dist, g[f"nndist_f{fl}"])
        assert np.array_equal(mig, g[f"migtype_f{fl}"])


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
@pytest.mark.parametrize("iname", list(INTEGRATORS))
def test_oracle_drivers_match_reference_golden(path, iname):
    g, s, bary, neb = load_case(path)
    log = g[f"{iname}_log"]
    o = Oracle(s, bary, neb)
    t, h = 0.0, (0.01 if iname == "rk4" else 0.05)
    for k in range(len(log)):
        r, t, h, hd, _, _ = o.step(INTEGRATORS[iname], t, h)
        assert r == 0
        assert (t, h, hd) == tuple(log[k]), f"step {k}"
        if k == 0:
            assert np.array_equal(o.array("y0"), g[f"{iname}_y0_first"])
    assert np.array_equal(o.array("y0"), g[f"{iname}_y0_last"])
    assert np.array_equal(o.side()[3], g[f"{iname}_migtype_last"])

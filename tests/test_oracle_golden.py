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


def test_phases_writer_matches_reference_golden():
    """(f) row 2: oracle_pack_phases / oracle_format_phases_text against bytes the reference's
    BinaryFileAdapter::SavePhases wrote (tests/golden/io/phases_writer.npz)."""
    from oraclelib import oracle_format_phases_text, oracle_pack_phases
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "io", "phases_writer.npz"))
    b, t = b"", b""
    for k in range(3):
        b += oracle_pack_phases(float(g[f"t_{k}"]), g[f"y_{k}"], g[f"id_{k}"])
        t += oracle_format_phases_text(float(g[f"t_{k}"]), g[f"y_{k}"], g[f"id_{k}"])
    assert b == g["binary"].tobytes()
    assert t == g["text"].tobytes()


def test_remove_body_matches_reference_golden():
    """(f) row 3: oracle_remove_body against states the reference's Simulator::RemoveBody produced
    (tests/golden/io/remove_body.npz), including the slots the reference does not move (cD, migStopAt)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "io", "remove_body.npz"))
    s = synth.System({k: g[k] for k in ("counts", "y0", "mass", "radius", "density", "cD", "gammaStokes", "gammaEpstein",
                                        "migStopAt", "type", "migType", "id")})
    s["n"] = int(s["counts"].sum())
    o = Oracle(s, False, default_nebula())
    for step, bid in enumerate(g["victim_ids"]):
        assert o.remove_body(int(bid)) == 0
        for k, v in o.params().items():
            assert np.array_equal(v, g[f"after{step}_{k}"]), (step, k)
        assert np.array_equal(o.array("y0"), g[f"after{step}_y0"])
        np.testing.assert_array_equal(o.compute(1.0, o.array("y0"), 7), g[f"after{step}_accel"])
    assert o.remove_body(123456789) == 2          # unknown id: rejected (the reference reads past the end)


def test_elements_to_phases_matches_reference_golden():
    """(f) row 4: oracle_elements_to_phases against Ephemeris::CalculatePhase outputs (tests/golden/io/elements.npz)."""
    from oraclelib import oracle_elements_to_phases
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "io", "elements.npz"))
    out, bad = oracle_elements_to_phases(g["mu"], g["elements"])
    assert bad == int(g["failed"])
    assert np.array_equal(out, g["phases"])

"""The oracle against the compiled reference, live (only where oracle/_ref exists, i.e. the build
container; the committed golden vectors cover the GPU box).  Bit-exact on every output."""
import numpy as np
import pytest

from solaris_b200 import synth
from oraclelib import NebulaPod, Oracle, Reference, default_nebula, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="oracle/_ref not built (no /root/reference)")


def test_default_nebula_equals_reference_constructor():
    d, r = default_nebula(), Reference.nebula_defaults()
    for f, _ in NebulaPod._fields_:
        assert getattr(d, f) == getattr(r, f), f


CASES = [
    ("ac-mixed-neb", lambda: synth.mixed([1, 2, 3, 5, 4, 20, 10], migration=True, seed=3), False, True),
    ("ac-mixed", lambda: synth.mixed([1, 2, 3, 5, 4, 20, 10], migration=False, seed=4), False, False),
    ("bc-mixed-neb", lambda: synth.to_barycentric(synth.mixed([1, 2, 0, 6, 4, 20, 10], migration=True, seed=5)), True, True),
    ("bc-disk", lambda: synth.to_barycentric(synth.massive_disk(60)), True, False),
    ("ac-trojans", lambda: synth.trojans(80), False, False),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_compute_bit_exact(case):
    _, make, bary, neb = case
    s = make()
    neb = default_nebula() if neb else None
    o, r = Oracle(s, bary, neb), Reference(s, bary, neb)
    for flags in (7, 1, 0, 6):
        y = s.y0 * (1.0 + 1e-4 * flags)
        assert np.array_equal(o.compute(3.0, y, flags), r.compute(3.0, y, flags))
        for a, b in zip(o.side(), r.side()):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("integ", [3, 1, 0])
def test_drivers_bit_exact(case, integ):
    name, make, bary, neb = case
    if bary and neb:
        pytest.skip("barycentric + non-massive bodies: the reference accumulates into uninitialised k-arrays "
                    "(Acceleration.cpp:593-631 never zeroes them), undefined behaviour")
    s = make()
    neb = default_nebula() if neb else None
    o, r = Oracle(s, bary, neb), Reference(s, bary, neb, integ)
    to = tr = 0.0
    ho = hr = 0.01 if integ == 1 else 300.0      # 300 d forces rejected attempts in the adaptive drivers
    for _ in range(15):
        ro, to, ho, hdo, _, _ = o.step(integ, to, ho)
        rr, tr, hr, hdr, _, _ = r.step(integ, tr, hr)
        assert (ro, to, ho, hdo) == (rr, tr, hr, hdr)
        if ro != 0:
            break
        assert np.array_equal(o.array("y0"), r.array("y0"))
        assert np.array_equal(o.array("y"), r.array("y"))
    for a, b in zip(o.side(), r.side()):
        assert np.array_equal(a, b)


def test_rows_oracle_equals_full_oracle():
    s = synth.massive_disk(300)
    for bary in (False, True):
        ss = synth.to_barycentric(s) if bary else s
        o = Oracle(ss, bary, None)
        full = o.compute(0.0, ss.y0, 0)
        rows = o.gravity_rows(ss.y0, 0, ss.n, threads=4)
        assert np.array_equal(full, rows)


def test_phases_writer_bytes_equal_reference(tmp_path):
    """(f) row 2: oracle_pack_phases / oracle_format_phases_text against BinaryFileAdapter::SavePhases."""
    from oraclelib import oracle_format_phases_text, oracle_pack_phases, reference_save_phases
    rng = np.random.default_rng(11)
    want_bin, want_txt = b"", b""
    for k, n in enumerate((0, 1, 7, 1000)):
        y = rng.normal(size=(n, 6)) * 10.0 ** rng.integers(-30, 30, size=(n, 6))
        if n:
            y[0] = 0.0
            y[-1, 2] = -0.0
        ids = rng.integers(0, 2 ** 31 - 1, size=n).astype(np.int32)
        t = 365.25 * k * 1.0e3 + 0.125
        reference_save_phases(str(tmp_path), "Phases.dat", t, y, ids, text=False)
        reference_save_phases(str(tmp_path), "Phases.dat", t, y, ids, text=True)
        want_bin += oracle_pack_phases(t, y, ids)
        want_txt += oracle_format_phases_text(t, y, ids)
    assert (tmp_path / "Phases.dat").read_bytes() == want_bin
    assert (tmp_path / "Phases.txt").read_bytes() == want_txt


def test_remove_body_bit_exact():
    """(f) row 3: oracle_remove_body against Simulator::RemoveBody, including what the reference does NOT move
    (cD, migStopAt) and a force evaluation on the shrunk system."""
    s = synth.mixed([1, 2, 3, 5, 4, 20, 10], migration=True, seed=9)
    s.id = (np.arange(s.n, dtype=np.int32) * 31 + 5)
    neb = default_nebula()
    o, r = Oracle(s, False, neb), Reference(s, False, neb)
    for victim in (s.n - 1, 1, 7, 20, 3):          # last body, first giant, a protoplanet, a planetesimal, a rocky planet
        bid = int(o.params()["id"][victim])
        assert o.remove_body(bid) == 0 and r.remove_body(bid) == 0
        po, pr = o.params(), r.params()
        for k in po:
            assert np.array_equal(po[k], pr[k]), k
        assert np.array_equal(o.array("y0"), r.array("y0"))
        y = o.array("y0")
        # (cD keeps its slot, so a planetesimal can inherit cD = 0 from a massive body: NaN drag in the transition
        #  regime, in the reference and in the oracle alike -> NaN-aware comparison)
        np.testing.assert_array_equal(o.compute(1.0, y, 7), r.compute(1.0, y, 7))
    assert o.n == s.n - 5


def elements_sample(n, seed):
    rng = np.random.default_rng(seed)
    el = np.column_stack([rng.uniform(0.3, 40.0, n), rng.uniform(0.0, 0.95, n), rng.uniform(0.0, 0.6, n),
                          rng.uniform(0, 2 * np.pi, n), rng.uniform(0, 2 * np.pi, n), rng.uniform(0, 2 * np.pi, n)])
    el[0, 1] = 0.0                      # circular: E = M shortcut
    el[1, 5] = 0.0                      # M = 0 shortcut
    el[2, 5] = 3.14159265358979323846   # M = pi shortcut
    el[3, 1] = 0.99                     # very eccentric
    el[4, 5] = 1.0e-9                   # tiny mean anomaly
    mu = 2.959122082855911025e-4 * (1.0 + rng.uniform(0.0, 1.0e-3, n))
    return mu, el


def test_elements_to_phases_bit_exact():
    """(f) row 4: oracle_elements_to_phases against Ephemeris::CalculatePhase (same libm -> bit-exact)."""
    from oraclelib import oracle_elements_to_phases, reference_elements_to_phases
    mu, el = elements_sample(20000, 3)
    o, bad_o = oracle_elements_to_phases(mu, el)
    r, bad_r = reference_elements_to_phases(mu, el)
    # the reference's Newton iteration does not converge for a few very eccentric orbits ("Could not compute the
    # excentric anomaly E!"): same bodies fail in both, their rows stay untouched
    assert bad_o == bad_r and 0 < bad_o < 20
    assert np.array_equal(o, r)
    assert int((np.abs(o).sum(axis=1) == 0).sum()) == bad_o


def test_event_records_bytes_equal_reference_writer(tmp_path):
    """oracle_event_records against the reference's own TwoBodyAffair constructor (running ids) and
    BinaryFileAdapter::SaveTwoBodyAffairs, for a scan that yields ejections and hit centrums."""
    s = synth.mixed([1, 2, 3, 5, 4, 20, 10], migration=False, seed=4)
    s.id = (np.arange(s.n, dtype=np.int32) * 13 + 7)
    o, r = Oracle(s, False, None), Reference(s, False, None)
    o.compute(0.0, s.y0, 0); r.compute(0.0, s.y0, 0)               # fills rm3
    dist = np.sqrt((s.y0[1:, :3] ** 2).sum(axis=1))
    ej, hc = float(np.quantile(dist, 0.8)), float(np.quantile(dist, 0.15))
    e_idx, h_idx, _ = o.detect_events(ej, hc, 0.0)
    assert len(e_idx) > 2 and len(h_idx) > 2
    merged = sorted([(int(i), 0) for i in e_idx] + [(int(i), 1) for i in h_idx])       # scan order over the bodies
    r.write_affairs(str(tmp_path), "TwoBodyAffair.dat", [k for _, k in merged], [i for i, _ in merged], 123.5, 40)
    rec, ne, nh = o.event_records(ej, hc, 123.5, 40)
    assert (ne, nh) == (len(e_idx), len(h_idx))
    assert (tmp_path / "TwoBodyAffair.dat").read_bytes() == rec

"""Whole-program parity: the reference program (oracle/_ref/solaris_ref) against the DROP-IN program
(solaris_b200/host/_build/solaris_b200_dropin = the reference's own main/Simulator/XML/output objects
linked with this repo's Acceleration / RungeKutta4 / RungeKuttaFehlberg78 / DormandPrince translation
units and libsolaris_b200.so) on the same input files.  Checks of BASELINE.json north_star:
energy and orbital elements within 1e-10 relative, identical event lists."""
import os
import re
import struct
import subprocess
import sys

import numpy as np
import pytest

import xmlgen
from helpers import orbital_elements_ae

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "solaris_ref")
DROPIN_BIN = os.path.join(ROOT, "solaris_b200", "host", "_build", "solaris_b200_dropin")
CASES = xmlgen.cases()


def read_phases(path):
    """Phases.dat: per snapshot double time, int n, n x {int id, double y[6]} (BinaryFileAdapter.cpp:107-180)."""
    out = []
    b = open(path, "rb").read()
    off = 0
    while off < len(b):
        t, n = struct.unpack_from("<di", b, off); off += 12
        ids = np.zeros(n, dtype=np.int32); y = np.zeros((n, 6))
        for k in range(n):
            ids[k] = struct.unpack_from("<i", b, off)[0]; off += 4
            y[k] = struct.unpack_from("<6d", b, off); off += 48
        out.append((t, ids, y))
    return out


def read_events(path):
    """TwoBodyAffair.dat: int id, type, body1Id, body2Id; double p1[6], p2[6], time = 120 B (BinaryFileAdapter.cpp:244-261)."""
    if not os.path.exists(path):
        return []
    b = open(path, "rb").read()
    assert len(b) % 120 == 0
    ev = []
    for off in range(0, len(b), 120):
        eid, typ, b1, b2 = struct.unpack_from("<4i", b, off)
        p1 = np.array(struct.unpack_from("<6d", b, off + 16)); p2 = np.array(struct.unpack_from("<6d", b, off + 64))
        t = struct.unpack_from("<d", b, off + 112)[0]
        ev.append((eid, typ, b1, b2, p1, p2, t))
    return ev


def read_integrals(path):
    """Integrals.dat: int len, header; per snapshot int 17, double time, double[16] (BinaryFileAdapter.cpp:186-219)."""
    b = open(path, "rb").read()
    ln = struct.unpack_from("<i", b, 0)[0]
    off = 4 + ln
    rows = []
    while off < len(b):
        k = struct.unpack_from("<i", b, off)[0]; off += 4
        rows.append(struct.unpack_from(f"<{k}d", b, off)); off += 8 * k
    return np.array(rows)


def run(binary, xml, workdir, extra_env=None, log=None):
    os.makedirs(workdir, exist_ok=True)
    p = os.path.join(workdir, "in.xml")
    open(p, "w").write(xml)
    env = dict(os.environ, OSTYPE="linux")
    env.update(extra_env or {})
    r = subprocess.run([binary, "-i", p], cwd=workdir, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    if log is not None:
        log.append(r.stderr)
    return workdir


def compare_outputs(d_ref, d_new, name, expect_events=False, barycentric=False):
    """north_star checks on the output files of the two programs: identical event lists, orbital elements and energy
    within 1e-10 relative at every snapshot."""
    # ---- identical event lists ----
    ev_r, ev_n = read_events(os.path.join(d_ref, "TwoBodyAffair.dat")), read_events(os.path.join(d_new, "TwoBodyAffair.dat"))
    assert [(e[1], e[2], e[3]) for e in ev_n] == [(e[1], e[2], e[3]) for e in ev_r], "event lists differ"
    for a, b in zip(ev_n, ev_r):
        assert abs(a[6] - b[6]) <= 1e-10 * max(abs(b[6]), 1e-300)
        np.testing.assert_allclose(a[4], b[4], rtol=1e-9, atol=1e-14)
        np.testing.assert_allclose(a[5], b[5], rtol=1e-9, atol=1e-14)
    if expect_events:
        assert len(ev_r) > 0

    # ---- snapshots: same count, same bodies, same times; orbital elements 1e-10 ----
    ph_r, ph_n = read_phases(os.path.join(d_ref, "Phases.dat")), read_phases(os.path.join(d_new, "Phases.dat"))
    # Snapshots are matched BY TIME: Simulator::DecisionMaking saves when the float sum of accepted steps
    # `lastSave` reaches `output` (Simulator.cpp:234-243); with step sequences that differ in the last
    # bits one program may need one extra ulp-sized step, i.e. one extra (duplicate-time) record.
    assert len(ph_r) >= 2 and abs(len(ph_n) - len(ph_r)) <= 1
    key = lambda t: round(t / 1e-6)   # noqa: E731
    by_time = {}
    for t_n, id_n, y_n in ph_n:
        by_time.setdefault(key(t_n), (t_n, id_n, y_n))
    matched = [(r, by_time[key(r[0])]) for r in ph_r if key(r[0]) in by_time]
    assert len(matched) >= len(ph_r) - 1
    assert key(ph_r[-1][0]) == key(ph_n[-1][0]), "final times differ"
    for (t_r, id_r, y_r), (t_n, id_n, y_n) in matched:
        assert np.array_equal(id_n, id_r)
        assert abs(t_n - t_r) <= 1e-9 * max(abs(t_r), 1.0)
        # a state that went NaN in the reference (UnifiedDragForce: log10(0) in the transition regime with cd = 0) must be
        # NaN here too, in the same places; everything below compares the finite bodies
        assert np.array_equal(np.isnan(y_n), np.isnan(y_r)), name
        fin = ~np.isnan(y_r).any(axis=1)
        if fin.sum() >= 2 and fin[0] and not barycentric:
            m = np.zeros(int(fin.sum())); m[0] = 1.0      # elements w.r.t. the star, test masses
            a_r, e_r = orbital_elements_ae(y_r[fin], m)
            a_n, e_n = orbital_elements_ae(y_n[fin], m)
            assert np.max(np.abs(a_n - a_r) / np.abs(a_r)) <= 1e-10, name
            # e^2 = 1 + 2 c^2 h / mu^2 is a cancellation: from a double-precision state e itself is only defined to
            # ~eps / e (1e-8 for a circular orbit), whoever computes it
            assert np.all(np.abs(e_n - e_r) <= 1e-10 + 5e-16 / np.maximum(e_r, 1e-12)), name
        for sl in (slice(0, 3), slice(3, 6)):           # |dr| <= 1e-8 |r|, |dv| <= 1e-8 |v| per body
            d = np.sqrt(((y_n[fin][:, sl] - y_r[fin][:, sl]) ** 2).sum(axis=1))
            nrm = np.sqrt((y_r[fin][:, sl] ** 2).sum(axis=1))
            assert np.all(d <= 1e-8 * nrm), name

    # ---- integrals: total energy (column 16 = T - U) 1e-10 relative ----
    pi_r, pi_n = os.path.join(d_ref, "Integrals.dat"), os.path.join(d_new, "Integrals.dat")
    assert os.path.exists(pi_r) == os.path.exists(pi_n)
    if os.path.exists(pi_r):
        in_r, in_n = read_integrals(pi_r), read_integrals(pi_n)
        by_time = {key(row[0]): row for row in in_n}
        rows = [(row, by_time[key(row[0])]) for row in in_r if key(row[0]) in by_time]
        assert len(rows) >= len(in_r) - 1
        for row_r, row_n in rows:
            assert (np.isnan(row_r[16]) and np.isnan(row_n[16])) or abs(row_n[16] - row_r[16]) <= 1e-10 * abs(row_r[16])
    return ph_r, ph_n


@pytest.mark.skipif(not (os.path.exists(REF_BIN) and os.path.exists(DROPIN_BIN)), reason="prebuilt reference / drop-in programs missing")
@pytest.mark.parametrize("name", list(CASES))
def test_program_parity(tmp_path, name):
    xml = CASES[name]
    d_ref = run(REF_BIN, xml, str(tmp_path / "ref"))
    d_new = run(DROPIN_BIN, xml, str(tmp_path / "b200"))
    compare_outputs(d_ref, d_new, name, expect_events=name in ("events_ejection_hitcentrum", "collisions", "late_collision"),
                    barycentric="bc" in name)


# ---- the reference's OWN shipped scenarios (tests/golden/testcases/, made by make_testcases.py from TestCases/*/*.xml) ----
TESTCASE_DIR = os.path.join(ROOT, "tests", "golden", "testcases")
SHIPPED = sorted(f[:-4] for f in os.listdir(TESTCASE_DIR) if f.endswith(".xml"))


@pytest.mark.skipif(not (os.path.exists(REF_BIN) and os.path.exists(DROPIN_BIN)), reason="prebuilt reference / drop-in programs missing")
@pytest.mark.parametrize("name", SHIPPED)
def test_shipped_testcase_parity(tmp_path, name):
    """BASELINE.json configs[0] / [1] (TestCases/SunJupiter, TestCases/SolarSystem) and the other scenarios the reference
    ships and can run: the reference program and the drop-in program on the same file."""
    xml = open(os.path.join(TESTCASE_DIR, name + ".xml")).read()
    d_ref = run(REF_BIN, xml, str(tmp_path / "ref"))
    d_new = run(DROPIN_BIN, xml, str(tmp_path / "b200"))
    ph_r, ph_n = compare_outputs(d_ref, d_new, name, expect_events=name in ("EjectionTest", "HitCentrumTest"))
    if name == "SolarSystem":
        # the one stored state of the reference (TestCases/SolarSystemWithBalint/SS.data): the initial phases of this
        # scenario in Simulator order, bit for bit - in both programs' first snapshot
        ss = np.array([[float(v) for v in ln.split()] for ln in open(os.path.join(TESTCASE_DIR, "SS.data")) if ln.strip()]).reshape(9, 6)
        assert np.array_equal(ph_r[0][2], ss) and np.array_equal(ph_n[0][2], ss)


STATS = re.compile(r"(resident|eager) synchronisation: (\d+) steps, (\d+) state downloads, (\d+) host event scans skipped, (\d+) event edits")


def stats(text):
    m = STATS.search(text)
    assert m, text[-500:]
    return (m.group(1),) + tuple(int(v) for v in m.groups()[1:])


@pytest.mark.skipif(not os.path.exists(DROPIN_BIN), reason="prebuilt drop-in program missing")
@pytest.mark.parametrize("name", list(CASES))
def test_resident_default_is_byte_identical_to_eager(tmp_path, name):
    """The default synchronisation keeps the state on the device: the hooks in front of Simulator::BodyListToBodyData
    and Simulator::CheckEvent (solaris_b200/host/SimulatorHooks.cpp) hand the thresholds of Settings to the bridge, the
    host arrays are refreshed only on event, snapshot and final steps, and the reference's CheckEvent scan runs only
    when the device flag reduction found a candidate.  It must write exactly the same output files as
    SOLARIS_B200_EAGER=1 (download after every step, the reference's host scan after every step), events included."""
    xml = CASES[name]
    elog, log = [], []
    d_eager = run(DROPIN_BIN, xml, str(tmp_path / "eager"), {"SOLARIS_B200_STATS": "1", "SOLARIS_B200_EAGER": "1"}, elog)
    d_res = run(DROPIN_BIN, xml, str(tmp_path / "resident"), {"SOLARIS_B200_STATS": "1"}, log)
    mode_e, steps_e, down_e, skip_e, edits_e = stats(elog[0])
    mode_r, steps_r, down_r, skip_r, edits_r = stats(log[0])
    assert mode_e == "eager" and mode_r == "resident"
    assert steps_e == steps_r == down_e and skip_e == 0 and down_r <= steps_r
    if name in ("events_ejection_hitcentrum", "collisions", "late_collision"):
        assert edits_e > 0 and edits_r == edits_e        # merges / removals were replayed on the device, not re-uploaded
        assert 0 < skip_r < steps_r
    else:
        assert down_r < steps_r // 2 and skip_r == steps_r
    for f in ("Phases.dat", "Integrals.dat", "TwoBodyAffair.dat"):
        pe, pr = os.path.join(d_eager, f), os.path.join(d_res, f)
        assert os.path.exists(pe) == os.path.exists(pr), f
        if os.path.exists(pe):
            assert open(pe, "rb").read() == open(pr, "rb").read(), f
    if name in ("events_ejection_hitcentrum", "collisions", "late_collision"):
        assert len(read_events(os.path.join(d_res, "TwoBodyAffair.dat"))) > 0


# ---- SOLARIS_B200_BODIES: the binary side loader for large populations (solaris_b200/host/side_loader.h) ----
def _particle_xml(k, btype, y, extra=""):
    y = [float(v) for v in y]
    return (f'        <Body type="{btype}" name="b{k}">\n          <Phase>\n'
            f'            <Position x="{y[0]!r}" y="{y[1]!r}" z="{y[2]!r}" unit="au" />\n'
            f'            <Velocity x="{y[3]!r}" y="{y[4]!r}" z="{y[5]!r}" unit="auday" />\n          </Phase>\n{extra}        </Body>\n')


@pytest.mark.skipif(not os.path.exists(DROPIN_BIN), reason="prebuilt drop-in program missing")
@pytest.mark.parametrize("kind", ["phases", "elements"])
def test_side_loaded_test_particles_equal_the_xml_loader(tmp_path, kind):
    """400 test particles once inside the XML (TinyXML DOM -> std::list<Body> -> BodyData) and once in a flat side file
    appended by the BodyListToBodyData hook: same Phases.dat, byte for byte when the file carries phases; to 1e-9 when it
    carries orbital elements (one batched device call instead of the host's per-body Kepler solves: device sin / cos
    make the initial phases differ by ~1e-11)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from write_bodies import write_bodies
    from solaris_b200 import synth
    s = synth.trojans(400)
    planets = [xmlgen.planet("Jupiter"), xmlgen.planet("Saturn")]
    ev = '    <Ejection value="15" unit="au" />\n'
    if kind == "phases":
        y = s.y0[3:]
        inside = [_particle_xml(k, "testparticle", y[k]) for k in range(len(y))]
        write_bodies(str(tmp_path / "bodies.bin"), 7, y, kind=0)
    else:
        rng = np.random.default_rng(5)
        el = np.column_stack([rng.uniform(5.05, 5.35, 400), rng.uniform(0, 0.15, 400), rng.uniform(0, 0.4, 400),
                              rng.uniform(0, 2 * np.pi, 400), rng.uniform(0, 2 * np.pi, 400), rng.uniform(0, 2 * np.pi, 400)])
        inside = [f'        <Body type="testparticle" name="b{k}">\n          <OrbitalElement a="{e[0]!r}" e="{e[1]!r}" incl="{e[2]!r}" '
                  f'peri="{e[3]!r}" node="{e[4]!r}" M="{e[5]!r}" distanceUnit="au" angleUnit="radian" />\n        </Body>\n'
                  for k, e in enumerate(el.tolist())]
        write_bodies(str(tmp_path / "bodies.bin"), 7, el, kind=1)
    xml_all = xmlgen.make("all in the XML", "DormandPrince", "12", "4", planets + inside, events=ev)
    xml_few = xmlgen.make("planets in the XML", "DormandPrince", "12", "4", planets, events=ev)
    d_all = run(DROPIN_BIN, xml_all, str(tmp_path / "all"))
    d_side = run(DROPIN_BIN, xml_few, str(tmp_path / "side"), {"SOLARIS_B200_BODIES": str(tmp_path / "bodies.bin")})
    pa, ps = os.path.join(d_all, "Phases.dat"), os.path.join(d_side, "Phases.dat")
    ph_a, ph_s = read_phases(pa), read_phases(ps)
    assert len(ph_a) == len(ph_s) >= 3 and len(ph_a[0][1]) == 403
    if kind == "phases":
        assert open(pa, "rb").read() == open(ps, "rb").read()
    else:
        for k, ((t_a, id_a, y_a), (t_s, id_s, y_s)) in enumerate(zip(ph_a, ph_s)):
            assert np.array_equal(id_a, id_s) and abs(t_a - t_s) <= 1e-9 * max(abs(t_a), 1.0)
            # the initial phases agree to the device's sin / cos (1e-11, as sol_elements_to_phases is tested); 12 years
            # next to Jupiter then amplify that difference in the initial conditions, as they would any other
            assert np.abs(y_a - y_s).max() <= (1e-10 if k == 0 else 1e-6) * np.abs(y_a).max()
    # ... and against the reference program on the all-XML input (events included)
    if os.path.exists(REF_BIN):
        d_ref = run(REF_BIN, xml_all, str(tmp_path / "ref"))
        if kind == "phases":
            compare_outputs(d_ref, d_side, "side-loaded " + kind)


@pytest.mark.skipif(not os.path.exists(DROPIN_BIN), reason="prebuilt drop-in program missing")
def test_side_loaded_planetesimals_with_drag(tmp_path):
    """Planetesimals that feel gas drag (mass, radius, density, cD from the file; gamma_Stokes / gamma_Epstein formed like
    Simulator::BodyListToBodyData does) against the same bodies inside the XML, RK4 with the step from ShortestPeriod."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from write_bodies import write_bodies
    from solaris_b200 import synth
    s = synth.planetesimal_drag(300)
    y, m, R, cd = s.y0[2:], s.mass[2:], s.radius[2:], s.cD[2:]
    dens = m / (4.0 / 3.0 * np.pi * R ** 3)
    inside = [_particle_xml(k, "planetesimal", y[k],
                            f'          <Characteristics cd="{float(cd[k])!r}">\n            <Mass value="{float(m[k])!r}" unit="solar" />\n'
                            f'            <Radius value="{float(R[k])!r}" unit="au" />\n          </Characteristics>\n') for k in range(len(y))]
    write_bodies(str(tmp_path / "bodies.bin"), 6, y, kind=0, mass=m, radius=R, density=dens, cD=cd)
    planets = [xmlgen.planet("Jupiter")]
    xml_all = xmlgen.make("all in the XML", "RungeKutta4", "0.05", "0.01", planets + inside, nebula=True)
    xml_few = xmlgen.make("planet in the XML", "RungeKutta4", "0.05", "0.01", planets, nebula=True)
    d_all = run(DROPIN_BIN, xml_all, str(tmp_path / "all"))
    d_side = run(DROPIN_BIN, xml_few, str(tmp_path / "side"), {"SOLARIS_B200_BODIES": str(tmp_path / "bodies.bin")})
    ph_a, ph_s = read_phases(os.path.join(d_all, "Phases.dat")), read_phases(os.path.join(d_side, "Phases.dat"))
    assert len(ph_a) == len(ph_s) >= 3 and len(ph_a[0][1]) == 302
    for (t_a, id_a, y_a), (t_s, id_s, y_s) in zip(ph_a, ph_s):
        assert np.array_equal(id_a, id_s) and t_a == t_s              # same h0 = P_min / 50000 (ShortestPeriod hook)
        assert np.abs(y_a - y_s).max() <= 1e-12 * np.abs(y_a).max()   # the XML loader recomputes the density from mass and radius
    assert np.abs(ph_a[-1][2] - ph_a[0][2]).max() > 0


@pytest.mark.skipif(not os.path.exists(DROPIN_BIN), reason="prebuilt drop-in program missing")
@pytest.mark.parametrize("name", ["sunjupiter_rkf78", "outer_dp", "inner_rk4", "events_ejection_hitcentrum"])
def test_run_ahead_mode(tmp_path, name):
    """SOLARIS_B200_RUN_AHEAD=1: for systems of <= 32 massive bodies the first Driver call of a stretch lets sol_run's
    persistent kernel take every step up to the next snapshot / event / end, and the following Driver calls hand the
    recorded steps to the unmodified Simulator one by one (solaris_b200/host/sol_bridge.h).  Same snapshots at the same
    times, same events; RK4 (no step-size formula) byte for byte, the adaptive integrators to 1e-8 (their step sizes
    come from the device's pow here)."""
    xml = CASES[name]
    log = []
    d_def = run(DROPIN_BIN, xml, str(tmp_path / "default"))
    d_run = run(DROPIN_BIN, xml, str(tmp_path / "ahead"), {"SOLARIS_B200_RUN_AHEAD": "1", "SOLARIS_B200_STATS": "1"}, log)
    m = re.search(r"(\d+) steps, (\d+) state downloads, .* (\d+) run-ahead launches", log[0])
    steps, downloads, batches = (int(v) for v in m.groups())
    assert 0 < batches < steps / 4 and downloads <= batches          # the device was asked a few times, not once per step
    ev_d, ev_r = read_events(os.path.join(d_def, "TwoBodyAffair.dat")), read_events(os.path.join(d_run, "TwoBodyAffair.dat"))
    assert [(e[1], e[2], e[3]) for e in ev_d] == [(e[1], e[2], e[3]) for e in ev_r]
    ph_d, ph_r = read_phases(os.path.join(d_def, "Phases.dat")), read_phases(os.path.join(d_run, "Phases.dat"))
    if name == "inner_rk4":
        assert open(os.path.join(d_def, "Phases.dat"), "rb").read() == open(os.path.join(d_run, "Phases.dat"), "rb").read()
    assert abs(len(ph_d) - len(ph_r)) <= 1
    for (t_d, id_d, y_d), (t_r, id_r, y_r) in zip(ph_d, ph_r):
        if abs(t_d - t_r) > 1e-6:
            break                                                      # one extra ulp-sized step before a snapshot (see above)
        assert np.array_equal(id_d, id_r)
        assert np.abs(y_d - y_r).max() <= 1e-8 * np.abs(y_d).max()

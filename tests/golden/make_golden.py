"""Generates tests/golden/*.npz from the COMPILED, UNMODIFIED reference (oracle/_ref/libref_harness.so,
built by oracle/build_ref.sh from /root/reference).  Run in the build container only:

    python tests/golden/make_golden.py           # all fixtures
    python tests/golden/make_golden.py phases    # only io/phases_writer.npz
    python tests/golden/make_golden.py remove    # only io/remove_body.npz
    python tests/golden/make_golden.py elements  # only io/elements.npz

Each fixture stores the complete input system plus the reference's outputs for
Acceleration::Compute (several flag combinations, side outputs) and for a sequence of Driver calls
of the three integrators.  The fixtures travel with the repository; /root/reference does not.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from solaris_b200 import synth  # noqa: E402
from oraclelib import Reference, default_nebula  # noqa: E402

INTEGRATORS = {"rkf78": 3, "rk4": 1, "dp": 0}


def case(name, system, bary, with_nebula, steps, h0, flags_list=(7, 1, 0)):
    neb = default_nebula() if with_nebula else None
    out = {k: np.asarray(system[k]) for k in ("counts", "y0", "mass", "radius", "density", "cD", "gammaStokes",
                                              "gammaEpstein", "migStopAt", "type", "migType", "id")}
    out["barycentric"] = np.array(int(bary))
    out["with_nebula"] = np.array(int(with_nebula))
    out["t_compute"] = np.array(12.5)
    ref = Reference(system, bary, neb)
    for fl in flags_list:
        out[f"compute_f{fl}"] = ref.compute(12.5, system.y0, fl)
        rm3, idx, dist, mig = ref.side()
        out[f"rm3_f{fl}"], out[f"nnidx_f{fl}"], out[f"nndist_f{fl}"], out[f"migtype_f{fl}"] = rm3, idx, dist, mig
    for iname, icode in INTEGRATORS.items():
        ref = Reference(system, bary, neb, icode)
        t, h = 0.0, (0.01 if iname == "rk4" else h0)
        log = []
        states = []
        for _ in range(steps):
            r, t, h, hd, _, _ = ref.step(icode, t, h)
            assert r == 0
            log.append((t, h, hd))
            states.append(ref.array("y0").copy())
        out[f"{iname}_log"] = np.array(log)
        out[f"{iname}_y0_first"] = states[0]
        out[f"{iname}_y0_last"] = states[-1]
        out[f"{iname}_migtype_last"] = ref.side()[3]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {k: v.shape for k, v in out.items() if k.endswith("_log")})


def phases_writer():
    """(f) row 2: bytes written by the reference's BinaryFileAdapter::SavePhases (BINARY and TEXT) for three
    snapshots of a small state -> tests/golden/io/phases_writer.npz."""
    import tempfile
    from oraclelib import reference_save_phases
    rng = np.random.default_rng(20240601)
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for k, n in enumerate((0, 5, 64)):
            y = rng.normal(size=(n, 6)) * 10.0 ** rng.integers(-20, 20, size=(n, 6))
            if n:
                y[0] = 0.0
                y[-1, 2] = -0.0
            ids = rng.integers(0, 2 ** 31 - 1, size=n).astype(np.int32)
            t = 3652.5 * k + 0.125
            out[f"y_{k}"], out[f"id_{k}"], out[f"t_{k}"] = y, ids, np.float64(t)
            reference_save_phases(d, "Phases.dat", t, y, ids, text=False)
            reference_save_phases(d, "Phases.dat", t, y, ids, text=True)
        out["binary"] = np.frombuffer(open(os.path.join(d, "Phases.dat"), "rb").read(), dtype=np.uint8)
        out["text"] = np.frombuffer(open(os.path.join(d, "Phases.txt"), "rb").read(), dtype=np.uint8)
    os.makedirs(os.path.join(HERE, "io"), exist_ok=True)
    np.savez_compressed(os.path.join(HERE, "io", "phases_writer.npz"), **out)
    print("wrote io/phases_writer", len(out["binary"]), len(out["text"]))


def remove_body():
    """(f) row 3: Simulator::RemoveBody on a 45-body mixed system, five victims in sequence, state after each
    -> tests/golden/io/remove_body.npz."""
    s = synth.mixed([1, 2, 3, 5, 4, 20, 10], migration=True, seed=9)
    s.id = (np.arange(s.n, dtype=np.int32) * 31 + 5)
    r = Reference(s, False, default_nebula())
    out = {k: np.asarray(s[k]) for k in ("counts", "y0", "mass", "radius", "density", "cD", "gammaStokes", "gammaEpstein",
                                         "migStopAt", "type", "migType", "id")}
    victims = []
    for step, victim in enumerate((s.n - 1, 1, 7, 20, 3)):
        bid = int(r.params()["id"][victim])
        victims.append(bid)
        assert r.remove_body(bid) == 0
        for k, v in r.params().items():
            out[f"after{step}_{k}"] = v
        out[f"after{step}_y0"] = r.array("y0")
        out[f"after{step}_accel"] = r.compute(1.0, r.array("y0"), 7)
    out["victim_ids"] = np.array(victims, dtype=np.int32)
    os.makedirs(os.path.join(HERE, "io"), exist_ok=True)
    np.savez_compressed(os.path.join(HERE, "io", "remove_body.npz"), **out)
    print("wrote io/remove_body", victims)


def elements():
    """(f) row 4: Ephemeris::CalculatePhase for 2000 element sets (edge cases included) -> tests/golden/io/elements.npz."""
    from oraclelib import reference_elements_to_phases
    from test_oracle_vs_reference import elements_sample
    mu, el = elements_sample(2000, 20240601)
    out, bad = reference_elements_to_phases(mu, el)
    os.makedirs(os.path.join(HERE, "io"), exist_ok=True)
    np.savez_compressed(os.path.join(HERE, "io", "elements.npz"), mu=mu, elements=el, phases=out, failed=np.int32(bad))
    print("wrote io/elements", bad, "non-converged")


if __name__ == "__main__":
    if "elements" in sys.argv[1:]:
        elements()
        sys.exit(0)
    if "phases" in sys.argv[1:] or "remove" in sys.argv[1:]:
        if "phases" in sys.argv[1:]:
            phases_writer()
        if "remove" in sys.argv[1:]:
            remove_body()
        sys.exit(0)
    case("sunjupiter_ac", synth.mixed([1, 1, 0, 0, 0, 0, 0], migration=False), False, False, 40, 0.05, flags_list=(0,))
    case("solar9_ac", synth.solar_system(), False, False, 30, 0.05, flags_list=(0,))
    case("mixed66_ac_nebula", synth.mixed([1, 2, 3, 5, 4, 20, 31], migration=True), False, True, 12, 0.05)
    case("solar9_bc", synth.to_barycentric(synth.solar_system()), True, False, 20, 0.05, flags_list=(0,))
    case("disk200_bc", synth.to_barycentric(synth.massive_disk(200)), True, False, 5, 0.05, flags_list=(0,))
    case("drag300_ac_nebula", synth.planetesimal_drag(300), False, True, 8, 0.05)
    phases_writer()
    remove_body()
    elements()

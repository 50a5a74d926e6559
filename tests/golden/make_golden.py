"""Generates tests/golden/*.npz from the COMPILED, UNMODIFIED reference (oracle/_ref/libref_harness.so,
built by oracle/build_ref.sh from /root/reference).  Run in the build container only:

    python tests/golden/make_golden.py

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


if __name__ == "__main__":
    case("sunjupiter_ac", synth.mixed([1, 1, 0, 0, 0, 0, 0], migration=False), False, False, 40, 0.05, flags_list=(0,))
    case("solar9_ac", synth.solar_system(), False, False, 30, 0.05, flags_list=(0,))
    case("mixed66_ac_nebula", synth.mixed([1, 2, 3, 5, 4, 20, 31], migration=True), False, True, 12, 0.05)
    case("solar9_bc", synth.to_barycentric(synth.solar_system()), True, False, 20, 0.05, flags_list=(0,))
    case("disk200_bc", synth.to_barycentric(synth.massive_disk(200)), True, False, 5, 0.05, flags_list=(0,))
    case("drag300_ac_nebula", synth.planetesimal_drag(300), False, True, 8, 0.05)

"""Secondary measurements on the BASELINE.json configs C1..C5 (parity-test configs, not the headline):
steps/s, pairs/s and achieved HBM GB/s of the stage / error / finalize kernel families against the
algorithmic byte counts of SURVEY.md §8(d).  Run under gpurun; prints one JSON object per config.

    python tools/bench_configs.py [c1 c2 c3 c4 c5]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

from solaris_b200 import capi, synth

if os.environ.get("SOLARIS_B200_LIB"):          # A/B builds from tools/build_variant.py
    capi.LIB_PATH = os.environ["SOLARIS_B200_LIB"]

PEAKS = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}

# algorithmic bytes per ATTEMPT and per body of the stage + solution/error kernels
ALG_BYTES = {capi.RUNGE_KUTTA_FEHLBERG78: 4464.0, capi.RUNGE_KUTTA4: 720.0, capi.DORMAND_PRINCE: 1800.0}


def shortest_period(s):
    r = np.sqrt((s.y0[1:, :3] ** 2).sum(axis=1)); v2 = (s.y0[1:, 3:] ** 2).sum(axis=1)
    mu = synth.GAUSS2 * (1.0 + s.mass[1:])
    a = 1.0 / (2.0 / r - v2 / mu)
    return float((2 * np.pi * np.sqrt(a ** 3 / mu)).min())


def run(name, s, integ, neb, nsteps, nn_mode=2, warm=5):
    ctx = capi.Context(0)
    ctx.set_frame(False); ctx.set_nn_tracking(nn_mode); ctx.set_bodies(s); ctx.set_nebula(neb)
    t, h = 0.0, shortest_period(s) / 50000.0
    for _ in range(warm):
        rc, t, h, hd, att, em, ev, pr = ctx.step(integ, t, h)
        assert rc == 0, ctx.last_error()
    ctx.profile_read(True); ctx.profile_enable(True)
    l0 = ctx.launch_count()
    t0 = time.perf_counter()
    att_tot = ev_tot = pr_tot = 0
    for _ in range(nsteps):
        rc, t, h, hd, att, em, ev, pr = ctx.step(integ, t, h)
        assert rc == 0, ctx.last_error()
        att_tot += att; ev_tot += ev; pr_tot += pr
    wall = time.perf_counter() - t0
    ms, cnt = ctx.profile_read(True)
    ctx.profile_enable(False)
    stage_ms = ms[2] + ms[3] + ms[4]      # the stage combinations are formed by the finalize kernel of the previous evaluation
    out = {"config": name, "n": int(s.n), "integrator": {0: "DormandPrince", 1: "RungeKutta4", 3: "RungeKuttaFehlberg78"}[integ],
           "steps": nsteps, "attempts": att_tot, "steps_per_s": nsteps / wall, "ms_per_step_wall": 1e3 * wall / nsteps,
           "pairs_per_s": pr_tot / wall, "launches_per_step": (ctx.launch_count() - l0) / nsteps,
           "kernel_ms_per_step": {k: v / nsteps for k, v in zip(("pair", "prep_indirect_fold", "finalize", "rk_stage", "solution_error", "misc"), ms)},
           "finalize_stage_error_GBps": ((ALG_BYTES[integ] * att_tot + (6 + 3 + 6) * 8.0 * ev_tot) * s.n / (stage_ms * 1e-3) / 1e9) if stage_ms > 0 else None,
           "hbm_peak_GBps": PEAKS.get("hbm_gbs")}
    print(json.dumps(out), flush=True)
    ctx.close()


if __name__ == "__main__":
    import oraclelib
    which = sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5"]
    neb = oraclelib.default_nebula()
    if "c1" in which:
        run("C1 SunJupiter", synth.mixed([1, 1, 0, 0, 0, 0, 0], migration=False), capi.RUNGE_KUTTA_FEHLBERG78, None, 2000)
    if "c2" in which:
        run("C2 SolarSystem", synth.solar_system(), capi.RUNGE_KUTTA_FEHLBERG78, None, 2000)
    if "c3" in which:
        run("C3 planetesimals+drag", synth.planetesimal_drag(100_000), capi.RUNGE_KUTTA4, neb, 300)
    if "c4" in which:
        run("C4 trojans", synth.trojans(1_000_000), capi.DORMAND_PRINCE, None, 100)
    if "c5" in which:
        run("C5 disk 2^18 + type I", synth.massive_disk(262_144, migration=True), capi.RUNGE_KUTTA_FEHLBERG78, neb, 3, warm=2)

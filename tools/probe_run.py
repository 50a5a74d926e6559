"""sol_run throughput on the small BASELINE configs (C1 SunJupiter, C2 SolarSystem): steps/s of the persistent one-warp
kernel beside the single-step path and the compiled reference on one host core.  python tools/probe_run.py [steps]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from solaris_b200 import capi, synth          # noqa: E402
import oraclelib                              # noqa: E402

nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
for name, s in (("C1 SunJupiter", synth.mixed([1, 1, 0, 0, 0, 0, 0], migration=False)), ("C2 SolarSystem", synth.solar_system())):
    for integ, iname in ((capi.RUNGE_KUTTA_FEHLBERG78, "RKF78"), (capi.RUNGE_KUTTA4, "RK4"), (capi.DORMAND_PRINCE, "RKN76")):
        ctx = capi.Context(0)
        ctx.set_frame(False); ctx.set_bodies(s); ctx.set_nebula(None)
        h0 = 0.05
        rc, a, _ = ctx.run(integ, 0.0, h0, 200)             # warm-up
        t0 = time.perf_counter()
        rc, a, _ = ctx.run(integ, a.time, a.h_next, nsteps)
        dt = time.perf_counter() - t0
        assert rc == 0 and a.steps == nsteps
        run_rate = nsteps / dt
        t, h = a.time, a.h_next
        t0 = time.perf_counter()
        for _ in range(2000):
            rc, t, h, *_ = ctx.step(integ, t, h)
        step_rate = 2000 / (time.perf_counter() - t0)
        ref_rate = None
        if oraclelib.reference_available():
            r = oraclelib.Reference(s, False, None, integ)
            tt, hh = 0.0, h0
            for _ in range(200):
                _, tt, hh, *_ = r.step(integ, tt, hh)
            t0 = time.perf_counter()
            for _ in range(5000):
                _, tt, hh, *_ = r.step(integ, tt, hh)
            ref_rate = 5000 / (time.perf_counter() - t0)
        print(f"{name:16s} {iname:6s} sol_run {run_rate:10.0f} steps/s ({1e6 / run_rate:6.2f} us/step, {a.attempts / nsteps:.3f} attempts/step) | "
              f"sol_step loop {step_rate:8.0f} | reference 1 core {ref_rate if ref_rate else float('nan'):9.0f} (incl. ctypes call)", flush=True)
        ctx.close()

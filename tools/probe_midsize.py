"""Mid-size self-gravitating systems (the general multi-launch path): steps/s, launches per step, time per launch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from solaris_b200 import capi, synth
ctx = capi.Context(0)
for n in [int(a) for a in (sys.argv[1:] or ["300", "1000", "2000", "4000", "8000"])]:
    s = synth.massive_disk(n)
    ctx.set_frame(False); ctx.set_bodies(s); ctx.set_nebula(None)
    t, h = 0.0, 0.05
    for _ in range(3):
        rc, t, h, *_ = ctx.step(capi.RUNGE_KUTTA_FEHLBERG78, t, h)
    l0 = ctx.launch_count(); steps = 30; att = 0
    t0 = time.perf_counter()
    for _ in range(steps):
        rc, t, h, hd, a, *_ = ctx.step(capi.RUNGE_KUTTA_FEHLBERG78, t, h)
        assert rc == 0
        att += a
    dt = time.perf_counter() - t0
    L = (ctx.launch_count() - l0) / steps
    print(f"N={n:6d}: {steps/dt:9.1f} steps/s  {dt/steps*1e3:8.3f} ms/step  {att/steps:.2f} attempts/step  {L:6.1f} launches/step  {dt/steps*1e6/L:6.2f} us/launch")
    ctx.profile_read(True); ctx.profile_enable(True)
    for _ in range(10):
        rc, t, h, *_ = ctx.step(capi.RUNGE_KUTTA_FEHLBERG78, t, h)
    ms, cnt = ctx.profile_read(True); ctx.profile_enable(False)
    names = ["pair", "prep/indirect/fold", "finalize", "rk_stage", "solution/error", "misc"]
    print("        per launch (us): " + ", ".join(f"{nm} {1e3*m/max(c,1):.1f} x{c//10}" for nm, m, c in zip(names, ms, cnt) if c))

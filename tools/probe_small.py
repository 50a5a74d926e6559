"""Small self-gravitating systems (33 ... 256 bodies): the single-CTA whole-attempt kernel against the general path
(graph replay), RKF78 steps/s.  python tools/probe_small.py [N ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from solaris_b200 import capi, synth
ctx = capi.Context(0)
for n in [int(a) for a in (sys.argv[1:] or ["40", "64", "100", "150", "200", "256"])]:
    s = synth.massive_disk(n)
    res = []
    for mode in (3, 0):
        ctx.set_frame(False); ctx.set_bodies(s); ctx.set_nebula(None)
        ctx.set_small_system_kernel(mode)
        t, h = 0.0, 0.05
        for _ in range(4):
            rc, t, h, *_ = ctx.step(capi.RUNGE_KUTTA_FEHLBERG78, t, h)
        steps, att = 60, 0
        t0 = time.perf_counter()
        for _ in range(steps):
            rc, t, h, hd, a, *_ = ctx.step(capi.RUNGE_KUTTA_FEHLBERG78, t, h)
            assert rc == 0
            att += a
        dt = time.perf_counter() - t0
        res.append((steps / dt, att / steps))
    ctx.set_small_system_kernel(3)
    print(f"N={n:4d}: single-CTA kernel {res[0][0]:8.1f} steps/s ({res[0][1]:.2f} att/step)   general path {res[1][0]:8.1f} steps/s ({res[1][1]:.2f} att/step)", flush=True)

"""Mid-size systems: the cooperative one-kernel-per-segment path (graph mode 2) against graph replay (1) and
launch-by-launch (0) - bit-identity of a few steps and steps/s."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from solaris_b200 import capi, synth
ctx = capi.Context(0)
sizes = [int(a) for a in (sys.argv[1:] or ["300", "1000", "2000", "4000", "8000"])]
for n in sizes:
    s = synth.massive_disk(n)
    res = {}
    for mode in (0, 1, 2):
        ctx.set_frame(False); ctx.set_bodies(s); ctx.set_nebula(None)
        ctx.set_graph_mode(mode)
        t, h = 0.0, 0.05
        log = []
        for _ in range(4):
            rc, t, h, hd, a, em, *_ = ctx.step(capi.RUNGE_KUTTA_FEHLBERG78, t, h)
            assert rc == 0, ctx.last_error()
            log.append((t, h, hd, a, em))
        y = ctx.download(capi.Y0)
        steps = 40; att = 0
        l0 = ctx.launch_count()
        t0 = time.perf_counter()
        for _ in range(steps):
            rc, t, h, hd, a, *_ = ctx.step(capi.RUNGE_KUTTA_FEHLBERG78, t, h)
            assert rc == 0, ctx.last_error()
            att += a
        dt = time.perf_counter() - t0
        res[mode] = (log, y, steps / dt, att / steps, (ctx.launch_count() - l0) / steps)
    same1 = res[0][0] == res[1][0] and np.array_equal(res[0][1], res[1][1])
    same2 = res[0][0] == res[2][0] and np.array_equal(res[0][1], res[2][1])
    if not same2:
        d = np.abs(res[0][1] - res[2][1])
        print("   MISMATCH mode 2: max |dy| = %.3e at %s; log0 %s log2 %s" % (np.nanmax(d), np.unravel_index(np.nanargmax(d), d.shape), res[0][0][:2], res[2][0][:2]))
    print(f"N={n:6d}: steps/s  launches {res[0][2]:8.1f}  graphs {res[1][2]:8.1f}  fused {res[2][2]:8.1f}   "
          f"({res[2][3]:.2f} attempts/step, {res[2][4]:.1f} launches/step fused)  identical: graphs {same1} fused {same2}", flush=True)
ctx.set_graph_mode(1)

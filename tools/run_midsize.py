"""A few RKF78 steps of a mid-size self-gravitating disk on the general path, launches issued one by one (for ncu).
python tools/run_midsize.py [N] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from solaris_b200 import capi, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = capi.Context(0)
ctx.set_frame(False); ctx.set_bodies(synth.massive_disk(n)); ctx.set_nebula(None)
ctx.set_graph_mode(0)
t, h = 0.0, 0.05
for _ in range(steps):
    rc, t, h, *_ = ctx.step(capi.RUNGE_KUTTA_FEHLBERG78, t, h)
    assert rc == 0, ctx.last_error()
print(n, steps, t, h, ctx.launch_count())

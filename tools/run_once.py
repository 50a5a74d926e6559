"""One sol_run call on a small config, for ncu captures of the persistent one-warp kernel.
python tools/run_once.py c1|c2 rkf78|rk4|dp [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from solaris_b200 import capi, synth          # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "c1"
integ = {"rkf78": capi.RUNGE_KUTTA_FEHLBERG78, "rk4": capi.RUNGE_KUTTA4, "dp": capi.DORMAND_PRINCE}[sys.argv[2] if len(sys.argv) > 2 else "rkf78"]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 300
s = synth.mixed([1, 1, 0, 0, 0, 0, 0], migration=False) if cfg == "c1" else synth.solar_system()
ctx = capi.Context(0)
ctx.set_frame(False); ctx.set_bodies(s); ctx.set_nebula(None)
rc, a, _ = ctx.run(integ, 0.0, 0.05, steps)
assert rc == 0
rc, a, _ = ctx.run(integ, a.time, a.h_next, steps)
print(cfg, steps, a.time, a.attempts)

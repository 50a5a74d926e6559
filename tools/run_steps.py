"""A few Driver steps of one BASELINE config on the device-resident path, for ncu captures.
python tools/run_steps.py h|c3|c4|c5 [steps] [bodies]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from solaris_b200 import capi, synth          # noqa: E402
import oraclelib                              # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "h"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
neb = None
if cfg == "h":
    s, integ = synth.massive_disk(int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000), capi.RUNGE_KUTTA_FEHLBERG78
elif cfg == "c3":
    s, integ, neb = synth.planetesimal_drag(100_000), capi.RUNGE_KUTTA4, oraclelib.default_nebula()
elif cfg == "c4":
    s, integ = synth.trojans(1_000_000), capi.DORMAND_PRINCE
else:
    s, integ, neb = synth.massive_disk(262_144, migration=True), capi.RUNGE_KUTTA_FEHLBERG78, oraclelib.default_nebula()
ctx = capi.Context(0)
ctx.set_frame(False); ctx.set_bodies(s); ctx.set_nebula(neb)
t, h = 0.0, 1.0e-3
for _ in range(steps):
    rc, t, h, *_ = ctx.step(integ, t, h)
    assert rc == 0, ctx.last_error()
print(cfg, s.n, steps, t, h, ctx.launch_count())

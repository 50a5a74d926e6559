"""Small workload touching every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from solaris_b200 import capi, synth
from oraclelib import default_nebula

ctx = capi.Context(0)
ctx.set_pair_algorithm(1)        # symmetric kernel from 4096 bodies (the default switches at 12288)
neb = default_nebula()
cases = [
    (synth.massive_disk(4700), False, None, capi.RUNGE_KUTTA4),                               # symmetric kernel (+ diag, half round, padding)
    (synth.to_barycentric(synth.massive_disk(4200)), True, None, capi.RUNGE_KUTTA4),
    (synth.mixed([1, 2, 3, 5, 4, 20, 31], migration=True), False, neb, capi.RUNGE_KUTTA_FEHLBERG78),   # single-CTA kernel
    (synth.trojans(700), False, None, capi.DORMAND_PRINCE),                                   # tracer kernel
    (synth.planetesimal_drag(600), False, neb, capi.RUNGE_KUTTA_FEHLBERG78),
    (synth.mixed([1, 3, 40, 300, 100, 600, 456], migration=True, seed=77), False, neb, capi.DORMAND_PRINCE),  # general multi-launch path
]
for s, bary, nb, integ in cases:
    for nn in (1, 2):
        ctx.set_frame(bary); ctx.set_nn_tracking(nn); ctx.set_bodies(s); ctx.set_nebula(nb)
        a = ctx.compute(0.0, s.y0, capi.EVAL_ALL)
        t, h = 0.0, 0.02
        for _ in range(2):
            rc, t, h, *_ = ctx.step(integ, t, h)
            assert rc == 0
        ctx.detect_events(5.5, 5.2, 3.0)
        ctx.integrals()
        ctx.flush_tiny()
        ctx.pack_phases(t)                                   # snapshot record kernel
    if s.n > 40:
        ctx.remove_bodies([s.n - 1, 5, 17])                  # compaction kernels, then the shrunk system steps on
        rc, t, h, *_ = ctx.step(integ, t, h)
        assert rc == 0
print("sanitize run ok")

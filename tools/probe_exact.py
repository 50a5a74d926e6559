"""Where do the two device pair algorithms and the reference's double row disagree at N = 10^6, and who is right?
Compares all three with the extended-precision row oracle on the rows of largest device-device disagreement, the
worst-conditioned rows and random rows.  Usage: python tools/probe_exact.py [N]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from solaris_b200 import capi, synth          # noqa: E402
from oraclelib import Oracle                  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
s = synth.massive_disk(n)
ctx = capi.Context(0)
ctx.set_frame(False); ctx.set_bodies(s); ctx.set_nebula(None)
ctx.set_pair_algorithm(1); a_sym = ctx.compute(0.0, s.y0, 0)
ctx.set_pair_algorithm(0); a_ord = ctx.compute(0.0, s.y0, 0)
nrm = np.sqrt((a_sym[1:, 3:] ** 2).sum(axis=1))
dd = np.abs(a_sym[1:, 3:] - a_ord[1:, 3:]).max(axis=1) / nrm
r2 = (s.y0[1:, :3] ** 2).sum(axis=1)
kep = synth.GAUSS2 * (s.mass[0] + s.mass[1:]) / r2
cond = kep / nrm
print("device-device: max %.3e  >1e-13: %d  median %.3e" % (dd.max(), (dd > 1e-13).sum(), np.median(dd)))
print("cond: max %.1f, >10: %d, >100: %d" % (cond.max(), (cond > 10).sum(), (cond > 100).sum()))
rng = np.random.default_rng(11)
rows = np.unique(np.concatenate([1 + np.argsort(-dd)[:24], 1 + np.argsort(-cond)[:24], rng.integers(1, n, 2048)])).astype(np.int32)
o = Oracle(s, False, None)
ex = o.gravity_rows_exact(s.y0, rows)
en = np.sqrt((ex ** 2).sum(axis=1))
es = np.abs(a_sym[rows, 3:] - ex).max(axis=1) / en
eo = np.abs(a_ord[rows, 3:] - ex).max(axis=1) / en
print("vs exact over %d rows: sym max %.3e (>1e-13: %d) median %.3e | ord max %.3e (>1e-13: %d) median %.3e" %
      (len(rows), es.max(), (es > 1e-13).sum(), np.median(es), eo.max(), (eo > 1e-13).sum(), np.median(eo)))
order = np.argsort(-np.maximum(es, eo))[:16]
print(" row      cond     sym-exact  ord-exact  ref-exact  nn_dist")
nnd = ctx.download(capi.NN_DISTANCE)
for k in order:
    i = int(rows[k])
    ref = o.gravity_rows(s.y0, i, i + 1, 1)[0, 3:]
    er = np.abs(ref - ex[k]).max() / en[k]
    print("%8d %8.1f  %.3e  %.3e  %.3e  %.3e" % (i, cond[i - 1], es[k], eo[k], er, nnd[i]))
# error relative to the largest term actually summed (Kepler term or nearest-neighbour pull)
big = np.maximum(kep[rows - 1], synth.GAUSS2 * s.mass.max() / np.maximum(nnd[rows], 1e-300) ** 2)
print("relative to the largest term: sym max %.3e, ord max %.3e" % ((es * en / big).max(), (eo * en / big).max()))

import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np
from solaris_b200 import capi, synth
from helpers import *
from oraclelib import Oracle, default_nebula
ctx = capi.Context(0)
for name, s, bary, integ in [("bc-disk", synth.to_barycentric(synth.massive_disk(300)), True, 0),
                       ("ac-planets", synth.mixed([1,4,4,0,0,0,0], migration=False), False, 3),
                       ("ac-planets", synth.mixed([1,4,4,0,0,0,0], migration=False), False, 0)]:
    configure(ctx, s, bary, None)
    o = Oracle(s, bary, None)
    t_g=t_o=0.0; h_g=h_o=0.05
    print(name, integ)
    for k in range(25):
        r_o, t_o, h_o, hd_o, att_o, em_o = o.step(integ, t_o, h_o)
        r_g, t_g, h_g, hd_g, att_g, em_g, ev, pr = ctx.step(integ, t_g, h_g)
        yg = ctx.download(capi.Y0); yo = o.array('y0')
        print(k, att_o, att_g, "hd %.6e %.2e  hn %.6e %.2e  em %.6e %.6e  staterr %.2e" % (hd_o, abs(hd_g-hd_o)/abs(hd_o), h_o, abs(h_g-h_o)/abs(h_o), em_o, em_g, rel_state_error(yg, yo)))

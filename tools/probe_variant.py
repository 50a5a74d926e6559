"""A/B helper: time the pair kernel of alternative builds of the library (SOLARIS_B200_LIB=<path>)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from solaris_b200 import capi, synth
if os.environ.get("SOLARIS_B200_LIB"):
    capi.LIB_PATH = os.environ["SOLARIS_B200_LIB"]
ctx = capi.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
s = synth.massive_disk(n)
ctx.set_frame(False); ctx.set_bodies(s); ctx.set_nebula(None)
if os.environ.get("SOL_ORDERED"):
    ctx.set_pair_algorithm(0)      # ordered pair kernel only
for nn in (0, 1):
    ctx.set_nn_tracking(nn)
    ms, pairs = ctx.time_gravity_kernel(3)
    print(f"{os.path.basename(capi.LIB_PATH)} N={n} nn={nn}: {ms:.2f} ms  {pairs/(ms*1e-3):.4e} pairs/s", flush=True)

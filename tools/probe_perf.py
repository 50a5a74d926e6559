"""Quick performance probe (run under gpurun): FP64 peak + pair-kernel rate at a few sizes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from solaris_b200 import capi, synth

ctx = capi.Context(0)
peak = ctx.measure_fp64_peak()
print(f"fp64 DFMA peak measured: {peak:.2f} TFLOP/s")
for n in [int(a) for a in (sys.argv[1:] or ["16384", "65536", "262144", "1000000"])]:
    t0 = time.time()
    s = synth.massive_disk(n)
    tg = time.time() - t0
    ctx.set_frame(False); ctx.set_bodies(s); ctx.set_nebula(None)
    for nn, alg in ((1, 1), (0, 1)):
        ctx.set_nn_tracking(nn); ctx.set_pair_algorithm(alg)
        reps = 3 if n >= 500000 else 10
        ms, pairs = ctx.time_gravity_kernel(reps)
        rate = pairs / (ms * 1e-3)
        print(f"N={n:8d} nn={nn} sym={alg} pair kernel {ms:10.3f} ms  {rate:.3e} pairs/s  {rate*20/1e12:.2f} TFLOP/s alg ({rate*20/1e12/peak*100:.1f}% of peak; pipe-instr {rate*16*2/1e12/peak*100:.1f}%)  [gen {tg:.1f}s]")

"""SURVEY.md §8(f) rank 2 - snapshot writer: sol_write_phases (record assembled on the device, one transfer, one
write) against the compiled reference's BinaryFileAdapter::SavePhases (oracle/_ref, 2 n stream writes) on the
same state, both appending to a file in a fresh temporary directory.  Run under gpurun; prints one JSON object.

    python tools/bench_phases.py [N]
"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

from solaris_b200 import capi, synth
import oraclelib


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    s = synth.trojans(n)
    ctx = capi.Context(0)
    ctx.set_frame(False); ctx.set_bodies(s); ctx.set_nebula(None)
    reps = 5
    res = {"bodies": int(s.n), "record_bytes": 12 + 52 * int(s.n)}
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "Phases.dat")
        ctx.write_phases(p, 0.0)                                   # warm-up: pinned buffer, file creation
        t0 = time.perf_counter()
        for k in range(reps):
            ctx.write_phases(p, 1.0 + k)
        res["device_ms_per_snapshot"] = (time.perf_counter() - t0) / reps * 1e3
        ctx.profile_read(True); ctx.profile_enable(True)
        for k in range(reps):
            ctx.pack_phases(0.0)
        prof = ctx.profile_read(True); ctx.profile_enable(False)
        res["pack_kernel_ms"] = prof[0][5] / reps
        res["pack_kernel_GBps"] = 2 * res["record_bytes"] / (res["pack_kernel_ms"] * 1e-3) / 1e9
        size_dev = os.path.getsize(p)
        if oraclelib.reference_available():
            q = "PhasesRef.dat"
            oraclelib.reference_save_phases(d, q, 0.0, s.y0, s.id)
            t0 = time.perf_counter()
            for k in range(reps):
                oraclelib.reference_save_phases(d, q, 1.0 + k, s.y0, s.id)
            res["reference_ms_per_snapshot"] = (time.perf_counter() - t0) / reps * 1e3
            res["files_identical"] = open(p, "rb").read() == open(os.path.join(d, q), "rb").read()
            res["speedup"] = res["reference_ms_per_snapshot"] / res["device_ms_per_snapshot"]
        res["file_bytes"] = size_dev
    print(json.dumps(res))


if __name__ == "__main__":
    main()

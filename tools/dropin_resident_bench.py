"""Whole-program timing of the DROP-IN (solaris_b200/host/_build/solaris_b200_dropin): eager host
synchronisation (default) against SOLARIS_B200_RESIDENT=1 on a tracer-dominated system (Sun + Jupiter + N test
particles between 2 and 3.2 au, RKN7(6), ejection radius set so that event detection is active every step).
The start-up cost (XML parse, BodyList construction, initial snapshot) is removed by running two lengths and
differencing.  (The reference program's own start-up is O(N^2): ~1.4 s at N = 8000, minutes at 10^5 - keep N modest.)
Run under gpurun; prints one JSON object.

    python tools/dropin_resident_bench.py [N]
"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import xmlgen

DROPIN = os.path.join(ROOT, "solaris_b200", "host", "_build", "solaris_b200_dropin")


def particles(n, seed=5):
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n):
        a, e = rng.uniform(2.0, 3.2), rng.uniform(0.0, 0.05)      # main-belt like: no close encounters, regular steps
        out.append(f'        <Body type="testparticle" name="t{k}">\n'
                   f'          <OrbitalElement a="{a!r}" e="{e!r}" incl="{rng.uniform(0, 5)!r}" peri="{rng.uniform(0, 360)!r}" '
                   f'node="{rng.uniform(0, 360)!r}" M="{rng.uniform(0, 360)!r}" distanceUnit="au" angleUnit="degree" />\n'
                   f'        </Body>\n')
    return out


def run(xml, env_extra):
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "in.xml")
        open(p, "w").write(xml)
        env = dict(os.environ, OSTYPE="linux"); env.update(env_extra)
        t0 = time.perf_counter()
        r = subprocess.run([DROPIN, "-i", p], cwd=d, env=env, capture_output=True, text=True, timeout=1200)
        dt = time.perf_counter() - t0
        assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
        phases = open(os.path.join(d, "Phases.dat"), "rb").read()
        note = [l for l in r.stderr.splitlines() if "state downloads" in l]
        return dt, phases, note


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    short, long_ = (20, 220) if n <= 50000 else (20, 120)
    parts = particles(n)
    ev = '    <Ejection value="100" unit="au" />\n'
    res = {"bodies": n + 2, "integrator": "DormandPrince"}
    out = {}
    for years in (short, long_):
        xml = xmlgen.make("tracers", "DormandPrince", str(years), str(years // 2), [xmlgen.planet("Jupiter")] + parts, events=ev)
        for mode, env in (("eager", {}), ("resident", {"SOLARIS_B200_RESIDENT": "1"})):
            run(xml, env) if years == short and mode == "eager" else None      # page the binary / driver in once
            dt, ph, note = run(xml, env)
            out[(years, mode)] = (dt, ph, note)
        assert out[(years, "eager")][1] == out[(years, "resident")][1], "resident and eager snapshots differ"
    for mode in ("eager", "resident"):
        res[mode + "_s_per_long_minus_short"] = out[(long_, mode)][0] - out[(short, mode)][0]
        res[mode + "_wall_s"] = [out[(short, mode)][0], out[(long_, mode)][0]]
    res["resident_note"] = out[(long_, "resident")][2]
    res["speedup"] = res["eager_s_per_long_minus_short"] / res["resident_s_per_long_minus_short"]
    res["snapshots_identical"] = True
    print(json.dumps(res))


if __name__ == "__main__":
    main()

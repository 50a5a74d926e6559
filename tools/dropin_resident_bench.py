"""Whole-program timing of the DROP-IN (solaris_b200/host/_build/solaris_b200_dropin): eager host
synchronisation (default) against SOLARIS_B200_RESIDENT=1 on a tracer-dominated system (Sun + Jupiter + N test
particles between 2 and 3.2 au, RKN7(6), ejection radius set so that event detection is active every step).
Reported: the seconds spent inside the Driver calls, by phase (the bridge's own timers, SOLARIS_B200_STATS), and the
program's wall time, which is dominated by the reference's XML loader at these sizes.  Run under gpurun; prints one
JSON object.

    python tools/dropin_resident_bench.py [N]
"""
import json
import os
import re
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


def driver_seconds(note):
    m = re.search(r"(\d+) steps, (\d+) state downloads, (\d+) event edits.*sync_in ([\d.]+), sol_step ([\d.]+), detect ([\d.]+), sync_out ([\d.]+)", note[0])
    steps, downloads, edits = (int(m.group(k)) for k in (1, 2, 3))
    parts = [float(m.group(k)) for k in (4, 5, 6, 7)]
    return {"steps": steps, "state_downloads": downloads, "sync_in_s": parts[0], "sol_step_s": parts[1], "detect_s": parts[2],
            "sync_out_s": parts[3], "driver_total_s": sum(parts)}


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    years = 220 if n <= 50000 else 120
    parts = particles(n)
    ev = '    <Ejection value="100" unit="au" />\n'
    xml = xmlgen.make("tracers", "DormandPrince", str(years), str(years // 2), [xmlgen.planet("Jupiter")] + parts, events=ev)
    res = {"bodies": n + 2, "integrator": "DormandPrince", "years": years}
    run(xml, {})                                                   # page the binary / driver in once
    out = {}
    for mode, env in (("eager", {"SOLARIS_B200_STATS": "1"}), ("resident", {"SOLARIS_B200_RESIDENT": "1"})):
        dt, ph, note = run(xml, env)
        out[mode] = ph
        res[mode] = driver_seconds(note)
        res[mode]["program_wall_s"] = dt                           # includes the reference's XML loader and start-up
    assert out["eager"] == out["resident"], "resident and eager snapshots differ"
    res["snapshots_identical"] = True
    res["driver_time_ratio"] = res["eager"]["driver_total_s"] / res["resident"]["driver_total_s"]
    print(json.dumps(res))


if __name__ == "__main__":
    main()

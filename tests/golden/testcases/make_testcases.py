"""Fixtures: the reference's OWN shipped scenarios (TestCases/*/*.xml) prepared for whole-program
parity runs of both binaries on a box that has no /root/reference.

    python tests/golden/testcases/make_testcases.py [/root/reference]

For every scenario the reference's parser accepts, the input file is copied with exactly these edits and nothing else:
  * `guid="..."` attributes removed  (Tools::GuidToCharArray overflows a heap buffer when one is saved),
  * `epoch="..."` attributes removed (DecisionMaking compares absolute Julian dates with `length`),
  * <TimeLine length/output> shortened where the shipped horizon is 10^5 - 10^6 years (recorded in MANIFEST.json),
so that a run takes seconds and ends inside the horizon over which two floating-point implementations of an adaptive
integrator can be compared at 1e-10.  TestCases/SolarSystemWithBalint/SS.data (the one stored state of the reference) is
copied as SS.data.  Scenarios the reference itself cannot run are listed with the reason and skipped."""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

# name -> (path under TestCases, new length or None, new output or None)
# Horizons: a few thousand accepted RKF78 steps each - beyond that two correct double-precision implementations of an
# adaptive integrator drift apart by more than the north star's 1e-10 (measured: 1.3e-10 in e after 5600 steps of
# Sun-Jupiter-Saturn, 4e-10 in a after 14000 steps of the migrating pair).
CASES = {
    "SunJupiter": ("SunJupiter/SunJupiter.xml", "1e3", None),                       # 2900 steps (BASELINE configs[0])
    "SolarSystem": ("SolarSystem/SolarSystem.xml", "20", "1"),                      # 2900 steps (BASELINE configs[1])
    "SunJupiterSaturn": ("SunJupiterSaturn/SunJupiterSaturn.xml", "1e3", "100"),
    "SJN": ("SJN/SJN.xml", "1e3", "100"),
    "OSS": ("OSS/OSS.xml", "1e3", "100"),
    "MigratingJupiter": ("MigratingJupiter/MigratingJupiter.xml", "5e3", None),
    "JupiterMigratingSaturn": ("JupiterMigratingSaturn/JupiterMigratingSaturn.xml", "1e3", None),
    # Jupiter and a Saturn-mass clone 0.1 au apart in a: close encounters scatter them at ~5, ~55 and ~175 yr (a jumps
    # 5.2 / 5.3 -> 5.4 / 4.7 -> 4.8 / 7.4 au); every encounter multiplies rounding differences by orders of magnitude
    "CollisionTest": ("CollisionTest/CollisionTest.xml", "40", None),
    "EjectionTest": ("EjectionTest/EjectionTest.xml", "1e3", None),
    "HitCentrumTest": ("HitCentrumTest/HitCentrumTest.xml", "1e3", None),
    "PlanetesimalWithDrag": ("PlanetesimalWithDrag/PlanetesimalWithDrag.xml", "1e2", "1"),
    # Epstein regime until ~2855 yr, then the transition regime with cd = 0: log10(0) -> NaN in the reference (and here)
    "UnifiedDragForce_r50AUR2m": ("UnifiedDragForce/r50AUR2m/PlConstant.xml", "3000", None),
}
SKIPPED = {
    "TypeIMigration/TypeIMigration.xml": "load error in the reference: unknown attribute `inc` (SURVEY.md Q20)",
    "UnifiedDragForce/r5AUR2m/PlConstant.xml": "load error in the reference: unknown attribute `path` (SURVEY.md Q20)",
    "Skeleton.xml": "template, XML syntax error",
    "JupiterSaturnWithJupiterTrojans/L4_T1e7/JupiterL4Trojans.xml": "the reference aborts in SaveConstantProperty (guid heap overflow, Q19); 1.3 MB",
    "JupiterSaturnWithJupiterTrojans/L5_T1e7/JupiterL5Trojans.xml": "same",
    "JupiterSaturnWithJupiterTrojans/JupiterSaturn.xml": "ends at t = length measured from JD 0 (Q18): one step",
    "OSSSynchron/OSS.xml": "finishes immediately (epoch / length interplay, Q18)",
    "PlanetesimalWithDragCopy*/PlanetesimalWithDrag.xml": "byte-identical copies of PlanetesimalWithDrag",
}

manifest = {"source": "suliaron/solaris TestCases/", "cases": {}, "skipped": SKIPPED}
for name, (rel, length, output) in CASES.items():
    raw = open(os.path.join(REF, "TestCases", rel), "rb").read().decode("utf-8-sig")
    edits = []
    xml, n = re.subn(r'\s+guid="[^"]*"', "", raw)
    if n:
        edits.append(f"removed {n} guid attribute(s)")
    xml, n = re.subn(r'\s+epoch="[^"]*"', "", xml)
    if n:
        edits.append(f"removed {n} epoch attribute(s)")
    m = re.search(r'<TimeLine\s+length="([^"]*)"\s+output="([^"]*)"', xml)
    assert m, name
    old_len, old_out = m.group(1), m.group(2)
    new_len, new_out = length or old_len, output or old_out
    if (new_len, new_out) != (old_len, old_out):
        xml = xml.replace(m.group(0), f'<TimeLine length="{new_len}" output="{new_out}"', 1)
        edits.append(f"TimeLine length {old_len} -> {new_len}, output {old_out} -> {new_out}")
    open(os.path.join(HERE, name + ".xml"), "w").write(xml)
    manifest["cases"][name] = {"from": "TestCases/" + rel, "edits": edits}
ss = open(os.path.join(REF, "TestCases", "SolarSystemWithBalint", "SS.data")).read()
open(os.path.join(HERE, "SS.data"), "w").write(ss)
manifest["SS.data"] = "TestCases/SolarSystemWithBalint/SS.data, verbatim: 9 bodies x (position line, velocity line)"
json.dump(manifest, open(os.path.join(HERE, "MANIFEST.json"), "w"), indent=1)
print("wrote", len(CASES), "scenarios +", "SS.data")

"""Small Solaris input files for whole-program parity runs (reference program vs. drop-in program).
Written from scratch following the grammar the reference's parser accepts; no `guid`
and no `epoch` attributes.  Orbital elements are public ephemeris values."""
import numpy as np

HEADER = """<Simulation name="{name}" description="{name}">
  <Settings enableDistinctStartTimes="False"{bary}>
    <Output>
      <Phases>Phases.dat</Phases>
      <Integrals>Integrals.dat</Integrals>
      <TwoBodyAffair>TwoBodyAffair.dat</TwoBodyAffair>
      <Log>Log.txt</Log>
    </Output>
    <Integrator xmlns:xsi="http://www.w3.org/2001/XMLSchema-instance" xmlns:xsd="http://www.w3.org/2001/XMLSchema" xsi:type="{integrator}">
      <Accuracy value="-10" />
    </Integrator>
    <TimeLine length="{length}" output="{output}" unit="year" />
{events}  </Settings>
  <BodyGroupList>
    <BodyGroup>
      <Items>
"""
FOOTER = """      </Items>
    </BodyGroup>
  </BodyGroupList>
{nebula}</Simulation>
"""

SUN = """        <Body type="centralbody" name="Sun">
          <Phase>
            <Position x="0" y="0" z="0" unit="au" />
            <Velocity x="0" y="0" z="0" unit="auday" />
          </Phase>
          <Characteristics>
            <Mass value="1" unit="solar" />
          </Characteristics>
        </Body>
"""

PLANETS = {
    "Jupiter": ("giantplanet", 5.20336301, 0.04839266, 1.3053, 274.1977, 100.55615, 19.65053, "jupiter"),
    "Saturn": ("giantplanet", 9.53707032, 0.0541506, 2.48446, 338.7169, 113.71504, 317.51238, "saturn"),
    "Uranus": ("giantplanet", 19.19126393, 0.04716771, 0.76986, 96.73436, 74.22988, 142.26794, "uranus"),
    "Neptune": ("giantplanet", 30.06896348, 0.00858587, 1.76917, 273.24966, 131.72169, 259.90868, "neptune"),
    "Earth": ("rockyplanet", 1.00000011, 0.01671022, 0.00005, 114.20783, 348.73936, 357.51716, "earth"),
    "Mars": ("rockyplanet", 1.52366231, 0.09341233, 1.85061, 286.4623, 49.57854, 19.41248, "mars"),
}


def body(name, btype, a, e, incl, peri, node, M, mass_value, mass_unit, extra="", migration=""):
    return f"""        <Body type="{btype}" name="{name}">
          <OrbitalElement a="{a!r}" e="{e!r}" incl="{incl!r}" peri="{peri!r}" node="{node!r}" M="{M!r}" distanceUnit="au" angleUnit="degree" />
          <Characteristics{extra[0] if extra else ""}>
            <Mass value="{mass_value!r}" unit="{mass_unit}" />
{extra[1] if extra else ""}          </Characteristics>
{migration}        </Body>
"""


def planet(name):
    t, a, e, i, w, O, M, unit = PLANETS[name]
    return body(name, t, a, e, i, w, O, M, 1, unit)


NEBULA = """  <Nebula name="MMSN">
    <GasComponent alpha="0.002" type="constant">
      <Eta c="0.0019" index="0.5" />
      <Tau c="0.6666666666666666" index="2" />
      <ScaleHeight c="0.02" index="1.25" />
    </GasComponent>
  </Nebula>
"""


def make(name, integrator, length, output, bodies, events="", nebula=False, barycentric=False):
    bary = ' baryCentric="True"' if barycentric else ""
    return (HEADER.format(name=name, integrator=integrator, length=length, output=output, events=events, bary=bary) + SUN +
            "".join(bodies) + FOOTER.format(nebula=NEBULA if nebula else ""))


def cases():
    rng = np.random.default_rng(20240601)
    out = {}
    out["sunjupiter_rkf78"] = make("SunJupiter", "RungeKutta78", "100", "10", [planet("Jupiter")])
    out["outer_dp"] = make("Outer planets", "DormandPrince", "200", "20", [planet(p) for p in ("Jupiter", "Saturn", "Uranus", "Neptune")])
    out["inner_rk4"] = make("Six planets RK4", "RungeKutta4", "0.2", "0.05", [planet(p) for p in ("Jupiter", "Saturn", "Uranus", "Neptune", "Earth", "Mars")])
    out["outer_bc_rkf78"] = make("Outer planets barycentric", "RungeKutta78", "100", "10",
                                 [planet(p) for p in ("Jupiter", "Saturn", "Uranus", "Neptune")], barycentric=True)
    # events: Saturn, Uranus, Neptune start outside a 7 au ejection radius; Earth inside a 1.2 au hit-centrum radius
    out["events_ejection_hitcentrum"] = make(
        "Ejection and hit centrum", "RungeKutta78", "50", "10",
        [planet(p) for p in ("Jupiter", "Saturn", "Uranus", "Neptune", "Earth", "Mars")],
        events='    <Ejection value="7" unit="au" />\n    <HitCentrum value="1.2" unit="au" />\n')
    # collisions: protoplanet pairs on nearly identical orbits with inflated radii
    protos = []
    for k in range(8):
        a = 2.0 + 0.35 * (k // 2)
        M = 40.0 * (k // 2) + (0.0 if k % 2 == 0 else 0.02)
        extra = ("", f'            <Radius value="{2.0e4 + 100.0 * k!r}" unit="km" />\n')
        protos.append(body(f"P{k}", "protoplanet", a, 0.01 + 0.001 * (k // 2), 0.5, 10.0 * (k // 2), 20.0, M, 0.05 + 0.01 * k, "earth", extra))
    out["collisions"] = make("Collisions", "RungeKutta78", "30", "5", [planet("Jupiter")] + protos,
                             events='    <Collision factor="5" />\n')
    # a collision that fires late: two protoplanets drifting together over ~2 yr, far from any snapshot, so that the
    # device-resident drop-in has gone many steps without refreshing the host arrays when Simulator::CheckEvent
    # builds the Collision record from bodyData.y (the state BEFORE the step, Simulator.cpp:709)
    big = ("", '            <Radius value="100000.0" unit="km" />\n')
    late = [body("L0", "protoplanet", 2.0, 0.0, 0.5, 0.0, 20.0, 0.0, 1.0e-5, "earth", big),
            body("L1", "protoplanet", 2.004, 0.0, 0.5, 0.0, 20.0, 1.0, 2.0e-5, "earth", big),
            body("L2", "protoplanet", 3.1, 0.02, 0.4, 50.0, 20.0, 200.0, 0.06, "earth", big)]
    out["late_collision"] = make("Late collision", "RungeKutta78", "10", "5", [planet("Jupiter")] + late,
                                 events='    <Collision factor="5" />\n')
    # gas drag: planetesimals with density + cd (so gammaStokes / gammaEpstein != 0)
    pls = []
    for k in range(12):
        a, e = float(rng.uniform(1.5, 3.5)), float(rng.uniform(0.0, 0.1))
        extra = (' cd="1.0"', f'            <Density value="{float(rng.uniform(1.0, 2.0))!r}" unit="gcm3" />\n')
        pls.append(body(f"pl{k}", "planetesimal", a, e, 1.0, float(rng.uniform(0, 360)), float(rng.uniform(0, 360)),
                        float(rng.uniform(0, 360)), float(10.0 ** rng.uniform(3.0, 16.0)), "kg", extra))
    out["gasdrag_rk4"] = make("Planetesimals with gas drag", "RungeKutta4", "0.1", "0.02", [planet("Jupiter")] + pls, nebula=True)
    # type I migration of protoplanets
    mig = '          <Migration type="I" stopAt="0.4" />\n'
    protos = [body(f"M{k}", "protoplanet", 1.0 + 0.4 * k, 0.02, 0.3, 25.0 * k, 10.0 * k, 33.0 * k, 0.5 + 0.2 * k, "earth", migration=mig)
              for k in range(10)]
    out["migration_typeI_rkf78"] = make("Type I migration", "RungeKutta78", "300", "50", protos, nebula=True)
    # test particles + an ejection radius that never fires: event detection active on every step, no event
    # (the host's initial rm3 array is all zeros, which WOULD fire if it were read before the first download)
    tps = []
    for k in range(40):
        tps.append(f'        <Body type="testparticle" name="t{k}">\n'
                   f'          <OrbitalElement a="{float(rng.uniform(2.0, 3.2))!r}" e="{float(rng.uniform(0.0, 0.05))!r}" '
                   f'incl="{float(rng.uniform(0, 5))!r}" peri="{float(rng.uniform(0, 360))!r}" node="{float(rng.uniform(0, 360))!r}" '
                   f'M="{float(rng.uniform(0, 360))!r}" distanceUnit="au" angleUnit="degree" />\n        </Body>\n')
    out["testparticles_idle_ejection_dp"] = make("Test particles", "DormandPrince", "30", "10", [planet("Jupiter")] + tps,
                                                 events='    <Ejection value="100" unit="au" />\n')
    return out

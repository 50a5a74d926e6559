"""The drop-in's resident mode must use exactly the event thresholds the reference's loader reads from the input
file (XmlFileAdapter::DeserializeSettings, Solaris/XmlFileAdapter.cpp:215-275; UnitTool::DistanceToAu, Solaris/Units.cpp:
76-100).  solaris_b200/host/resident_config.h is free of the reference's headers, so it is compiled here into a tiny
driver and checked on the shapes the loader accepts and on the inputs it must NOT be fooled by; whole-program runs in
both modes are in tests/test_gpu_dropin_program.py."""
import os
import subprocess

import pytest

import xmlgen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = r'''
#include <cstdio>
#include <fstream>
#include <sstream>
#include "resident_config.h"
int main(int argc, char **argv)
{
	std::ifstream f(argv[1]);
	std::stringstream ss; ss << f.rdbuf();
	const solb200::UnitFactors uf = {1.0 / 1.495978707e11, 1.0 / 1.495978707e8, 1.0 / 215.094};   // Constants.h:52-69
	const solb200::EventThresholds t = solb200::read_event_thresholds(ss.str(), uf);
	printf("%d %d %.17g %.17g %.17g\n", (int)t.parsed, (int)t.doubt, t.ejection, t.hitCentrum, t.collisionFactor);
	return 0;
}
'''


@pytest.fixture(scope="module")
def reader(tmp_path_factory):
    d = tmp_path_factory.mktemp("resident_config")
    src = d / "driver.cpp"
    src.write_text(DRIVER)
    exe = d / "driver"
    subprocess.check_call(["g++", "-std=gnu++11", "-O1", "-I", os.path.join(ROOT, "solaris_b200", "host"), str(src), "-o", str(exe)])

    def run(xml):
        p = d / "in.xml"
        p.write_text(xml)
        out = subprocess.check_output([str(exe), str(p)], text=True).split()
        return bool(int(out[0])), bool(int(out[1])), float(out[2]), float(out[3]), float(out[4])
    return run


def settings(events):
    return xmlgen.make("t", "RungeKutta78", "10", "5", [xmlgen.planet("Jupiter")], events=events)


def test_generated_cases(reader):
    cases = xmlgen.cases()
    assert reader(cases["events_ejection_hitcentrum"]) == (True, False, 7.0, 1.2, 0.0)
    assert reader(cases["collisions"]) == (True, False, 0.0, 0.0, 5.0)
    assert reader(cases["sunjupiter_rkf78"]) == (True, False, 0.0, 0.0, 0.0)          # no event elements: all criteria off


def test_units_and_unit_inheritance(reader):
    au_km = 1.495978707e8
    ok, doubt, ej, hc, cf = reader(settings('    <Ejection value="1.495978707e9" unit="km" />\n    <HitCentrum value="10" unit="SolarRadius" />\n'))
    assert (ok, doubt) == (True, False)
    assert ej == pytest.approx(1.495978707e9 * (1.0 / au_km), rel=1e-15) and hc == pytest.approx(10.0 / 215.094, rel=1e-15)
    # the loader keeps ONE `unit` variable: a HitCentrum without its own unit inherits the Ejection's
    ok, doubt, ej, hc, cf = reader(settings('    <Ejection value="3e9" unit="km" />\n    <HitCentrum value="1.5e8" />\n'))
    assert hc == pytest.approx(1.5e8 / au_km, rel=1e-15)
    # attribute names are case-insensitive in the loader; unknown / missing units mean au
    ok, doubt, ej, hc, cf = reader(settings('    <Ejection VALUE="30" Unit="AU" />\n'))
    assert (ok, doubt, ej) == (True, False, 30.0)


def test_not_fooled_by_comments_and_other_places(reader):
    base = settings('    <Ejection value="7" unit="au" />\n')
    decoy = base.replace("    <Output>", '    <!-- <Ejection value="1000" unit="au" /> <Collision factor="9" /> -->\n    <Output>', 1)
    assert reader(decoy) == (True, False, 7.0, 0.0, 0.0)
    # an element of that name outside <Settings> is not the loader's threshold
    outside = settings("").replace("</Simulation>", '<Ejection value="3" unit="au" />\n</Simulation>')
    assert reader(outside) == (True, False, 0.0, 0.0, 0.0)
    # an attribute whose name merely ends in "value"
    odd = settings('    <Ejection maxvalue="99" value="7" unit="au" />\n')
    assert reader(odd)[2] == 7.0


def test_doubt_switches_the_mode_off(reader):
    # TinyXML accepts blanks around '=', this reader does not try to: it reports doubt and the bridge stays eager
    ok, doubt, *_ = reader(settings("    <Ejection unit='au'\n value = '7' />\n"))
    assert ok and doubt
    ok, doubt, *_ = reader(settings('    <Collision />\n'))
    assert ok and doubt
    ok, doubt, *_ = reader("<Simulation><BodyGroupList/></Simulation>")
    assert not ok

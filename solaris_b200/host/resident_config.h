// Reads the three event thresholds the reference's loader takes from the <Settings> element of an input file
// (XmlFileAdapter::DeserializeSettings, Solaris/XmlFileAdapter.cpp:215-275): <Ejection value unit>, <HitCentrum value
// unit>, <Collision factor>.  Used by the drop-in's resident mode (sol_bridge.cpp), which must use exactly the values
// the host program uses; anything this reader is not sure about is reported as `doubt`, and the caller then stays in
// the default (eager) mode.  No dependency on the reference's headers: the unit factors are passed in, so the same
// code is unit-tested on its own (tests/test_resident_config.py).
#pragma once
#include <cctype>
#include <cstdlib>
#include <string>

namespace solb200 {

struct UnitFactors { double meterToAu, kilometerToAu, solarRadiusToAu; };
struct EventThresholds {
	bool parsed;      // a <Settings> block was found
	bool doubt;       // an event element is there but its value could not be read
	double ejection, hitCentrum, collisionFactor;   // au, au, factor; 0 = criterion off (the loader's defaults)
};

inline std::string rc_lower(std::string v)
{
	for (size_t i = 0; i < v.size(); i++) v[i] = (char)tolower((unsigned char)v[i]);
	return v;
}

// value of attribute `name` (case-insensitive, like the loader) inside the first <tag ...> element of `xml`; "" if absent
inline std::string rc_attribute(const std::string &xml, const std::string &tag, const std::string &name)
{
	size_t p = xml.find("<" + tag);
	while (p != std::string::npos && p + tag.size() + 1 < xml.size() && isalnum((unsigned char)xml[p + tag.size() + 1]))
		p = xml.find("<" + tag, p + 1);
	if (p == std::string::npos) return "";
	const size_t e = xml.find('>', p);
	if (e == std::string::npos) return "";
	const std::string el = xml.substr(p, e - p), low = rc_lower(el);
	const std::string key = rc_lower(name) + "=";
	size_t a = low.find(key);
	while (a != std::string::npos && a > 0 && (isalnum((unsigned char)low[a - 1]) || low[a - 1] == '_')) a = low.find(key, a + 1);   // not the tail of another name
	if (a == std::string::npos) return "";
	a += key.size();
	if (a >= el.size()) return "";
	const char q = el[a];
	if (q != '"' && q != '\'') return "";
	const size_t z = el.find(q, a + 1);
	if (z == std::string::npos) return "";
	return el.substr(a + 1, z - a - 1);
}

inline bool rc_has_element(const std::string &xml, const std::string &tag)
{
	size_t p = xml.find("<" + tag);
	while (p != std::string::npos && p + tag.size() + 1 < xml.size() && isalnum((unsigned char)xml[p + tag.size() + 1]))
		p = xml.find("<" + tag, p + 1);
	return p != std::string::npos;
}

// UnitTool::DistanceToAu (Solaris/Units.cpp:76-100): unknown / empty units leave the value as it is (au)
inline double rc_distance_to_au(double v, const std::string &unit, const UnitFactors &f)
{
	const std::string u = rc_lower(unit);
	if (u == "m" || u == "meter") return v * f.meterToAu;
	if (u == "km" || u == "kilometer") return v * f.kilometerToAu;
	if (u == "solarradius") return v * f.solarRadiusToAu;
	return v;
}

inline EventThresholds read_event_thresholds(std::string xml, const UnitFactors &f)
{
	EventThresholds r = {false, false, 0.0, 0.0, 0.0};
	for (size_t c0 = xml.find("<!--"); c0 != std::string::npos; c0 = xml.find("<!--", c0)) {   // comments out
		const size_t c1 = xml.find("-->", c0 + 4);
		xml.erase(c0, c1 == std::string::npos ? std::string::npos : c1 + 3 - c0);
	}
	const size_t s0 = xml.find("<Settings"), s1 = xml.find("</Settings>");
	if (s0 == std::string::npos || s1 == std::string::npos || s1 < s0) return r;
	const std::string st = xml.substr(s0, s1 - s0);
	r.parsed = true;
	std::string unit;                       // the loader reuses ONE `unit` variable for both elements (XmlFileAdapter.cpp:215-262)
	if (rc_has_element(st, "Ejection")) {
		const std::string v = rc_attribute(st, "Ejection", "value"), u = rc_attribute(st, "Ejection", "unit");
		if (v.empty()) r.doubt = true;
		if (!u.empty()) unit = u;
		r.ejection = rc_distance_to_au(atof(v.c_str()), unit, f);
	}
	if (rc_has_element(st, "HitCentrum")) {
		const std::string v = rc_attribute(st, "HitCentrum", "value"), u = rc_attribute(st, "HitCentrum", "unit");
		if (v.empty()) r.doubt = true;
		if (!u.empty()) unit = u;
		r.hitCentrum = rc_distance_to_au(atof(v.c_str()), unit, f);
	}
	if (rc_has_element(st, "Collision")) {
		const std::string v = rc_attribute(st, "Collision", "factor");
		if (v.empty()) r.doubt = true;
		r.collisionFactor = atof(v.c_str());
	}
	return r;
}

}  // namespace solb200

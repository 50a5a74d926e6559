// Binary side loader for large populations of planetesimals / test particles.
//
// The reference's loader parses one XML DOM (TinyXML) into std::list<Body> objects and solves Kepler's equation body by
// body (XmlFileAdapter.cpp:697-760, Simulation.cpp:131-172): at 10^5 - 10^6 bodies that is gigabytes of DOM and minutes
// of start-up before the first step.  With SOLARIS_B200_BODIES=<file> the drop-in program reads the small bodies from a
// flat little-endian file instead and appends them to BodyData right where Simulator::BodyListToBodyData builds it; the
// XML then only carries the settings, the star and the planets.
//
//   char   magic[8]  = "SOLB200B"
//   int32  version   = 1
//   int32  n                      number of bodies
//   int32  type                   6 = planetesimal, 7 = test particle            (Body.h:14-24)
//   int32  kind                   0 = phases {x,y,z,vx,vy,vz} [au, au/day], 1 = orbital elements {a,e,incl,peri,node,M} [au, rad]
//   double state[n][6]
//   double mass[n], radius[n], density[n], cD[n]      (solar mass, au, solar mass / au^3, -; ignored for test particles)
//
// kind 1 is converted with ONE batched device call (sol_elements_to_phases: the statements of Ephemeris::CalculatePhase,
// mu = G (m0 + m) for planetesimals and G m0 for test particles as in Simulation::SetPhasesRadiiDensity).
// Only these two body types can be side-loaded: they are never the surviving body of a merger (their nearest
// neighbour is always one of the XML's massive bodies), so no code of the reference ever looks them up in its body list.
// They get the ids following the XML's bodies and are not listed in ConstantProperties.dat.
#pragma once
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace solb200 {

struct SideBodies {
	int n = 0, type = 0, kind = 0;
	std::vector<double> state, mass, radius, density, cD;
};

// 0 ok, 1 error (message in err)
inline int read_side_bodies(const char *path, SideBodies &out, std::string &err)
{
	FILE *f = fopen(path, "rb");
	if (!f) { err = std::string("SOLARIS_B200_BODIES: cannot open '") + path + "'"; return 1; }
	char magic[8];
	int hdr[4];
	bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, "SOLB200B", 8) == 0 && fread(hdr, sizeof(int), 4, f) == 4;
	if (!ok || hdr[0] != 1 || hdr[1] < 0 || (hdr[2] != 6 && hdr[2] != 7) || (hdr[3] != 0 && hdr[3] != 1)) {
		fclose(f);
		err = std::string("SOLARIS_B200_BODIES: '") + path + "' is not a version-1 body file of planetesimals or test particles";
		return 1;
	}
	out.n = hdr[1]; out.type = hdr[2]; out.kind = hdr[3];
	const size_t n = (size_t)out.n;
	out.state.resize(6 * n); out.mass.resize(n); out.radius.resize(n); out.density.resize(n); out.cD.resize(n);
	ok = fread(out.state.data(), sizeof(double), 6 * n, f) == 6 * n && fread(out.mass.data(), sizeof(double), n, f) == n &&
	     fread(out.radius.data(), sizeof(double), n, f) == n && fread(out.density.data(), sizeof(double), n, f) == n &&
	     fread(out.cD.data(), sizeof(double), n, f) == n;
	fclose(f);
	if (!ok) { err = std::string("SOLARIS_B200_BODIES: '") + path + "' is truncated"; return 1; }
	return 0;
}

}  // namespace solb200

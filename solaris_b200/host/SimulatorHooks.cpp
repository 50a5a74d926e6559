// Hooks in front of TWO private members of the reference's Simulator (class declared by the reference's own
// Simulator.h; every other member stays the reference's code, linked unchanged):
//
//   int Simulator::BodyListToBodyData()     Solaris/Simulator.cpp:525-593   (start of every Integrate call)
//   int Simulator::CheckEvent(double)       Solaris/Simulator.cpp:621-735   (after every accepted step)
//
// build_dropin.sh renames those two symbols in a COPY of the reference's Simulator.o (objcopy --redefine-sym), so the
// reference's functions stay callable under the names below and the members defined here take their place.
//
// Why: the Driver is the only thing Simulator calls per step, and it is told neither the event thresholds nor that
// CheckEvent is about to scan rm3 / indexOfNN / distanceOfNN.  With these hooks the device-resident state is the
// default: BodyListToBodyData hands Settings::ejection / hitCentrum / collision->factor to the bridge, the Driver ends
// with the device flag reduction (sol_detect_events), and CheckEvent runs the reference's scan + merge + removal code
// - on freshly downloaded arrays - only on steps where a count is non-zero.  On all other steps the reference's loops
// would find nothing either (same values, same comparisons), so skipping them changes no output.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "Acceleration.h"
#include "BodyData.h"
#include "Constants.h"
#include "Error.h"
#include "EventCondition.h"
#include "Settings.h"
#include "Simulation.h"
#include "Simulator.h"
#include "side_loader.h"
#include "sol_bridge.h"

using namespace solb200;

// the reference's members under their new names (`this` is the first argument in the ABI)
extern "C" int solb200_reference_Simulator_BodyListToBodyData(Simulator *self);
extern "C" int solb200_reference_Simulator_CheckEvent(Simulator *self, double timeOfEvent);
extern "C" double solb200_reference_Simulator_ShortestPeriod(Simulator *self);

// Appends the side-loaded bodies (side_loader.h) to the BodyData the reference has just built from its body list: the
// arrays are re-allocated for the merged counts and refilled in BodyType order (list bodies of a type first, then the
// file's), with the statements of Simulator::BodyListToBodyData for the per-body drag coefficients (Simulator.cpp:548-559).
// the side file is read once per process (ShortestPeriod needs it before BodyListToBodyData does)
static int side_bodies(const char *path, const SideBodies **out)
{
	static SideBodies cache;
	static std::string cached_path;
	if (cached_path != path) {
		std::string err;
		SideBodies sb;
		if (read_side_bodies(path, sb, err) == 1) { Error::_errMsg = err; Error::PushLocation(__FILE__, __FUNCTION__, __LINE__); return 1; }
		cache = sb;
		cached_path = path;
	}
	*out = &cache;
	return 0;
}

static int append_side_bodies(BodyData &bd, sol_ctx *ctx, const char *path)
{
	const SideBodies *cached = 0;
	if (side_bodies(path, &cached) == 1) return 1;
	SideBodies sb = *cached;
	if (sb.n == 0) return 0;
	if (bd.nBodies.centralBody != 1) { Error::_errMsg = "SOLARIS_B200_BODIES needs the central body in the XML input"; Error::PushLocation(__FILE__, __FUNCTION__, __LINE__); return 1; }
	if (sb.kind == 1) {
		// Simulation::SetPhasesRadiiDensity's per-body Ephemeris::CalculatePhase as one device batch
		std::vector<double> mu(sb.n), ph(6 * (size_t)sb.n, 0.0);
		const double gm0 = Constants::Gauss2 * bd.mass[0];
		for (int i = 0; i < sb.n; i++) mu[i] = gm0 + (sb.type != TestParticle ? Constants::Gauss2 * sb.mass[i] : 0.0);
		int bad = 0;
		if (sol_elements_to_phases(ctx, sb.n, &mu[0], &sb.state[0], &ph[0], &bad) != SOL_OK) {
			Error::_errMsg = sol_last_error(ctx);
			Error::PushLocation(__FILE__, __FUNCTION__, __LINE__);
			return 1;
		}
		sb.state.swap(ph);
	}
	const int n0 = bd.nBodies.total, n1 = n0 + sb.n;
	// keep what the reference filled in
	std::vector<int> id(bd.id, bd.id + n0), type(bd.type, bd.type + n0), migType(bd.migType, bd.migType + n0);
	std::vector<double> migStop(bd.migStopAt, bd.migStopAt + n0), mass(bd.mass, bd.mass + n0), radius(bd.radius, bd.radius + n0),
	    density(bd.density, bd.density + n0), cD(bd.cD, bd.cD + n0), gS(bd.gammaStokes, bd.gammaStokes + n0),
	    gE(bd.gammaEpstein, bd.gammaEpstein + n0), y0(bd.y0, bd.y0 + 6 * (size_t)n0);
	int max_id = 0;
	for (int i = 0; i < n0; i++) if (id[i] > max_id) max_id = id[i];
	NBodies nb = bd.nBodies;
	bd.Free();
	if (sb.type == Planetesimal) nb.planetsimal += sb.n; else nb.testParticle += sb.n;
	nb.total = n1;
	bd.nBodies = nb;
	if (bd.Allocate() == 1) { Error::PushLocation(__FILE__, __FUNCTION__, __LINE__); return 1; }
	// position of the file's block: behind the list's bodies of the same type
	int at = 0;
	while (at < n0 && type[at] <= sb.type) at++;
	for (int k = 0; k < n1; k++) {
		const bool from_file = k >= at && k < at + sb.n;
		if (!from_file) {
			const int i = k < at ? k : k - sb.n;
			bd.id[k] = id[i]; bd.type[k] = type[i]; bd.migType[k] = migType[i]; bd.migStopAt[k] = migStop[i];
			bd.mass[k] = mass[i]; bd.radius[k] = radius[i]; bd.density[k] = density[i]; bd.cD[k] = cD[i];
			bd.gammaStokes[k] = gS[i]; bd.gammaEpstein[k] = gE[i];
			memcpy(&bd.y0[6 * (size_t)k], &y0[6 * (size_t)i], 6 * sizeof(double));
			continue;
		}
		const int q = k - at;
		bd.id[k] = max_id + 1 + q; bd.type[k] = sb.type; bd.migType[k] = No; bd.migStopAt[k] = 0.0;
		if (sb.type != TestParticle) {
			bd.mass[k] = sb.mass[q]; bd.radius[k] = sb.radius[q]; bd.density[k] = sb.density[q]; bd.cD[k] = sb.cD[q];
			if (bd.radius[k] > 0) {
				bd.gammaEpstein[k] = 1.0 / (bd.density[k] * bd.radius[k]);                                   // Characteristics.cpp:75-78
				bd.gammaStokes[k] = bd.cD[k] > 0 ? (3.0 / 8.0) * bd.cD[k] / (bd.density[k] * bd.radius[k]) : 0.0;   // :84-87
			} else {
				bd.gammaEpstein[k] = 0.0; bd.gammaStokes[k] = 0.0;
			}
		} else {
			bd.mass[k] = bd.radius[k] = bd.density[k] = bd.cD[k] = bd.gammaStokes[k] = bd.gammaEpstein[k] = 0.0;
		}
		memcpy(&bd.y0[6 * (size_t)k], &sb.state[6 * (size_t)q], 6 * sizeof(double));
	}
	return 0;
}

int Simulator::BodyListToBodyData()
{
	if (solb200_reference_Simulator_BodyListToBodyData(this) == 1) return 1;
	const char *side = getenv("SOLARIS_B200_BODIES");
	if (side != 0 && side[0] != 0) {
		if (_simulation->settings.baryCentric) {
			Error::_errMsg = "SOLARIS_B200_BODIES: side-loaded bodies are astrocentric; the barycentric frame is not supported";
			Error::PushLocation(__FILE__, __FUNCTION__, __LINE__);
			return 1;
		}
		Bridge *bs = _acceleration != 0 ? bridge_of(_acceleration) : 0;
		if (bs == 0 || append_side_bodies(bodyData, bs->ctx, side) == 1) return 1;
	}
	Bridge *b = _acceleration != 0 ? bridge_of(_acceleration) : 0;
	if (b != 0) {
		const Settings &s = _simulation->settings;
		// the same three values Simulator::CheckEvent reads (Simulator.cpp:626-627,690-692)
		bridge_set_thresholds(b, s.ejection, s.hitCentrum, s.collision != 0 ? s.collision->factor : 0.0);
	}
	return 0;
}

// Simulator::ShortestPeriod (Simulator.cpp:504-517) walks the body list; the side-loaded bodies take part with the same
// formula (Body::CalculateOrbitalPeriod, Body.cpp:89-98, on a phase; 2 pi sqrt(a^3 / mu) on elements), so that the initial
// step h0 = P_min / 50000 of MainIntegration does not depend on which way a body was loaded.
double Simulator::ShortestPeriod()
{
	double period = solb200_reference_Simulator_ShortestPeriod(this);
	const char *side = getenv("SOLARIS_B200_BODIES");
	const SideBodies *sb = 0;
	if (side == 0 || side[0] == 0 || side_bodies(side, &sb) == 1) return period;
	const double centralGm = _simulation->bodyList.front()->GetGm();
	for (int i = 0; i < sb->n; i++) {
		const double mu = centralGm + (sb->type == TestParticle ? 0.0 : Constants::Gauss2 * sb->mass[i]);
		const double *q = &sb->state[6 * (size_t)i];
		double a;
		if (sb->kind == 1) {
			a = q[0];
		} else {
			const double kin = (q[3] * q[3] + q[4] * q[4] + q[5] * q[5]) / 2.0;            // Ephemeris.cpp:213-226
			const double pot = -mu / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
			const double h = kin + pot;
			if (h > 0.0) continue;
			a = -mu / (2.0 * h);
		}
		const double p = 2.0 * Constants::Pi * sqrt((a * a * a) / mu);
		if (p > 0 && p < period) period = p;
	}
	return period;
}

int Simulator::CheckEvent(double timeOfEvent)
{
	Bridge *b = _acceleration != 0 ? bridge_lookup(_acceleration) : 0;
	if (b != 0 && !b->event_pending) {
		b->host_scans_skipped++;
		return 0;
	}
	return solb200_reference_Simulator_CheckEvent(this, timeOfEvent);
}

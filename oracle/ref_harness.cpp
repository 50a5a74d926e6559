// TEST INFRASTRUCTURE ONLY.  C-ABI harness over the UNMODIFIED reference classes
// (Acceleration, RungeKutta4, RungeKuttaFehlberg78, DormandPrince, BodyData, Nebula), compiled by
// oracle/build_ref.sh against the headers where they lie under /root/reference/Solaris and linked
// with oracle/_ref/libsolaris_ref.a into oracle/_ref/libref_harness.so.
//
// It exists so that tests/, the golden-vector generator and bench.py's reference arm can call the
// reference's own code on an in-memory BodyData (bypassing XML, guid and epoch problems, SURVEY.md
// Q18-Q20).  Nothing in solaris_b200/ may load it.
//
// Reference entry points exercised:
//   Acceleration::Compute                 Solaris/Acceleration.cpp:60
//   RungeKutta4::Driver                   Solaris/RungeKutta4.cpp:20
//   RungeKuttaFehlberg78::Driver          Solaris/RungeKuttaFehlberg78.cpp:66
//   DormandPrince::Driver                 Solaris/DormandPrince.cpp:126
//   Tools::CheckAgainstSmallestNumber     Solaris/Tools.cpp:39
#include <cstring>
#include <list>
#include <cstdio>
#include <chrono>

#include "Acceleration.h"
#include "BodyData.h"
#include "BinaryFileAdapter.h"
#include "Calculate.h"
#include "Output.h"
#include "Constants.h"
#include "DormandPrince.h"
#include "Ephemeris.h"
#include "Error.h"
#include "OrbitalElement.h"
#include "Phase.h"
#include "IntegratorType.h"
#include "Nebula.h"
#include "RungeKutta4.h"
// Simulator::RemoveBody is a private member; the harness needs to call it on a BodyData of its own
#define private public
#include "Simulator.h"
#include "TwoBodyAffair.h"
#undef private
#include "RungeKuttaFehlberg78.h"
#include "TimeLine.h"
#include "Tools.h"

extern "C" {

// Plain-data snapshot of GasComponent (Solaris/GasComponent.h:8-52); same field order as
// include/solaris_b200.h:sol_nebula_pod so tests can share one ctypes.Structure.
struct ref_nebula_pod {
	double alpha;
	double mean_molecular_weight;
	double particle_diameter;
	int    decrease_type;   // GasDecreaseType: 0 CONSTANT, 1 LINEAR, 2 EXPONENTIAL
	int    _pad;
	double time_scale, t0, t1;
	double inner_edge;
	double eta_c, eta_index;
	double tau_c, tau_index;
	double scale_height_c, scale_height_index;
	double density_c, density_index;
	double mean_free_path_c, mean_free_path_index;
};

struct ref_handle {
	BodyData              bd;
	Nebula               *nebula;
	Acceleration         *acc;
	TimeLine              tl;
	RungeKutta4           rk4;
	RungeKuttaFehlberg78  rkf78;
	DormandPrince         dp;
};

// Values of a default-constructed GasComponent (Solaris/GasComponent.cpp:9-34).
void ref_nebula_defaults(ref_nebula_pod *p)
{
	GasComponent g;
	p->alpha = g.alpha;
	p->mean_molecular_weight = g.meanMolecularWeight;
	p->particle_diameter = g.particleDiameter;
	p->decrease_type = (int)g.type;
	p->_pad = 0;
	p->time_scale = g.timeScale; p->t0 = g.t0; p->t1 = g.t1;
	p->inner_edge = g.innerEdge;
	p->eta_c = g.eta.c; p->eta_index = g.eta.index;
	p->tau_c = g.tau.c; p->tau_index = g.tau.index;
	p->scale_height_c = g.scaleHeight.c; p->scale_height_index = g.scaleHeight.index;
	p->density_c = g.density.c; p->density_index = g.density.index;
	p->mean_free_path_c = g.meanFreePath.c; p->mean_free_path_index = g.meanFreePath.index;
}

// counts[7] = centralBody, giantPlanet, rockyPlanet, protoPlanet, superPlanetsimal, planetsimal, testParticle
// integrator: IntegratorType enum value (0 DORMAND_PRINCE, 1 RUNGE_KUTTA4, 3 RUNGE_KUTTA_FEHLBERG78)
ref_handle *ref_create(const int counts[7], const double *y0, const double *mass, const double *radius,
                       const double *density, const double *cD, const double *gammaStokes,
                       const double *gammaEpstein, const double *migStopAt, const int *type,
                       const int *migType, const int *id, int barycentric, const ref_nebula_pod *neb,
                       int integrator)
{
	ref_handle *h = new ref_handle();
	NBodies &nb = h->bd.nBodies;
	nb.centralBody = counts[0]; nb.giantPlanet = counts[1]; nb.rockyPlanet = counts[2];
	nb.protoPlanet = counts[3]; nb.superPlanetsimal = counts[4]; nb.planetsimal = counts[5];
	nb.testParticle = counts[6];
	nb.total = counts[0] + counts[1] + counts[2] + counts[3] + counts[4] + counts[5] + counts[6];
	if (h->bd.Allocate() == 1) { delete h; return 0; }
	int n = nb.total;
	memcpy(h->bd.y0, y0, 6 * n * sizeof(double));
	memset(h->bd.y, 0, 6 * n * sizeof(double));
	memset(h->bd.yBetterEst, 0, 6 * n * sizeof(double));
	memset(h->bd.yscale, 0, 6 * n * sizeof(double));
	memset(h->bd.accel, 0, 6 * n * sizeof(double));
	memset(h->bd.error, 0, 6 * n * sizeof(double));
	memcpy(h->bd.mass, mass, n * sizeof(double));
	memcpy(h->bd.radius, radius, n * sizeof(double));
	memcpy(h->bd.density, density, n * sizeof(double));
	memcpy(h->bd.cD, cD, n * sizeof(double));
	memcpy(h->bd.gammaStokes, gammaStokes, n * sizeof(double));
	memcpy(h->bd.gammaEpstein, gammaEpstein, n * sizeof(double));
	memcpy(h->bd.migStopAt, migStopAt, n * sizeof(double));
	memcpy(h->bd.type, type, n * sizeof(int));
	memcpy(h->bd.migType, migType, n * sizeof(int));
	memcpy(h->bd.id, id, n * sizeof(int));
	for (int i = 0; i < n; i++) { h->bd.indexOfNN[i] = -1; h->bd.distanceOfNN[i] = 0.0; }

	h->nebula = 0;
	if (neb != 0) {
		// Same effect as XmlFileAdapter filling a default-constructed GasComponent: fields are
		// overwritten, ctor-time meanFreePath / a are whatever the pod carries.
		h->nebula = new Nebula();
		GasComponent &g = h->nebula->gasComponent;
		g.alpha = neb->alpha;
		g.meanMolecularWeight = neb->mean_molecular_weight;
		g.particleDiameter = neb->particle_diameter;
		g.type = (GasDecreaseType)neb->decrease_type;
		g.timeScale = neb->time_scale; g.t0 = neb->t0; g.t1 = neb->t1;
		g.innerEdge = neb->inner_edge;
		g.eta = PowerLaw(neb->eta_c, neb->eta_index);
		g.tau = PowerLaw(neb->tau_c, neb->tau_index);
		g.scaleHeight = PowerLaw(neb->scale_height_c, neb->scale_height_index);
		g.density = PowerLaw(neb->density_c, neb->density_index);
		g.meanFreePath = PowerLaw(neb->mean_free_path_c, neb->mean_free_path_index);
	}
	h->acc = new Acceleration((IntegratorType)integrator, barycentric != 0, &h->bd, h->nebula);
	// The reference sizes accelMigrationTypeI by protoPlanet only but indexes it over rocky+proto
	// (Acceleration.cpp:108-110,117-127 / :200-204,209-220) - a heap overflow whenever rocky > 0.
	// The array is a public member and is allocated lazily only when still null, so the harness
	// pre-allocates it at the size the loops actually use.  No arithmetic changes.
	if (h->nebula != 0 && nb.protoPlanet > 0 && nb.rockyPlanet > 0) {
		int len = 3 * (nb.rockyPlanet + nb.protoPlanet);
		h->acc->accelMigrationTypeI = new double[len];
		memset(h->acc->accelMigrationTypeI, 0, len * sizeof(double));
	}
	return h;
}

void ref_destroy(ref_handle *h)
{
	if (h == 0) return;
	delete h->acc;
	// The reference's ComputeBaryCentric may already have deleted the nebula (Acceleration.cpp:152-160);
	// leaking it here is the safe choice for a test harness.
	delete h;
}

int ref_n(ref_handle *h) { return h->bd.nBodies.total; }

// eval_flags: bit0 evaluateGasDrag, bit1 evaluateTypeIMigration, bit2 evaluateTypeIIMigration
int ref_compute(ref_handle *h, double t, const double *y, double *dydt, unsigned eval_flags)
{
	h->acc->evaluateGasDrag         = (eval_flags & 1u) != 0;
	h->acc->evaluateTypeIMigration  = (eval_flags & 2u) != 0;
	h->acc->evaluateTypeIIMigration = (eval_flags & 4u) != 0;
	return h->acc->Compute(t, const_cast<double *>(y), dydt);
}

// Side outputs of the last Compute. Any pointer may be null.
void ref_get_side(ref_handle *h, double *rm3, int *indexOfNN, double *distanceOfNN, int *migType)
{
	int n = h->bd.nBodies.total;
	if (rm3) {
		if (h->acc->rm3) memcpy(rm3, h->acc->rm3, n * sizeof(double));
		else memset(rm3, 0, n * sizeof(double));
	}
	if (indexOfNN) memcpy(indexOfNN, h->bd.indexOfNN, n * sizeof(int));
	if (distanceOfNN) memcpy(distanceOfNN, h->bd.distanceOfNN, n * sizeof(double));
	if (migType) memcpy(migType, h->bd.migType, n * sizeof(int));
}

// One Driver call. integrator: 0 DormandPrince, 1 RungeKutta4, 3 RungeKuttaFehlberg78.
// time / hNext are in-out, hDid is out, exactly the TimeLine fields the Driver touches.
int ref_step(ref_handle *h, int integrator, double *time, double *hNext, double *hDid)
{
	h->tl.time = *time;
	h->tl.hNext = *hNext;
	int r;
	switch (integrator) {
	case DORMAND_PRINCE:         r = h->dp.Driver(&h->bd, h->acc, &h->tl); break;
	case RUNGE_KUTTA4:           r = h->rk4.Driver(&h->bd, h->acc, &h->tl); break;
	case RUNGE_KUTTA_FEHLBERG78: r = h->rkf78.Driver(&h->bd, h->acc, &h->tl); break;
	default: return 1;
	}
	*time = h->tl.time;
	*hNext = h->tl.hNext;
	*hDid = h->tl.hDid;
	return r;
}

// what: 0 y0, 1 y, 2 accel, 3 error, 4 yscale  (6n doubles each)
void ref_get_array(ref_handle *h, int what, double *out)
{
	int n6 = 6 * h->bd.nBodies.total;
	const double *src = 0;
	switch (what) {
	case 0: src = h->bd.y0; break;
	case 1: src = h->bd.y; break;
	case 2: src = h->bd.accel; break;
	case 3: src = h->bd.error; break;
	case 4: src = h->bd.yscale; break;
	}
	if (src) memcpy(out, src, n6 * sizeof(double));
}

void ref_set_y0(ref_handle *h, const double *y0)
{
	memcpy(h->bd.y0, y0, 6 * h->bd.nBodies.total * sizeof(double));
}

void ref_flush_tiny(ref_handle *h)
{
	int n6 = 6 * h->bd.nBodies.total;
	Tools::CheckAgainstSmallestNumber(n6, h->bd.y);
	Tools::CheckAgainstSmallestNumber(n6, h->bd.y0);
}

// Times `reps` calls of Acceleration::Compute(t, y0, accel) with steady_clock; returns the median
// seconds per call (bench.py --impl reference and cpu_baseline use this).
double ref_time_compute(ref_handle *h, double t, int reps)
{
	double best[64];
	if (reps > 64) reps = 64;
	if (reps < 1) reps = 1;
	h->acc->evaluateGasDrag = h->acc->evaluateTypeIMigration = h->acc->evaluateTypeIIMigration = true;
	for (int r = 0; r < reps; r++) {
		auto a = std::chrono::steady_clock::now();
		h->acc->Compute(t, h->bd.y0, h->bd.accel);
		auto b = std::chrono::steady_clock::now();
		best[r] = std::chrono::duration<double>(b - a).count();
	}
	for (int i = 1; i < reps; i++) {   // insertion sort, then median
		double v = best[i]; int j = i - 1;
		while (j >= 0 && best[j] > v) { best[j + 1] = best[j]; j--; }
		best[j + 1] = v;
	}
	return best[reps / 2];
}

// Calculate::Integrals on the current y0 (Solaris/Calculate.cpp:43-63): the 16 values written to Integrals.dat
void ref_integrals(ref_handle *h, double *out16)
{
	Calculate::Integrals(&h->bd);
	memcpy(out16, h->bd.integrals, 16 * sizeof(double));
}

// BinaryFileAdapter::SavePhases (Solaris/BinaryFileAdapter.cpp:107-158): appends one snapshot to <dir>/<file>
// (type 0 = BINARY) or <dir>/<file without extension>.txt (type 1 = TEXT) with the reference's own writer.
void ref_save_phases(const char *dir, const char *file, double time, int n, double *y, int *id, int type)
{
	Output out;
	Output::directory = dir;
	Output::directorySeparator = '/';
	out.phases = file;
	BinaryFileAdapter adapter(&out);
	adapter.SavePhases(time, n, y, id, type == 0 ? BinaryFileAdapter::BINARY : BinaryFileAdapter::TEXT);
}

// Simulator::RemoveBody (Solaris/Simulator.cpp:737-771) applied to this handle's BodyData: a Simulator object
// borrows the arrays for the call (bitwise copy in and out; the borrowed Simulator is never destroyed).
int ref_remove_body(ref_handle *h, int bodyId)
{
	static Simulator *sim = new Simulator(0);
	memcpy((void *)&sim->bodyData, (void *)&h->bd, sizeof(BodyData));
	const int rc = sim->RemoveBody(bodyId);
	memcpy((void *)&h->bd, (void *)&sim->bodyData, sizeof(BodyData));
	return rc;
}

void ref_get_params(ref_handle *h, int counts[7], double *mass, double *radius, double *density, double *cD,
                    double *gammaStokes, double *gammaEpstein, double *migStopAt, int *type, int *migType, int *id)
{
	NBodies &nb = h->bd.nBodies;
	const int c[7] = {nb.centralBody, nb.giantPlanet, nb.rockyPlanet, nb.protoPlanet, nb.superPlanetsimal, nb.planetsimal, nb.testParticle};
	memcpy(counts, c, sizeof(c));
	const size_t n = (size_t)nb.total;
	memcpy(mass, h->bd.mass, n * sizeof(double)); memcpy(radius, h->bd.radius, n * sizeof(double));
	memcpy(density, h->bd.density, n * sizeof(double)); memcpy(cD, h->bd.cD, n * sizeof(double));
	memcpy(gammaStokes, h->bd.gammaStokes, n * sizeof(double)); memcpy(gammaEpstein, h->bd.gammaEpstein, n * sizeof(double));
	memcpy(migStopAt, h->bd.migStopAt, n * sizeof(double));
	memcpy(type, h->bd.type, n * sizeof(int)); memcpy(migType, h->bd.migType, n * sizeof(int)); memcpy(id, h->bd.id, n * sizeof(int));
}

// Ephemeris::CalculatePhase (Solaris/Ephemeris.cpp:141-176) for n bodies; el6 = {a, e, incl, peri, node, M} per body.
// Returns the number of bodies for which the reference reported an error.
int ref_elements_to_phases(int n, const double *mu, const double *el6, double *out6)
{
	int bad = 0;
	for (int i = 0; i < n; i++) {
		const double *el = el6 + 6 * (size_t)i;
		OrbitalElement oe(el[0], el[1], el[2], el[3], el[4], el[5]);
		Phase ph(0);
		if (Ephemeris::CalculatePhase(mu[i], &oe, &ph) == 1) { bad++; continue; }
		double *o = out6 + 6 * (size_t)i;
		o[0] = ph.position.x; o[1] = ph.position.y; o[2] = ph.position.z;
		o[3] = ph.velocity.x; o[4] = ph.velocity.y; o[5] = ph.velocity.z;
	}
	return bad;
}

// The reference's TwoBodyAffair constructor (TwoBodyAffair.cpp:9-21, running id) and its writer
// BinaryFileAdapter::SaveTwoBodyAffairs (BinaryFileAdapter.cpp:225-261) for the events of one CheckEvent scan: the
// caller names the bodies (kind 0 = Ejection, 1 = HitCentrum, in scan order); phases and ids come from this
// handle's BodyData exactly as Simulator::CheckEvent passes them (Simulator.cpp:636,643).  Appends to dir/file.
void ref_write_affairs(ref_handle *h, const char *dir, const char *file, int n, const int *kind, const int *index,
                       double time, int first_event_id)
{
	TwoBodyAffair::_eventId = first_event_id;
	std::list<TwoBodyAffair> ejections, hits;
	for (int k = 0; k < n; k++) {
		const int i = index[k];
		TwoBodyAffair affair(kind[k] == 0 ? Ejection : HitCentrum, time, 0, i, h->bd.id[0], h->bd.id[i], h->bd.y0, &(h->bd.y0[6 * i]));
		(kind[k] == 0 ? ejections : hits).push_back(affair);
	}
	Output out;
	Output::directory = dir;
	Output::directorySeparator = '/';
	out.twoBodyAffair = file;
	BinaryFileAdapter adapter(&out);
	if (ejections.size() > 0) adapter.SaveTwoBodyAffairs(ejections);
	if (hits.size() > 0) adapter.SaveTwoBodyAffairs(hits);
}

const char *ref_last_error() { return Error::_errMsg.c_str(); }

double ref_gauss2() { return Constants::Gauss2; }

} // extern "C"

// See sol_bridge.h.  Compiled against the reference headers (BodyData.h, Acceleration.h, Nebula.h, Error.h).
#include <cstring>
#include <map>

#include "Acceleration.h"
#include "BodyData.h"
#include "Error.h"
#include "Nebula.h"
#include "sol_bridge.h"

namespace solb200 {

static std::map<Acceleration *, Bridge *> &table()
{
	static std::map<Acceleration *, Bridge *> t;
	return t;
}

static int fail(Bridge *b, const char *where)
{
	Error::_errMsg = std::string("solaris_b200: ") + where + ": " + sol_last_error(b ? b->ctx : 0);
	Error::PushLocation(__FILE__, __FUNCTION__, __LINE__);
	return 1;
}

Bridge *bridge_of(Acceleration *acc)
{
	std::map<Acceleration *, Bridge *>::iterator it = table().find(acc);
	if (it != table().end()) return it->second;
	Bridge *b = new Bridge();
	if (sol_create(0, &b->ctx) != SOL_OK) {
		Error::_errMsg = std::string("solaris_b200: ") + sol_last_error(0);
		Error::PushLocation(__FILE__, __FUNCTION__, __LINE__);
		delete b;
		return 0;
	}
	table()[acc] = b;
	return b;
}

Bridge *bridge_of_bodydata(BodyData *bd, Acceleration **acc_out)
{
	for (std::map<Acceleration *, Bridge *>::iterator it = table().begin(); it != table().end(); ++it) {
		if (it->first->bodyData == bd) {
			if (acc_out) *acc_out = it->first;
			return it->second;
		}
	}
	return 0;
}

void bridge_release(Acceleration *acc)
{
	std::map<Acceleration *, Bridge *>::iterator it = table().find(acc);
	if (it == table().end()) return;
	sol_destroy(it->second->ctx);
	delete it->second;
	table().erase(it);
}

static void nebula_to_pod(Nebula *neb, sol_nebula_pod *p)
{
	GasComponent &g = neb->gasComponent;
	memset(p, 0, sizeof(*p));
	p->alpha = g.alpha;
	p->mean_molecular_weight = g.meanMolecularWeight;
	p->particle_diameter = g.particleDiameter;
	p->decrease_type = (int)g.type;
	p->time_scale = g.timeScale; p->t0 = g.t0; p->t1 = g.t1;
	p->inner_edge = g.innerEdge;
	p->eta_c = g.eta.c; p->eta_index = g.eta.index;
	p->tau_c = g.tau.c; p->tau_index = g.tau.index;
	p->scale_height_c = g.scaleHeight.c; p->scale_height_index = g.scaleHeight.index;
	p->density_c = g.density.c; p->density_index = g.density.index;
	p->mean_free_path_c = g.meanFreePath.c; p->mean_free_path_index = g.meanFreePath.index;   // ctor-time law, Q14
}

template <typename T>
static bool same(const std::vector<T> &shadow, const T *host, int n)
{
	return (int)shadow.size() == n && (n == 0 || memcmp(&shadow[0], host, n * sizeof(T)) == 0);
}

int sync_in(Bridge *b, Acceleration *acc, BodyData *bd)
{
	NBodies &nb = bd->nBodies;
	const int counts[7] = {nb.centralBody, nb.giantPlanet, nb.rockyPlanet, nb.protoPlanet, nb.superPlanetsimal, nb.planetsimal, nb.testParticle};
	const int n = nb.total;
	const bool maybe_changed = nb.removed != b->removed_seen || (const void *)bd->mass != b->mass_ptr;
	b->removed_seen = nb.removed;
	b->mass_ptr = bd->mass;
	bool params_same = (n == b->n) && memcmp(counts, b->counts, sizeof(counts)) == 0;
	if (params_same && maybe_changed)
		params_same = same(b->mass, bd->mass, n) &&
	                   same(b->radius, bd->radius, n) && same(b->density, bd->density, n) && same(b->cD, bd->cD, n) &&
	                   same(b->gS, bd->gammaStokes, n) && same(b->gE, bd->gammaEpstein, n) && same(b->migStop, bd->migStopAt, n) &&
	                   same(b->type, bd->type, n) && same(b->migType, bd->migType, n) && same(b->id, bd->id, n);
	if (!params_same) {
		if (sol_set_bodies(b->ctx, counts, bd->y0, bd->mass, bd->radius, bd->density, bd->cD, bd->gammaStokes, bd->gammaEpstein,
		                   bd->migStopAt, bd->type, bd->migType, bd->id) != SOL_OK)
			return fail(b, "sol_set_bodies");
		memcpy(b->counts, counts, sizeof(counts));
		b->n = n;
		b->y0.assign(bd->y0, bd->y0 + 6 * n);
		b->mass.assign(bd->mass, bd->mass + n); b->radius.assign(bd->radius, bd->radius + n);
		b->density.assign(bd->density, bd->density + n); b->cD.assign(bd->cD, bd->cD + n);
		b->gS.assign(bd->gammaStokes, bd->gammaStokes + n); b->gE.assign(bd->gammaEpstein, bd->gammaEpstein + n);
		b->migStop.assign(bd->migStopAt, bd->migStopAt + n);
		b->type.assign(bd->type, bd->type + n); b->migType.assign(bd->migType, bd->migType + n); b->id.assign(bd->id, bd->id + n);
		// rm3 is sized once in the reference (Acceleration.cpp:65-69) and never shrunk; keep that
		if (acc->rm3 == 0) {
			acc->rm3 = new double[n];
			memset(acc->rm3, 0, n * sizeof(double));
		}
		b->nebula_set = false;   // the gas constants depend on mass[0]
	} else if (!same(b->y0, bd->y0, 6 * n)) {
		if (sol_upload(b->ctx, SOL_Y0, bd->y0) != SOL_OK) return fail(b, "sol_upload(y0)");
		b->y0.assign(bd->y0, bd->y0 + 6 * n);
	}
	if (!b->nebula_set) {
		if (acc->nebula != 0) {
			sol_nebula_pod pod;
			nebula_to_pod(acc->nebula, &pod);
			if (sol_set_nebula(b->ctx, &pod) != SOL_OK) return fail(b, "sol_set_nebula");
		} else if (sol_set_nebula(b->ctx, 0) != SOL_OK) {
			return fail(b, "sol_set_nebula");
		}
		b->nebula_set = true;
	}
	return 0;
}

int sync_out(Bridge *b, Acceleration *acc, BodyData *bd, double *dst_y)
{
	const int n = b->n;
	if (dst_y != 0) {
		if (sol_download(b->ctx, SOL_Y0, dst_y) != SOL_OK) return fail(b, "sol_download(y0)");
		b->y0.assign(dst_y, dst_y + 6 * n);
	}
	if (acc->rm3 != 0 && sol_download(b->ctx, SOL_RM3, acc->rm3) != SOL_OK) return fail(b, "sol_download(rm3)");
	if (sol_download(b->ctx, SOL_NN_INDEX, bd->indexOfNN) != SOL_OK) return fail(b, "sol_download(indexOfNN)");
	if (sol_download(b->ctx, SOL_NN_DISTANCE, bd->distanceOfNN) != SOL_OK) return fail(b, "sol_download(distanceOfNN)");
	if (acc->nebula != 0) {
		// type-I/II bodies that crossed migStopAt were flipped to `No` on the device (Acceleration.cpp:439-443)
		if (sol_download(b->ctx, SOL_MIGTYPE, bd->migType) != SOL_OK) return fail(b, "sol_download(migType)");
		b->migType.assign(bd->migType, bd->migType + n);
	}
	return 0;
}

int run_driver(int integrator, BodyData *bd, Acceleration *acc, double *time, double *hNext, double *hDid,
               const char *file, const char *function, long line, const char *step_error_message)
{
	Bridge *b = bridge_of(acc);
	if (b == 0) return 1;
	if (sync_in(b, acc, bd) == 1) return 1;
	double info[4] = {0, 0, 0, 0};
	if (sol_step(b->ctx, integrator, time, hNext, hDid, info) != SOL_OK) {
		const char *msg = sol_last_error(b->ctx);
		Error::_errMsg = (msg != 0 && msg[0] != 0) ? msg : step_error_message;
		Error::PushLocation(file, function, line);
		return 1;
	}
	// the new state lands in the host array that becomes y0 after the caller's std::swap
	if (sync_out(b, acc, bd, bd->y) == 1) return 1;
	return 0;
}

}  // namespace solb200

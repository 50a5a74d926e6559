// See sol_bridge.h.  Compiled against the reference headers (BodyData.h, Acceleration.h, Nebula.h, Error.h).
#include <cstring>
#include <map>

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "Acceleration.h"
#include "BodyData.h"
#include "Constants.h"
#include "Error.h"
#include "Nebula.h"
#include "TimeLine.h"
#include "sol_bridge.h"

namespace solb200 {

static double now_s()
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double)ts.tv_sec + 1.0e-9 * (double)ts.tv_nsec;
}

// SOLARIS_B200_EAGER=1 forces the download-every-step synchronisation (A/B tests of the resident default)
static bool eager_forced()
{
	static const bool on = getenv("SOLARIS_B200_EAGER") != 0 && std::string(getenv("SOLARIS_B200_EAGER")) == "1";
	return on;
}

namespace {
struct Table : std::map<Acceleration *, Bridge *> {
	// Simulator never deletes its Acceleration, so report the resident-mode counters of live bridges at exit
	~Table()
	{
		const bool stats = getenv("SOLARIS_B200_STATS") != 0;
		for (iterator it = begin(); it != end(); ++it)
			if (stats)
				fprintf(stderr, "solaris_b200: %s synchronisation: %ld steps, %ld state downloads, %ld host event scans skipped, "
				                "%ld event edits replayed on the device, %ld run-ahead launches; "
				                "seconds in the Driver: sync_in %.3f, sol_step %.3f, detect %.3f, sync_out %.3f\n",
				        (it->second->thresholds_known && !eager_forced()) ? "resident" : "eager",
				        it->second->steps_done, it->second->downloads, it->second->host_scans_skipped, it->second->edits_replayed,
				        it->second->batches, it->second->t_sync_in, it->second->t_step, it->second->t_detect, it->second->t_sync_out);
	}
};
}  // namespace

static std::map<Acceleration *, Bridge *> &table()
{
	static Table t;
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
	// SOLARIS_B200_GPUS=N: one handle over N devices of this (single-threaded) host program, sinks sharded over them
	const char *g = getenv("SOLARIS_B200_GPUS");
	const int n_gpus = g != 0 ? atoi(g) : 1;
	if ((n_gpus > 1 ? sol_create_multi(n_gpus, &b->ctx) : sol_create(0, &b->ctx)) != SOL_OK) {
		Error::_errMsg = std::string("solaris_b200: ") + sol_last_error(0);
		Error::PushLocation(__FILE__, __FUNCTION__, __LINE__);
		delete b;
		return 0;
	}
	table()[acc] = b;
	return b;
}

Bridge *bridge_lookup(Acceleration *acc)
{
	std::map<Acceleration *, Bridge *>::iterator it = table().find(acc);
	return it != table().end() ? it->second : 0;
}

void bridge_set_thresholds(Bridge *b, double ejection, double hitCentrum, double collisionFactor)
{
	b->thresholds_known = true;
	b->ejection = ejection; b->hitCentrum = hitCentrum; b->collisionFactor = collisionFactor;
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

Bridge *bridge_of_state(const double *y0, const int *id, int n, Acceleration **acc_out)
{
	for (std::map<Acceleration *, Bridge *>::iterator it = table().begin(); it != table().end(); ++it) {
		BodyData *bd = it->first->bodyData;
		if (bd != 0 && bd->y0 == y0 && bd->id == id && bd->nBodies.total == n) {
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

// After Simulator::CheckEvent merged / removed bodies (Simulator.cpp:648-735) the host arrays differ from the device
// state by a few removed bodies and a few edited survivors.  Instead of re-uploading every per-body array, find the
// removed set by walking the id arrays (order is preserved by RemoveBody), check that nothing else changed that the
// device cannot patch, and replay the edit on the device (sol_remove_bodies + sol_patch_body).
// Returns 1 if the device now equals the host, 0 if the caller has to fall back to the full upload, -1 on error.
static int replay_edit_on_device(Bridge *b, BodyData *bd, const int counts[7], int n)
{
	if (b->n <= 0 || n >= b->n || n < 1 || !b->host_fresh) return 0;
	std::vector<int> removed, from(n);
	int j = 0;
	for (int i = 0; i < b->n; i++) {
		if (j < n && b->id[i] == bd->id[j]) from[j++] = i;
		else removed.push_back(i);
	}
	if (j != n || (int)removed.size() != b->n - n || removed.empty() || removed[0] == 0) return 0;
	std::vector<int> patches;
	for (j = 0; j < n; j++) {
		const int i = from[j];
		if (bd->type[j] != b->type[i] || bd->migType[j] != b->migType[i] ||
		    memcmp(&bd->gammaStokes[j], &b->gS[i], sizeof(double)) != 0 || memcmp(&bd->gammaEpstein[j], &b->gE[i], sizeof(double)) != 0)
			return 0;
		// cD and migStopAt keep their SLOTS in RemoveBody (Simulator.cpp:757-769), on the device as well
		if (memcmp(&bd->cD[j], &b->cD[j], sizeof(double)) != 0 || memcmp(&bd->migStopAt[j], &b->migStop[j], sizeof(double)) != 0) return 0;
		if (memcmp(&bd->mass[j], &b->mass[i], sizeof(double)) != 0 || memcmp(&bd->radius[j], &b->radius[i], sizeof(double)) != 0 ||
		    memcmp(&bd->density[j], &b->density[i], sizeof(double)) != 0 || memcmp(&bd->y0[6 * j], &b->y0[6 * (size_t)i], 6 * sizeof(double)) != 0)
			patches.push_back(j);
	}
	if (patches.size() > 64) return 0;          // not a merger replay (e.g. a bulk edit): the full upload is cheaper
	int expect[7];
	memcpy(expect, b->counts, sizeof(expect));
	for (size_t m = 0; m < removed.size(); m++) {
		const int t = b->type[removed[m]];
		if (t < 1 || t > 7) return 0;
		expect[t - 1]--;
	}
	if (memcmp(expect, counts, sizeof(expect)) != 0) return 0;
	if (sol_remove_bodies(b->ctx, &removed[0], (int)removed.size()) != SOL_OK) return -1;
	for (size_t m = 0; m < patches.size(); m++) {
		const int k = patches[m];
		if (sol_patch_body(b->ctx, k, &bd->y0[6 * k], bd->mass[k], bd->radius[k], bd->density[k]) != SOL_OK) return -1;
	}
	return 1;
}

int sync_in(Bridge *b, Acceleration *acc, BodyData *bd)
{
	NBodies &nb = bd->nBodies;
	const int counts[7] = {nb.centralBody, nb.giantPlanet, nb.rockyPlanet, nb.protoPlanet, nb.superPlanetsimal, nb.planetsimal, nb.testParticle};
	const int n = nb.total;
	const bool same_arrays = (const void *)bd->mass == b->mass_ptr;      // BodyData was not re-allocated
	const bool maybe_changed = nb.removed != b->removed_seen || !same_arrays;
	b->removed_seen = nb.removed;
	b->mass_ptr = bd->mass;
	bool params_same = (n == b->n) && memcmp(counts, b->counts, sizeof(counts)) == 0;
	if (params_same && maybe_changed)
		params_same = same(b->mass, bd->mass, n) &&
	                   same(b->radius, bd->radius, n) && same(b->density, bd->density, n) && same(b->cD, bd->cD, n) &&
	                   same(b->gS, bd->gammaStokes, n) && same(b->gE, bd->gammaEpstein, n) && same(b->migStop, bd->migStopAt, n) &&
	                   same(b->type, bd->type, n) && same(b->migType, bd->migType, n) && same(b->id, bd->id, n);
	if (!params_same) {
		const int replayed = same_arrays ? replay_edit_on_device(b, bd, counts, n) : 0;
		if (replayed < 0) return fail(b, "sol_remove_bodies / sol_patch_body");
		if (replayed == 0 &&
		    sol_set_bodies(b->ctx, counts, bd->y0, bd->mass, bd->radius, bd->density, bd->cD, bd->gammaStokes, bd->gammaEpstein,
		                   bd->migStopAt, bd->type, bd->migType, bd->id) != SOL_OK)
			return fail(b, "sol_set_bodies");
		if (replayed == 1) b->edits_replayed++;
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
		if (replayed == 0) b->nebula_set = false;   // the gas constants depend on mass[0] (sol_patch_body refreshes them itself)
		b->host_fresh = true;
	} else if (b->host_fresh && !same(b->y0, bd->y0, 6 * n)) {
		// (when the host copy is stale - resident mode - the device state is the authority)
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

int run_driver(int integrator, BodyData *bd, Acceleration *acc, TimeLine *tl, double *time, double *hNext, double *hDid,
               const char *file, const char *function, long line, const char *step_error_message)
{
	Bridge *b = bridge_of(acc);
	if (b == 0) return 1;
	// resident synchronisation needs the CheckEvent hook (which brings the thresholds); without it every step ends
	// with a download, which is always correct
	const bool resident = b->thresholds_known && !eager_forced();
	double t0 = now_s();
	if (sync_in(b, acc, bd) == 1) return 1;
	b->t_sync_in += now_s() - t0;
	// small systems: the device runs ahead of the host program (sol_bridge.h) and does the flushes of the steps it takes itself
	// (opt-in, SOLARIS_B200_RUN_AHEAD=1: the persistent kernel takes its step sizes from the DEVICE's pow(), so its time grid
	//  differs from the host drivers' in the last bits and - RKF78's own global error being ~1e-8 at epsilon = 1e-10 - the
	//  trajectories then agree with the reference to ~1e-10 over a few hundred steps, not over tens of thousands)
	static const bool run_ahead = getenv("SOLARIS_B200_RUN_AHEAD") != 0 && std::string(getenv("SOLARIS_B200_RUN_AHEAD")) == "1";
	const bool small = run_ahead && resident && b->n <= 32 && b->n == b->counts[0] + b->counts[1] + b->counts[2] + b->counts[3];
	if (resident && !small && !b->host_fresh && b->steps_done > 0 && b->steps_done % Constants::CheckForSM == 0) {
		// Simulator just flushed its (stale) host copies (Simulator.cpp:159-162); do the real one on the device
		if (sol_flush_tiny(b->ctx, Constants::SmallestNumber) != SOL_OK) return fail(b, "sol_flush_tiny");
	}
	if (small) {
		if (b->ahead_pos >= b->ahead_count) {
			sol_run_args A;
			memset(&A, 0, sizeof(A));
			const int K = 1024;
			b->ahead.resize(4 * (size_t)K);
			A.integrator = integrator; A.max_steps = K; A.time = *time; A.h_next = *hNext;
			A.millenium_days = 1000.0 * Constants::YearToDay * tl->millenium;
			A.length = tl->length; A.output = tl->output; A.last_save = tl->lastSave;
			A.ejection = b->ejection; A.hit_centrum = b->hitCentrum; A.collision_factor = b->collisionFactor;
			A.step_counter = b->steps_done; A.flush_every = Constants::CheckForSM; A.flush_threshold = Constants::SmallestNumber;
			A.records = &b->ahead[0];
			t0 = now_s();
			const int rc = sol_run(b->ctx, &A);
			b->t_step += now_s() - t0;
			b->batches++;
			b->ahead_count = A.steps; b->ahead_pos = 0; b->ahead_stop = A.stop_reason; b->ahead_time_in = *time;
			b->host_fresh = false;
			if (rc != SOL_OK && A.steps == 0) {
				const char *msg = sol_last_error(b->ctx);
				Error::_errMsg = (msg != 0 && msg[0] != 0) ? msg : step_error_message;
				Error::PushLocation(file, function, line);
				return 1;
			}
			// (a Driver failure after some good steps surfaces when the host program asks for the failing step)
		}
		const double *rec = &b->ahead[4 * (size_t)b->ahead_pos];
		if (memcmp(time, &b->ahead_time_in, sizeof(double)) != 0 || memcmp(hNext, &rec[3], sizeof(double)) != 0) {
			Error::_errMsg = "solaris_b200: the host program entered a Driver with a time / trial step the device did not run ahead with";
			Error::PushLocation(file, function, line);
			return 1;
		}
		*time = rec[0]; *hDid = rec[1]; *hNext = rec[2];
		b->ahead_time_in = rec[0];
		b->ahead_pos++;
		b->steps_done++;
		const bool last = b->ahead_pos == b->ahead_count;
		b->event_pending = last && b->ahead_stop == SOL_RUN_EVENT;
		if (last && b->ahead_stop == SOL_RUN_ERROR) { b->ahead_count = b->ahead_pos = 0; }   // the next call re-runs the failing step and reports it
		t0 = now_s();
		if (last && (b->ahead_stop == SOL_RUN_EVENT || b->ahead_stop == SOL_RUN_END || b->ahead_stop == SOL_RUN_SAVE)) {
			if (b->event_pending && sol_download(b->ctx, SOL_Y, bd->y0) != SOL_OK) return fail(b, "sol_download(y)");
			if (sync_out(b, acc, bd, bd->y) == 1) return 1;
			b->host_fresh = true;
			b->downloads++;
		}
		b->t_sync_out += now_s() - t0;
		return 0;
	}
	double info[4] = {0, 0, 0, 0};
	t0 = now_s();
	if (sol_step(b->ctx, integrator, time, hNext, hDid, info) != SOL_OK) {
		const char *msg = sol_last_error(b->ctx);
		Error::_errMsg = (msg != 0 && msg[0] != 0) ? msg : step_error_message;
		Error::PushLocation(file, function, line);
		return 1;
	}
	b->t_step += now_s() - t0;
	b->steps_done++;
	bool download = true;
	b->event_pending = true;           // eager: the host scan decides
	t0 = now_s();
	if (resident) {
		int counts[3] = {0, 0, 0};
		if (sol_detect_events(b->ctx, b->ejection, b->hitCentrum, b->collisionFactor, counts) != SOL_OK) return fail(b, "sol_detect_events");
		b->event_pending = counts[0] > 0 || counts[1] > 0 || counts[2] > 0;
		// the two predicates of Simulator::DecisionMaking that make the host read the state (Simulator.cpp:197,219,234)
		const double actualTime = 1000.0 * Constants::YearToDay * tl->millenium + *time;
		const bool will_end = fabs(actualTime) >= fabs(tl->length);
		const bool will_save = fabs(tl->lastSave + *hDid) >= fabs(tl->output);
		download = b->event_pending || will_end || will_save;
	}
	b->t_detect += now_s() - t0;
	t0 = now_s();
	if (download) {
		// Simulator::CheckEvent builds its Collision records from bodyData.y, the state BEFORE this step
		// (Simulator.cpp:709): after the caller's std::swap that is the buffer bd->y0 points to now.  When the
		// host copy was stale it has to be refreshed too (the device keeps the previous state in SOL_Y).
		if (!b->host_fresh && b->event_pending && sol_download(b->ctx, SOL_Y, bd->y0) != SOL_OK) return fail(b, "sol_download(y)");
		// the new state lands in the host array that becomes y0 after the caller's std::swap
		if (sync_out(b, acc, bd, bd->y) == 1) return 1;
		b->host_fresh = true;
		b->downloads++;
	} else {
		b->host_fresh = false;
	}
	b->t_sync_out += now_s() - t0;
	return 0;
}

}  // namespace solb200

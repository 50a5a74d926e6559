// Host-side bridge between the reference's C++ class interface and the C-ABI (include/solaris_b200.h).
//
// The drop-in translation units in this directory (Acceleration.cpp, RungeKutta4.cpp,
// RungeKuttaFehlberg78.cpp, DormandPrince.cpp) are compiled AGAINST THE REFERENCE'S OWN HEADERS
// (-I<reference>/Solaris) so the class layouts seen by the unchanged Simulator.cpp are identical; all
// extra state lives in a side table keyed by the Acceleration object (SURVEY.md §8b "Dispatch").
//
// RESIDENT MODE (opt-in: SOLARIS_B200_RESIDENT=1).  The Driver is not told the event thresholds or the output
// cadence, so by default every step ends with a download.  In resident mode the bridge learns the three
// thresholds itself - from SOLARIS_B200_EJECTION / _HITCENTRUM / _COLLISION_FACTOR, or by reading the
// <Ejection>, <HitCentrum>, <Collision> elements of the input file named on the command line - runs the
// device flag reduction (sol_detect_events) after each step, and replicates the two predicates of
// Simulator::DecisionMaking that make the host read the state (end of integration, snapshot due;
// Simulator.cpp:219,234).  Only then - and on the step after an event, so that no stale firing value is left in
// the host's rm3 / NN arrays - are y0, rm3, the NN arrays and migType copied back: "only event records leave
// the device".  The flush-to-zero of every 100th step (Simulator.cpp:159-162) is done on the device.
//
// Synchronisation policy ("eager", correct for an unmodified Simulator): the host BodyData stays the
// authority between Driver calls.  On entry the bridge compares the host arrays with its shadow of
// what the device holds and re-uploads what the host changed (collision merges, body removal, the
// flush-to-zero every 100 steps); on exit it downloads the new y0, rm3, nearest-neighbour arrays and
// migType, which is everything Simulator::DecisionMaking / CheckEvent read (Simulator.cpp:181-248,
// 621-735).  That is 72 N bytes of PCIe traffic per step, negligible for the O(N * N_src) configs.
#pragma once
#include <vector>

#include "../../include/solaris_b200.h"

class Acceleration;
class BodyData;
class Nebula;
class TimeLine;

namespace solb200 {

struct Bridge {
	sol_ctx *ctx = nullptr;
	bool failed = false;
	// shadow of the device-resident system
	int counts[7] = {0, 0, 0, 0, 0, 0, 0};
	int n = 0;
	std::vector<double> y0, mass, radius, density, cD, gS, gE, migStop;
	std::vector<int> type, migType, id;
	bool nebula_set = false;
	// cheap change detection for the per-body parameters: they can only change when Simulator removes a
	// body (collision / ejection / hit centrum: NBodies::removed grows, Simulator.cpp:737-771) or rebuilds
	// BodyData (new allocation).  The full array compare runs only then.
	int removed_seen = -1;
	const void *mass_ptr = 0;
	// resident mode (opt-in, see sol_bridge.cpp): host arrays are refreshed only when Simulator can observe them
	bool host_fresh = true;      // host y0 == device y0
	bool side_hot = true;        // host rm3 / NN arrays may hold values that fire an event (initially: zeros / unset)
	long downloads = 0;          // state downloads after a step
	long edits_replayed = 0;     // event steps whose merge / removal was replayed on the device instead of re-uploaded
	long steps_done = 0;         // successful Driver calls (== Simulator's counter.succededStep)
	double t_sync_in = 0, t_step = 0, t_detect = 0, t_sync_out = 0;   // seconds spent inside run_driver, by phase
};

// Finds (or creates) the bridge of an Acceleration object; NULL + Error::_errMsg on failure.
Bridge *bridge_of(Acceleration *acc);
void bridge_release(Acceleration *acc);
// The bridge whose Acceleration object works on this BodyData (for Calculate::Integrals); NULL if none.
Bridge *bridge_of_bodydata(BodyData *bd, Acceleration **acc_out);
// the bridge whose BodyData currently exposes exactly these arrays as y0 / id (0 if none)
Bridge *bridge_of_state(const double *y0, const int *id, int n, Acceleration **acc_out);

// Makes the device system equal to the host BodyData (uploads only what differs). 0 / 1.
int sync_in(Bridge *b, Acceleration *acc, BodyData *bd);
// After a successful sol_step: new state into `dst_y` (6n), side outputs into acc->rm3,
// bd->indexOfNN / distanceOfNN / migType; refreshes the shadow. 0 / 1.
int sync_out(Bridge *b, Acceleration *acc, BodyData *bd, double *dst_y);

// Runs one Driver on the device: shared body of the three drop-in Drivers.
int run_driver(int integrator, BodyData *bd, Acceleration *acc, TimeLine *timeLine, double *time, double *hNext, double *hDid,
               const char *file, const char *function, long line, const char *step_error_message);

}  // namespace solb200

// Host-side bridge between the reference's C++ class interface and the C-ABI (include/solaris_b200.h).
//
// The drop-in translation units in this directory (Acceleration.cpp, RungeKutta4.cpp,
// RungeKuttaFehlberg78.cpp, DormandPrince.cpp, Calculate.cpp, SavePhases.cpp, SimulatorHooks.cpp) are compiled
// AGAINST THE REFERENCE'S OWN HEADERS (-I<reference>/Solaris) so the class layouts seen by the unchanged
// Simulator.cpp are identical; all extra state lives in a side table keyed by the Acceleration object
// (SURVEY.md §8b "Dispatch").
//
// RESIDENT SYNCHRONISATION (the default).  The state lives on the device; the host arrays of BodyData are
// refreshed only when the host program is about to read them.  Who reads them, and how the bridge knows:
//   * Simulator::CheckEvent (Simulator.cpp:621-735) reads rm3, the nearest-neighbour arrays, radius, y0 and y.
//     SimulatorHooks.cpp puts a hook in front of it: Simulator::BodyListToBodyData hands the three thresholds of
//     Settings (ejection, hitCentrum, collision->factor) to the bridge, every Driver call ends with the device flag
//     reduction (sol_detect_events), and the hooked CheckEvent runs the reference's own function - on freshly
//     downloaded arrays - only when a count is non-zero.  "Only event records leave the device."
//   * Simulator::DecisionMaking (Simulator.cpp:219,234) ends the integration (UpdateBodyListAfterIntegration
//     reads y0) or saves a snapshot; the Driver is handed the TimeLine, so it evaluates the same two predicates
//     and downloads on those steps.
//   * The flush-to-zero of every 100th step (Simulator.cpp:159-162) is repeated on the device when the host
//     copy it ran on was stale.
// Between such steps nothing crosses the bus but the 8-byte error norm and the event counts.
//
// RUNNING AHEAD (opt-in: SOLARIS_B200_RUN_AHEAD=1; systems of at most 32 bodies, all massive - SunJupiter, SolarSystem).  A 2- or 9-body step is a few
// microseconds of dependent arithmetic; one launch + synchronise per Driver call costs several times that.  For such a
// system the first Driver call of a stretch hands the whole TimeLine to sol_run, whose persistent kernel repeats what
// Simulator::Integrate would make happen next - Driver, DecisionMaking's step-size clamps, the event tests, the
// 100-step flush - until a step ends the integration, makes a snapshot due or fires an event (or 1024 steps are done).
// The following Driver calls each hand out one recorded step without touching the device, after checking that the
// host program enters them with exactly the time and trial step the device assumed (it always does: the clamps are the
// reference's own formulas; a mismatch is reported as an error, never papered over).  Opt-in because sol_run's
// persistent kernel evaluates the step-size formulas with the device's pow(): see run_driver.
//
// EAGER SYNCHRONISATION (SOLARIS_B200_EAGER=1, and whenever the hooks are not linked - e.g. a host program that
// uses the integrator classes without Simulator): after every step y0, rm3, the nearest-neighbour arrays and
// migType are downloaded (72 N bytes), and on entry the host y0 is compared with the bridge's shadow and
// re-uploaded if the host edited it.  Kept for A/B tests of the resident mode.
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
	// event thresholds of Settings, handed over by the Simulator::BodyListToBodyData hook
	bool thresholds_known = false;
	double ejection = 0.0, hitCentrum = 0.0, collisionFactor = 0.0;
	// resident synchronisation state
	bool host_fresh = true;      // host y0 == device y0
	bool event_pending = false;  // the device flag reduction of the last step found at least one candidate
	// Steps the device has already taken ahead of the host program (small systems, see run_driver): records of sol_run
	// {time, hDid, hNext as the Driver proposed it, trial step it was entered with} that the next Driver calls hand out
	std::vector<double> ahead;
	int ahead_count = 0, ahead_pos = 0, ahead_stop = 0;
	double ahead_time_in = 0.0;  // TimeLine::time the next handed-out step must be entered with
	long batches = 0;            // sol_run launches
	long downloads = 0;          // state downloads after a step
	long edits_replayed = 0;     // event steps whose merge / removal was replayed on the device instead of re-uploaded
	long steps_done = 0;         // successful Driver calls (== Simulator's counter.succededStep)
	long host_scans_skipped = 0; // CheckEvent calls answered by the device flag reduction alone
	double t_sync_in = 0, t_step = 0, t_detect = 0, t_sync_out = 0;   // seconds spent inside run_driver, by phase
};

// Finds (or creates) the bridge of an Acceleration object; NULL + Error::_errMsg on failure.
Bridge *bridge_of(Acceleration *acc);
// Finds the bridge of an Acceleration object without creating one (NULL if none).
Bridge *bridge_lookup(Acceleration *acc);
void bridge_release(Acceleration *acc);
// The bridge whose Acceleration object works on this BodyData (for Calculate::Integrals); NULL if none.
Bridge *bridge_of_bodydata(BodyData *bd, Acceleration **acc_out);
// the bridge whose BodyData currently exposes exactly these arrays as y0 / id (0 if none)
Bridge *bridge_of_state(const double *y0, const int *id, int n, Acceleration **acc_out);

// Settings::ejection / hitCentrum / collision->factor (0 = criterion off), from the BodyListToBodyData hook.
void bridge_set_thresholds(Bridge *b, double ejection, double hitCentrum, double collisionFactor);

// Makes the device system equal to the host BodyData (uploads only what differs). 0 / 1.
int sync_in(Bridge *b, Acceleration *acc, BodyData *bd);
// After a successful sol_step: new state into `dst_y` (6n), side outputs into acc->rm3,
// bd->indexOfNN / distanceOfNN / migType; refreshes the shadow. 0 / 1.
int sync_out(Bridge *b, Acceleration *acc, BodyData *bd, double *dst_y);

// Runs one Driver on the device: shared body of the three drop-in Drivers.
int run_driver(int integrator, BodyData *bd, Acceleration *acc, TimeLine *timeLine, double *time, double *hNext, double *hDid,
               const char *file, const char *function, long line, const char *step_error_message);

}  // namespace solb200

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
#include "Acceleration.h"
#include "BodyData.h"
#include "Error.h"
#include "EventCondition.h"
#include "Settings.h"
#include "Simulation.h"
#include "Simulator.h"
#include "sol_bridge.h"

using namespace solb200;

// the reference's members under their new names (`this` is the first argument in the ABI)
extern "C" int solb200_reference_Simulator_BodyListToBodyData(Simulator *self);
extern "C" int solb200_reference_Simulator_CheckEvent(Simulator *self, double timeOfEvent);

int Simulator::BodyListToBodyData()
{
	if (solb200_reference_Simulator_BodyListToBodyData(this) == 1) return 1;
	Bridge *b = _acceleration != 0 ? bridge_of(_acceleration) : 0;
	if (b != 0) {
		const Settings &s = _simulation->settings;
		// the same three values Simulator::CheckEvent reads (Simulator.cpp:626-627,690-692)
		bridge_set_thresholds(b, s.ejection, s.hitCentrum, s.collision != 0 ? s.collision->factor : 0.0);
	}
	return 0;
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

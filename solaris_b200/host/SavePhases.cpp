// Drop-in replacement for ONE member function of the reference's BinaryFileAdapter:
//   void BinaryFileAdapter::SavePhases(double time, int n, double *y, int *id, OutputType type)
//   (Solaris/BinaryFileAdapter.cpp:107-158).
// The reference appends a snapshot to Phases.dat with 2 n small stream writes.  Here the BINARY record
// (double time, int n, n x {int id, double y[6]}, no padding) is assembled on the device from the resident
// state and appended with one write (sol_write_phases).  Everything else - the TEXT format, a `y` that is not
// the y0 array of a BodyData the device holds - is handed to the reference's own function, which
// build_dropin.sh keeps under another symbol name (objcopy --redefine-sym on a copy of the reference's
// object file); every other member of BinaryFileAdapter stays the reference's code as well.
#include <cstdlib>
#include <string>

#include "Acceleration.h"
#include "BinaryFileAdapter.h"
#include "BodyData.h"
#include "Output.h"
#include "sol_bridge.h"

using namespace solb200;

// the reference's BinaryFileAdapter::SavePhases under its new name (`this` is the first argument in the ABI)
extern "C" void solb200_reference_SavePhases(BinaryFileAdapter *self, double time, int n, double *y, int *id,
                                             BinaryFileAdapter::OutputType type);

void BinaryFileAdapter::SavePhases(double time, int n, double *y, int *id, OutputType type)
{
	Acceleration *acc = 0;
	Bridge *b = (type == BINARY) ? bridge_of_state(y, id, n, &acc) : 0;
	// sync_in re-uploads whatever the host edited since the last step, so the device record equals (y, id)
	if (b == 0 || sync_in(b, acc, acc->bodyData) == 1) {
		solb200_reference_SavePhases(this, time, n, y, id, type);
		return;
	}
	const std::string path = output->GetPath(output->phases);
	if (sol_write_phases(b->ctx, path.c_str(), time) != SOL_OK) {
		Log(std::string(sol_last_error(b->ctx)), true);   // (the reference logs and exits on a failed open as well)
		exit(1);
	}
}

// Drop-in replacement for THREE members of the reference's Calculate class (declared by its Calculate.h):
// Integrals, PotentialEnergy, Energy.  The O(n) members (TotalMass, PhaseOfBC, PhaseWithRespectToBC, AngularMomentum,
// KineticEnergy, TransformTo*) stay the reference's own code: build_dropin.sh renames the three symbols in a copy of
// the reference's object file (objcopy --redefine-sym) and links that copy.
// Calculate::Integrals runs at every snapshot and its PotentialEnergy is O(n^2)
// over ALL bodies (Calculate.cpp:139-159) - 10^12 pair evaluations at n = 10^6, which would dwarf the
// accelerated step loop.  Integrals / PotentialEnergy / Energy go to the device (sol_integrals).
#include <cmath>
#include <cstring>

#include "Calculate.h"
#include "Acceleration.h"
#include "BodyData.h"
#include "Constants.h"
#include "Error.h"
#include "Phase.h"
#include "Vector.h"
#include "sol_bridge.h"

using namespace solb200;

static int device_integrals(BodyData *bodyData, double *out16)
{
	Acceleration *acc = 0;
	Bridge *b = bridge_of_bodydata(bodyData, &acc);
	if (b == 0) {
		Error::_errMsg = "solaris_b200: Calculate called before an Acceleration object exists for this BodyData";
		Error::PushLocation(__FILE__, __FUNCTION__, __LINE__);
		return 1;
	}
	if (sync_in(b, acc, bodyData) == 1) return 1;
	if (sol_integrals(b->ctx, out16) != SOL_OK) {
		Error::_errMsg = sol_last_error(b->ctx);
		Error::PushLocation(__FILE__, __FUNCTION__, __LINE__);
		return 1;
	}
	return 0;
}

int Calculate::Integrals(BodyData *bodyData)
{
	return device_integrals(bodyData, bodyData->integrals);
}

int Calculate::PotentialEnergy(BodyData *bodyData, double &result)
{
	double tmp[16];
	if (device_integrals(bodyData, tmp) == 1) return 1;
	result = tmp[14];
	return 0;
}

int Calculate::Energy(BodyData *bodyData, double &result)
{
	double tmp[16];
	if (device_integrals(bodyData, tmp) == 1) return 1;
	result = tmp[15];
	return 0;
}

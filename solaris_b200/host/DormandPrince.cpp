// Drop-in replacement for Solaris/DormandPrince.cpp (class declared by the reference's header).
// Driver contract: DormandPrince.cpp:126-170 (RKN7(6), Step2 form).
#include <cmath>
#include <algorithm>

#include "DormandPrince.h"
#include "Acceleration.h"
#include "BodyData.h"
#include "Error.h"
#include "TimeLine.h"
#include "sol_bridge.h"

DormandPrince::DormandPrince()
{
	name          = "Dormand-Prince 7(6) (B200 device build)";
	reference     = "New Runge-Kutta Algorithms for Numerical Simulation in Dynamical Astronomy, Celestial Mechanics, Vol. 18(1978), 223-232.";
	accuracy      = -10.0;
	epsilon       = pow(10, accuracy);
	maxIter       = 10;
	sizeHeightRKD = 9;
}

int DormandPrince::Driver(BodyData *bodyData, Acceleration *acceleration, TimeLine *timeLine)
{
	bodyData->time = timeLine->time;
	acceleration->evaluateGasDrag = true;
	double time = timeLine->time, hNext = timeLine->hNext, hDid = 0.0;
	if (solb200::run_driver(SOL_DORMAND_PRINCE, bodyData, acceleration, timeLine, &time, &hNext, &hDid, __FILE__, __FUNCTION__, __LINE__,
	                        "An error occurred during Prince-Dormand step!") == 1)
		return 1;
	acceleration->evaluateTypeIMigration  = false;
	acceleration->evaluateTypeIIMigration = false;
	bodyData->h     = hDid;          // the last trial h, :152
	timeLine->hDid  = hDid;
	timeLine->hNext = hNext;
	timeLine->time  = time;
	bodyData->time  = time;
	std::swap(bodyData->y0, bodyData->y);
	return 0;
}

static int fused(const char *what)
{
	Error::_errMsg = std::string("solaris_b200: DormandPrince::") + what + " is fused into Driver() on the device";
	Error::PushLocation(__FILE__, __FUNCTION__, __LINE__);
	return 1;
}

int DormandPrince::Step(BodyData *, Acceleration *) { return fused("Step"); }
int DormandPrince::Step2(BodyData *, Acceleration *) { return fused("Step2"); }

double DormandPrince::GetErrorMax(int n, const double *yerr)
{   // DormandPrince.cpp:493-503
	double errorMax = 0.0;
	for (int i = 0; i < n; i++) {
		double error = fabs(yerr[i]);
		if (error > errorMax) errorMax = error;
	}
	return errorMax;
}

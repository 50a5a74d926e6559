// Drop-in replacement for Solaris/RungeKuttaFehlberg78.cpp (class declared by the reference's header).
// Driver contract: RungeKuttaFehlberg78.cpp:66-140.  The 13 stages, the solution, the error max-norm
// and the accept/reject loop run in sol_step(); the tableau members of the class stay unused.
#include <cmath>
#include <algorithm>

#include "RungeKuttaFehlberg78.h"
#include "Acceleration.h"
#include "BodyData.h"
#include "Error.h"
#include "TimeLine.h"
#include "sol_bridge.h"

RungeKuttaFehlberg78::RungeKuttaFehlberg78()
{
	name      = "Runge-Kutta-Fehlberg 7(8) with stepsize control (B200 device build)";
	reference = "NASA Technical Reports R-381, by Erwin Fehlberg, 1972.";
	accuracy  = -10.0;
	epsilon   = pow(10, accuracy);
}

int RungeKuttaFehlberg78::Driver(BodyData *bodyData, Acceleration *acceleration, TimeLine *timeLine)
{
	bodyData->time = timeLine->time;
	bodyData->h    = timeLine->hNext;
	acceleration->evaluateGasDrag = true;
	double time = timeLine->time, hNext = timeLine->hNext, hDid = 0.0;
	if (solb200::run_driver(SOL_RUNGE_KUTTA_FEHLBERG78, bodyData, acceleration, timeLine, &time, &hNext, &hDid, __FILE__, __FUNCTION__,
	                        __LINE__, "An error occurred during Runge-Kutta-Fehlberg7(8) step!") == 1)
		return 1;
	acceleration->evaluateTypeIMigration  = false;
	acceleration->evaluateTypeIIMigration = false;
	timeLine->hDid  = hDid;
	timeLine->time  = time;          // == time + hDid, :126
	bodyData->time  = time;
	timeLine->hNext = hNext;
	bodyData->h     = hNext;         // :130
	std::swap(bodyData->y0, bodyData->y);
	return 0;
}

int RungeKuttaFehlberg78::Step(BodyData *, Acceleration *)
{
	Error::_errMsg = "solaris_b200: RungeKuttaFehlberg78::Step is fused into Driver() on the device";
	Error::PushLocation(__FILE__, __FUNCTION__, __LINE__);
	return 1;
}

// Pure helper of the public interface (RungeKuttaFehlberg78.cpp:252-262); the device computes the same
// max-norm inside sol_step.
double RungeKuttaFehlberg78::GetErrorMax(const int n, const double *yerr, const double *yscale)
{
	double errorMax = 0.0;
	for (int i = 0; i < n; i++) {
		double err = fabs(yerr[i] / yscale[i]);
		if (err > errorMax) errorMax = err;
	}
	return errorMax / epsilon;
}

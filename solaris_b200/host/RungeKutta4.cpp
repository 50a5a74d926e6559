// Drop-in replacement for Solaris/RungeKutta4.cpp (class declared by the reference's RungeKutta4.h).
// Driver keeps the reference's contract (RungeKutta4.cpp:20-56): TimeLine::hDid/time/hNext,
// BodyData::time/h, std::swap(y0, y); the step itself runs on the device (sol_step).
#include <cmath>
#include <algorithm>

#include "RungeKutta4.h"
#include "Acceleration.h"
#include "BodyData.h"
#include "Error.h"
#include "TimeLine.h"
#include "sol_bridge.h"

RungeKutta4::RungeKutta4()
{
	name      = "The classical Runge-Kutta method (B200 device build)";
	reference = "";
	accuracy  = -10.0;
	epsilon   = pow(10, accuracy);
}

int RungeKutta4::Driver(BodyData *bodyData, Acceleration *acceleration, TimeLine *timeLine)
{
	bodyData->time = timeLine->time;
	bodyData->h    = timeLine->hNext;
	acceleration->evaluateGasDrag = true;
	double time = timeLine->time, hNext = timeLine->hNext, hDid = 0.0;
	if (solb200::run_driver(SOL_RUNGE_KUTTA4, bodyData, acceleration, timeLine, &time, &hNext, &hDid, __FILE__, __FUNCTION__, __LINE__,
	                        "An error occurred during Runge-Kutta4 step!") == 1)
		return 1;
	acceleration->evaluateTypeIMigration  = false;   // state the reference leaves behind (:36-37)
	acceleration->evaluateTypeIIMigration = false;
	timeLine->hDid  = hDid;
	timeLine->time  = time;
	bodyData->time  = time;
	timeLine->hNext = hNext;
	std::swap(bodyData->y0, bodyData->y);
	return 0;
}

int RungeKutta4::Step(BodyData *, Acceleration *)
{
	Error::_errMsg = "solaris_b200: RungeKutta4::Step is fused into Driver() on the device";
	Error::PushLocation(__FILE__, __FUNCTION__, __LINE__);
	return 1;
}

// Drop-in replacement for Solaris/Calculate.cpp (class declared by the reference's Calculate.h).
// Calculate::Integrals runs at every snapshot and its PotentialEnergy is O(n^2)
// over ALL bodies (Calculate.cpp:139-159) - 10^12 pair evaluations at n = 10^6, which would dwarf the
// accelerated step loop.  Integrals / PotentialEnergy / Energy go to the device (sol_integrals); the O(n)
// helpers stay host loops.
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

void Calculate::TransformToAC(Phase *, Phase *) {}     // unimplemented in the reference as well
void Calculate::TransformToAC(double *, double *) {}

void Calculate::TransformToBC(Phase *phase, Phase *phaseOfBC)
{
	phase->position.x -= phaseOfBC->position.x; phase->position.y -= phaseOfBC->position.y; phase->position.z -= phaseOfBC->position.z;
	phase->velocity.x -= phaseOfBC->velocity.x; phase->velocity.y -= phaseOfBC->velocity.y; phase->velocity.z -= phaseOfBC->velocity.z;
}

void Calculate::TransformToBC(double *y, double *bc)
{
	for (int j = 0; j < 6; j++) y[j] -= bc[j];
}

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

int Calculate::TotalMass(BodyData *bodyData, double &result)
{
	result = 0.0;
	const int nm = bodyData->nBodies.NOfMassive();
	for (int i = 0; i < nm; i++) result += bodyData->mass[i];
	return 0;
}

int Calculate::PhaseOfBC(BodyData *bodyData, double *bc)
{
	double M = 0.0;
	TotalMass(bodyData, M);
	for (int j = 0; j < 6; j++) bc[j] = 0.0;
	const int n = bodyData->nBodies.total;
	for (int i = 0; i < n; i++)
		for (int j = 0; j < 6; j++) bc[j] += bodyData->mass[i] * bodyData->y0[6 * i + j];
	for (int j = 0; j < 6; j++) bc[j] /= M;
	return 0;
}

int Calculate::PhaseWithRespectToBC(BodyData *bodyData, double *bc)
{
	const int n = bodyData->nBodies.total;
	for (int i = 0; i < n; i++)
		for (int j = 0; j < 6; j++) bodyData->y0[6 * i + j] -= bc[j];
	return 0;
}

int Calculate::AngularMomentum(BodyData *bodyData, Vector *result)
{
	double lx = 0.0, ly = 0.0, lz = 0.0;
	const int n = bodyData->nBodies.total;
	for (int i = 0; i < n; i++) {
		const double *r = &bodyData->y0[6 * i], *v = r + 3;
		const double m = bodyData->mass[i];
		lx += m * (r[1] * v[2] - r[2] * v[1]);
		ly += m * (r[2] * v[0] - r[0] * v[2]);
		lz += m * (r[0] * v[1] - r[1] * v[0]);
	}
	result->x = lx; result->y = ly; result->z = lz;
	return 0;
}

int Calculate::KineticEnergy(BodyData *bodyData, double &result)
{
	result = 0.0;
	const int n = bodyData->nBodies.total;
	for (int i = 0; i < n; i++) {
		const double *v = &bodyData->y0[6 * i + 3];
		result += 0.5 * bodyData->mass[i] * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
	}
	return 0;
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

// Drop-in replacement for Solaris/Acceleration.cpp: same class (declared by the reference's own
// Acceleration.h), force evaluation on the B200 through the C-ABI.  Link this INSTEAD of the
// reference's Acceleration.cpp; Simulator.cpp and everything else stay unchanged.
//
// Only the members other translation units use are meaningful here:
//   Acceleration(IntegratorType, bool, BodyData*, Nebula*)   Simulator.cpp:82
//   Compute(t, y, totalAccel)                                 seam B (Acceleration.h:19)
//   rm3, evaluate* flags, accel* caches                       Simulator.cpp:633,640; the Drivers
// The per-term member functions (GravityAC, GasDragAC, ...) have no callers outside the reference's
// own Acceleration.cpp; they report an error instead of silently computing on the CPU.
#include <cstring>

#include "Acceleration.h"
#include "BodyData.h"
#include "Error.h"
#include "Nebula.h"
#include "sol_bridge.h"

using namespace solb200;

Acceleration::Acceleration(IntegratorType iType, bool baryCentric, BodyData *bD, Nebula *n)
{
	_integratorType         = iType;
	_baryCentric            = baryCentric;
	evaluateGasDrag         = true;
	evaluateTypeIMigration  = true;
	evaluateTypeIIMigration = true;
	bodyData                = bD;
	nebula                  = n;
	rm3                     = 0;
	accelGasDrag            = 0;
	accelMigrationTypeI     = 0;
	accelMigrationTypeII    = 0;
	Bridge *b = bridge_of(this);          // creates the device context; a failure surfaces at first use
	if (b != 0) sol_set_frame(b->ctx, baryCentric ? 1 : 0);
}

Acceleration::~Acceleration()
{
	bridge_release(this);
	delete[] rm3;
	delete[] accelGasDrag;
	delete[] accelMigrationTypeI;
	delete[] accelMigrationTypeII;
}

int Acceleration::Compute(double t, double *y, double *totalAccel)
{
	Bridge *b = bridge_of(this);
	if (b == 0) return 1;
	if (sync_in(b, this, bodyData) == 1) return 1;
	unsigned flags = (evaluateGasDrag ? SOL_EVAL_GAS_DRAG : 0u) | (evaluateTypeIMigration ? SOL_EVAL_MIG_TYPE1 : 0u) |
	                 (evaluateTypeIIMigration ? SOL_EVAL_MIG_TYPE2 : 0u);
	if (sol_compute(b->ctx, t, y, totalAccel, flags) != SOL_OK) {
		Error::_errMsg = sol_last_error(b->ctx);
		Error::PushLocation(__FILE__, __FUNCTION__, __LINE__);
		return 1;
	}
	if (sync_out(b, this, bodyData, 0) == 1) return 1;
	if (nebula != 0) {
		// the public caches (Acceleration.h:46-48), lazily allocated like the reference does
		int npl = bodyData->nBodies.NOfPlAndSpl();
		int nm1 = bodyData->nBodies.rockyPlanet + bodyData->nBodies.protoPlanet;
		int nm2 = bodyData->nBodies.giantPlanet;
		if (accelGasDrag == 0 && npl > 0) accelGasDrag = new double[3 * npl];
		if (accelMigrationTypeI == 0 && bodyData->nBodies.protoPlanet > 0) accelMigrationTypeI = new double[3 * nm1];
		if (accelMigrationTypeII == 0 && nm2 > 0) accelMigrationTypeII = new double[3 * nm2];
		if (accelGasDrag != 0) sol_download(b->ctx, SOL_ACCEL_GASDRAG, accelGasDrag);
		if (accelMigrationTypeI != 0) sol_download(b->ctx, SOL_ACCEL_MIGTYPE1, accelMigrationTypeI);
		if (accelMigrationTypeII != 0) sol_download(b->ctx, SOL_ACCEL_MIGTYPE2, accelMigrationTypeII);
	}
	return 0;
}

int Acceleration::ComputeAstroCentric(double t, double *y, double *totalAccel) { return Compute(t, y, totalAccel); }
int Acceleration::ComputeBaryCentric(double t, double *y, double *totalAccel) { return Compute(t, y, totalAccel); }

static int fused_away(const char *what)
{
	Error::_errMsg = std::string("solaris_b200: Acceleration::") + what + " is fused into the device force kernel; call Compute()";
	Error::PushLocation(__FILE__, __FUNCTION__, __LINE__);
	return 1;
}

int Acceleration::GravityAC(double, double *, double *) { return fused_away("GravityAC"); }
int Acceleration::GasDragAC(double, double *, double *) { return fused_away("GasDragAC"); }
int Acceleration::MigrationTypeIAC(double, double *, double *) { return fused_away("MigrationTypeIAC"); }
int Acceleration::MigrationTypeIIAC(double, double *, double *) { return fused_away("MigrationTypeIIAC"); }
int Acceleration::GravityBC(double, double *, double *) { return fused_away("GravityBC"); }
int Acceleration::GravityBC_SelfInteracting(double, double *, double *) { return fused_away("GravityBC_SelfInteracting"); }
int Acceleration::GravityBC_NonSelfInteracting(double, double *, double *) { return fused_away("GravityBC_NonSelfInteracting"); }
int Acceleration::GasDragBC(double, double *, double *) { return fused_away("GasDragBC"); }
int Acceleration::MigrationTypeIBC(double, double *, double *) { return fused_away("MigrationTypeIBC"); }
int Acceleration::MigrationTypeIIBC(double, double *, double *) { return fused_away("MigrationTypeIIBC"); }
double Acceleration::TypeIMigrationTime(const double, const double, const double, const double, const double) { fused_away("TypeIMigrationTime"); return 0.0; }
double Acceleration::TypeIEccentricityDampingTime(const double, const double, const double, const double, const double) { fused_away("TypeIEccentricityDampingTime"); return 0.0; }
double Acceleration::TauNu(double, double) { fused_away("TauNu"); return 0.0; }

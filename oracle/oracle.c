/* ORACLE - TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-file CPU restatement of the hot path of suliaron/solaris:
 * force evaluation (Acceleration::Compute), the RK4 / RKF7(8) / Dormand-Prince RKN7(6)
 * drivers and the per-step event detection.  Every function cites the reference file:line it
 * restates.  Arithmetic is written in the reference's operation order so that, compiled for
 * baseline x86-64 without FMA contraction (oracle/Makefile), it reproduces the compiled reference
 * (oracle/_ref, built by oracle/build_ref.sh) BIT FOR BIT; tests/test_oracle_vs_reference.py and the
 * committed vectors under tests/golden/ pin that.  => parity is PINNED against the compiled
 * reference, not against reference unit tests (the reference has none for this path).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library, and only as the checker.  Nothing under solaris_b200/ links or dlopens it.
 *
 * Differences from the reference that are deliberate and documented:
 *  - function-local statics (GasComponent::Temperature_CMU `pT`, MeanThermalSpeed_CMU `Cvth`,
 *    Solaris/GasComponent.cpp:223-224,239) are evaluated per oracle_sys from that system's nebula;
 *    identical whenever a process uses one nebula, which is the only defined use of the reference.
 *  - ComputeBaryCentric's "gas gone" block (Solaris/Acceleration.cpp:152-160) deletes the nebula and
 *    leaves a dangling pointer; here the nebula is simply switched off.
 *  - the abs() -> fabs() semantics of the MSVC build are used (oracle/absfix.h).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>

#define SQR(a)   ((a)*(a))
#define CUBE(a)  ((a)*(a)*(a))
#define FORTH(a) ((a)*(a)*(a)*(a))
#define FIFTH(a) ((a)*(a)*(a)*(a)*(a))

/* ---- Solaris/Constants.h:16-90, same expressions so the doubles are identical ---- */
static const double K_PI            = 3.14159265358979323846;
static const double K_SQRT_TWO_PI   = 2.50662827463100024161;
static const double K_BOLTZMAN_SI   = 1.3806488e-23;
static const double K_PROTONMASS_SI = 1.672621777e-27;
static const double K_GAUSS         = 1.720209895e-2;
static const double K_GAUSS2        = 2.959122082855911025e-4;
static const double K_SOLAR_TO_KG   = 1.98911e30;
static const double K_AU_TO_METER   = 1.495978707e11;
static const double K_DAY_TO_SECOND = 86400.0;

static double k_boltzman_cmu(void)
{   /* Constants.h:87 */
	double KilogramToSolar = 1.0 / K_SOLAR_TO_KG;
	double MeterToAu = 1.0 / K_AU_TO_METER;
	double SecondToDay = 1.0 / K_DAY_TO_SECOND;
	return K_BOLTZMAN_SI * (KilogramToSolar * SQR(MeterToAu)) / (SQR(SecondToDay));
}
static double k_protonmass_cmu(void)
{   /* Constants.h:88 */
	double KilogramToSolar = 1.0 / K_SOLAR_TO_KG;
	return K_PROTONMASS_SI * KilogramToSolar;
}

/* Body.h:14-24 */
enum { T_UNDEF = 0, T_CENTRAL = 1, T_GIANT = 2, T_ROCKY = 3, T_PROTO = 4, T_SUPERPL = 5, T_PL = 6, T_TEST = 7 };
/* Body.h:35-39 */
enum { MIG_NO = 0, MIG_I = 1, MIG_II = 2 };
/* IntegratorType.h:12-17 */
enum { INT_DP = 0, INT_RK4 = 1, INT_RKF78 = 3 };

/* Same layout as include/solaris_b200.h:sol_nebula_pod and oracle/ref_harness.cpp:ref_nebula_pod */
typedef struct {
	double alpha;
	double mean_molecular_weight;
	double particle_diameter;
	int    decrease_type;
	int    _pad;
	double time_scale, t0, t1;
	double inner_edge;
	double eta_c, eta_index;
	double tau_c, tau_index;
	double scale_height_c, scale_height_index;
	double density_c, density_index;
	double mean_free_path_c, mean_free_path_index;
} oracle_nebula_pod;

typedef struct {
	int     counts[7];
	int     n;
	int     n_alloc;
	double *mass, *radius, *density, *cD, *gammaStokes, *gammaEpstein, *migStopAt;
	int    *type, *migType, *id;
	int    *indexOfNN;
	double *distanceOfNN;
	double *rm3;            /* Acceleration::rm3, zero-initialised (Acceleration.cpp:65-69) */
	double *accelGasDrag;   /* 3*NOfPlAndSpl */
	double *accelMigI;      /* 3*(rocky+proto) (the reference under-allocates) */
	double *accelMigII;     /* 3*giant */
	int     barycentric;
	int     has_nebula;
	oracle_nebula_pod neb;
	int     evalDrag, evalMigI, evalMigII;
	/* BodyData arrays used by the drivers */
	double *y0, *y, *accel, *error, *yscale;
	double  time, h;
} oracle_sys;

static int n_massive(const oracle_sys *s) { return s->counts[0] + s->counts[1] + s->counts[2] + s->counts[3]; } /* NBodies.cpp:21 */
static int n_pl_spl(const oracle_sys *s)  { return s->counts[4] + s->counts[5]; }                               /* NBodies.cpp:26 */

static double powerlaw(double c, double index, double x) { return c * pow(x, index); } /* PowerLaw.cpp:17-20 */

oracle_sys *oracle_create(const int counts[7], const double *y0, const double *mass, const double *radius,
                          const double *density, const double *cD, const double *gammaStokes,
                          const double *gammaEpstein, const double *migStopAt, const int *type,
                          const int *migType, const int *id, int barycentric, const oracle_nebula_pod *neb)
{
	oracle_sys *s = (oracle_sys *)calloc(1, sizeof(oracle_sys));
	int n = 0;
	for (int k = 0; k < 7; k++) { s->counts[k] = counts[k]; n += counts[k]; }
	s->n = s->n_alloc = n;
#define DUPD(dst, src, cnt) do { s->dst = (double *)calloc((cnt) > 0 ? (cnt) : 1, sizeof(double)); if (src) memcpy(s->dst, src, (cnt) * sizeof(double)); } while (0)
#define DUPI(dst, src, cnt) do { s->dst = (int *)calloc((cnt) > 0 ? (cnt) : 1, sizeof(int)); if (src) memcpy(s->dst, src, (cnt) * sizeof(int)); } while (0)
	DUPD(mass, mass, n); DUPD(radius, radius, n); DUPD(density, density, n); DUPD(cD, cD, n);
	DUPD(gammaStokes, gammaStokes, n); DUPD(gammaEpstein, gammaEpstein, n); DUPD(migStopAt, migStopAt, n);
	DUPI(type, type, n); DUPI(migType, migType, n); DUPI(id, id, n);
	DUPI(indexOfNN, (const int *)0, n);
	DUPD(distanceOfNN, (const double *)0, n);
	DUPD(rm3, (const double *)0, n);
	DUPD(accelGasDrag, (const double *)0, 3 * n_pl_spl(s));
	DUPD(accelMigI, (const double *)0, 3 * (s->counts[2] + s->counts[3]));
	DUPD(accelMigII, (const double *)0, 3 * s->counts[1]);
	DUPD(y0, y0, 6 * n); DUPD(y, (const double *)0, 6 * n); DUPD(accel, (const double *)0, 6 * n);
	DUPD(error, (const double *)0, 6 * n); DUPD(yscale, (const double *)0, 6 * n);
	for (int i = 0; i < n; i++) s->indexOfNN[i] = -1;
	s->barycentric = barycentric;
	s->has_nebula = neb != 0;
	if (neb) s->neb = *neb;
	s->evalDrag = s->evalMigI = s->evalMigII = 1;
	return s;
}

void oracle_destroy(oracle_sys *s)
{
	if (!s) return;
	free(s->mass); free(s->radius); free(s->density); free(s->cD); free(s->gammaStokes); free(s->gammaEpstein);
	free(s->migStopAt); free(s->type); free(s->migType); free(s->id); free(s->indexOfNN); free(s->distanceOfNN);
	free(s->rm3); free(s->accelGasDrag); free(s->accelMigI); free(s->accelMigII);
	free(s->y0); free(s->y); free(s->accel); free(s->error); free(s->yscale);
	free(s);
}

int oracle_n(const oracle_sys *s) { return s->n; }

/* ------------------------------------------------------------------------------------------
 * Gravity, astrocentric.  Solaris/Acceleration.cpp:248-329.
 * oracle_gravity_ac_rows evaluates sinks [ib, ie) only (row-subset form of the second loop,
 * :268-326) and requires rm3[] of all sources to be current (first loop, :256-264).
 * ------------------------------------------------------------------------------------------ */
static void gravity_ac_rm3_pass(oracle_sys *s, const double *y)
{
	s->indexOfNN[0] = -1;             /* :253-254 */
	s->distanceOfNN[0] = 0.0;
	for (int i = 1; i < s->n; i++) {  /* :256-264 */
		int i0 = 6 * i;
		double r2 = SQR(y[i0 + 0]) + SQR(y[i0 + 1]) + SQR(y[i0 + 2]);
		double r = sqrt(r2);
		s->rm3[i] = 1.0 / (r2 * r);
		s->indexOfNN[i] = -1;
		s->distanceOfNN[i] = 0.0;
	}
}

static void gravity_ac_row(oracle_sys *s, const double *y, double *accel, int i)
{
	double rMin = 1.0e10;             /* :269 */
	double ax = 0.0, ay = 0.0, az = 0.0;
	double mu = K_GAUSS2 * (s->mass[0] + s->mass[i]);   /* :272 */
	int i0 = 6 * i;
	accel[i0 + 0] = y[i0 + 3];
	accel[i0 + 1] = y[i0 + 4];
	accel[i0 + 2] = y[i0 + 5];
	accel[i0 + 3] = -mu * s->rm3[i] * y[i0 + 0];        /* :281-283, Kepler term stored first */
	accel[i0 + 4] = -mu * s->rm3[i] * y[i0 + 1];
	accel[i0 + 5] = -mu * s->rm3[i] * y[i0 + 2];
	int nsrc = n_massive(s);
	if (s->type[i] <= T_PROTO) nsrc = n_massive(s) + s->counts[4];   /* :285-289 */
	for (int j = 1; j < nsrc; j++) {                    /* :294-317 */
		if (j == i) continue;
		int j0 = 6 * j;
		double xij = y[j0 + 0] - y[i0 + 0];
		double yij = y[j0 + 1] - y[i0 + 1];
		double zij = y[j0 + 2] - y[i0 + 2];
		double rij2 = SQR(xij) + SQR(yij) + SQR(zij);
		double rij = sqrt(rij2);
		double rijm3 = 1.0 / (rij2 * rij);
		if (rij < rMin) {
			rMin = rij;
			s->indexOfNN[i] = j;
			s->distanceOfNN[i] = rij;
		}
		double Gmj = K_GAUSS2 * s->mass[j];
		ax += Gmj * (xij * rijm3 - y[j0 + 0] * s->rm3[j]);
		ay += Gmj * (yij * rijm3 - y[j0 + 1] * s->rm3[j]);
		az += Gmj * (zij * rijm3 - y[j0 + 2] * s->rm3[j]);
	}
	accel[i0 + 3] += ax;              /* :323-325 */
	accel[i0 + 4] += ay;
	accel[i0 + 5] += az;
}

static void gravity_ac(oracle_sys *s, const double *y, double *accel)
{
	gravity_ac_rm3_pass(s, y);
	accel[0] = accel[1] = accel[2] = accel[3] = accel[4] = accel[5] = 0.0;   /* :266 */
	for (int i = 1; i < s->n; i++) gravity_ac_row(s, y, accel, i);
}

/* ------------------------------------------------------------------------------------------
 * Gravity, barycentric.  Solaris/Acceleration.cpp:541-587 (massive sinks) and :589-634 (the
 * rest) have the same loop body; one row function serves both.
 * ------------------------------------------------------------------------------------------ */
static void gravity_bc_row(oracle_sys *s, const double *y, double *accel, int i)
{
	int nMassive = n_massive(s);
	double rMin = 1.0e10;
	int i0 = 6 * i;
	s->indexOfNN[i] = -1;
	s->distanceOfNN[i] = 0.0;
	accel[i0 + 0] = y[i0 + 3];
	accel[i0 + 1] = y[i0 + 4];
	accel[i0 + 2] = y[i0 + 5];
	/* :554 zeroes the acceleration part for massive sinks; :589-634 does NOT zero it for the
	 * non-massive sinks and accumulates into whatever the caller's array held.  Every caller in
	 * the reference passes freshly new[]-ed (uninitialised) or reused k-arrays, so the defined
	 * behaviour is "caller supplies zeros"; the oracle zeroes explicitly. */
	accel[i0 + 3] = accel[i0 + 4] = accel[i0 + 5] = 0.0;
	for (int j = nMassive - 1; j >= 0; j--) {   /* lightest first, :558 / :604 */
		if (j == i) continue;
		int j0 = 6 * j;
		double dxij = y[j0 + 0] - y[i0 + 0];
		double dyij = y[j0 + 1] - y[i0 + 1];
		double dzij = y[j0 + 2] - y[i0 + 2];
		double rij2 = SQR(dxij) + SQR(dyij) + SQR(dzij);
		double rij = sqrt(rij2);
		if (rij < rMin) {
			rMin = rij;
			s->indexOfNN[i] = j;
			s->distanceOfNN[i] = rij;
		}
		double c = s->mass[j] * 1.0 / (rij2 * rij);   /* :576 / :622 */
		accel[i0 + 3] += c * dxij;
		accel[i0 + 4] += c * dyij;
		accel[i0 + 5] += c * dzij;
	}
	accel[i0 + 3] *= K_GAUSS2;
	accel[i0 + 4] *= K_GAUSS2;
	accel[i0 + 5] *= K_GAUSS2;
}

static void gravity_bc(oracle_sys *s, const double *y, double *accel)
{
	for (int i = 0; i < s->n; i++) gravity_bc_row(s, y, accel, i);
}

/* Row-subset gravity for large N (SURVEY.md §8c "Large-N oracle"): sinks [ib, ie) against the full
 * source set, optionally over `threads` POSIX threads (interleaved blocks of 16 sink rows; rows are
 * independent).  Writes dydt rows [ib, ie) (6 doubles each, indexed by absolute i) and the NN side
 * outputs. */
typedef struct { oracle_sys *s; const double *y; double *dydt; int ib, ie, tid, nthreads; } rows_job;

static void *rows_worker(void *arg)
{
	rows_job *jb = (rows_job *)arg;
	oracle_sys *s = jb->s;
	const int blk = 16;
	for (int b = jb->ib + jb->tid * blk; b < jb->ie; b += jb->nthreads * blk) {
		int e = b + blk < jb->ie ? b + blk : jb->ie;
		for (int i = b; i < e; i++) {
			if (s->barycentric) {
				gravity_bc_row(s, jb->y, jb->dydt, i);
			} else if (i == 0) {
				jb->dydt[0] = jb->dydt[1] = jb->dydt[2] = jb->dydt[3] = jb->dydt[4] = jb->dydt[5] = 0.0;
			} else {
				gravity_ac_row(s, jb->y, jb->dydt, i);
			}
		}
	}
	return 0;
}

int oracle_gravity_rows(oracle_sys *s, const double *y, double *dydt, int ib, int ie, int threads)
{
	if (ib < 0 || ie > s->n || ib > ie) return 1;
	if (!s->barycentric) gravity_ac_rm3_pass(s, y);   /* rm3 of every body (first pass, cheap, serial) */
	if (threads < 1) threads = 1;
	if (threads > 256) threads = 256;
	rows_job jobs[256];
	pthread_t th[256];
	for (int t = 0; t < threads; t++) {
		jobs[t].s = s; jobs[t].y = y; jobs[t].dydt = dydt; jobs[t].ib = ib; jobs[t].ie = ie;
		jobs[t].tid = t; jobs[t].nthreads = threads;
	}
	for (int t = 1; t < threads; t++) pthread_create(&th[t], 0, rows_worker, &jobs[t]);
	rows_worker(&jobs[0]);
	for (int t = 1; t < threads; t++) pthread_join(th[t], 0);
	return 0;
}

/* ------------------------------------------------------------------------------------------
 * EXTENDED-PRECISION rows: the mathematically exact value of the sum the reference evaluates in double.
 *
 * At N ~ 10^6 the reference's own sequential double sum (Acceleration.cpp:294-317) carries ~1e-12 of
 * rounding noise relative to |a_i| on the few bodies whose Kepler term is nearly cancelled by the disk,
 * so "within 1e-13 of the reference" cannot be decided against the reference's double result there.
 * These functions evaluate the SAME expression (A.1 / A.2 of SURVEY.md: same source sets, same j != i
 * exclusion, no softening) from the same double inputs, but
 *   - oracle_gravity_rows_exact: in x87 long double (64-bit significand) with Neumaier-compensated
 *     accumulation: every term is good to ~1e-19 relative, the sum to ~1e-19 of sum |terms|;
 *   - oracle_gravity_row_quad: in IEEE binary128 (libquadmath, 113-bit significand), plain summation:
 *     ~1e-34 per term; used by the CPU tests to validate the long-double version, which is ~100x faster.
 * Results are rounded to double once at the end.  out3 receives 3 doubles per listed row (the
 * acceleration part only; the velocity part of dy/dt is a copy).
 * ------------------------------------------------------------------------------------------ */
typedef struct { long double s, c; } nsum;   /* Neumaier: running sum + compensation */
static inline void nsum_add(nsum *a, long double x)
{
	long double t = a->s + x;
	if (fabsl(a->s) >= fabsl(x)) a->c += (a->s - t) + x;
	else                         a->c += (x - t) + a->s;
	a->s = t;
}

static void gravity_row_exact(const oracle_sys *s, const double *y, int i, double *out3)
{
	const int M = n_massive(s);
	const long double k2 = (long double)K_GAUSS2;
	const int i0 = 6 * i;
	const long double xi = y[i0 + 0], yi = y[i0 + 1], zi = y[i0 + 2];
	nsum ax = {0.0L, 0.0L}, ay = {0.0L, 0.0L}, az = {0.0L, 0.0L};
	if (s->barycentric) {
		for (int j = 0; j < M; j++) {
			if (j == i) continue;
			const int j0 = 6 * j;
			const long double dx = (long double)y[j0 + 0] - xi, dy = (long double)y[j0 + 1] - yi, dz = (long double)y[j0 + 2] - zi;
			const long double r2 = dx * dx + dy * dy + dz * dz;
			const long double w = k2 * (long double)s->mass[j] / (r2 * sqrtl(r2));
			nsum_add(&ax, w * dx); nsum_add(&ay, w * dy); nsum_add(&az, w * dz);
		}
	} else {
		if (i == 0) { out3[0] = out3[1] = out3[2] = 0.0; return; }
		int nsrc = M;
		if (s->type[i] <= T_PROTO) nsrc = M + s->counts[4];
		const long double ri2 = xi * xi + yi * yi + zi * zi;
		const long double kep = -k2 * ((long double)s->mass[0] + (long double)s->mass[i]) / (ri2 * sqrtl(ri2));
		nsum_add(&ax, kep * xi); nsum_add(&ay, kep * yi); nsum_add(&az, kep * zi);
		for (int j = 1; j < nsrc; j++) {
			if (j == i) continue;
			const int j0 = 6 * j;
			const long double xj = y[j0 + 0], yj = y[j0 + 1], zj = y[j0 + 2];
			const long double dx = xj - xi, dy = yj - yi, dz = zj - zi;
			const long double r2 = dx * dx + dy * dy + dz * dz;
			const long double rj2 = xj * xj + yj * yj + zj * zj;
			const long double gm = k2 * (long double)s->mass[j];
			const long double w = gm / (r2 * sqrtl(r2)), wj = gm / (rj2 * sqrtl(rj2));
			nsum_add(&ax, w * dx); nsum_add(&ay, w * dy); nsum_add(&az, w * dz);
			nsum_add(&ax, -wj * xj); nsum_add(&ay, -wj * yj); nsum_add(&az, -wj * zj);
		}
	}
	out3[0] = (double)(ax.s + ax.c); out3[1] = (double)(ay.s + ay.c); out3[2] = (double)(az.s + az.c);
}

typedef struct { const oracle_sys *s; const double *y; const int *rows; int nrows; double *out; int tid, nthreads; } exact_job;
static void *exact_worker(void *arg)
{
	exact_job *jb = (exact_job *)arg;
	for (int k = jb->tid; k < jb->nrows; k += jb->nthreads) gravity_row_exact(jb->s, jb->y, jb->rows[k], jb->out + 3 * (size_t)k);
	return 0;
}

int oracle_gravity_rows_exact(const oracle_sys *s, const double *y, const int *rows, int nrows, double *out3, int threads)
{
	for (int k = 0; k < nrows; k++) if (rows[k] < 0 || rows[k] >= s->n) return 1;
	if (threads < 1) threads = 1;
	if (threads > 256) threads = 256;
	exact_job jobs[256];
	pthread_t th[256];
	for (int t = 0; t < threads; t++) {
		jobs[t].s = s; jobs[t].y = y; jobs[t].rows = rows; jobs[t].nrows = nrows; jobs[t].out = out3;
		jobs[t].tid = t; jobs[t].nthreads = threads;
	}
	for (int t = 1; t < threads; t++) pthread_create(&th[t], 0, exact_worker, &jobs[t]);
	exact_worker(&jobs[0]);
	for (int t = 1; t < threads; t++) pthread_join(th[t], 0);
	return 0;
}

#include <quadmath.h>
int oracle_gravity_row_quad(const oracle_sys *s, const double *y, int i, double *out3)
{
	if (i < 0 || i >= s->n) return 1;
	const int M = n_massive(s);
	const __float128 k2 = (__float128)K_GAUSS2;
	const int i0 = 6 * i;
	const __float128 xi = y[i0 + 0], yi = y[i0 + 1], zi = y[i0 + 2];
	__float128 ax = 0, ay = 0, az = 0;
	if (s->barycentric) {
		for (int j = 0; j < M; j++) {
			if (j == i) continue;
			const int j0 = 6 * j;
			const __float128 dx = (__float128)y[j0 + 0] - xi, dy = (__float128)y[j0 + 1] - yi, dz = (__float128)y[j0 + 2] - zi;
			const __float128 r2 = dx * dx + dy * dy + dz * dz;
			const __float128 w = k2 * (__float128)s->mass[j] / (r2 * sqrtq(r2));
			ax += w * dx; ay += w * dy; az += w * dz;
		}
	} else {
		if (i == 0) { out3[0] = out3[1] = out3[2] = 0.0; return 0; }
		int nsrc = M;
		if (s->type[i] <= T_PROTO) nsrc = M + s->counts[4];
		const __float128 ri2 = xi * xi + yi * yi + zi * zi;
		const __float128 kep = -k2 * ((__float128)s->mass[0] + (__float128)s->mass[i]) / (ri2 * sqrtq(ri2));
		ax = kep * xi; ay = kep * yi; az = kep * zi;
		for (int j = 1; j < nsrc; j++) {
			if (j == i) continue;
			const int j0 = 6 * j;
			const __float128 xj = y[j0 + 0], yj = y[j0 + 1], zj = y[j0 + 2];
			const __float128 dx = xj - xi, dy = yj - yi, dz = zj - zi;
			const __float128 r2 = dx * dx + dy * dy + dz * dz;
			const __float128 rj2 = xj * xj + yj * yj + zj * zj;
			const __float128 gm = k2 * (__float128)s->mass[j];
			const __float128 w = gm / (r2 * sqrtq(r2)), wj = gm / (rj2 * sqrtq(rj2));
			ax += w * dx - wj * xj; ay += w * dy - wj * yj; az += w * dz - wj * zj;
		}
	}
	out3[0] = (double)ax; out3[1] = (double)ay; out3[2] = (double)az;
	return 0;
}

/* ---- Gas model.  Solaris/GasComponent.cpp ---- */
static double reduction_factor(const oracle_nebula_pod *g, double t)
{   /* GasComponent.cpp:36-61 */
	switch (g->decrease_type) {
	case 0: return 1.0;
	case 1:
		if (t <= g->t0) return 1.0;
		else if (t > g->t0 && t <= g->t1) return 1.0 - (t - g->t0) / (g->t1 - g->t0);
		else return 0.0;
	case 2: return exp(-t / g->time_scale);
	default: return 1.0;
	}
}

static double midplane_density(const oracle_nebula_pod *g, double r)
{   /* GasComponent.cpp:63-69 */
	double a1 = powerlaw(g->density_c, g->density_index, r);
	double a2 = powerlaw(g->scale_height_c, g->scale_height_index, r);
	double a3 = a1 * a2 * K_SQRT_TWO_PI;
	return a3;
}

static void circular_velocity(double mu, double x, double yy, double *vx, double *vy)
{   /* GasComponent.cpp:97-126 */
	*vx = 0.0; *vy = 0.0;
	double r = sqrt(SQR(x) + SQR(yy));
	double vc = sqrt(mu / r);
	double p;
	if (x == 0.0 && yy == 0.0) {
		return;
	} else if (yy == 0.0) {
		*vy = x > 0.0 ? vc : -vc;
	} else if (x == 0.0) {
		*vx = yy > 0.0 ? -vc : vc;
	} else if (x >= yy) {
		p = yy / x;
		*vy = x >= 0 ? vc / sqrt(1.0 + SQR(p)) : -vc / sqrt(1.0 + SQR(p));
		*vx = -(*vy) * p;
	} else {
		p = x / yy;
		*vx = yy >= 0 ? -vc / sqrt(1.0 + SQR(p)) : vc / sqrt(1.0 + SQR(p));
		*vy = -(*vx) * p;
	}
}

static void gas_velocity(const oracle_nebula_pod *g, double mu, double x, double yy, double *vx, double *vy)
{   /* GasComponent.cpp:128-138 */
	circular_velocity(mu, x, yy, vx, vy);
	double r = sqrt(SQR(x) + SQR(yy));
	double v = sqrt(1.0 - 2.0 * powerlaw(g->eta_c, g->eta_index, r));
	*vx *= v;
	*vy *= v;
}

static double gas_density_at(const oracle_nebula_pod *g, double x, double yy, double z)
{   /* GasComponent.cpp:141-157 */
	double result = 0.0;
	double r = sqrt(SQR(x) + SQR(yy));
	double h = powerlaw(g->scale_height_c, g->scale_height_index, r);
	double arg = SQR(z / h);
	if (g->inner_edge < r) {
		result = powerlaw(g->density_c, g->density_index, r) * exp(-arg);
	} else {
		double a = g->density_c * pow(g->inner_edge, g->density_index - 4.0);
		result = a * SQR(SQR(r)) * exp(-arg);
	}
	return result;
}

static double temperature_cmu(const oracle_nebula_pod *g, double mC, double r)
{   /* GasComponent.cpp:221-229 */
	double ProtonMassBoltzman_CMU = 1.0 / (k_boltzman_cmu() / k_protonmass_cmu());   /* Constants.h:89-90 */
	double cTp = K_GAUSS2 * ProtonMassBoltzman_CMU;
	double pT = 2.0 * g->scale_height_index - 3.0;
	double cT = SQR(g->scale_height_c) * mC * g->mean_molecular_weight * cTp;
	double result = cT * pow(r, pT);
	return result;
}

static double mean_thermal_speed_cmu(const oracle_nebula_pod *g, double mC, double r)
{   /* GasComponent.cpp:238-243; NOTE the swapped arguments */
	double Cvth = sqrt((8.0 * k_boltzman_cmu()) / (K_PI * g->mean_molecular_weight * k_protonmass_cmu()));
	double result = Cvth * sqrt(temperature_cmu(g, r, mC));
	return result;
}

/* ---- Gas drag.  Solaris/Acceleration.cpp:331-422 ---- */
static void gas_drag(oracle_sys *s, double t, const double *y, double *accel)
{
	const oracle_nebula_pod *g = &s->neb;
	double factor = reduction_factor(g, t);
	int lower = n_massive(s);
	int upper = lower + s->counts[5] + s->counts[4];
	for (int i = lower; i < upper; i++) {
		int i0 = 6 * i;
		int j0 = 3 * (i - lower);
		double r = sqrt(SQR(y[i0 + 0]) + SQR(y[i0 + 1]) + SQR(y[i0 + 2]));
		double C = 0.0;
		double vgx, vgy;
		gas_velocity(g, K_GAUSS2 * s->mass[0], y[i0 + 0], y[i0 + 1], &vgx, &vgy);
		double ux = y[i0 + 3] - vgx, uy = y[i0 + 4] - vgy, uz = y[i0 + 5] - 0.0;
		double rhoGas = factor * gas_density_at(g, y[i0 + 0], y[i0 + 1], y[i0 + 2]);
		double lambda = powerlaw(g->mean_free_path_c, g->mean_free_path_index, r);
		if (s->radius[i] <= 0.1 * lambda) {            /* Epstein, :361-369 */
			double vth = mean_thermal_speed_cmu(g, s->mass[0], r);
			C = s->gammaEpstein[i] * vth * rhoGas;
		} else if (s->radius[i] >= 10.0 * lambda) {    /* Stokes, :371-379 */
			double uLength = sqrt(ux * ux + uy * uy + uz * uz);
			C = s->gammaStokes[i] * uLength * rhoGas;
		} else {                                       /* transition, :381-397 */
			double lambda1 = 0.1 * lambda;
			double lambda2 = 10.0 * lambda;
			double gammaE = 1.0 / (s->density[i] * lambda1);
			double gammaS = 3.0 / 8.0 * s->cD[i] / (s->density[i] * lambda2);
			double vth = mean_thermal_speed_cmu(g, s->mass[0], r);
			double K = gammaS * sqrt(ux * ux + uy * uy + uz * uz) / (gammaE * vth);
			double eta = lambda2 / lambda1;
			double kappa = log10(K) / log10(eta);
			double gamma = gammaE * vth * pow(lambda1, -kappa);
			C = gamma * pow(s->radius[i], kappa) * rhoGas;
		}
		accel[j0 + 0] = -C * ux;
		accel[j0 + 1] = -C * uy;
		accel[j0 + 2] = -C * uz;
	}
}

/* Ephemeris::CalculateOrbitalElement(mu, phase, &a, &e), Solaris/Ephemeris.cpp:10-41 (+213-226) */
static int orbital_element_ae(double mu, const double *rv, double *a, double *e)
{
	const double sq3 = 1.0e-14;
	double kin = (rv[3] * rv[3] + rv[4] * rv[4] + rv[5] * rv[5]) / 2.0;
	double pot = -mu / sqrt(rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2]);
	double h = kin + pot;
	if (h >= 0.0) return 1;
	double cx = rv[1] * rv[5] - rv[2] * rv[4];
	double cy = rv[2] * rv[3] - rv[0] * rv[5];
	double cz = rv[0] * rv[4] - rv[1] * rv[3];
	double e2 = 1.0 + 2.0 * (cx * cx + cy * cy + cz * cz) * h / (mu * mu);
	if (fabs(e2) < sq3) e2 = 0.0;
	*e = sqrt(e2);
	*a = -mu / (2.0 * h);
	return 0;
}

/* Acceleration::TypeIMigrationTime / TypeIEccentricityDampingTime, Acceleration.cpp:766-789 */
static double type1_migration_time(const oracle_nebula_pod *g, double C, double O, double ar, double er, double h)
{
	double Cm = 2.0 / (2.7 + 1.1 * fabs(g->density_index)) / O;
	double er1 = er / (1.3 * h);
	double er2 = er / (1.1 * h);
	double frac = (1.0 + FIFTH(er1)) / (1.0 - FORTH(er2));
	return Cm * C * SQR(ar) * frac;
}
static double type1_ecc_damping_time(double C, double O, double ar, double er, double h)
{
	const double Q = 0.78;
	double Ce = 0.1 / (Q * O);
	double frac = 1.0 + 0.25 * CUBE(er / h);
	return Ce * C * FORTH(ar) * frac;
}

/* MigrationTypeIAC (:424-485) == MigrationTypeIBC (:645-713) because TransformToAC is empty (Q16) */
static void migration_type1(oracle_sys *s, double t, const double *y, double *accel)
{
	const oracle_nebula_pod *g = &s->neb;
	double factor = reduction_factor(g, t);
	int lower = s->counts[0] + s->counts[1];
	int upper = n_massive(s);
	for (int i = lower; i < upper; i++) {
		if (s->migType[i] != MIG_I) continue;
		int i0 = 6 * i;
		int j0 = 3 * (i - lower);
		double r2 = SQR(y[i0 + 0]) + SQR(y[i0 + 1]) + SQR(y[i0 + 2]);
		double r = sqrt(r2);
		if (r <= s->migStopAt[i]) {
			accel[j0 + 0] = accel[j0 + 1] = accel[j0 + 2] = 0.0;
			s->migType[i] = MIG_NO;
			continue;
		}
		double m = s->mass[i];
		double mc = s->mass[0];
		double a = 0.0, e = 0.0;
		double mu = K_GAUSS2 * (mc + m);
		orbital_element_ae(mu, &y[i0], &a, &e);
		double O = K_GAUSS * sqrt((mc + m) / CUBE(a));
		double C = SQR(mc) / (m * midplane_density(g, r) * a * a);
		double h = powerlaw(g->scale_height_c, g->scale_height_index, r);
		double ar = h / r;
		double er = e * r;
		double tm = 0.0;
		if (e < 1.1 * h / r) {
			tm = type1_migration_time(g, C, O, ar, er, h);
			tm = 1.0 / tm;
		}
		double te = type1_ecc_damping_time(C, O, ar, er, h);
		double ti = te;
		double vr = y[i0 + 0] * y[i0 + 3] + y[i0 + 1] * y[i0 + 4] + y[i0 + 2] * y[i0 + 5];
		te = 2.0 * vr / (r2 * te);
		ti = 2.0 / ti;
		accel[j0 + 0] = -factor * (tm * y[i0 + 3] + te * y[i0 + 0]);
		accel[j0 + 1] = -factor * (tm * y[i0 + 4] + te * y[i0 + 1]);
		accel[j0 + 2] = -factor * (tm * y[i0 + 5] + te * y[i0 + 2] + ti * y[i0 + 5]);
	}
}

/* Acceleration::TauNu, Acceleration.cpp:831-849 */
static double tau_nu(const oracle_nebula_pod *g, double r, double O)
{
	double index = g->tau_index;
	double c = g->tau_c;
	double h = powerlaw(g->scale_height_c, g->scale_height_index, r);
	if (index == 2) return c * SQR(r / h) / (g->alpha * O);
	return c * pow(r / h, index) / (g->alpha * O);
}

/* MigrationTypeIIAC (:487-528) / MigrationTypeIIBC (:716-764): AC uses 1/TauNu, BC TauNu (Q16) */
static void migration_type2(oracle_sys *s, double t, const double *y, double *accel)
{
	const oracle_nebula_pod *g = &s->neb;
	double factor = reduction_factor(g, t);
	int lower = s->counts[0];
	int upper = s->counts[0] + s->counts[1];
	for (int i = lower; i < upper; i++) {
		if (s->migType[i] != MIG_II) continue;
		int i0 = 6 * i;
		int j0 = 3 * (i - lower);
		double r2 = SQR(y[i0 + 0]) + SQR(y[i0 + 1]) + SQR(y[i0 + 2]);
		double r = sqrt(r2);
		if (r <= s->migStopAt[i]) {
			s->migType[i] = MIG_NO;
			accel[j0 + 0] = accel[j0 + 1] = accel[j0 + 2] = 0.0;
			continue;
		}
		double m = s->mass[i];
		double mc = s->mass[0];
		double a = 0.0, e = 0.0;
		double mu = K_GAUSS2 * (mc + m);
		orbital_element_ae(mu, &y[i0], &a, &e);
		double O = K_GAUSS * sqrt((mc + m) / CUBE(a));
		double c0 = s->barycentric ? tau_nu(g, r, O) : 1.0 / tau_nu(g, r, O);
		double vr = y[i0 + 3] * y[i0 + 0] + y[i0 + 4] * y[i0 + 1] + y[i0 + 5] * y[i0 + 2];
		double c1 = vr / r2;
		accel[j0 + 0] = -factor * (c0 * (0.5 * y[i0 + 3] + 50 * (c1 * y[i0 + 0])));
		accel[j0 + 1] = -factor * (c0 * (0.5 * y[i0 + 4] + 50 * (c1 * y[i0 + 1])));
		accel[j0 + 2] = -factor * (c0 * (0.5 * y[i0 + 5] + 50 * (c1 * y[i0 + 2]) + y[i0 + 5]));
	}
}

/* Gas terms shared by ComputeAstroCentric (:176-243) and ComputeBaryCentric (:88-161) */
static void add_gas_terms(oracle_sys *s, double t, const double *y, double *total)
{
	if (!s->has_nebula) return;
	if (s->evalDrag && n_pl_spl(s) > 0) gas_drag(s, t, y, s->accelGasDrag);
	int lower = n_massive(s);
	int upper = lower + n_pl_spl(s);
	for (int i = lower; i < upper; i++) {
		int i0 = 6 * i, j0 = 3 * (i - lower);
		total[i0 + 3] += s->accelGasDrag[j0 + 0];
		total[i0 + 4] += s->accelGasDrag[j0 + 1];
		total[i0 + 5] += s->accelGasDrag[j0 + 2];
	}
	/* The reference allocates (and therefore evaluates / adds) the type-I cache only when
	 * nBodies.protoPlanet > 0 (:108-111 / :200-204); with protoPlanet == 0 and a migrating rocky
	 * planet it dereferences a null pointer.  The oracle treats that case as "no type-I term". */
	if (s->counts[3] > 0) {
		if (s->evalMigI) migration_type1(s, t, y, s->accelMigI);
		lower = s->counts[0] + s->counts[1];
		upper = n_massive(s);
		for (int i = lower; i < upper; i++) {
			if (s->migType[i] == MIG_NO) continue;
			int i0 = 6 * i, j0 = 3 * (i - lower);
			total[i0 + 3] += s->accelMigI[j0 + 0];
			total[i0 + 4] += s->accelMigI[j0 + 1];
			total[i0 + 5] += s->accelMigI[j0 + 2];
		}
	}
	if (s->counts[1] > 0) {
		if (s->evalMigII) migration_type2(s, t, y, s->accelMigII);
		lower = s->counts[0];
		upper = s->counts[0] + s->counts[1];
		for (int i = lower; i < upper; i++) {
			if (s->migType[i] == MIG_NO) continue;
			int i0 = 6 * i, j0 = 3 * (i - lower);
			total[i0 + 3] += s->accelMigII[j0 + 0];
			total[i0 + 4] += s->accelMigII[j0 + 1];
			total[i0 + 5] += s->accelMigII[j0 + 2];
		}
	}
	if (s->barycentric && reduction_factor(&s->neb, t) < 1.0e-5) {   /* :152-160 */
		s->evalDrag = s->evalMigI = s->evalMigII = 0;
		s->has_nebula = 0;
	}
}

/* Acceleration::Compute, Solaris/Acceleration.cpp:60-81 */
static int compute(oracle_sys *s, double t, const double *y, double *total)
{
	if (s->barycentric) gravity_bc(s, y, total);
	else gravity_ac(s, y, total);
	add_gas_terms(s, t, y, total);
	return 0;
}

int oracle_compute(oracle_sys *s, double t, const double *y, double *dydt, unsigned eval_flags)
{
	s->evalDrag  = (eval_flags & 1u) != 0;
	s->evalMigI  = (eval_flags & 2u) != 0;
	s->evalMigII = (eval_flags & 4u) != 0;
	return compute(s, t, y, dydt);
}

void oracle_get_side(oracle_sys *s, double *rm3, int *indexOfNN, double *distanceOfNN, int *migType)
{
	if (rm3) memcpy(rm3, s->rm3, s->n * sizeof(double));
	if (indexOfNN) memcpy(indexOfNN, s->indexOfNN, s->n * sizeof(int));
	if (distanceOfNN) memcpy(distanceOfNN, s->distanceOfNN, s->n * sizeof(double));
	if (migType) memcpy(migType, s->migType, s->n * sizeof(int));
}

/* what: 0 y0, 1 y, 2 accel, 3 error, 4 yscale */
void oracle_get_array(oracle_sys *s, int what, double *out)
{
	const double *src = what == 0 ? s->y0 : what == 1 ? s->y : what == 2 ? s->accel : what == 3 ? s->error : s->yscale;
	memcpy(out, src, 6 * s->n * sizeof(double));
}

void oracle_set_y0(oracle_sys *s, const double *y0) { memcpy(s->y0, y0, 6 * s->n * sizeof(double)); }

/* ------------------------------------------------------------------------------------------
 * RungeKutta4.  Solaris/RungeKutta4.cpp:20-56 (Driver), :58-130 (Step)
 * ------------------------------------------------------------------------------------------ */
static void swap_ptr(double **a, double **b) { double *t = *a; *a = *b; *b = t; }

static int rk4_driver(oracle_sys *s, double *time, double *hNext, double *hDid)
{
	int nVar = 6 * s->n;
	s->time = *time;
	s->h = *hNext;
	s->evalDrag = s->evalMigI = s->evalMigII = 1;
	compute(s, *time, s->y0, s->accel);
	s->evalMigI = s->evalMigII = 0;

	double *fk[4];
	fk[0] = s->accel;
	for (int i = 1; i < 4; i++) fk[i] = (double *)calloc(nVar, sizeof(double));
	double *yTemp = (double *)calloc(nVar, sizeof(double));
	double h = s->h, t = s->time;
	const double a21 = 1.0 / 2.0, a32 = 1.0 / 2.0, a43 = 1.0;
	const double b1 = 1.0 / 6.0, b2 = 1.0 / 3.0, b3 = 1.0 / 3.0, b4 = 1.0 / 6.0;
	const double c2 = 1.0 / 2.0, c3 = 1.0 / 2.0, c4 = 1.0;
	for (int i = 0; i < nVar; i++) yTemp[i] = s->y0[i] + h * (a21 * fk[0][i]);
	compute(s, t + c2 * h, yTemp, fk[1]);
	for (int i = 0; i < nVar; i++) yTemp[i] = s->y0[i] + h * (a32 * fk[1][i]);
	compute(s, t + c3 * h, yTemp, fk[2]);
	for (int i = 0; i < nVar; i++) yTemp[i] = s->y0[i] + h * (a43 * fk[2][i]);
	compute(s, t + c4 * h, yTemp, fk[3]);
	for (int i = 0; i < nVar; i++)
		s->y[i] = s->y0[i] + h * (b1 * fk[0][i] + b2 * fk[1][i] + b3 * fk[2][i] + b4 * fk[3][i]);
	for (int i = 1; i < 4; i++) free(fk[i]);
	free(yTemp);

	*hDid = s->h;
	*time += *hDid;
	s->time = *time;
	*hNext = s->h;
	swap_ptr(&s->y0, &s->y);
	return 0;
}

/* ------------------------------------------------------------------------------------------
 * RungeKuttaFehlberg78.  Solaris/RungeKuttaFehlberg78.cpp:33-57 (tableau), :66-140 (Driver),
 * :147-250 (Step), :252-262 (GetErrorMax)
 * ------------------------------------------------------------------------------------------ */
static void rkf78_step(oracle_sys *s)
{
	const double D1_0 = 41.0 / 840.0, D1_5 = 34.0 / 105.0, D1_6 = 9.0 / 35.0, D1_8 = 9.0 / 280.0, D1_10 = 41.0 / 840.0;
	const double D_1_0 = 2.0 / 27.0, D_2_0 = 1.0 / 36.0, D_3_0 = 1.0 / 24.0, D_4_0 = 5.0 / 12.0;
	const double D_5_0 = 1.0 / 20.0, D_6_0 = -25.0 / 108.0, D_7_0 = 31.0 / 300.0, D_8_0 = 2.0;
	const double D_9_0 = -91.0 / 108.0, D_10_0 = 2383.0 / 4100.0, D_11_0 = 3.0 / 205.0, D_12_0 = -1777.0 / 4100.0;
	const double D_2_1 = 1.0 / 12.0;
	const double D_3_2 = 1.0 / 8.0, D_4_2 = -25.0 / 16.0;
	const double D_4_3 = 25.0 / 16.0, D_5_3 = 1.0 / 4.0, D_6_3 = 125.0 / 108.0, D_8_3 = -53.0 / 6.0, D_9_3 = 23.0 / 108.0, D_10_3 = -341.0 / 164.0, D_12_3 = -341.0 / 164.0;
	const double D_5_4 = 1.0 / 5.0, D_6_4 = -65.0 / 27.0, D_7_4 = 61.0 / 225.0, D_8_4 = 704.0 / 45.0, D_9_4 = -976.0 / 135.0, D_10_4 = 4496.0 / 1025.0, D_12_4 = 4496.0 / 1025.0;
	const double D_6_5 = 125.0 / 54.0, D_7_5 = -2.0 / 9.0, D_8_5 = -107.0 / 9.0, D_9_5 = 311.0 / 54.0, D_10_5 = -301.0 / 82.0, D_11_5 = -6.0 / 41.0, D_12_5 = -289.0 / 82.0;
	const double D_7_6 = 13.0 / 900.0, D_8_6 = 67.0 / 90.0, D_9_6 = -19.0 / 60.0, D_10_6 = 2133.0 / 4100.0, D_11_6 = -3.0 / 205.0, D_12_6 = 2193.0 / 4100.0;
	const double D_8_7 = 3.0, D_9_7 = 17.0 / 6.0, D_10_7 = 45.0 / 82.0, D_11_7 = -3.0 / 41.0, D_12_7 = 51.0 / 82.0;
	const double D_9_8 = -1.0 / 12.0, D_10_8 = 45.0 / 164.0, D_11_8 = 3.0 / 41.0, D_12_8 = 33.0 / 164.0;
	const double D_10_9 = 18.0 / 41.0, D_11_9 = 6.0 / 41.0, D_12_9 = 12.0 / 41.0;
	const double D_12_11 = 1.0;

	int nVar = 6 * s->n;
	double *fk[13];
	for (int i = 1; i < 13; i++) fk[i] = (double *)calloc(nVar, sizeof(double));
	double *yTemp = (double *)calloc(nVar, sizeof(double));
	fk[0] = s->accel;
	double h = s->h, t = s->time;
	const double *y0 = s->y0;
#define STAGE(k, expr) do { for (int i = 0; i < nVar; i++) yTemp[i] = y0[i] + h * (expr); compute(s, t, yTemp, fk[k]); } while (0)
	STAGE(1, D_1_0 * fk[0][i]);
	STAGE(2, D_2_0 * fk[0][i] + D_2_1 * fk[1][i]);
	STAGE(3, D_3_0 * fk[0][i] + D_3_2 * fk[2][i]);
	STAGE(4, D_4_0 * fk[0][i] + D_4_2 * fk[2][i] + D_4_3 * fk[3][i]);
	STAGE(5, D_5_0 * fk[0][i] + D_5_3 * fk[3][i] + D_5_4 * fk[4][i]);
	STAGE(6, D_6_0 * fk[0][i] + D_6_3 * fk[3][i] + D_6_4 * fk[4][i] + D_6_5 * fk[5][i]);
	STAGE(7, D_7_0 * fk[0][i] + D_7_4 * fk[4][i] + D_7_5 * fk[5][i] + D_7_6 * fk[6][i]);
	STAGE(8, D_8_0 * fk[0][i] + D_8_3 * fk[3][i] + D_8_4 * fk[4][i] + D_8_5 * fk[5][i] + D_8_6 * fk[6][i] + D_8_7 * fk[7][i]);
	STAGE(9, D_9_0 * fk[0][i] + D_9_3 * fk[3][i] + D_9_4 * fk[4][i] + D_9_5 * fk[5][i] + D_9_6 * fk[6][i] + D_9_7 * fk[7][i] + D_9_8 * fk[8][i]);
	STAGE(10, D_10_0 * fk[0][i] + D_10_3 * fk[3][i] + D_10_4 * fk[4][i] + D_10_5 * fk[5][i] + D_10_6 * fk[6][i] + D_10_7 * fk[7][i] + D_10_8 * fk[8][i] + D_10_9 * fk[9][i]);
	STAGE(11, D_11_0 * fk[0][i] + D_11_5 * fk[5][i] + D_11_6 * fk[6][i] + D_11_7 * fk[7][i] + D_11_8 * fk[8][i] + D_11_9 * fk[9][i]);
	STAGE(12, D_12_0 * fk[0][i] + D_12_3 * fk[3][i] + D_12_4 * fk[4][i] + D_12_5 * fk[5][i] + D_12_6 * fk[6][i] + D_12_7 * fk[7][i] + D_12_8 * fk[8][i] + D_12_9 * fk[9][i] + D_12_11 * fk[11][i]);
#undef STAGE
	for (int i = 0; i < nVar; i++)
		s->y[i] = y0[i] + h * (D1_0 * fk[0][i] + D1_5 * fk[5][i] + D1_6 * (fk[6][i] + fk[7][i]) + D1_8 * (fk[8][i] + fk[9][i]) + D1_10 * fk[10][i]);
	for (int i = 0; i < nVar; i++)
		s->error[i] = h * fabs(fk[0][i] + fk[10][i] - fk[11][i] - fk[12][i]) * 41.0 / 840.0;
	for (int i = 1; i < 13; i++) free(fk[i]);
	free(yTemp);
}

static double rkf78_error_max(int n, const double *yerr, const double *yscale, double epsilon)
{
	double errorMax = 0.0;
	for (int i = 0; i < n; i++) {
		double err = fabs(yerr[i] / yscale[i]);
		if (err > errorMax) errorMax = err;
	}
	return errorMax / epsilon;
}

/* info[0] = number of Step attempts, info[1] = last errorMax */
static int rkf78_driver(oracle_sys *s, double *time, double *hNext, double *hDid, double *info)
{
	const double SAFETY = 0.9, PGROW = -0.2, PSHRNK = -0.25, ERRCON = 1.89e-4, TINY = 1.0e-30;
	const double epsilon = pow(10, -10.0);          /* :38-39 */
	int nVar = 6 * s->n;
	int result = 0;
	s->time = *time;
	s->h = *hNext;
	s->evalDrag = s->evalMigI = s->evalMigII = 1;
	compute(s, *time, s->y0, s->accel);
	for (int i = 0; i < nVar; i++)
		s->yscale[i] = fabs(s->y0[i]) + fabs(s->h * s->accel[i]) + TINY;
	s->evalMigI = s->evalMigII = 0;
	double errorMax = 0.0;
	int attempts = 0;
	for (;;) {
		rkf78_step(s);
		attempts++;
		errorMax = rkf78_error_max(nVar, s->error, s->yscale, epsilon);
		if (errorMax < 1.0) {
			*hDid = s->h;
			result = 0;
			break;
		}
		double hTemp = SAFETY * s->h * pow(errorMax, PSHRNK);
		s->h = fabs(hTemp) > fabs(0.1 * s->h) ? hTemp : 0.1 * s->h;
		double tNew = *time + s->h;
		if (tNew == *time) { result = 1; break; }
	}
	if (result == 0) {
		*time += *hDid;
		s->time = *time;
		*hNext = errorMax > ERRCON ? (SAFETY * s->h * pow(errorMax, PGROW)) : (5.0 * s->h);
		s->h = *hNext;
		swap_ptr(&s->y0, &s->y);
	}
	if (info) { info[0] = attempts; info[1] = errorMax; }
	return result;
}

/* ------------------------------------------------------------------------------------------
 * DormandPrince RKN7(6).  Solaris/DormandPrince.cpp:26-124 (coefficients), :126-170 (Driver),
 * :246-491 (Step2), :493-503 (GetErrorMax)
 * ------------------------------------------------------------------------------------------ */
typedef struct { double b[9], bdh[9], c[9], a[9][8]; } dp_tableau;

static void dp_init(dp_tableau *T)
{
	memset(T, 0, sizeof(*T));
	double sQ = sqrt(21.0);
	T->b[0] = 1.0 / 20.0; T->b[4] = 8.0 / 45.0; T->b[5] = 7.0 * (7.0 + sQ) / 360.0; T->b[6] = 7.0 * (7.0 - sQ) / 360.0;
	T->b[7] = -1.0 / 20.0; T->b[8] = 1.0 / 20.0;
	T->bdh[0] = 1.0 / 20.0; T->bdh[4] = 16.0 / 45.0; T->bdh[5] = 49.0 / 180.0; T->bdh[6] = 49.0 / 180.0; T->bdh[7] = 1.0 / 20.0;
	T->c[0] = 0.0; T->c[1] = 1.0 / 10.0; T->c[2] = 1.0 / 5.0; T->c[3] = 3.0 / 8.0; T->c[4] = 1.0 / 2.0;
	T->c[5] = (7.0 - sQ) / 14.0; T->c[6] = (7.0 + sQ) / 14.0; T->c[7] = 1.0; T->c[8] = 1.0;
	T->a[1][0] = 1.0 / 200.0;
	T->a[2][0] = 1.0 / 150.0; T->a[2][1] = 1.0 / 75.0;
	T->a[3][0] = 171.0 / 8192.0; T->a[3][1] = 45.0 / 4096.0; T->a[3][2] = 315.0 / 8192.0;
	T->a[4][0] = 5.0 / 288.0; T->a[4][1] = 25.0 / 528.0; T->a[4][2] = 25.0 / 672.0; T->a[4][3] = 16.0 / 693.0;
	T->a[5][0] = (1003.0 - 205.0 * sQ) / 12348.0; T->a[5][1] = -25.0 * (751.0 - 173.0 * sQ) / 90552.0;
	T->a[5][2] = 25.0 * (624.0 - 137.0 * sQ) / 43218.0; T->a[5][3] = -128.0 * (361.0 - 79.0 * sQ) / 237699.0;
	T->a[5][4] = (3411.0 - 745.0 * sQ) / 24696.0;
	T->a[6][0] = (793.0 + 187.0 * sQ) / 12348.0; T->a[6][1] = -25.0 * (331.0 + 113.0 * sQ) / 90552.0;
	T->a[6][2] = 25.0 * (1044.0 + 247.0 * sQ) / 43218.0; T->a[6][3] = -128.0 * (14885.0 + 3779.0 * sQ) / 9745659.0;
	T->a[6][4] = (3327.0 + 797.0 * sQ) / 24696.0; T->a[6][5] = -(581.0 + 127.0 * sQ) / 1722.0;
	T->a[7][0] = -(157.0 - 3.0 * sQ) / 378.0; T->a[7][1] = 25.0 * (143.0 - 10.0 * sQ) / 2772.0;
	T->a[7][2] = -25.0 * (876.0 + 55.0 * sQ) / 3969.0; T->a[7][3] = 1280.0 * (913.0 + 18.0 * sQ) / 596673.0;
	T->a[7][4] = -(1353.0 + 26.0 * sQ) / 2268.0; T->a[7][5] = 7.0 * (1777.0 + 377.0 * sQ) / 4428.0;
	T->a[7][6] = 7.0 * (5.0 - sQ) / 36.0;
	T->a[8][0] = 1.0 / 20.0; T->a[8][4] = 8.0 / 45.0; T->a[8][5] = 7.0 * (7.0 + sQ) / 360.0; T->a[8][6] = 7.0 * (7.0 - sQ) / 360.0;
}

static void dp_step2(oracle_sys *s, const dp_tableau *T)
{
	int nVar = 6 * s->n, n_total = s->n;
	double *f[9];
	for (int i = 1; i < 9; i++) f[i] = (double *)calloc(nVar, sizeof(double));
	double *yTemp = (double *)calloc(nVar, sizeof(double));
	double h = s->h, h2 = h * h;
	const double *y0 = s->y0;
	const double (*a)[8] = T->a;
	const double *c = T->c;
	f[0] = s->accel;
	for (int k = 1; k <= 8; k++) {
		double ttemp = s->time + c[k] * h;
		for (int i = 0; i < n_total; i++) {
			int i0 = 6 * i;
			for (int j = 0; j < 3; j++) {
				int n = i0 + j;
				double var;
				switch (k) {   /* sums exactly as unrolled in Step2, :274-409 */
				case 1: var = a[k][0] * f[0][n + 3]; break;
				case 2: var = a[k][0] * f[0][n + 3] + a[k][1] * f[1][n + 3]; break;
				case 3: var = a[k][0] * f[0][n + 3] + a[k][1] * f[1][n + 3] + a[k][2] * f[2][n + 3]; break;
				case 4: var = a[k][0] * f[0][n + 3] + a[k][1] * f[1][n + 3] + a[k][2] * f[2][n + 3] + a[k][3] * f[3][n + 3]; break;
				case 5: var = a[k][0] * f[0][n + 3] + a[k][1] * f[1][n + 3] + a[k][2] * f[2][n + 3] + a[k][3] * f[3][n + 3] + a[k][4] * f[4][n + 3]; break;
				case 6: var = a[k][0] * f[0][n + 3] + a[k][1] * f[1][n + 3] + a[k][2] * f[2][n + 3] + a[k][3] * f[3][n + 3] + a[k][4] * f[4][n + 3] + a[k][5] * f[5][n + 3]; break;
				case 7: var = a[k][0] * f[0][n + 3] + a[k][1] * f[1][n + 3] + a[k][2] * f[2][n + 3] + a[k][3] * f[3][n + 3] + a[k][4] * f[4][n + 3] + a[k][5] * f[5][n + 3] + a[k][6] * f[6][n + 3]; break;
				default: var = a[k][0] * f[0][n + 3] + a[k][4] * f[4][n + 3] + a[k][5] * f[5][n + 3] + a[k][6] * f[6][n + 3]; break;
				}
				yTemp[n] = y0[n] + c[k] * h * y0[n + 3] + h2 * (var);
				yTemp[n + 3] = y0[n + 3] + h * (var);
			}
		}
		compute(s, ttemp, yTemp, f[k]);
	}
	const double *b = T->b, *bdh = T->bdh;
	for (int i = 0; i < n_total; i++) {   /* :471-483 */
		int i0 = 6 * i;
		for (int j = 0; j < 3; j++) {
			int n = i0 + j;
			s->y[n] = y0[n] + h * y0[n + 3] + h2 * (b[0] * f[0][n + 3] + b[4] * f[4][n + 3] + b[5] * f[5][n + 3] +
			                                        b[6] * f[6][n + 3] + b[7] * f[7][n + 3] + b[8] * f[8][n + 3]);
			s->error[n] = h2 * fabs(f[7][n + 3] - f[8][n + 3]) / 20.0;
			s->y[n + 3] = y0[n + 3] + h * (bdh[0] * f[0][n + 3] + bdh[4] * f[4][n + 3] + bdh[5] * f[5][n + 3] +
			                               bdh[6] * f[6][n + 3] + bdh[7] * f[7][n + 3]);
			s->error[n + 3] = 0.0;
		}
	}
	for (int i = 1; i < 9; i++) free(f[i]);
	free(yTemp);
}

static int dp_driver(oracle_sys *s, double *time, double *hNext, double *hDid, double *info)
{
	static dp_tableau T;
	static int T_ready = 0;
	if (!T_ready) { dp_init(&T); T_ready = 1; }
	const double epsilon = pow(10, -10.0);
	const int maxIter = 10;
	s->time = *time;
	s->evalDrag = s->evalMigI = s->evalMigII = 1;
	compute(s, *time, s->y0, s->accel);
	s->evalMigI = s->evalMigII = 0;
	int iter = 0;
	double errorMax = 0.0;
	do {
		iter++;
		s->h = *hNext;
		dp_step2(s, &T);
		errorMax = 0.0;
		int nVar = 6 * s->n;
		for (int i = 0; i < nVar; i++) {
			double e = fabs(s->error[i]);
			if (e > errorMax) errorMax = e;
		}
		*hDid = s->h;
		*hNext = errorMax < 1.0e-20 ? 2.0 * s->h : 0.9 * s->h * pow(epsilon / errorMax, 1.0 / 7.0);
	} while (errorMax > epsilon && iter <= maxIter);
	if (info) { info[0] = iter; info[1] = errorMax; }
	if (iter > maxIter) return 1;
	*time += *hDid;
	s->time = *time;
	swap_ptr(&s->y0, &s->y);
	return 0;
}

/* One Driver call.  integrator: 0 DormandPrince, 1 RungeKutta4, 3 RungeKuttaFehlberg78
 * (IntegratorType.h:12-17).  info (nullable, 2 doubles): attempts, last errorMax. */
int oracle_step(oracle_sys *s, int integrator, double *time, double *hNext, double *hDid, double *info)
{
	switch (integrator) {
	case INT_RK4:   if (info) { info[0] = 1; info[1] = 0; } return rk4_driver(s, time, hNext, hDid);
	case INT_RKF78: return rkf78_driver(s, time, hNext, hDid, info);
	case INT_DP:    return dp_driver(s, time, hNext, hDid, info);
	}
	return 1;
}

/* Tools::CheckAgainstSmallestNumber on y and y0, Solaris/Tools.cpp:39-46, Simulator.cpp:159-162 */
void oracle_flush_tiny(oracle_sys *s)
{
	int n6 = 6 * s->n;
	for (int i = 0; i < n6; i++) if (fabs(s->y[i]) < 1.0e-50) s->y[i] = 0.0;
	for (int i = 0; i < n6; i++) if (fabs(s->y0[i]) < 1.0e-50) s->y0[i] = 0.0;
}

/* ------------------------------------------------------------------------------------------
 * Event DETECTION of Simulator::CheckEvent, Solaris/Simulator.cpp:621-646 (ejection / hit
 * centrum scan) and :690-695 (collision criterion on the nearest-neighbour arrays).  The merge /
 * removal replay (:648-735) is host logic kept verbatim by the drop-in and is out of the
 * oracle's detection scope.  Outputs are index lists in scan order.
 *   ej_idx / hc_idx: body indices i (1..n-1) that fire; col_idx: indices i (0..n-1) whose
 *   criterion  factor*(R_i+R_j) > distanceOfNN[i]  holds on the UNMODIFIED NN arrays.
 * Returns the three counts through n_out[3].
 * ------------------------------------------------------------------------------------------ */
void oracle_detect_events(oracle_sys *s, double ejection, double hitCentrum, double collisionFactor,
                          int *ej_idx, int *hc_idx, int *col_idx, int *n_out)
{
	double e3 = ejection > 0 ? 1.0 / (ejection * ejection * ejection) : 0.0;
	double h3 = hitCentrum > 0 ? 1.0 / (hitCentrum * hitCentrum * hitCentrum) : 0.0;
	int ne = 0, nh = 0, nc = 0;
	for (int i = 1; i < s->n; i++) {
		if (ejection > 0 && s->rm3[i] < e3) ej_idx[ne++] = i;
		if (hitCentrum > 0 && s->rm3[i] > h3) hc_idx[nh++] = i;
	}
	if (collisionFactor > 0) {
		for (int i = 0; i < s->n; i++) {
			int j = s->indexOfNN[i];
			if (j >= 0 && collisionFactor * (s->radius[i] + s->radius[j]) > s->distanceOfNN[i]) col_idx[nc++] = i;
		}
	}
	n_out[0] = ne; n_out[1] = nh; n_out[2] = nc;
}

/* Wall-clock helper for bench.py's cpu_baseline: median seconds of `reps` gravity-row sweeps. */
#include <time.h>
double oracle_time_gravity_rows(oracle_sys *s, double *dydt, int ib, int ie, int threads, int reps)
{
	double tms[64];
	if (reps > 64) reps = 64;
	if (reps < 1) reps = 1;
	for (int r = 0; r < reps; r++) {
		struct timespec a, b;
		clock_gettime(CLOCK_MONOTONIC, &a);
		oracle_gravity_rows(s, s->y0, dydt, ib, ie, threads);
		clock_gettime(CLOCK_MONOTONIC, &b);
		tms[r] = (b.tv_sec - a.tv_sec) + 1e-9 * (b.tv_nsec - a.tv_nsec);
	}
	for (int i = 1; i < reps; i++) { double v = tms[i]; int j = i - 1; while (j >= 0 && tms[j] > v) { tms[j + 1] = tms[j]; j--; } tms[j + 1] = v; }
	return tms[reps / 2];
}

/* ---- test hooks for the reference's own known-answer vectors:
 * src/Solaris.NBody.Cuda.Test/unit_test.cpp:558-699 and Test/Test.cpp:421-532 ---- */
void oracle_circular_velocity(double mu, double x, double y, double *out2) { circular_velocity(mu, x, y, &out2[0], &out2[1]); }
void oracle_gas_velocity(const oracle_nebula_pod *g, double mu, double x, double y, double *out2) { gas_velocity(g, mu, x, y, &out2[0], &out2[1]); }
double oracle_gas_density_at(const oracle_nebula_pod *g, double x, double y, double z) { return gas_density_at(g, x, y, z); }
double oracle_temperature_cmu(const oracle_nebula_pod *g, double mC, double r) { return temperature_cmu(g, mC, r); }
double oracle_mean_thermal_speed_cmu(const oracle_nebula_pod *g, double mC, double r) { return mean_thermal_speed_cmu(g, mC, r); }
double oracle_mean_free_path(const oracle_nebula_pod *g, double r) { return powerlaw(g->mean_free_path_c, g->mean_free_path_index, r); }
double oracle_reduction_factor(const oracle_nebula_pod *g, double t) { return reduction_factor(g, t); }
int oracle_orbital_element_ae(double mu, const double *rv, double *a, double *e) { return orbital_element_ae(mu, rv, a, e); }

/* ------------------------------------------------------------------------------------------
 * (f) next row 1: Calculate::Integrals, Solaris/Calculate.cpp:43-63 with TotalMass :64-71,
 * PhaseOfBC :73-92, AngularMomentum :106-124, PotentialEnergy :139-159 (O(n^2) over ALL bodies),
 * KineticEnergy :161-172.  out[16] = mass, bc position (3), bc velocity (3), |bc r|, |bc v|,
 * L (3), |L|, kinetic, potential, kinetic - potential.  Evaluated on the current y0.
 * ------------------------------------------------------------------------------------------ */
void oracle_integrals(oracle_sys *s, double *out)
{
	const double *y0 = s->y0;
	double M = 0.0;
	for (int i = 0; i < n_massive(s); i++) M += s->mass[i];
	out[0] = M;
	double bc[6] = {0, 0, 0, 0, 0, 0};
	for (int i = 0; i < s->n; i++)
		for (int j = 0; j < 6; j++) bc[j] += s->mass[i] * y0[6 * i + j];
	for (int j = 0; j < 6; j++) { bc[j] /= M; out[1 + j] = bc[j]; }
	out[7] = sqrt(SQR(out[1]) + SQR(out[2]) + SQR(out[3]));
	out[8] = sqrt(SQR(out[4]) + SQR(out[5]) + SQR(out[6]));
	double cx = 0.0, cy = 0.0, cz = 0.0;
	for (int i = 0; i < s->n; i++) {
		const double *r = &y0[6 * i], *v = &y0[6 * i + 3];
		double lx = r[1] * v[2] - r[2] * v[1], ly = r[2] * v[0] - r[0] * v[2], lz = r[0] * v[1] - r[1] * v[0];
		cx += s->mass[i] * lx; cy += s->mass[i] * ly; cz += s->mass[i] * lz;
	}
	out[9] = cx; out[10] = cy; out[11] = cz;
	out[12] = sqrt(cx * cx + cy * cy + cz * cz);
	double kin = 0.0;
	for (int i = 0; i < s->n; i++) {
		double v2 = SQR(y0[6 * i + 3]) + SQR(y0[6 * i + 4]) + SQR(y0[6 * i + 5]);
		kin += 0.5 * s->mass[i] * v2;
	}
	double pot = 0.0;
	for (int i = 0; i < s->n; i++) {
		for (int j = 0; j < s->n; j++) {
			if (i == j) continue;
			double dx = y0[6 * j + 0] - y0[6 * i + 0], dy = y0[6 * j + 1] - y0[6 * i + 1], dz = y0[6 * j + 2] - y0[6 * i + 2];
			double rij = sqrt(dx * dx + dy * dy + dz * dz);
			pot += s->mass[i] * s->mass[j] / rij;
		}
	}
	pot *= 0.5 * K_GAUSS2;
	out[13] = kin; out[14] = pot; out[15] = kin - pot;
}

/* ------------------------------------------------------------------------------------------
 * (f) next row 2: the snapshot record of Phases.dat, BinaryFileAdapter::SavePhases / SavePhase
 * (Solaris/BinaryFileAdapter.cpp:107-122,161-169): double time, int n, then per body int id + 6 doubles,
 * no padding (12 + 52 n bytes, host byte order).  Returns the number of bytes written to out.
 * ------------------------------------------------------------------------------------------ */
size_t oracle_pack_phases(double time, int n, const double *y, const int *id, unsigned char *out)
{
	unsigned char *p = out;
	memcpy(p, &time, sizeof(time)); p += sizeof(time);
	memcpy(p, &n, sizeof(n)); p += sizeof(n);
	for (int i = 0; i < n; i++) {
		memcpy(p, &id[i], sizeof(int)); p += sizeof(int);
		memcpy(p, &y[6 * i], 6 * sizeof(double)); p += 6 * sizeof(double);
	}
	return (size_t)(p - out);
}

/* The TEXT variant (BinaryFileAdapter.cpp:133-142,171-175): iostream default float format is printf's %g,
 * so setw(15) << setprecision(10) == "%15.10g", setw(8) << int == "%8d", setw(15) << setprecision(6) ==
 * "%15.6g"; one line per snapshot.  Returns the number of characters written (cap must be >= 24 + 98 n + 2). */
size_t oracle_format_phases_text(double time, int n, const double *y, const int *id, char *out, size_t cap)
{
	size_t k = 0;
	k += (size_t)snprintf(out + k, cap - k, "%15.10g%8d", time, n);
	for (int i = 0; i < n; i++) {
		k += (size_t)snprintf(out + k, cap - k, "%8d", id[i]);
		for (int c = 0; c < 6; c++) k += (size_t)snprintf(out + k, cap - k, "%15.6g", y[6 * i + c]);
	}
	k += (size_t)snprintf(out + k, cap - k, "\n");
	return k;
}

/* ------------------------------------------------------------------------------------------
 * (f) next row 3: Simulator::RemoveBody (Solaris/Simulator.cpp:737-771) with NBodies::UpdateAfterRemove
 * (Solaris/NBodies.cpp:80-113).  The body with this id leaves; id, type, migType, mass, radius, density,
 * gammaStokes, gammaEpstein and y0 of the bodies behind it move down one slot.  cD and migStopAt are NOT
 * moved by the reference (they keep their slots, :757-769), nor are y, rm3 and the nearest-neighbour arrays.
 * Returns 1 for an unknown body type like the reference; an id that does not exist is not guarded by the
 * reference (reads one past the end) and is rejected here.
 * ------------------------------------------------------------------------------------------ */
int oracle_remove_body(oracle_sys *s, int bodyId)
{
	int index = 0;
	for (; index < s->n; index++) if (s->id[index] == bodyId) break;
	if (index >= s->n) return 2;
	const int t = s->type[index];          /* BodyType: CentralBody = 1 ... TestParticle = 7 */
	if (t < 1 || t > 7) return 1;          /* "Unknown or undefined Body Type!" */
	s->counts[t - 1]--;
	s->n--;
	for (int i = index; i < s->n; i++) {
		s->id[i] = s->id[i + 1];
		s->type[i] = s->type[i + 1];
		s->migType[i] = s->migType[i + 1];
		s->mass[i] = s->mass[i + 1];
		s->radius[i] = s->radius[i + 1];
		s->density[i] = s->density[i + 1];
		s->gammaStokes[i] = s->gammaStokes[i + 1];
		s->gammaEpstein[i] = s->gammaEpstein[i + 1];
		memcpy(&s->y0[6 * i], &s->y0[6 * i + 6], 6 * sizeof(double));
	}
	return 0;
}

/* current per-body parameter arrays and counts (n = sum of counts entries are valid) */
void oracle_get_params(const oracle_sys *s, int counts[7], double *mass, double *radius, double *density, double *cD,
                       double *gammaStokes, double *gammaEpstein, double *migStopAt, int *type, int *migType, int *id)
{
	memcpy(counts, s->counts, 7 * sizeof(int));
	const size_t n = (size_t)s->n;
	memcpy(mass, s->mass, n * sizeof(double)); memcpy(radius, s->radius, n * sizeof(double));
	memcpy(density, s->density, n * sizeof(double)); memcpy(cD, s->cD, n * sizeof(double));
	memcpy(gammaStokes, s->gammaStokes, n * sizeof(double)); memcpy(gammaEpstein, s->gammaEpstein, n * sizeof(double));
	memcpy(migStopAt, s->migStopAt, n * sizeof(double));
	memcpy(type, s->type, n * sizeof(int)); memcpy(migType, s->migType, n * sizeof(int)); memcpy(id, s->id, n * sizeof(int));
}

/* ------------------------------------------------------------------------------------------
 * (f) next row 4 (loader): orbital elements -> phase, Ephemeris::CalculatePhase (Solaris/Ephemeris.cpp:141-176)
 * with Ephemeris::KeplerEquationSolver (:187-213; start value and Newton steps from B. Erdi, eps = 1e-14, at most
 * 26 iterations), as Simulation::SetPhasesRadiiDensity applies it to every body given by elements
 * (Solaris/Simulation.cpp:131-172; mu = G(m0 + m), m = 0 for test particles).  el = {a, e, incl, peri, node, M}.
 * Returns 1 when the Kepler solver does not converge ("Could not compute the excentric anomaly E!").
 * ------------------------------------------------------------------------------------------ */
static int kepler_solve(double e, double m, double eps, double *Eout)
{
	if (e == 0.0 || m == 0.0 || m == 3.14159265358979323846) { *Eout = m; return 0; }   /* Constants::Pi */
	double E = m + e * (sin(m)) / (1.0 - sin(m + e) + sin(m));
	double E1 = 0.0, error;
	int step = 0;
	do {
		E1 = E - (E - e * sin(E) - m) / (1.0 - e * cos(E));
		error = fabs(E1 - E);                 /* abs() with MSVC / absfix semantics, SURVEY.md Q12 */
		E = E1;
		step++;
	} while (error > eps && step <= 25);
	*Eout = E;
	return step > 25 ? 1 : 0;
}

int oracle_elements_to_phase(double mu, const double *el, double *out)
{
	const double a = el[0], e = el[1], incl = el[2], peri = el[3], node = el[4], M = el[5];
	double E = 0;
	if (kepler_solve(e, M, 1.0e-14, &E) == 1) return 1;
	double v = 2.0 * atan(sqrt((1.0 + e) / (1.0 - e)) * tan(E / 2.0));
	double p = a * (1.0 - e * e);
	double r = p / (1.0 + e * cos(v));
	double kszi = r * cos(v);
	double eta = r * sin(v);
	double vKszi = -sqrt(mu / p) * sin(v);
	double vEta = sqrt(mu / p) * (e + cos(v));
	double cw = cos(peri), sw = sin(peri), cO = cos(node), sO = sin(node), ci = cos(incl), si = sin(incl);
	double P[3] = {cw * cO - sw * sO * ci, cw * sO + sw * cO * ci, sw * si};
	double Q[3] = {-sw * cO - cw * sO * ci, -sw * sO + cw * cO * ci, cw * si};
	for (int c = 0; c < 3; c++) {
		out[c] = kszi * P[c] + eta * Q[c];        /* Vector operator*(double, Vector) then operator+ (Vector.cpp) */
		out[3 + c] = vKszi * P[c] + vEta * Q[c];
	}
	return 0;
}

/* batch form; returns the number of bodies whose Kepler equation did not converge (their rows are left untouched) */
int oracle_elements_to_phases(int n, const double *mu, const double *el6, double *out6)
{
	int bad = 0;
	for (int i = 0; i < n; i++) bad += oracle_elements_to_phase(mu[i], el6 + 6 * (size_t)i, out6 + 6 * (size_t)i);
	return bad;
}

/* ------------------------------------------------------------------------------------------
 * Event RECORDS of the ejection / hit-centrum scan, Simulator::CheckEvent (Solaris/Simulator.cpp:631-646):
 *   TwoBodyAffair(Ejection | HitCentrum, timeOfEvent, 0, i, id[0], id[i], y0, &y0[6 i])
 * (TwoBodyAffair.cpp:9-21: id = running static counter, one per constructed affair, in scan order over i with the
 * ejection test before the hit-centrum test) in the byte layout BinaryFileAdapter::SaveTwoBodyAffair writes
 * (BinaryFileAdapter.cpp:244-261): int id, int type, int body1Id, int body2Id, double body1Phase[6],
 * double body2Phase[6], double time = 120 bytes.  Output: all ejection records, then all hit-centrum records, like the
 * two SaveTwoBodyAffairs calls (:648-649, :661-662).  Uses rm3 of the last evaluation and the current y0.
 * Returns the number of records; n_out[0] / n_out[1] = ejections / hit centrums.
 * ------------------------------------------------------------------------------------------ */
static void put_record(unsigned char *p, int id, int type, int b1, int b2, const double *p1, const double *p2, double time)
{
	memcpy(p, &id, 4); memcpy(p + 4, &type, 4); memcpy(p + 8, &b1, 4); memcpy(p + 12, &b2, 4);
	memcpy(p + 16, p1, 48); memcpy(p + 64, p2, 48); memcpy(p + 112, &time, 8);
}

int oracle_event_records(const oracle_sys *s, double ejection, double hitCentrum, double time, int first_event_id,
                         unsigned char *out, int *n_out)
{
	double e3 = ejection > 0 ? 1.0 / (ejection * ejection * ejection) : 0.0;
	double h3 = hitCentrum > 0 ? 1.0 / (hitCentrum * hitCentrum * hitCentrum) : 0.0;
	int ne = 0, nh = 0;
	for (int i = 1; i < s->n; i++) {
		if (ejection > 0 && s->rm3[i] < e3) ne++;
		if (hitCentrum > 0 && s->rm3[i] > h3) nh++;
	}
	int id = first_event_id, ke = 0, kh = 0;
	for (int i = 1; i < s->n; i++) {
		if (ejection > 0 && s->rm3[i] < e3)
			put_record(out + 120 * (size_t)(ke++), id++, 0, s->id[0], s->id[i], s->y0, s->y0 + 6 * (size_t)i, time);
		if (hitCentrum > 0 && s->rm3[i] > h3)
			put_record(out + 120 * (size_t)(ne + kh++), id++, 1, s->id[0], s->id[i], s->y0, s->y0 + 6 * (size_t)i, time);
	}
	n_out[0] = ne; n_out[1] = nh;
	return ne + nh;
}

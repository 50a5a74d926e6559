/* solaris_b200 - C-ABI of the B200-native force-evaluation + integrator hot path of Solaris.
 *
 * This header is the drop-in boundary.  The reference has no FFI: its seam is the
 * C++ class interface `Simulator` compiles against.  The drop-in translation units under
 * solaris_b200/host/ (Acceleration.cpp, RungeKutta4.cpp, RungeKuttaFehlberg78.cpp, DormandPrince.cpp)
 * keep those class declarations and forward to the entry points below; each entry point names the
 * reference interface it replaces (paths relative to the reference root).  Plain pointers and sizes
 * only - no C++, CUDA or torch types cross this boundary.
 *
 * Host array layout is the reference's: bodies sorted by BodyType
 * (Solaris/Body.h:14-24), state and derivative arrays AoS, 6 doubles per body (x,y,z,vx,vy,vz),
 * body 0 = central body.  All functions return SOL_OK (0) or SOL_ERR (1) like the reference's
 * 0/1 convention (Solaris/Error.h:7-16); sol_last_error() gives the message.
 *
 * There is NO CPU fallback: every compute entry point fails with SOL_ERR when no CUDA device is
 * usable.
 */
#ifndef SOLARIS_B200_H_
#define SOLARIS_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOL_OK  0
#define SOL_ERR 1

typedef struct sol_ctx sol_ctx;

/* Snapshot of GasComponent (Solaris/GasComponent.h:8-52) as the reference holds it when
 * Acceleration is constructed, INCLUDING the constructor-time meanFreePath power law that XML
 * overrides of the density do not refresh. */
typedef struct sol_nebula_pod {
	double alpha;                  /* GasComponent::alpha */
	double mean_molecular_weight;  /* GasComponent::meanMolecularWeight */
	double particle_diameter;      /* GasComponent::particleDiameter [m] */
	int    decrease_type;          /* GasDecreaseType.h:4-9  0 CONSTANT, 1 LINEAR, 2 EXPONENTIAL */
	int    _pad;
	double time_scale, t0, t1;     /* GasComponent::timeScale, t0, t1 [day] */
	double inner_edge;             /* GasComponent::innerEdge [AU] */
	double eta_c, eta_index;                       /* GasComponent::eta */
	double tau_c, tau_index;                       /* GasComponent::tau */
	double scale_height_c, scale_height_index;     /* GasComponent::scaleHeight */
	double density_c, density_index;               /* GasComponent::density */
	double mean_free_path_c, mean_free_path_index; /* GasComponent::meanFreePath */
} sol_nebula_pod;

/* IntegratorType (Solaris/IntegratorType.h:12-17) */
#define SOL_DORMAND_PRINCE          0
#define SOL_RUNGE_KUTTA4            1
#define SOL_RUNGE_KUTTA_FEHLBERG78  3

/* eval_flags = Acceleration::evaluateGasDrag / evaluateTypeIMigration / evaluateTypeIIMigration
 * (Solaris/Acceleration.h:54-56) */
#define SOL_EVAL_GAS_DRAG   1u
#define SOL_EVAL_MIG_TYPE1  2u
#define SOL_EVAL_MIG_TYPE2  4u
#define SOL_EVAL_ALL        7u

/* Arrays addressable by sol_download / sol_upload. */
#define SOL_Y0           0   /* BodyData::y0            double[6n] AoS */
#define SOL_Y            1   /* BodyData::y             double[6n] AoS */
#define SOL_ACCEL        2   /* BodyData::accel (k0)    double[6n] AoS; valid after sol_compute_device, and after
                                 sol_step on the general path (the single-CTA / tracer kernels keep k0 on chip) */
#define SOL_YSCALE       3   /* BodyData::yscale        double[6n] AoS */
#define SOL_RM3          4   /* Acceleration::rm3       double[n]      */
#define SOL_NN_INDEX     5   /* BodyData::indexOfNN     int[n]         */
#define SOL_NN_DISTANCE  6   /* BodyData::distanceOfNN  double[n]      */
#define SOL_MIGTYPE      7   /* BodyData::migType       int[n]         */
#define SOL_MASS         8   /* BodyData::mass          double[n]      */
#define SOL_RADIUS       9   /* BodyData::radius        double[n]      */
#define SOL_ACCEL_GASDRAG   10 /* Acceleration::accelGasDrag         double[3*NOfPlAndSpl]    */
#define SOL_ACCEL_MIGTYPE1  11 /* Acceleration::accelMigrationTypeI  double[3*(rocky+proto)]  */
#define SOL_ACCEL_MIGTYPE2  12 /* Acceleration::accelMigrationTypeII double[3*giant]          */
#define SOL_DENSITY      13  /* BodyData::density       double[n]      */
#define SOL_CD           14  /* BodyData::cD            double[n]      */
#define SOL_GAMMA_STOKES 15  /* BodyData::gammaStokes   double[n]      */
#define SOL_GAMMA_EPSTEIN 16 /* BodyData::gammaEpstein  double[n]      */
#define SOL_MIGSTOPAT    17  /* BodyData::migStopAt     double[n]      */
#define SOL_TYPE         18  /* BodyData::type          int[n]         */
#define SOL_ID           19  /* BodyData::id            int[n]         */

/* ---- lifecycle -------------------------------------------------------------------------- */

/* Creates a context bound to CUDA device `device` (one context per process per GPU).
 * Replaces: construction of Acceleration (Solaris/Acceleration.cpp:35-50, Simulator.cpp:82). */
int  sol_create(int device, sol_ctx **out);
/* One handle over the first n_gpus devices of this process (SURVEY.md §8b proposed sol_create(int n_gpus, ...)): the
 * reference's host program is ONE single-threaded process (Simulator::Integrate, Solaris/Simulator.cpp:121-178), so this is
 * the form the C++ drop-in uses to run a large system on all GPUs of a box.  Sinks are sharded over the devices and
 * the sources exchanged over NVLink exactly as with one process per GPU (sol_dist_init below); internally one worker
 * thread per device runs the same single-device code, and every entry point of this header called on the handle acts
 * on all of them and returns when all are done.  Host arrays passed to sol_download / sol_compute are filled rank by
 * rank, each rank writing the rows of its own sinks.  n_gpus == 1 is sol_create(0). */
int  sol_create_multi(int n_gpus, sol_ctx **out);
void sol_destroy(sol_ctx *ctx);
/* Message of the last failure on this context (Error::_errMsg, Solaris/Error.h:12).  ctx may be
 * NULL to read the message of a failed sol_create. */
const char *sol_last_error(const sol_ctx *ctx);
/* Run all work of this context on an existing CUDA stream (a cudaStream_t passed as void*), so a
 * host program can time or order it with its own events.  Default: a private stream.  (No counterpart in the
 * reference, which is synchronous host code.) */
int  sol_set_stream(sol_ctx *ctx, void *cuda_stream);

/* ---- configuration ---------------------------------------------------------------------- */

/* Uploads a complete BodyData and (re)builds the device-resident SoA mirror.
 * counts[7] = NBodies::centralBody, giantPlanet, rockyPlanet, protoPlanet, superPlanetsimal,
 * planetsimal, testParticle (Solaris/NBodies.h:21-27).  Must be called again whenever the host
 * removes or merges bodies.  Replaces: Simulator::BodyListToBodyData (Solaris/Simulator.cpp:525-593)
 * + BodyData::Allocate (Solaris/BodyData.cpp:66-187) on the device side. */
int sol_set_bodies(sol_ctx *ctx, const int counts[7], const double *y0_aos6,
                   const double *mass, const double *radius, const double *density, const double *cD,
                   const double *gammaStokes, const double *gammaEpstein, const double *migStopAt,
                   const int *type, const int *migType, const int *id);
/* Settings::baryCentric (Solaris/Settings.cpp:7): 0 astrocentric (default), 1 barycentric. */
int sol_set_frame(sol_ctx *ctx, int barycentric);
/* Simulation::nebula (Solaris/Simulator.cpp:82); NULL = no nebula. */
int sol_set_nebula(sol_ctx *ctx, const sol_nebula_pod *nebula);
/* track_nn: 2 (default) = indexOfNN / distanceOfNN are produced by the LAST evaluation of every Driver
 * call and by every sol_compute call - exactly the values that are observable in the reference, whose
 * earlier stages' NN arrays are overwritten before anything can read them;
 * 1 = by every evaluation (the reference's literal habit; same observable results, more work);
 * 0 = never (legal only when Settings::collision == 0).
 * The arrays are the ones Acceleration::GravityAC / GravityBC_* fill per sink (Solaris/Acceleration.cpp:301-311,
 * :563-571, :610-618) and Simulator::CheckEvent reads (Solaris/Simulator.cpp:690-695). */
int sol_set_nn_tracking(sol_ctx *ctx, int track_nn);

/* Pair-interaction algorithm for the self-gravitating block (sinks == sources): the symmetric kernel evaluates
 * each unordered pair once (Newton's third law), the ordered kernel every (sink, source) pair like the reference's
 * double loop.  2 (default) = symmetric kernel from 12288 bodies, where it overtakes the ordered one (it needs
 * ~nb^2/2 CTAs of 512 x 512 bodies to fill the chip); 1 = symmetric kernel from 4096 bodies; 0 = always the ordered
 * kernel.  Results agree to rounding (different summation order). */
int sol_set_pair_algorithm(sol_ctx *ctx, int mode);

/* Systems of at most 256 bodies on one GPU can run every Driver attempt as ONE kernel launch (a single
 * CTA walks all stages with block barriers; arithmetic identical to the multi-launch path); by default that kernel is
 * chosen up to 160 bodies, where the multi-launch path with graph replay overtakes it (modes 2 and 3: up to 256).  When the bodies the
 * single CTA integrates are at most 32 and all massive, a one-warp variant runs instead (k-vectors in registers, warp
 * barriers and shuffles; identical arithmetic again).  sol_run additionally has a component-parallel one-warp kernel for
 * at most 10 massive bodies (one lane per body AND coordinate; identical arithmetic).  1 (default) = all of them,
 * 3 = without the component-parallel kernel, 2 = single-CTA kernel only, 0 = always the multi-launch path. */
int sol_set_small_system_kernel(sol_ctx *ctx, int on);

/* Systems with at most 64 massive bodies, no super-planetesimals and any number of planetesimals /
 * test particles: the massive bodies run in the single-CTA kernel (which records their trial positions
 * at every evaluation), and every non-source body's WHOLE attempt runs privately in one thread of
 * tracer_attempt_kernel (k-vectors in thread-local storage; y0 read once, y written once).
 * 1 (default) = on, 0 = general multi-launch path.  Bit-identical results. */
int sol_set_tracer_kernel(sol_ctx *ctx, int on);

/* Systems on the general multi-launch path with at most 32768 bodies on one GPU (257 ... a few 10^4 self-gravitating
 * bodies) are bound by the chain of ~40 short dependent kernels per RKF78 attempt, not by their work.
 *   mode 1 (default): the launches of a Driver call replay from CUDA graphs captured once per integrator;
 *   mode 2: they run as phases of ONE cooperative kernel per segment (k0 evaluation / rest of the attempt), separated
 *           by grid barriers instead of kernel boundaries - measured within 10 % of mode 1 (a phase is
 *           bound by its own dependent chain, a grid barrier costs what a graph edge costs); where it does not apply
 *           (symmetric pair kernel in use, no cooperative launch) it behaves like mode 1;
 *   mode 0: every launch is issued from the host.
 * In modes 1 and 2 the per-attempt scalars (h, c_k h, the gas reduction factors) are read by the kernels from device
 * memory.  Same device code over the same block decomposition in all three modes: bit-identical results.
 * (No counterpart in the reference.) */
int sol_set_graph_mode(sol_ctx *ctx, int mode);

/* ---- seam B: one force evaluation --------------------------------------------------------- */

/* Replaces: int Acceleration::Compute(double t, double *y, double *totalAccel)
 * (Solaris/Acceleration.h:19, Solaris/Acceleration.cpp:60-81).  y_host / dydt_host are host AoS
 * arrays of 6n doubles; host<->device copies are part of the call.  Side outputs (rm3, nearest
 * neighbour, migType, the three cached gas-term arrays) stay on the device until sol_download.  On a context that is
 * one rank of several (sol_dist_init) the call is collective - every rank passes the full y - and fills the rows of this
 * rank's own sinks only; a sol_create_multi handle fills the whole array. */
int sol_compute(sol_ctx *ctx, double t, const double *y_host, double *dydt_host, unsigned eval_flags);
/* Same evaluation (Acceleration::Compute, Solaris/Acceleration.h:19) on the device-resident state: k0 = f(t, y0).
 * No host traffic. */
int sol_compute_device(sol_ctx *ctx, double t, unsigned eval_flags);

/* ---- seam A: one integrator step ---------------------------------------------------------- */

/* Replaces: RungeKutta4::Driver (Solaris/RungeKutta4.cpp:20-56),
 *           RungeKuttaFehlberg78::Driver (Solaris/RungeKuttaFehlberg78.cpp:66-140),
 *           DormandPrince::Driver (Solaris/DormandPrince.cpp:126-170)
 * on the device-resident y0.  time / h_next are in-out and h_did is out with the meaning of
 * TimeLine::time / hNext / hDid; on success y0 holds the new state and y the previous one (the
 * reference's std::swap).  info (nullable, 4 doubles): [0] attempts (Step calls), [1] last
 * errorMax, [2] force evaluations, [3] ordered pair interactions evaluated. */
int sol_step(sol_ctx *ctx, int integrator, double *time, double *h_next, double *h_did, double *info);

/* ---- seam A, many steps per call ------------------------------------------------------------ */

/* Replaces: the step loop of Simulator::Integrate (Solaris/Simulator.cpp:131-170) between two moments at which the host
 * has to look at the state: Driver after Driver, each followed by what Simulator::DecisionMaking does on a step without
 * consequences (Simulator.cpp:181-248: lastSave += hDid, the event tests of CheckEvent :631-646,690-695, end of the
 * integration, hNext clamped to `length`, snapshot due, hNext clamped to `output`) and by the flush of every
 * flush_every-th step (:159-162).  The call returns after max_steps steps or as soon as a step
 *   SOL_RUN_END    reached |millenium_days + time| >= |length|,
 *   SOL_RUN_SAVE   made a snapshot due (|lastSave + hDid| >= |output|),
 *   SOL_RUN_EVENT  left side outputs that fire an ejection / hit-centrum / collision test (event_counts; the indices and
 *                  records are available through sol_event_indices / sol_event_records as after sol_detect_events).
 * For those three the DecisionMaking of the LAST step is left to the caller (which has to save, merge bodies or stop
 * anyway): last_save and h_next come back as they were before it, i.e. h_next is the Driver's own proposal; and the
 * caller applies the flush of that step if it is due.  With SOL_RUN_MAX_STEPS everything is applied.
 * Systems of at most 32 bodies, all massive (SunJupiter, SolarSystem), run the whole call as ONE persistent kernel
 * launch with the state in registers; the step-size formulas then use the device's pow() (<= 2 ulp) instead of the host
 * libm's, so step sizes may differ from sol_step's in the last bits.  All other systems are stepped from the host with
 * one Driver and one flag reduction per step, bit-identical to sol_step + sol_detect_events.
 * records (nullable, 4 * max_steps doubles): per step the time reached, hDid, the Driver's own hNext proposal and the trial
 * step the Driver was entered with (TimeLine::hNext after the previous step's clamps). */
#define SOL_RUN_MAX_STEPS 0
#define SOL_RUN_END       1
#define SOL_RUN_SAVE      2
#define SOL_RUN_EVENT     3
#define SOL_RUN_ERROR     4
typedef struct sol_run_args {
	int       integrator;          /* in      SOL_RUNGE_KUTTA4 / SOL_RUNGE_KUTTA_FEHLBERG78 / SOL_DORMAND_PRINCE */
	int       max_steps;           /* in      >= 1 */
	double    time;                /* in/out  TimeLine::time */
	double    h_next;              /* in/out  TimeLine::hNext */
	double    h_did;               /* out     TimeLine::hDid of the last step */
	double    millenium_days;      /* in      1000 * Constants::YearToDay * TimeLine::millenium (Simulator.cpp:197) */
	double    length;              /* in      TimeLine::length */
	double    output;              /* in      TimeLine::output */
	double    last_save;           /* in/out  TimeLine::lastSave */
	double    ejection;            /* in      Settings::ejection [AU], <= 0 off */
	double    hit_centrum;         /* in      Settings::hitCentrum [AU], <= 0 off */
	double    collision_factor;    /* in      Settings::collision->factor, <= 0 off */
	long long step_counter;        /* in/out  Counter::succededStep */
	int       flush_every;         /* in      Constants::CheckForSM (100); 0 = never */
	double    flush_threshold;     /* in      Constants::SmallestNumber (1e-50) */
	int       steps;               /* out     accepted steps of this call */
	int       stop_reason;         /* out     SOL_RUN_* */
	int       event_counts[3];     /* out     ejection, hit-centrum, collision candidates of the last step */
	long long attempts;            /* out     Step calls */
	double    err_max;             /* out     errorMax of the last attempt */
	double   *records;             /* in      NULL or room for 4 * max_steps doubles */
} sol_run_args;
int sol_run(sol_ctx *ctx, sol_run_args *args);

/* ---- events ------------------------------------------------------------------------------- */

/* Device flag reduction for Simulator::CheckEvent's three detections (Solaris/Simulator.cpp:631-646,
 * 690-695) on the side outputs of the last evaluation.  ejection / hit_centrum: Settings values in
 * AU (<= 0 disables); collision_factor: Collision::factor (<= 0 disables).  counts_out[3] =
 * number of ejection, hit-centrum and collision candidates.  Only when a count is non-zero does
 * the host need to download rm3 / NN arrays and replay the reference's merge logic. */
int sol_detect_events(sol_ctx *ctx, double ejection, double hit_centrum, double collision_factor,
                      int counts_out[3]);
/* Body indices (scan order of the loops at Solaris/Simulator.cpp:631-646 and :690-695) of the candidates found by the
 * last sol_detect_events.
 * kind: 0 ejection, 1 hit centrum, 2 collision.  Writes at most cap indices, returns the count
 * through n_out. */
int sol_event_indices(sol_ctx *ctx, int kind, int *idx_out, int cap, int *n_out);
/* Replaces: the record construction of the same scan - TwoBodyAffair(Ejection | HitCentrum, timeOfEvent, 0, i, id[0],
 * id[i], y0, &y0[6 i]) (Solaris/Simulator.cpp:636,643; Solaris/TwoBodyAffair.cpp:9-21) - for the events the last
 * sol_detect_events found, in the byte layout BinaryFileAdapter::SaveTwoBodyAffair writes (Solaris/BinaryFileAdapter.cpp:
 * 244-261; 120 bytes: int id, type, body1Id, body2Id; double body1Phase[6], body2Phase[6], time).  The phases are
 * gathered on the device, so only the records cross the bus.  Order: all ejections, then all hit centrums, each in scan
 * order (= the two SaveTwoBodyAffairs calls); ids count up from first_event_id in the order the reference constructs
 * the objects (one scan, ejection test first).  records == NULL only reports the count.  Collision records depend on
 * the host's merge logic and are not built here.  Single-GPU contexts and sol_create_multi handles (which first gather
 * the accepted state); not on one rank of a multi-process job. */
int sol_event_records(sol_ctx *ctx, double time, int first_event_id, void *records, int capacity, int *n_records);

/* ---- diagnostics (first "next" row) -------------------------------------------- */

/* Replaces: Calculate::Integrals (Solaris/Calculate.cpp:43-63) on the device-resident y0, including the
 * O(n^2) PotentialEnergy over all bodies (:139-159).  out[16] = total mass of the massive bodies,
 * barycentre position (3) and velocity (3), their norms, angular momentum (3) and norm, kinetic energy,
 * potential energy, kinetic - potential: the record written to Integrals.dat. */
int sol_integrals(sol_ctx *ctx, double out[16]);

/* Replaces: BinaryFileAdapter::SavePhases(time, n, y0, id, BINARY) (Solaris/BinaryFileAdapter.cpp:107-122 with
 * SavePhase :161-169), which appends one snapshot to Phases.dat through 2 n small stream writes.  The record
 * (double time, int n, n x {int id, double y[6]}, no padding: 12 + 52 n bytes) is assembled on the device from
 * the resident y0 and id arrays and leaves it in one transfer.
 *   sol_pack_phases: record into a caller buffer; host == NULL only reports the size in *nbytes.
 *   sol_write_phases: record appended to `path` (created if missing) with a single write().
 * Multi-GPU: the accepted state is gathered first (sol_gather_state); every rank then holds the full record. */
int sol_pack_phases(sol_ctx *ctx, double time, void *host, size_t capacity, size_t *nbytes);
int sol_write_phases(sol_ctx *ctx, const char *path, double time);

/* Replaces: the per-body Ephemeris::CalculatePhase calls of Simulation::SetPhasesRadiiDensity (Solaris/Simulation.cpp:
 * 131-172; Solaris/Ephemeris.cpp:141-176 with the Kepler solver :187-213) for a batch: elements[6 i ..] = {a, e, incl,
 * peri, node, M} (au, radians), mu[i] = G (m0 + m_i) (m_i = 0 for test particles), phases[6 i ..] = {x, y, z, vx, vy, vz}.
 * Independent of the loaded system.  Bodies whose Kepler iteration does not converge (the reference's error return)
 * keep their row of `phases`; their number is stored in *n_failed (may be NULL) and the call returns SOL_ERR with the
 * reference's message.  Agreement with the reference is to rounding (device sin / cos / tan / atan), not bit-exact. */
int sol_elements_to_phases(sol_ctx *ctx, int n, const double *mu, const double *elements, double *phases, int *n_failed);

/* Replaces: Simulator::RemoveBody (Solaris/Simulator.cpp:737-771) + NBodies::UpdateAfterRemove
 * (Solaris/NBodies.cpp:80-113) for `count` bodies at once, on the device-resident arrays: the bodies at the given
 * CURRENT indices (distinct, any order, never 0) leave; id, type, migType, mass, radius, density, gammaStokes,
 * gammaEpstein and y0 of the others close the gaps in order; the per-type counts shrink.  As in the reference,
 * cD and migStopAt keep their slots and y, rm3 and the nearest-neighbour arrays are not moved.  The result equals
 * `count` successive RemoveBody calls.  The cached gas-drag / migration terms are stale afterwards (in the reference
 * too: its caches are indexed relative to a class start and are not moved) until the next evaluation that
 * recomputes them - the first stage of every Driver call does.  Multi-GPU: call on every rank with the same indices. */
int sol_remove_bodies(sol_ctx *ctx, const int *indices, int count);
/* Replaces the host writes of Simulator::CalculatePhaseAfterCollision / CalculateCharacteristicsAfterCollision
 * (Solaris/Simulator.cpp:801-899) into BodyData for the surviving body of a merger: new y0[6], mass, radius, density
 * (computed by the caller, as the host code does). */
int sol_patch_body(sol_ctx *ctx, int index, const double y0[6], double mass, double radius, double density);

/* ---- transfers ---------------------------------------------------------------------------- */

/* One array of BodyData / Acceleration (Solaris/BodyData.h:8-51, Solaris/Acceleration.h:46-52; ids SOL_Y0 ... above) in the
 * reference's host layout, from / to the device-resident copy. */
int sol_download(sol_ctx *ctx, int what, void *host);
int sol_upload(sol_ctx *ctx, int what, const void *host);
/* Tools::CheckAgainstSmallestNumber on y and y0 (Solaris/Tools.cpp:39-46, Simulator.cpp:159-162). */
int sol_flush_tiny(sol_ctx *ctx, double threshold);
/* NBodies::total of the device-resident system (Solaris/NBodies.h:9-35). */
int sol_body_count(const sol_ctx *ctx);

/* ---- multi-GPU (one process per GPU, sinks sharded, sources replicated) ------
 * The reference is a single-process CPU program: nothing in this group replaces reference code. */

/* Fills 128 bytes with an NCCL unique id (rank 0 calls it, the launcher distributes the bytes). */
int sol_nccl_unique_id(void *out128);
/* Joins the communicator; after this call sol_set_bodies shards sinks contiguously over ranks
 * and every evaluation all-gathers the source bodies' trial positions over NVLink. */
int sol_dist_init(sol_ctx *ctx, int rank, int nranks, const void *unique_id128);
/* The partition rule itself (pure function, no device needed): contiguous chunks of
 * ceil(n / nranks) rounded up to a multiple of 32 bodies; trailing ranks may be empty. */
int sol_shard_of(int n, int nranks, int rank, int *lo, int *hi);
/* Schedule of the symmetric pair kernel (pure function, no device needed): in round `round`
 * (0 .. nb/2) CTA `p` owns the block pair (p, q) with q = (p + round) mod nb.  Returns 1 and writes q
 * when the CTA has work, 0 when it is idle (second half of the half round of an even block count), -1 on
 * bad arguments.  Over all rounds every unordered pair of distinct blocks appears exactly once and every
 * diagonal pair (p, p) exactly once (round 0). */
int sol_sym_round_pair(int nb, int round, int p, int *q);
/* Whole rounds [lo, hi) per rank: the split used until round 2; kept as a pure helper.  The library now deals the CTAs of
 * all rounds in pieces of equal cost: */
int sol_sym_rounds_of_rank(int nb, int nranks, int rank, int *lo, int *hi);
/* Share of `rank` in the symmetric kernel's work (pure function): out4 = {round_first, p_first_lo, round_last, p_last_hi}
 * - the CTAs p >= p_first_lo of round round_first, all CTAs of the rounds in between, and the CTAs p < p_last_hi of round
 * round_last (round_last < round_first: nothing).  Over all ranks every (round, CTA) with work appears exactly once. */
int sol_sym_work_of_rank(int nb, int nranks, int rank, int out4[4]);
/* Launch plan of the ordered pair kernel (pure function): `sinks` sinks of this context against `sources` sources;
 * sinks_all = the sinks of the same launch on an unsharded context (0: same as sinks).  out3 = {sinks per thread,
 * source chunks (partial sums per sink, at most 32), sources per chunk}.  A mid-size launch is cut into chunks from
 * sinks_all alone, so that every rank count sums in the same order; at most 256 sources are never cut (the order of the
 * single-CTA kernel). */
int sol_plan_pairs(int sinks, int sources, int sinks_all, int out3[3]);
/* Sink range [lo, hi) this rank integrates (whole range on one GPU). */
int sol_shard_range(const sol_ctx *ctx, int *lo, int *hi);
/* All-gathers y0 so that every rank holds the full accepted state (before output / events). */
int sol_gather_state(sol_ctx *ctx);

/* ---- measurement (bench.py, tools/; no counterpart in the reference) ----------------------- */

/* Times `reps` launches of the pair-interaction kernel alone at the current y0 with CUDA events on
 * the context's stream; ms_out = mean milliseconds per launch, pairs_out = ordered pairs per launch. */
int sol_time_gravity_kernel(sol_ctx *ctx, int reps, float *ms_out, double *pairs_out);
/* Dependent-free DFMA stream on every SM: measured FP64 FMA peak of this GPU in TFLOP/s
 * (2 flops per DFMA).  Used as the roofline denominator of the gravity kernel. */
int sol_measure_fp64_peak(sol_ctx *ctx, double *tflops_out);

/* Self-test of the device code's straight-line copies of the library's double-precision sqrt / reciprocal fast paths
 * (used by the persistent small-system kernel of sol_run so that the scheduler can overlap them with the pair sums):
 * `samples` pseudo-random arguments over the whole range in which the copies are used, compared bit for bit with
 * sqrt(x), 1.0 / x and 1.0 / (x * sqrt(x)) evaluated by the library on the device.  *mismatches_out must come back 0.
 * (No counterpart in the reference.) */
int sol_selftest_fast_paths(sol_ctx *ctx, unsigned long long seed, long long samples, unsigned long long *mismatches_out);
/* Kernel launches issued by this context since creation (bench.py's gpu_launches). */
long long sol_launch_count(const sol_ctx *ctx);
/* Accumulated device time [ms] and launch count per kernel family since the last reset, measured
 * with CUDA events when profiling is enabled (sol_profile_enable(ctx,1)); families:
 * 0 pair kernel, 1 source prep + indirect sum, 2 per-body finalize (+gas terms), 3 RK stage
 * combinations, 4 solution + error norm, 5 misc. */
int sol_profile_enable(sol_ctx *ctx, int on);
int sol_profile_read(sol_ctx *ctx, double ms_out[6], long long launches_out[6], int reset);

#ifdef __cplusplus
}
#endif
#endif /* SOLARIS_B200_H_ */

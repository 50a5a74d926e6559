// Internal declarations shared by the translation units of libsolaris_b200.so.
//
//   gravity.cu      pair-interaction kernel family (K1), source staging, indirect-term reduction   [FMA on]
//   elementwise.cu  per-body finalize incl. gas drag / type-I / type-II terms (K2), RK stage
//                   combinations (K3), solution + error max-norm (K4), event flags (K5), layout
//                   transposes                                                                     [-fmad=false]
//   api.cu          context, drivers (RK4 / RKF78 / RKN76), C-ABI, NCCL plumbing
//
// Device layout ("planes"): a state or derivative array holds 6 planes x,y,z,vx,vy,vz of `ld`
// doubles each (ld = n rounded up to 64), element (c,i) at base[c*ld + i].  Per-body parameters are
// separate arrays.  Source bodies are additionally packed as double4 {x,y,z,m} (`src4`) so a j-tile
// is one contiguous bulk copy into shared memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/solaris_b200.h"

namespace sol {

constexpr double kGauss  = 1.720209895e-2;            // Solaris/Constants.h:28
constexpr double kGauss2 = 2.959122082855911025e-4;   // Solaris/Constants.h:29

// BodyType / MigrationType values (Solaris/Body.h:14-24, 35-39)
enum { T_CENTRAL = 1, T_GIANT = 2, T_ROCKY = 3, T_PROTO = 4, T_SUPERPL = 5, T_PL = 6, T_TEST = 7 };
enum { MIG_NO = 0, MIG_I = 1, MIG_II = 2 };

constexpr int kMaxSplit = 32;      // max j-splits of the pair kernel
constexpr int kTileJ    = 256;     // sources per shared-memory tile (8 KB of double4)
constexpr int kPairThreads = 128;  // threads per CTA of the pair kernel
constexpr int kIndirectBlocks = 64;
constexpr int kSymB      = 512;    // symmetric pair kernel: bodies per block = warps x 32 lanes x sinks per lane (4x4 or 2x8)
constexpr int kSymRounds = 32;     //   rounds (partial-sum slots) per launch
constexpr int kSymMinBodies = 8 * kSymB;    // the symmetric kernel can run from this many self-gravitating bodies (mode 1)
constexpr int kSymAutoBodies = 24 * kSymB;  // ... and is faster than the ordered kernel from about here (mode 2, default):
                                            // measured 4.7e11 vs 7.9e11 pairs/s at 8192, 1.04e12 vs 0.90e12 at 16384

// Everything the gas-term device code needs, precomputed on the host with the reference's own
// expression order (so the constants are bit-identical to the reference's).
struct GasParams {
	int    enabled;
	int    decrease_type;
	double time_scale, t0, t1;
	double inner_edge;
	double eta_c, eta_index;
	double tau_c, tau_index;
	double sh_c, sh_index;
	double rho_c, rho_index;
	double mfp_c, mfp_index;
	double alpha;
	double a_inner;        // density.c * pow(innerEdge, density.index - 4)   GasComponent.cpp:153
	double Cvth;           // sqrt(8 kB / (pi mu mp))                         GasComponent.cpp:240
	double cTp;            // Gauss2 * ProtonMassBoltzman_CMU                 GasComponent.cpp:223
	double mmw;            // meanMolecularWeight
	double pow_m0_pT;      // pow(mass[0], 2*sh_index - 3)  (argument swap)
	double abs_rho_index;  // fabs(density.index)                              Acceleration.cpp:770
};

// The scalars of an attempt that change from launch to launch.  Kernels captured in a CUDA graph (mid-size systems, api.cu)
// read them from device memory, so that ONE graph serves every step; everything else a captured kernel gets by value
// is fixed for a given system and configuration.
struct StepScalars {
	double h, h2;          // trial step, h * h (DormandPrince.cpp:266)
	double ckh[13];        // RKN: c_k * h per evaluation
	double factor[13];     // GasComponent::ReductionFactor at each evaluation's time
};

struct Counts {
	int c, g, r, p, s, l, t;   // central, giant, rocky, proto, superpl, planetesimal, test
	int n;                     // total
	int M;                     // NOfMassive
	__host__ __device__ int nsrc_max() const { return M + s; }
};

struct PairLaunch {
	// sinks [i_lo, i_hi) against sources [j_lo, j_hi); partial sums written for split index
	// blockIdx.y into part[(split*3 + c)*ld + i]
	int i_lo, i_hi, j_lo, j_hi;
	int splits;          // gridDim.y
	int split_offset;    // first partial-sum slot this launch writes (slot 0 may hold the symmetric kernel's sums)
	int chunk;           // sources per split (whole tiles of kTileJ, or any count for a mid-size launch: plan_pairs)
	int sinks_per_thread;
	int track_nn;
	int tie_prefers_larger_j;   // barycentric: descending j + strict '<' == largest j among ties
};

struct SymLaunch {
	int r0, nR;          // bodies [r0, r0+nR) interact among themselves
	int nb;              // blocks of kSymB
	int round_begin, nrounds;
	int track_nn, tie_ge;
	// multi-GPU: a rank's share of the block pairs is a contiguous range of the (round, CTA) sequence, so its first and
	// last round can be partial: CTAs p >= p_first_lo of round `round_first`, p < p_last_hi of round `round_last`
	int round_first, p_first_lo, round_last, p_last_hi;
};

// Control block / result of the device-resident multi-step driver (warp_run_kernel, sol_run)
struct RunCtl {
	int max_steps;
	double time, h_next;
	double millenium_days, length, output, last_save;        // TimeLine fields Simulator::DecisionMaking reads
	double e3, h3;                                           // 1/ejection^3, 1/hitCentrum^3 (Simulator.cpp:626-629)
	int ej_on, hc_on;
	double col_factor;
	long long step_counter;                                  // Counter::succededStep before the first step
	int flush_every;                                         // Constants::CheckForSM, 0 = never
	double tiny;                                             // Constants::SmallestNumber
	double eps;                                              // pow(10, -10.0) of the host libm (the drivers' epsilon)
	double cstage[13];                                       // stage abscissae c_q as the host drivers use them
	int time_dependent_factor;                               // GasComponent LINEAR: ReductionFactor(t) per evaluation
	double *rec;                                             // [max_steps][4] time, hDid, hNext (raw), trial step per step, or null
};
struct RunOut {
	double time, h_next, h_did, last_save, err_max;
	long long step_counter, attempts, evals;
	int steps, stop_reason, err_code;
	int ev[3];
};

struct Ctx {
	int device = 0;
	cudaStream_t stream = nullptr;
	bool own_stream = false;
	std::string err;

	Counts cnt{};
	int ld = 0;               // plane stride
	int barycentric = 0;
	int nn_mode = 2;
	GasParams gas{};
	sol_nebula_pod neb{};
	bool has_nebula = false;
	double mass0 = 0.0;

	// shard (single GPU: [0,n))
	int rank = 0, nranks = 1;
	int lo = 0, hi = 0;
	void *nccl = nullptr;     // ncclComm_t

	// state planes
	double *y0 = nullptr, *y = nullptr, *ytmp = nullptr, *yscale = nullptr;
	double *k[13] = {};
	// parameters
	double *mass = nullptr, *radius = nullptr, *density = nullptr, *cD = nullptr;
	double *gS = nullptr, *gE = nullptr, *migStop = nullptr;
	int *type = nullptr, *migType = nullptr, *id = nullptr;
	// side outputs
	double *rm3 = nullptr, *nnDist = nullptr;
	int *nnIdx = nullptr;
	// gas caches, 3 planes each, stride ld
	double *aGas = nullptr, *aMig1 = nullptr, *aMig2 = nullptr;
	// gravity scratch
	double4 *src4 = nullptr;          // nsrc_max (+pad) sources
	double *part = nullptr;           // [kMaxSplit][3][ld] partial sums
	double *partR2 = nullptr;         // [kMaxSplit][ld] nearest-neighbour r^2 partials
	int *partIdx = nullptr;           // [kMaxSplit][ld]
	// symmetric-kernel round slots: [kSymRounds][3][ld] (+ NN candidates [kSymRounds][ld])
	double *symPI = nullptr, *symPJ = nullptr, *symPIr2 = nullptr, *symPJr2 = nullptr;
	int *symPIidx = nullptr, *symPJidx = nullptr;
	int warp_mode = 1;                // one-warp attempt kernel for <= 32 massive bodies (sol_set_small_system_kernel bit 1)
	int cp_mode = 1;                  // sol_run: component-parallel one-warp kernel for <= 10 massive bodies (0: body-per-lane kernel)
	int *symThr = nullptr;            // [ld] nearest-neighbour filter thresholds (high word of d^2), reset per evaluation
	int sym_mode = 2;                 // 0 never, 1 from kSymMinBodies, 2 from kSymAutoBodies (where it starts to win)
	int tracer_mode = 1;              // 1: few massive bodies + many non-source bodies use the tracer attempt kernel
	double4 *stageSrc = nullptr;      // [13][kSmallMax] per-evaluation source snapshots (tracer path)
	double *stageS6 = nullptr;        // [13][6]
	int small_mode = 1;               // 1: systems of <= kSmallAuto bodies use the whole-attempt kernel, 2: up to kSmallMax, 0: never
	double *indPart = nullptr;        // [kIndirectBlocks][6] indirect-term partials
	double *indirect = nullptr;       // [6]: S over j<M (x,y,z), S over j<M+s (x,y,z)
	unsigned *indCounter = nullptr;
	double *integralsPart = nullptr;  // [kIndirectBlocks][12]
	double *integralsDev = nullptr;   // [12]
	double *integralsHost = nullptr;  // pinned [12]
	// reductions / events
	unsigned long long *errBits = nullptr;   // max-norm accumulator (bit pattern of a non-negative double)
	unsigned long long *errBitsHost = nullptr;   // pinned
	int *evCount = nullptr;           // [4]
	int *evIdx = nullptr;             // [3][ld]
	int *evCountHost = nullptr;       // pinned
	RunOut *runOut = nullptr, *runOutHost = nullptr;   // sol_run result (device / pinned)
	StepScalars *ssDev = nullptr, *ssHost = nullptr;   // per-attempt scalars of the graph path (device / pinned)
	bool capturing = false;           // launches go into a CUDA graph: kernels read h / c_k h / factors from ssDev
	int graph_mode = 1;               // mid-size systems on the general path: 1 = replay of captured CUDA graphs (default), 2 = one
	                                  // cooperative kernel per segment, 0 = every launch issued from the host (sol_set_graph_mode)
	unsigned long long cfg_epoch = 0; // bumped by every call that changes what a captured kernel gets by value
	struct GraphEntry { int integrator, kind; const double *y0; unsigned long long epoch; cudaGraphExec_t exec; int launches; };
	std::vector<GraphEntry> graphs;
	bool stage_in_finalize = true;        // (SOLARIS_B200_STAGE_IN_FINALIZE=0 switches the fusion off, for A/B runs)
	const double *src4_state = nullptr;   // the trial state src4 already mirrors (staged by the previous evaluation's finalize kernel)
	cudaStream_t side = nullptr;          // second capture stream: the indirect-term reduction runs beside the pair kernel
	cudaEvent_t evFork = nullptr, evJoin = nullptr;
	// mid-size systems, graph_mode 2: the launches of a segment as phases of ONE cooperative kernel (fused_attempt_kernel,
	// elementwise.cu).  While `rec` is set the launch_* functions append to the program being recorded instead of launching.
	void *rec = nullptr;
	unsigned *fusedBar = nullptr;     // grid-barrier counter of the fused kernel (zero between launches)
	unsigned long long *fusedTrace = nullptr;   // SOLARIS_B200_FUSED_TRACE=1: phase timestamps of CTA 0
	struct FusedEntry { int integrator, kind; const double *y0; unsigned long long epoch; void *program; int ops; };
	std::vector<FusedEntry> fused;
	double *runRec = nullptr; size_t runRecCap = 0;    // per-step records of sol_run (device)
	// staging for seam B
	double *stage_aos = nullptr;      // 6n doubles, device
	size_t stage_cap = 0;
	void *pin = nullptr;              // pinned host staging for sol_write_phases
	size_t pin_cap = 0;
	int alloc_n = 0;

	long long launches = 0;
	double evals = 0, pairs = 0;

	// profiling
	bool prof = false;
	double prof_ms[6] = {};
	long long prof_n[6] = {};
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	std::vector<cudaEvent_t> ev_pool;      // pairs (begin,end)
	std::vector<int> ev_fam;               // family of each recorded pair
	size_t ev_used = 0;                    // recorded pairs
};

// ---- gravity.cu ----
void launch_prep_sources(Ctx &c, const double *state, int j_lo, int j_hi);
void launch_indirect(Ctx &c);
void launch_prep_indirect(Ctx &c, const double *state);
void launch_pairs(Ctx &c, const double *state, const PairLaunch &pl);
void launch_fp64_peak(Ctx &c, double *out_dev, int iters, int blocks, int threads);
void launch_sym_phase(Ctx &c, const SymLaunch &L, bool first);
void launch_integrals(Ctx &c);
void launch_sym_merge_nn(Ctx &c, int i_lo, int i_hi, int tie_ge);

// ---- elementwise.cu: recording of a fused program (see fused_attempt_kernel) ----
bool fused_recording(const Ctx &c);
void *fused_begin_record(Ctx &c);
int fused_end_record(Ctx &c, void **program_dev_out, int *ops_out);   // SOL_OK / SOL_ERR / 1 = segment cannot be fused
int fused_grid_size(Ctx &c);
int launch_fused(Ctx &c, const void *program_dev);
void fused_rec_pack(Ctx &c, const double *state, int j_lo, int j_hi);
void fused_rec_indirect(Ctx &c);
void fused_rec_pairs(Ctx &c, const double *state, const PairLaunch &pl);
void fused_rec_zero_err(Ctx &c);
void fused_rec_unsupported(Ctx &c, const char *what);

// ---- elementwise.cu ----
struct StageArgs {
	int nterms;
	double coef[9];
	const double *k[9];
};
// The trial state of the NEXT stage, formed by the finalize kernel right after it has stored this evaluation's
// derivative (each body only needs its own k-values): one launch less per stage.  kind 0 = none, 1 = Runge-Kutta
// (rk_stage_kernel's statement), 2 = Runge-Kutta-Nystrom (rkn_stage_kernel's).
struct NextStage {
	int kind;
	int self_term;         // index of the term whose k-array is the derivative this very kernel produces (-1: none): that
	                       // term comes from registers instead of being stored and read back through L2
	StageArgs st;
	const double *y0;
	double *out;
	double h, h2, ckh;
};
struct FinalizeArgs {
	const double *state;   // trial state planes
	double *kout;          // derivative planes
	int q, qnext;          // index of this evaluation / of the stage whose trial state is formed (graph path: StepScalars slots)
	double t;
	unsigned eval_flags;
	int splits_massive;    // partial-sum splits used for sinks < M
	int splits_rest;       // ... for sinks >= M
	int track_nn;
	int write_velocity;    // 0: only the acceleration planes are needed (RKN stages)
	int pack_hi;           // > 0: the finalize kernel also stages the next trial state's sources [0, pack_hi) into src4
	NextStage next;
};
void launch_finalize(Ctx &c, const FinalizeArgs &a);

// out = y0 + h*(sum coef_j * k_j), all six planes of sinks [lo,hi)
// (while c.capturing is set, h comes from c.ssDev instead of the argument - see StepScalars)
void launch_rk_stage(Ctx &c, const double *y0, double h, const StageArgs &s, double *out);
void launch_yscale(Ctx &c, const double *y0, const double *k0, double h, double *yscale);
// RKF78: y = y0 + h*(...), errBits = max |err/yscale| (bit pattern)
void launch_rkf78_final(Ctx &c, const double *y0, double h, double *const *k, const double *yscale, double *y);
// RKN7(6) stage k (1..8): x = x0 + c_k h v0 + h^2 S, v = v0 + h S, S = sum a_kl f_l(accel)
void launch_rkn_stage(Ctx &c, const double *y0, double h, double ck, const StageArgs &s, double *out);
void launch_rkn_final(Ctx &c, const double *y0, double h, const double *b, const double *bd, double *const *f, double *y);
void launch_aos_to_planes(Ctx &c, const double *aos, double *planes, int n);
void launch_planes_to_aos(Ctx &c, const double *planes, double *aos, int n);
void launch_flush_tiny(Ctx &c, double *planes, double threshold);
void launch_pack_phases(Ctx &c, const double *planes, double time, void *out);
void launch_event_records(Ctx &c, const int *table, void *out, int m, double time);
void launch_elements_to_phases(Ctx &c, const double *mu, const double *el, double *out, int *failed, int n);
void launch_compact(Ctx &c, const double *in, double *out, int n_new, int planes, const int *adj, int count);
void launch_compact(Ctx &c, const int *in, int *out, int n_new, const int *adj, int count);
void launch_detect_events(Ctx &c, double e3, double h3, int ej_on, int hc_on, double col_factor);

// Whole-attempt kernel for small systems (n <= kSmallMax): ONE CTA runs every stage of an RK4 / RKF78 /
// RKN7(6) attempt - trial state, source staging, indirect sum, pair loop, finalize, solution and error
// norm - with block barriers instead of ~66 launches.  Arithmetic is statement-for-statement the
// multi-launch path's, so both paths give bit-identical results.
constexpr int kSmallMax = 256;
constexpr int kSmallAuto = 160;          // ... and is faster than the general path (graph replay) up to about here: measured
                                        // 3310 vs 3300 RKF78 steps/s at 150 bodies, 2640 vs 2900 at 200, 2220 vs 2640 at 256
constexpr int kTracerMaxSources = 64;   // tracer path: at most this many massive bodies (their per-evaluation snapshots sit in shared memory)
struct SmallEval {
	int nterms;            // 0: state = y0
	int kidx[9];
	double coef[9];
	int out;               // k index that receives dy/dt
	double factor;         // GasComponent::ReductionFactor at this evaluation's time
	unsigned flags;
	int last;              // last stage of the step (nearest-neighbour outputs in mode 2)
	double ckh;            // RKN: c_k * h
};
struct SmallPlan {
	int integrator;        // SOL_RUNGE_KUTTA4 / SOL_RUNGE_KUTTA_FEHLBERG78 / SOL_DORMAND_PRINCE
	int nevals;
	int first;             // ev[0] is the k0 = f(t, y0) evaluation of the Driver (and yscale is (re)computed)
	int n_active;          // bodies [0, n_active) are integrated by the single-CTA kernel (all, or the massive ones)
	double h;
	double h_first;        // first trial step of this Driver call (RKF78 yscale)
	SmallEval ev[13];
	double b[9], bd[9];    // RKN weights
};
bool warp_run_eligible(const Ctx &c);
void launch_warp_run(Ctx &c, const SmallPlan &plan, const RunCtl &ctl, RunOut *out_dev);
void launch_small_attempt(Ctx &c, const SmallPlan &plan);
void launch_tracer_attempt(Ctx &c, const SmallPlan &plan);
double reduction_factor_host(const sol_nebula_pod &g, double t);
int selftest_fast_paths(Ctx &c, unsigned long long seed, long long samples, unsigned long long *mismatches_out);

// ---- device helpers shared by the pair kernels (gravity.cu) and the small-system kernel (elementwise.cu) ----
#ifdef __CUDACC__
// MUFU.RSQ64H: ~2^-22 relative seed of 1/sqrt(x), refined by the callers
__device__ __forceinline__ double rsqrt_seed(double x)
{
	double y;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
	return y;
}

// r^2 >= 0, so the IEEE bit patterns order like integers: the nearest-neighbour compare runs on the
// integer ALU (2 ISETP) instead of taking a DSETP slot on the saturated FP64 pipe.  NaN (coincident
// bodies) has the largest pattern and never wins, like `rij < rMin` in the reference.
template <bool TIE_GE>
__device__ __forceinline__ bool closer_than(double r2, double r2min)
{
	const long long a = __double_as_longlong(r2), b = __double_as_longlong(r2min);
	return TIE_GE ? (a <= b) : (a < b);
}

// m_j / |d|^3 from d^2: seed + one third-order correction applied to m*y^3 (7 FP64 instructions).
__device__ __forceinline__ double mass_over_r3(double r2, double m)
{
	const double y0 = rsqrt_seed(r2);
	const double c2 = y0 * y0;
	const double e = fma(-r2, c2, 1.0);
	const double my = m * y0;
	const double c3m = c2 * my;
	const double p = fma(1.875, e, 1.5);
	const double pe = p * e;
	return fma(c3m, pe, c3m);
}
#endif

// ---- helpers ----
#define SOL_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { \
	c.err = std::string(#call) + ": " + cudaGetErrorString(e__); return SOL_ERR; } } while (0)

// Per-family device timing with CUDA events on the launching stream.  Events are taken from a pool
// and only RESOLVED in sol_profile_read, so profiling adds no host synchronisation to the timed region.
void prof_begin(Ctx &c, int fam);
void prof_end(Ctx &c, int fam);
struct ProfScope {
	Ctx &c; int fam;
	ProfScope(Ctx &ctx, int family) : c(ctx), fam(family) { if (c.prof) prof_begin(c, fam); }
	~ProfScope() { c.prof_n[fam]++; if (c.prof) prof_end(c, fam); }
};

}  // namespace sol

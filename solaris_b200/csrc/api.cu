// Context, force-evaluation orchestration, the three integrator drivers and the C-ABI
// (include/solaris_b200.h).  Host-side control flow follows the reference drivers statement by
// statement (RungeKutta4.cpp:20-56, RungeKuttaFehlberg78.cpp:66-140, DormandPrince.cpp:126-170); the
// step-size formulas (pow) stay on the host so they use the same libm as the reference, fed by ONE
// 8-byte read-back per attempt.
#include <dlfcn.h>
#include <fcntl.h>
#include <unistd.h>
#include <math.h>
#include <nccl.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

using namespace sol;

struct MultiCtx;
struct sol_ctx {
	Ctx c;
	MultiCtx *multi = nullptr;   // sol_create_multi: this handle fans every call out to one sol_ctx per GPU
};

// ---------------------------------------------------------------------------------------------
// Single-process multi-GPU (sol_create_multi): one sol_ctx and one host WORKER THREAD per device, joined into one NCCL
// communicator.  Every entry point called on the front handle runs the same single-rank code on all workers at once
// (the per-rank code contains blocking collectives and stream synchronisations, so each rank needs its own thread) and
// returns when all have finished.  The ranks live in one address space: results that are sharded over the ranks
// (states, side outputs) are written by each rank straight into its slice of the caller's host array.
// ---------------------------------------------------------------------------------------------
struct MultiCtx {
	std::vector<sol_ctx *> ranks;
	std::vector<std::thread> threads;
	std::mutex m;
	std::condition_variable cv_job, cv_done;
	std::function<int(sol_ctx *, int)> job;
	unsigned long long epoch = 0;
	int pending = 0;
	bool quit = false;
	std::vector<int> rc;
};

static void multi_worker(MultiCtx *M, int rank)
{
	unsigned long long seen = 0;
	for (;;) {
		std::function<int(sol_ctx *, int)> job;
		{
			std::unique_lock<std::mutex> lk(M->m);
			M->cv_job.wait(lk, [&] { return M->quit || M->epoch != seen; });
			if (M->quit) return;
			seen = M->epoch;
			job = M->job;
		}
		const int r = job(M->ranks[rank], rank);
		{
			std::lock_guard<std::mutex> lk(M->m);
			M->rc[rank] = r;
			if (--M->pending == 0) M->cv_done.notify_all();
		}
	}
}

// runs f(rank handle, rank) on every worker; SOL_ERR (and the first failing rank's message) if any rank failed
static int fan_out(sol_ctx *h, const std::function<int(sol_ctx *, int)> &f)
{
	MultiCtx *M = h->multi;
	{
		std::unique_lock<std::mutex> lk(M->m);
		M->job = f;
		M->pending = (int)M->ranks.size();
		M->epoch++;
		M->cv_job.notify_all();
		M->cv_done.wait(lk, [&] { return M->pending == 0; });
	}
	for (size_t r = 0; r < M->ranks.size(); r++)
		if (M->rc[r] != SOL_OK) { h->c.err = "rank " + std::to_string(r) + ": " + M->ranks[r]->c.err; return SOL_ERR; }
	return SOL_OK;
}
#define SOL_FANOUT(h, expr) if ((h)->multi) return fan_out(h, [&](sol_ctx *r, int rank) -> int { (void)rank; return (expr); })

static std::string g_create_error;

// ---------------------------------------------------------------------------------------------
// NCCL through dlopen: single-GPU users never need libnccl.
// ---------------------------------------------------------------------------------------------
namespace {
struct NcclApi {
	void *lib = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*ReduceScatter)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
	bool load(std::string &err)
	{
		if (lib) return true;
		const char *names[] = {"libnccl.so.2", "libnccl.so"};
		for (const char *nm : names) { lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
		if (!lib) { err = std::string("cannot load libnccl: ") + dlerror(); return false; }
#define LOAD(sym) sym = (decltype(sym))dlsym(lib, "nccl" #sym); if (!sym) { err = "libnccl lacks nccl" #sym; return false; }
		LOAD(GetUniqueId) LOAD(CommInitRank) LOAD(CommDestroy) LOAD(AllReduce) LOAD(Broadcast) LOAD(AllGather) LOAD(ReduceScatter) LOAD(GroupStart) LOAD(GroupEnd)
		LOAD(GetErrorString)
#undef LOAD
		return true;
	}
};
NcclApi g_nccl;
}  // namespace

#define SOL_NCCL(call) do { ncclResult_t r__ = (call); if (r__ != ncclSuccess) { \
	c.err = std::string(#call) + ": " + g_nccl.GetErrorString(r__); return SOL_ERR; } } while (0)

// ---------------------------------------------------------------------------------------------
// profiling events (see ProfScope in common.cuh)
// ---------------------------------------------------------------------------------------------
namespace sol {
static void prof_resolve(Ctx &c)
{
	if (c.ev_used == 0) return;
	cudaStreamSynchronize(c.stream);
	for (size_t q = 0; q < c.ev_used; q++) {
		float ms = 0.f;
		if (cudaEventElapsedTime(&ms, c.ev_pool[2 * q], c.ev_pool[2 * q + 1]) == cudaSuccess) c.prof_ms[c.ev_fam[q]] += ms;
	}
	c.ev_used = 0;
}
void prof_begin(Ctx &c, int fam)
{
	if (c.ev_used >= 8192) prof_resolve(c);
	if (c.ev_pool.size() < 2 * (c.ev_used + 1)) {
		cudaEvent_t a, b;
		cudaEventCreate(&a); cudaEventCreate(&b);
		c.ev_pool.push_back(a); c.ev_pool.push_back(b);
		c.ev_fam.push_back(fam);
	}
	c.ev_fam[c.ev_used] = fam;
	cudaEventRecord(c.ev_pool[2 * c.ev_used], c.stream);
}
void prof_end(Ctx &c, int)
{
	cudaEventRecord(c.ev_pool[2 * c.ev_used + 1], c.stream);
	c.ev_used++;
}
}  // namespace sol

// ---------------------------------------------------------------------------------------------
// memory
// ---------------------------------------------------------------------------------------------
static void free_bodies(Ctx &c)
{
	auto F = [](auto *&p) { if (p) { cudaFree(p); p = nullptr; } };
	F(c.y0); F(c.y); F(c.ytmp); F(c.yscale);
	for (auto &k : c.k) F(k);
	F(c.mass); F(c.radius); F(c.density); F(c.cD); F(c.gS); F(c.gE); F(c.migStop);
	F(c.type); F(c.migType); F(c.id);
	F(c.rm3); F(c.nnDist); F(c.nnIdx);
	F(c.aGas); F(c.aMig1); F(c.aMig2);
	F(c.src4); F(c.part); F(c.partR2); F(c.partIdx);
	F(c.symPI); F(c.symPJ); F(c.symPIr2); F(c.symPJr2); F(c.symPIidx); F(c.symPJidx); F(c.symThr);
	F(c.evIdx);
	F(c.stage_aos); c.stage_cap = 0;
	c.alloc_n = 0;
}

template <typename T>
static int dalloc(Ctx &c, T *&p, size_t count)
{
	SOL_CUDA(cudaMalloc((void **)&p, std::max<size_t>(count, 1) * sizeof(T)));
	SOL_CUDA(cudaMemsetAsync(p, 0, std::max<size_t>(count, 1) * sizeof(T), c.stream));
	return SOL_OK;
}

static int alloc_sym(Ctx &c)
{
	if (c.symPI) return SOL_OK;
	size_t ld = (size_t)c.ld;
	if (dalloc(c, c.symPI, (size_t)kSymRounds * 3 * ld) != SOL_OK) return SOL_ERR;
	if (dalloc(c, c.symPJ, (size_t)kSymRounds * 3 * ld) != SOL_OK) return SOL_ERR;
	if (dalloc(c, c.symPIr2, (size_t)kSymRounds * ld) != SOL_OK) return SOL_ERR;
	if (dalloc(c, c.symPJr2, (size_t)kSymRounds * ld) != SOL_OK) return SOL_ERR;
	if (dalloc(c, c.symPIidx, (size_t)kSymRounds * ld) != SOL_OK) return SOL_ERR;
	if (dalloc(c, c.symPJidx, (size_t)kSymRounds * ld) != SOL_OK) return SOL_ERR;
	if (dalloc(c, c.symThr, ld) != SOL_OK) return SOL_ERR;
	return SOL_OK;
}

static void shard_of(int n, int nranks, int r, int &lo, int &hi);

static int alloc_bodies(Ctx &c, int n)
{
	free_bodies(c);
	c.ld = (n + 63) / 64 * 64;
	if (c.nranks > 1) {
		// the reduce-scatter of the symmetric kernel's partial sums treats a plane as [rank][chunk]: the plane stride has
		// to cover nranks equal chunks (the rows past n are padding that stays zero)
		int lo, hi;
		shard_of(n, c.nranks, 0, lo, hi);
		const int chunk = hi - lo;
		c.ld = std::max(c.ld, (c.nranks * chunk + 63) / 64 * 64);
	}
	size_t ld = (size_t)c.ld;
#define A(p, cnt) if (dalloc(c, p, (cnt)) != SOL_OK) return SOL_ERR;
	A(c.y0, 6 * ld) A(c.y, 6 * ld) A(c.ytmp, 6 * ld) A(c.yscale, 6 * ld)
	for (auto &k : c.k) A(k, 6 * ld)
	A(c.mass, ld) A(c.radius, ld) A(c.density, ld) A(c.cD, ld) A(c.gS, ld) A(c.gE, ld) A(c.migStop, ld)
	A(c.type, ld) A(c.migType, ld) A(c.id, ld)
	A(c.rm3, ld) A(c.nnDist, ld) A(c.nnIdx, ld)
	A(c.aGas, 3 * ld) A(c.aMig1, 3 * ld) A(c.aMig2, 3 * ld)
	A(c.src4, ld + kTileJ)
	A(c.part, (size_t)kMaxSplit * 3 * ld) A(c.partR2, (size_t)kMaxSplit * ld) A(c.partIdx, (size_t)kMaxSplit * ld)
	A(c.evIdx, 3 * ld)
#undef A
	c.alloc_n = n;
	return SOL_OK;
}

static int ensure_stage(Ctx &c, size_t doubles)
{
	if (c.stage_cap >= doubles) return SOL_OK;
	if (c.stage_aos) cudaFree(c.stage_aos);
	c.stage_aos = nullptr; c.stage_cap = 0;
	SOL_CUDA(cudaMalloc((void **)&c.stage_aos, doubles * sizeof(double)));
	c.stage_cap = doubles;
	return SOL_OK;
}

// ---------------------------------------------------------------------------------------------
// gas constants, with the reference's expression order (Constants.h:87-90, GasComponent.cpp)
// ---------------------------------------------------------------------------------------------
static void refresh_gas(Ctx &c)
{
	GasParams &g = c.gas;
	memset(&g, 0, sizeof(g));
	if (!c.has_nebula) return;
	const sol_nebula_pod &p = c.neb;
	const double Pi = 3.14159265358979323846;
	const double Boltzman_SI = 1.3806488e-23, ProtonMass_SI = 1.672621777e-27;
	const double SolarToKilogram = 1.98911e30, AuToMeter = 1.495978707e11, DayToSecond = 86400.0;
	const double KilogramToSolar = 1.0 / SolarToKilogram, MeterToAu = 1.0 / AuToMeter, SecondToDay = 1.0 / DayToSecond;
	const double Boltzman_CMU = Boltzman_SI * (KilogramToSolar * (MeterToAu * MeterToAu)) / (SecondToDay * SecondToDay);
	const double ProtonMass_CMU = ProtonMass_SI * KilogramToSolar;
	const double BoltzmanProtonMass_CMU = Boltzman_CMU / ProtonMass_CMU;
	const double ProtonMassBoltzman_CMU = 1.0 / BoltzmanProtonMass_CMU;
	g.enabled = 1;
	g.decrease_type = p.decrease_type;
	g.time_scale = p.time_scale; g.t0 = p.t0; g.t1 = p.t1;
	g.inner_edge = p.inner_edge;
	g.eta_c = p.eta_c; g.eta_index = p.eta_index;
	g.tau_c = p.tau_c; g.tau_index = p.tau_index;
	g.sh_c = p.scale_height_c; g.sh_index = p.scale_height_index;
	g.rho_c = p.density_c; g.rho_index = p.density_index;
	g.mfp_c = p.mean_free_path_c; g.mfp_index = p.mean_free_path_index;
	g.alpha = p.alpha;
	g.a_inner = p.density_c * pow(p.inner_edge, p.density_index - 4.0);
	g.Cvth = sqrt((8.0 * Boltzman_CMU) / (Pi * p.mean_molecular_weight * ProtonMass_CMU));
	g.cTp = kGauss2 * ProtonMassBoltzman_CMU;
	g.mmw = p.mean_molecular_weight;
	g.pow_m0_pT = pow(c.mass0, 2.0 * p.scale_height_index - 3.0);
	g.abs_rho_index = fabs(p.density_index);
}

// ---------------------------------------------------------------------------------------------
// one force evaluation:  kout = f(t, state)           (Acceleration::Compute)
// ---------------------------------------------------------------------------------------------
// ni_all: the sinks of the same launch on an unsharded context (a sharded one passes its own ni and the global count, so
// that a mid-size system is cut into the same source chunks - and sums in the same order - on any number of GPUs)
static void plan_pairs(int ni, int nj, PairLaunch &pl, int max_splits = kMaxSplit, int ni_all = 0)
{
	// sinks per thread: amortise the shared-memory tile reads once there are enough sinks to fill
	// the chip (148 SMs x >= 4 CTAs of 128 threads)
	int I = 1;
	if (ni >= 148 * 4 * kPairThreads * 4) I = 4;
	else if (ni >= 148 * 4 * kPairThreads * 2) I = 2;
	int iblocks = (ni + kPairThreads * I - 1) / (kPairThreads * I);
	int tiles = (nj + kTileJ - 1) / kTileJ;
	// aim at >= ~16 CTAs per SM in total so the tail wave is small, at least 2 tiles per CTA (so that the second
	// tile's bulk copy overlaps the first tile's pairs) - unless the whole launch is so small that it would not even
	// give every SM a few CTAs: then one tile per CTA, the launch is latency-bound anyway
	const int min_tiles_per_cta = ((long long)iblocks * tiles >= 148 * 8) ? 2 : 1;
	int want = (148 * 16 + iblocks - 1) / iblocks;
	int splits = std::max(1, std::min({want, max_splits, std::max(1, tiles / min_tiles_per_cta)}));
	int chunk_tiles = (tiles + splits - 1) / splits;
	splits = (tiles + chunk_tiles - 1) / chunk_tiles;
	pl.sinks_per_thread = I;
	pl.splits = std::max(1, splits);
	pl.chunk = chunk_tiles * kTileJ;
	// A few thousand sinks against a few thousand sources: with whole tiles the launch has about one CTA - four warps - per
	// SM and every warp walks 256 sources alone, its dependent FP64 chain unhidden (measured 22 us per launch at N = 2000,
	// five times the pair work).  Such a launch is cut into source chunks so that its CTAs are ONE wave of the cooperative
	// kernel that runs mid-size attempts (two CTAs per SM, a few left for the indirect-term reduction that shares the
	// phase): every CTA then walks the shortest chain the GPU allows, and none waits for a second round.  At least 32
	// sources per chunk - finalize adds the chunks' partial sums one by one.  Only from two tiles on: a single tile
	// keeps the summation order of the single-CTA kernel, which is asserted to be bit-identical.
	const int iblocks_all = ni_all > ni ? (ni_all + kPairThreads * I - 1) / (kPairThreads * I) : iblocks;
	if (tiles >= 2 && (long long)iblocks_all * tiles < 148 * 8) {
		const int wave = 148 * 2 - 8;
		int s = std::max(1, std::min({wave / iblocks_all, max_splits, nj / 32}));
		const int chunk = (nj + s - 1) / s;
		pl.chunk = chunk;
		pl.splits = (nj + chunk - 1) / chunk;
	}
}

static double pairs_per_eval(const Ctx &c)
{
	const Counts &n = c.cnt;
	if (c.barycentric) return (double)n.n * n.M - n.M;
	double p = 0;
	if (n.M >= 1) p += (double)(n.M - 1) * std::max(0, n.M + n.s - 2);
	p += (double)(n.n - n.M) * std::max(0, n.M - 1);
	return p;
}

static int exchange_sources(Ctx &c, int src_hi);

// Share of rank `rank` in the symmetric kernel's work: the CTAs of all rounds form one sequence (round 0 = the nb diagonal
// blocks, cheaper: ordered evaluation with the self pair masked, no j-side sums; then nb CTAs per round, nb / 2 in the half
// round of an even block count) that is cut into nranks pieces of equal COST.  Dealing whole rounds left the ranks up to one
// round apart - 0.8 % at N = 10^6 on 8 GPUs, which every rank then waits for at the exchange.
// out: round_first, p_first_lo, round_last, p_last_hi (see SymLaunch); round_last < round_first = no work.
static void sym_work_of_rank(int nb, int nranks, int rank, int out[4])
{
	const int rounds_total = nb / 2 + 1;
	const double kDiag = 0.75;                                          // cost of a diagonal CTA relative to a full block pair
	auto ctas = [&](int r) { return (2 * r == nb) ? nb / 2 : nb; };
	auto cost = [&](int r) { return r == 0 ? kDiag : 1.0; };
	double total = 0.0;
	for (int r = 0; r < rounds_total; r++) total += cost(r) * ctas(r);
	// position (round, p) of the boundary at cumulative cost w: the first CTA whose start is >= w
	auto boundary = [&](double w, int &br, int &bp) {
		double acc = 0.0;
		for (int r = 0; r < rounds_total; r++) {
			const double rc = cost(r) * ctas(r);
			if (w < acc + rc) { br = r; bp = std::min(ctas(r), (int)ceil((w - acc) / cost(r) - 1e-9)); return; }
			acc += rc;
		}
		br = rounds_total; bp = 0;
	};
	int r0, p0, r1, p1;
	if (rank == 0) { r0 = 0; p0 = 0; } else boundary(total * rank / nranks, r0, p0);
	if (rank == nranks - 1) { r1 = rounds_total; p1 = 0; } else boundary(total * (rank + 1) / nranks, r1, p1);
	// [ (r0, p0), (r1, p1) ) as first / last round with CTA bounds
	if (r0 < rounds_total && p0 >= ctas(r0)) { r0++; p0 = 0; }
	int rl = r1, pl = p1;
	if (pl == 0) { rl = r1 - 1; pl = rl >= 0 ? nb : 0; }               // ends exactly at a round boundary: the whole previous round
	out[0] = r0; out[1] = p0; out[2] = rl; out[3] = pl;
	if (rl < r0 || (rl == r0 && pl <= p0)) { out[0] = 0; out[1] = 0; out[2] = -1; out[3] = 0; }
}

static int eval_force(Ctx &c, const double *state, double *kout, double t, unsigned flags, bool last_stage, bool write_velocity,
                      const NextStage *next = nullptr, int q = 0)
{
	const Counts &n = c.cnt;
	const bool bary = c.barycentric != 0;
	const bool track = c.nn_mode == 1 || (c.nn_mode == 2 && last_stage);
	const int jlo = bary ? 0 : 1;
	const int nsrcA = bary ? n.M : n.M + n.s;   // sources seen by sinks < M   (Acceleration.cpp:285-289)
	const int nsrcB = n.M;                      // sources seen by the rest
	const int src_hi = std::max(nsrcA, nsrcB);

	// The previous evaluation's finalize kernel may already have staged this trial state's sources (FinalizeArgs::pack_hi).
	const bool staged = c.nranks == 1 && c.src4_state == state;
	c.src4_state = nullptr;
	bool joined = true;
	if (staged) {
		if (!bary && src_hi > 1) {
			// only the indirect sums are left - beside the pair kernel when the launches go into a CUDA graph (a second
			// capture stream forks here and joins before finalize: the graph gets two parallel branches)
			const bool fork = c.capturing && !fused_recording(c) && c.side != nullptr;
			if (fork) {
				SOL_CUDA(cudaEventRecord(c.evFork, c.stream));
				SOL_CUDA(cudaStreamWaitEvent(c.side, c.evFork, 0));
				cudaStream_t main_stream = c.stream;
				c.stream = c.side;
				launch_indirect(c);
				c.stream = main_stream;
				SOL_CUDA(cudaEventRecord(c.evJoin, c.side));
				joined = false;
			} else {
				launch_indirect(c);
			}
		} else if (!bary) {
			SOL_CUDA(cudaMemsetAsync(c.indirect, 0, 6 * sizeof(double), c.stream));
		}
	} else if (c.nranks == 1 && !bary && src_hi > 1 && src_hi == n.M + n.s) {
		launch_prep_indirect(c, state);          // staging + astrocentric indirect sums in one launch
	} else {
		launch_prep_sources(c, state, std::max(c.lo, 0), std::min(c.hi, src_hi));
		if (c.nranks > 1 && exchange_sources(c, src_hi) != SOL_OK) return SOL_ERR;
		if (!bary && src_hi > 1) launch_indirect(c);
		// no source besides the star (e.g. after the last planet was removed): finalize_sink still subtracts the
		// indirect sums, so they must not keep the previous evaluation's values
		else if (!bary) SOL_CUDA(cudaMemsetAsync(c.indirect, 0, 6 * sizeof(double), c.stream));
	}

	FinalizeArgs fa{};
	fa.state = state; fa.kout = kout; fa.t = t; fa.eval_flags = flags;
	fa.q = q; fa.qnext = q + 1;
	fa.track_nn = track ? 1 : 0; fa.write_velocity = write_velocity ? 1 : 0;
	if (next) {
		fa.next = *next;
		fa.next.self_term = -1;
		// this evaluation's own derivative, when the next stage uses it, is the last term of that stage's sum in every
		// tableau here (a_{s+1,s} k_s): the finalize kernel then adds it from registers
		const int last = fa.next.st.nterms - 1;
		if (last >= 0 && fa.next.st.k[last] == kout) fa.next.self_term = last;
	}

	PairLaunch pl{};
	pl.track_nn = track ? 1 : 0;
	pl.tie_prefers_larger_j = bary ? 1 : 0;
	const int sink_lo_all = bary ? 0 : 1;
	const int sink_lo = std::max(c.lo, sink_lo_all);
	// (i_lo_all, i_hi_all: the same sink range on an unsharded context)
	auto ordered = [&](int i_lo, int i_hi, int j_lo, int j_hi, int split_offset, int i_lo_all, int i_hi_all) -> int {
		pl.i_lo = i_lo; pl.i_hi = i_hi; pl.j_lo = j_lo; pl.j_hi = j_hi; pl.split_offset = split_offset;
		if (pl.i_hi <= pl.i_lo || pl.j_hi <= pl.j_lo) return 0;
		plan_pairs(pl.i_hi - pl.i_lo, pl.j_hi - pl.j_lo, pl, kMaxSplit - split_offset, i_hi_all - i_lo_all);
		launch_pairs(c, state, pl);
		return pl.splits;
	};
	// The square block "massive sinks x massive sources" goes to the symmetric kernel (each unordered
	// pair once) when it is large enough and this rank owns all of it.
	const int sq_lo = jlo, sq_hi = n.M;
	const bool use_sym = c.sym_mode != 0 && (sq_hi - sq_lo) >= (c.sym_mode == 1 ? kSymMinBodies : kSymAutoBodies);
	if (use_sym) {
		if (alloc_sym(c) != SOL_OK) return SOL_ERR;
		SymLaunch L{};
		L.r0 = sq_lo; L.nR = sq_hi - sq_lo; L.nb = (L.nR + kSymB - 1) / kSymB;
		L.track_nn = track ? 1 : 0; L.tie_ge = bary ? 1 : 0;
		// multi-GPU: the CTAs of all rounds are dealt to the ranks in pieces of equal cost (sym_work_of_rank; every round
		// touches every block once as i-block and once as j-block, so each rank produces partial sums for all bodies);
		// the partial sums are then combined over NVLink.  Single GPU: all rounds, no collective.
		int share[4];
		sym_work_of_rank(L.nb, c.nranks, c.rank, share);
		L.round_first = share[0]; L.p_first_lo = share[1]; L.round_last = share[2]; L.p_last_hi = share[3];
		const int r_lo = share[0], r_hi = share[2] + 1;
		if (r_hi <= r_lo) {
			SOL_CUDA(cudaMemsetAsync(c.part, 0, 3 * (size_t)c.ld * sizeof(double), c.stream));
			if (track) {
				SOL_CUDA(cudaMemsetAsync(c.partIdx, 0xff, (size_t)c.ld * sizeof(int), c.stream));
				SOL_CUDA(cudaMemsetAsync(c.partR2, 0, (size_t)c.ld * sizeof(double), c.stream));
			}
		}
		// nearest-neighbour filter thresholds start just above the reference's cutoff rMin^2 = 1e20 (high word 0x4415af1d;
		// 0x44444444 ~ 7.5e20 is the closest byte pattern above it): farther candidates can never be the neighbour
		if (track) SOL_CUDA(cudaMemsetAsync(c.symThr, 0x44, (size_t)c.ld * sizeof(int), c.stream));
		for (int rb = r_lo; rb < r_hi; rb += kSymRounds) {
			L.round_begin = rb; L.nrounds = std::min(kSymRounds, r_hi - rb);
			launch_sym_phase(c, L, rb == r_lo);
		}
		if (c.nranks > 1) {
			// every rank holds partial sums for ALL bodies (its rounds touch every block) but finalizes only its own sinks:
			// a reduce-scatter per plane, in place (recv = send + rank * chunk), moves half the bytes of an all-reduce
			int lo0, hi0;
			shard_of(n.n, c.nranks, 0, lo0, hi0);
			const size_t chunk = (size_t)(hi0 - lo0);
			{
				ProfScope ps(c, 5);   // (scopes do not nest)
				SOL_NCCL(g_nccl.GroupStart());
				for (int pl3 = 0; pl3 < 3; pl3++) {
					double *plane = c.part + (size_t)pl3 * c.ld;
					SOL_NCCL(g_nccl.ReduceScatter(plane, plane + (size_t)c.rank * chunk, chunk, ncclDouble, ncclSum, (ncclComm_t)c.nccl, c.stream));
				}
				SOL_NCCL(g_nccl.GroupEnd());
			}
			if (track) {
				{
					ProfScope ps(c, 5);
					SOL_NCCL(g_nccl.AllGather(c.partR2, c.symPIr2, (size_t)c.ld, ncclDouble, (ncclComm_t)c.nccl, c.stream));
					SOL_NCCL(g_nccl.AllGather(c.partIdx, c.symPIidx, (size_t)c.ld, ncclInt, (ncclComm_t)c.nccl, c.stream));
				}
				launch_sym_merge_nn(c, std::max(c.lo, sq_lo), std::min(c.hi, sq_hi), bary ? 1 : 0);
			}
		}
		// massive sinks also see the super-planetesimals (astrocentric, Acceleration.cpp:285-289)
		fa.splits_massive = 1 + ordered(sink_lo, std::min(c.hi, n.M), n.M, nsrcA, 1, sink_lo_all, n.M);
		fa.splits_rest = ordered(std::max(c.lo, n.M), c.hi, jlo, nsrcB, 0, n.M, n.n);
	} else if (nsrcA == nsrcB) {
		fa.splits_massive = fa.splits_rest = ordered(sink_lo, c.hi, jlo, nsrcA, 0, sink_lo_all, n.n);
	} else {
		fa.splits_massive = ordered(sink_lo, std::min(c.hi, n.M), jlo, nsrcA, 0, sink_lo_all, n.M);
		fa.splits_rest = ordered(std::max(c.lo, n.M), c.hi, jlo, nsrcB, 0, n.M, n.n);
	}
	// the next stage's sources are staged by this finalize kernel - unless it reads other bodies' entries of src4 itself
	// (nearest-neighbour distance) or the sources are exchanged between ranks first
	if (next != nullptr && c.nranks == 1 && !track && c.stage_in_finalize) {
		fa.pack_hi = src_hi;
		c.src4_state = next->out;
	}
	if (!joined) SOL_CUDA(cudaStreamWaitEvent(c.stream, c.evJoin, 0));
	launch_finalize(c, fa);
	c.evals += 1;
	c.pairs += pairs_per_eval(c);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { c.err = std::string("kernel launch: ") + cudaGetErrorString(e); return SOL_ERR; }
	return SOL_OK;
}

// ---------------------------------------------------------------------------------------------
// multi-GPU plumbing: sinks are sharded contiguously, every rank needs every source's {x,y,z,m}.
// Each rank broadcasts the slice of src4 it owns (an all-gather with ragged counts).
// ---------------------------------------------------------------------------------------------
static void shard_of(int n, int nranks, int r, int &lo, int &hi)
{
	int chunk = ((n + nranks - 1) / nranks + 31) / 32 * 32;
	lo = std::min(n, r * chunk);
	hi = std::min(n, lo + chunk);
}

static int exchange_sources(Ctx &c, int src_hi)
{
	ProfScope ps(c, 5);   // family 5 on a sharded context: the collectives (their device time includes the wait for the slowest rank)
	SOL_NCCL(g_nccl.GroupStart());
	for (int r = 0; r < c.nranks; r++) {
		int lo, hi;
		shard_of(c.cnt.n, c.nranks, r, lo, hi);
		hi = std::min(hi, src_hi);
		if (hi <= lo) continue;
		SOL_NCCL(g_nccl.Broadcast(c.src4 + lo, c.src4 + lo, (size_t)(hi - lo) * 4, ncclDouble, r, (ncclComm_t)c.nccl, c.stream));
	}
	SOL_NCCL(g_nccl.GroupEnd());
	return SOL_OK;
}

static int read_error_max(Ctx &c, double &out)
{
	if (c.nranks > 1)
		SOL_NCCL(g_nccl.AllReduce(c.errBits, c.errBits, 1, ncclUint64, ncclMax, (ncclComm_t)c.nccl, c.stream));
	SOL_CUDA(cudaMemcpyAsync(c.errBitsHost, c.errBits, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c.stream));
	SOL_CUDA(cudaStreamSynchronize(c.stream));
	long long bits = (long long)*c.errBitsHost;
	memcpy(&out, &bits, sizeof(double));
	return SOL_OK;
}

// ---------------------------------------------------------------------------------------------
// drivers
// ---------------------------------------------------------------------------------------------
namespace {
// Fehlberg 7(8) coupling coefficients a_sj as (stage, j, value), in the summation order of
// RungeKuttaFehlberg78.cpp:170-232 (values :41-56).
struct Term { int j; double a; };
const std::vector<std::vector<Term>> &rkf78_tableau()
{
	static const std::vector<std::vector<Term>> T = {
	    {},
	    {{0, 2.0 / 27.0}},
	    {{0, 1.0 / 36.0}, {1, 1.0 / 12.0}},
	    {{0, 1.0 / 24.0}, {2, 1.0 / 8.0}},
	    {{0, 5.0 / 12.0}, {2, -25.0 / 16.0}, {3, 25.0 / 16.0}},
	    {{0, 1.0 / 20.0}, {3, 1.0 / 4.0}, {4, 1.0 / 5.0}},
	    {{0, -25.0 / 108.0}, {3, 125.0 / 108.0}, {4, -65.0 / 27.0}, {5, 125.0 / 54.0}},
	    {{0, 31.0 / 300.0}, {4, 61.0 / 225.0}, {5, -2.0 / 9.0}, {6, 13.0 / 900.0}},
	    {{0, 2.0}, {3, -53.0 / 6.0}, {4, 704.0 / 45.0}, {5, -107.0 / 9.0}, {6, 67.0 / 90.0}, {7, 3.0}},
	    {{0, -91.0 / 108.0}, {3, 23.0 / 108.0}, {4, -976.0 / 135.0}, {5, 311.0 / 54.0}, {6, -19.0 / 60.0}, {7, 17.0 / 6.0}, {8, -1.0 / 12.0}},
	    {{0, 2383.0 / 4100.0}, {3, -341.0 / 164.0}, {4, 4496.0 / 1025.0}, {5, -301.0 / 82.0}, {6, 2133.0 / 4100.0}, {7, 45.0 / 82.0}, {8, 45.0 / 164.0}, {9, 18.0 / 41.0}},
	    {{0, 3.0 / 205.0}, {5, -6.0 / 41.0}, {6, -3.0 / 205.0}, {7, -3.0 / 41.0}, {8, 3.0 / 41.0}, {9, 6.0 / 41.0}},
	    {{0, -1777.0 / 4100.0}, {3, -341.0 / 164.0}, {4, 4496.0 / 1025.0}, {5, -289.0 / 82.0}, {6, 2193.0 / 4100.0}, {7, 51.0 / 82.0}, {8, 33.0 / 164.0}, {9, 12.0 / 41.0}, {11, 1.0}},
	};
	return T;
}

// Dormand-Prince RKN7(6) coefficients, DormandPrince.cpp:37-123, summation order of Step2 :274-409.
struct RknTableau {
	double b[9] = {}, bd[9] = {}, c[9] = {};
	std::vector<std::vector<Term>> a;
	RknTableau()
	{
		const double sQ = sqrt(21.0);
		b[0] = 1.0 / 20.0; b[4] = 8.0 / 45.0; b[5] = 7.0 * (7.0 + sQ) / 360.0; b[6] = 7.0 * (7.0 - sQ) / 360.0;
		b[7] = -1.0 / 20.0; b[8] = 1.0 / 20.0;
		bd[0] = 1.0 / 20.0; bd[4] = 16.0 / 45.0; bd[5] = 49.0 / 180.0; bd[6] = 49.0 / 180.0; bd[7] = 1.0 / 20.0;
		c[1] = 1.0 / 10.0; c[2] = 1.0 / 5.0; c[3] = 3.0 / 8.0; c[4] = 1.0 / 2.0;
		c[5] = (7.0 - sQ) / 14.0; c[6] = (7.0 + sQ) / 14.0; c[7] = 1.0; c[8] = 1.0;
		a = {
		    {},
		    {{0, 1.0 / 200.0}},
		    {{0, 1.0 / 150.0}, {1, 1.0 / 75.0}},
		    {{0, 171.0 / 8192.0}, {1, 45.0 / 4096.0}, {2, 315.0 / 8192.0}},
		    {{0, 5.0 / 288.0}, {1, 25.0 / 528.0}, {2, 25.0 / 672.0}, {3, 16.0 / 693.0}},
		    {{0, (1003.0 - 205.0 * sQ) / 12348.0}, {1, -25.0 * (751.0 - 173.0 * sQ) / 90552.0}, {2, 25.0 * (624.0 - 137.0 * sQ) / 43218.0},
		     {3, -128.0 * (361.0 - 79.0 * sQ) / 237699.0}, {4, (3411.0 - 745.0 * sQ) / 24696.0}},
		    {{0, (793.0 + 187.0 * sQ) / 12348.0}, {1, -25.0 * (331.0 + 113.0 * sQ) / 90552.0}, {2, 25.0 * (1044.0 + 247.0 * sQ) / 43218.0},
		     {3, -128.0 * (14885.0 + 3779.0 * sQ) / 9745659.0}, {4, (3327.0 + 797.0 * sQ) / 24696.0}, {5, -(581.0 + 127.0 * sQ) / 1722.0}},
		    {{0, -(157.0 - 3.0 * sQ) / 378.0}, {1, 25.0 * (143.0 - 10.0 * sQ) / 2772.0}, {2, -25.0 * (876.0 + 55.0 * sQ) / 3969.0},
		     {3, 1280.0 * (913.0 + 18.0 * sQ) / 596673.0}, {4, -(1353.0 + 26.0 * sQ) / 2268.0}, {5, 7.0 * (1777.0 + 377.0 * sQ) / 4428.0},
		     {6, 7.0 * (5.0 - sQ) / 36.0}},
		    {{0, 1.0 / 20.0}, {4, 8.0 / 45.0}, {5, 7.0 * (7.0 + sQ) / 360.0}, {6, 7.0 * (7.0 - sQ) / 360.0}},
		};
	}
};
const RknTableau &rkn_tableau() { static const RknTableau T; return T; }

StageArgs make_stage(const std::vector<Term> &terms, double *const *k)
{
	StageArgs s{};
	s.nterms = (int)terms.size();
	for (int q = 0; q < s.nterms; q++) { s.coef[q] = terms[q].a; s.k[q] = k[terms[q].j]; }
	return s;
}

// next-stage descriptors for the finalize kernel (see NextStage)
NextStage next_rk(const Ctx &c, const std::vector<Term> &terms, double h)
{
	NextStage n{};
	n.kind = 1; n.st = make_stage(terms, c.k); n.y0 = c.y0; n.out = c.ytmp; n.h = h;
	return n;
}
NextStage next_rkn(const Ctx &c, const std::vector<Term> &terms, double h, double ck)
{
	NextStage n{};
	n.kind = 2; n.st = make_stage(terms, c.k); n.y0 = c.y0; n.out = c.ytmp; n.h = h;
	n.h2 = h * h;        // DormandPrince.cpp:266
	n.ckh = ck * h;      // as launch_rkn_stage
	return n;
}
}  // namespace

// ---- small systems: one launch per attempt (see launch_small_attempt) ----
// 1: whole system in the single-CTA kernel; 2: massive bodies in the single-CTA kernel + every tracer's
// whole attempt in tracer_attempt_kernel; 0: general multi-launch path
static int attempt_path(const Ctx &c)
{
	if (c.small_mode != 0 && c.nranks == 1 && c.cnt.n <= (c.small_mode == 1 ? kSmallAuto : kSmallMax)) return 1;
	if (c.tracer_mode != 0 && c.cnt.s == 0 && c.cnt.M <= kTracerMaxSources && c.cnt.n > c.cnt.M) return 2;
	return 0;
}
static bool use_small(const Ctx &c) { return attempt_path(c) != 0; }

static void launch_attempt(Ctx &c, SmallPlan &P, double h_first)
{
	const int path = attempt_path(c);
	P.h_first = h_first;
	P.n_active = path == 2 ? c.cnt.M : c.cnt.n;
	launch_small_attempt(c, P);
	if (path == 2) launch_tracer_attempt(c, P);
}

static SmallEval small_eval(Ctx &c, const std::vector<Term> &terms, int out, double t, unsigned flags, bool last, double ckh)
{
	SmallEval e{};
	e.nterms = (int)terms.size();
	for (int q = 0; q < e.nterms; q++) { e.kidx[q] = terms[q].j; e.coef[q] = terms[q].a; }
	e.out = out;
	e.factor = c.has_nebula ? reduction_factor_host(c.neb, t) : 1.0;
	e.flags = flags;
	e.last = last ? 1 : 0;
	e.ckh = ckh;
	return e;
}

static void small_account(Ctx &c, int evals)
{
	c.evals += evals;
	c.pairs += evals * pairs_per_eval(c);
}

// ---- general (multi-launch) path: the launches of a Driver call, in two segments ----
// kind 0: the k0 = f(t, y0) evaluation that opens a Driver call (+ yscale for RKF78); for RK4 the whole step.
// kind 1: the remaining evaluations of one attempt + solution / error kernel.
static int issue_segment_body(Ctx &c, int integrator, int kind, double t, double h);
static int issue_segment(Ctx &c, int integrator, int kind, double t, double h)
{
	// (staging of the sources by a finalize kernel only carries over between the evaluations of one segment)
	c.src4_state = nullptr;
	const int rc = issue_segment_body(c, integrator, kind, t, h);
	c.src4_state = nullptr;
	return rc;
}
static int issue_segment_body(Ctx &c, int integrator, int kind, double t, double h)
{
	const unsigned flags = SOL_EVAL_GAS_DRAG;   // type-I/II terms frozen for the rest of the step
	// every evaluation's finalize kernel also forms the next stage's trial state (NextStage): no separate stage launches
	if (integrator == SOL_RUNGE_KUTTA4) {
		const double a21 = 1.0 / 2.0, a32 = 1.0 / 2.0, a43 = 1.0;
		const double b1 = 1.0 / 6.0, b2 = 1.0 / 3.0, b3 = 1.0 / 3.0, b4 = 1.0 / 6.0;
		const double c2 = 1.0 / 2.0, c3 = 1.0 / 2.0, c4 = 1.0;
		const NextStage n1 = next_rk(c, {{0, a21}}, h), n2 = next_rk(c, {{1, a32}}, h), n3 = next_rk(c, {{2, a43}}, h);
		if (eval_force(c, c.y0, c.k[0], t, SOL_EVAL_ALL, false, true, &n1, 0) != SOL_OK) return SOL_ERR;
		if (eval_force(c, c.ytmp, c.k[1], t + c2 * h, flags, false, true, &n2, 1) != SOL_OK) return SOL_ERR;
		if (eval_force(c, c.ytmp, c.k[2], t + c3 * h, flags, false, true, &n3, 2) != SOL_OK) return SOL_ERR;
		if (eval_force(c, c.ytmp, c.k[3], t + c4 * h, flags, true, true, nullptr, 3) != SOL_OK) return SOL_ERR;
		launch_rk_stage(c, c.y0, h, make_stage({{0, b1}, {1, b2}, {2, b3}, {3, b4}}, c.k), c.y);
		return SOL_OK;
	}
	if (integrator == SOL_RUNGE_KUTTA_FEHLBERG78) {
		const auto &T = rkf78_tableau();
		if (kind == 0) {
			const NextStage n1 = next_rk(c, T[1], h);
			if (eval_force(c, c.y0, c.k[0], t, SOL_EVAL_ALL, false, true, &n1, 0) != SOL_OK) return SOL_ERR;
			launch_yscale(c, c.y0, c.k[0], h, c.yscale);   // once, with the first trial h (:87-89)
			return SOL_OK;
		}
		for (int s = 1; s <= 12; s++) {
			const NextStage nx = s < 12 ? next_rk(c, T[s + 1], h) : NextStage{};
			// NOTE: every stage is evaluated at the SAME time t
			if (eval_force(c, c.ytmp, c.k[s], t, flags, s == 12, true, s < 12 ? &nx : nullptr, s) != SOL_OK) return SOL_ERR;
		}
		if (fused_recording(c)) fused_rec_zero_err(c);
		else SOL_CUDA(cudaMemsetAsync(c.errBits, 0, sizeof(unsigned long long), c.stream));
		launch_rkf78_final(c, c.y0, h, c.k, c.yscale, c.y);
		return SOL_OK;
	}
	const RknTableau &T = rkn_tableau();
	if (kind == 0) {
		const NextStage n1 = next_rkn(c, T.a[1], h, T.c[1]);
		return eval_force(c, c.y0, c.k[0], t, SOL_EVAL_ALL, false, true, &n1, 0);
	}
	for (int k = 1; k <= 8; k++) {
		const NextStage nx = k < 8 ? next_rkn(c, T.a[k + 1], h, T.c[k + 1]) : NextStage{};
		if (eval_force(c, c.ytmp, c.k[k], t + T.c[k] * h, flags, k == 8, false, k < 8 ? &nx : nullptr, k) != SOL_OK) return SOL_ERR;
	}
	if (fused_recording(c)) fused_rec_zero_err(c);
	else SOL_CUDA(cudaMemsetAsync(c.errBits, 0, sizeof(unsigned long long), c.stream));
	launch_rkn_final(c, c.y0, h, T.b, T.bd, c.k, c.y);
	return SOL_OK;
}

// ---- CUDA graphs for mid-size systems ----
// A system of a few hundred to a few thousand bodies on the general path is launch-bound: ~40 launches of a few
// microseconds per RKF78 attempt, each preceded by the host's launch latency.  The launches of a segment are therefore
// captured ONCE per (integrator, segment, state buffer) into a CUDA graph and replayed; what changes from attempt to
// attempt - h, c_k h, the reduction factors - is read by the kernels from c.ssDev (StepScalars), refreshed by one small
// copy per attempt.  Same kernels, same arguments, same order: bit-identical to issuing the launches one by one.
constexpr int kGraphMaxBodies = 32768;
static bool graph_ok(const Ctx &c)
{
	return c.graph_mode != 0 && c.nranks == 1 && !c.prof && attempt_path(c) == 0 && c.cnt.n <= kGraphMaxBodies && c.ssDev != nullptr;
}

// the scalars of the attempt that starts at time t with step h, for every evaluation index (same expressions as the
// by-value path: next_rkn's c_k * h, FinalizeDev::factor)
static int stage_scalars(Ctx &c, int integrator, double t, double h)
{
	StepScalars &S = *c.ssHost;
	S.h = h; S.h2 = h * h;
	double tq[13];
	for (int q = 0; q < 13; q++) { tq[q] = t; S.ckh[q] = 0.0; }
	if (integrator == SOL_RUNGE_KUTTA4) {
		const double c2 = 1.0 / 2.0, c3 = 1.0 / 2.0, c4 = 1.0;
		tq[1] = t + c2 * h; tq[2] = t + c3 * h; tq[3] = t + c4 * h;
	} else if (integrator == SOL_DORMAND_PRINCE) {
		const RknTableau &T = rkn_tableau();
		for (int k = 1; k <= 8; k++) { tq[k] = t + T.c[k] * h; S.ckh[k] = T.c[k] * h; }
	}
	for (int q = 0; q < 13; q++) S.factor[q] = c.has_nebula ? reduction_factor_host(c.neb, tq[q]) : 1.0;
	SOL_CUDA(cudaMemcpyAsync(c.ssDev, c.ssHost, sizeof(StepScalars), cudaMemcpyHostToDevice, c.stream));
	return SOL_OK;
}

// graph_mode 2: the segment as ONE cooperative kernel (fused_attempt_kernel, elementwise.cu).  Same device code over the
// same block decomposition as the launches it stands for: bit-identical.  Not for segments that use the symmetric pair
// kernel, a memset between launches (no source besides the star) or several sinks per thread - those replay a graph.
constexpr int kFusedMaxBodies = 32768;
static bool fused_ok(Ctx &c)
{
	if (!(c.graph_mode == 2 && graph_ok(c) && c.cnt.n <= kFusedMaxBodies && c.fusedBar != nullptr)) return false;
	const Counts &n = c.cnt;
	const bool bary = c.barycentric != 0;
	const int src_hi = bary ? n.M : n.M + n.s;
	if (!bary && src_hi <= 1) return false;
	const int sq_n = n.M - (bary ? 0 : 1);
	if (c.sym_mode != 0 && sq_n >= (c.sym_mode == 1 ? kSymMinBodies : kSymAutoBodies)) return false;
	return fused_grid_size(c) > 0;
}

static int run_segment(Ctx &c, int integrator, int kind, double t, double h)
{
	if (!graph_ok(c)) return issue_segment(c, integrator, kind, t, h);
	if (fused_ok(c)) {
		Ctx::FusedEntry *f = nullptr;
		for (auto &e : c.fused)
			if (e.integrator == integrator && e.kind == kind && e.y0 == c.y0 && e.epoch == c.cfg_epoch) f = &e;
		if (f == nullptr) {
			for (size_t k = 0; k < c.fused.size();) {
				if (c.fused[k].epoch != c.cfg_epoch) { if (c.fused[k].program) cudaFree(c.fused[k].program); c.fused.erase(c.fused.begin() + k); }
				else k++;
			}
			const long long l0 = c.launches;
			const double ev0 = c.evals, pr0 = c.pairs;
			fused_begin_record(c);
			c.capturing = true;
			const int rc = issue_segment(c, integrator, kind, t, h);
			c.capturing = false;
			Ctx::FusedEntry e{};
			e.integrator = integrator; e.kind = kind; e.y0 = c.y0; e.epoch = c.cfg_epoch;
			const int fr = fused_end_record(c, &e.program, &e.ops);
			c.launches = l0; c.evals = ev0; c.pairs = pr0;
			if (rc != SOL_OK || fr == SOL_ERR) return SOL_ERR;
			c.fused.push_back(e);                 // (program == null: this segment is not fusable, remembered)
			f = &c.fused.back();
		}
		if (f->program != nullptr) {
			if (launch_fused(c, f->program) != SOL_OK) return SOL_ERR;
			const int ne = integrator == SOL_RUNGE_KUTTA4 ? 4 : (kind == 0 ? 1 : (integrator == SOL_RUNGE_KUTTA_FEHLBERG78 ? 12 : 8));
			c.evals += ne; c.pairs += ne * pairs_per_eval(c);
			return SOL_OK;
		}
	}
	Ctx::GraphEntry *g = nullptr;
	for (auto &e : c.graphs)
		if (e.integrator == integrator && e.kind == kind && e.y0 == c.y0 && e.epoch == c.cfg_epoch) g = &e;
	if (g == nullptr) {
		// drop graphs of older configurations, then capture this segment
		for (size_t k = 0; k < c.graphs.size();) {
			if (c.graphs[k].epoch != c.cfg_epoch) { cudaGraphExecDestroy(c.graphs[k].exec); c.graphs.erase(c.graphs.begin() + k); }
			else k++;
		}
		const int sq_n = c.cnt.M - (c.barycentric ? 0 : 1);
		if (c.sym_mode != 0 && sq_n >= (c.sym_mode == 1 ? kSymMinBodies : kSymAutoBodies) && alloc_sym(c) != SOL_OK) return SOL_ERR;
		const long long l0 = c.launches;
		const double ev0 = c.evals, pr0 = c.pairs;
		cudaGraph_t graph = nullptr;
		SOL_CUDA(cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeRelaxed));
		c.capturing = true;
		const int rc = issue_segment(c, integrator, kind, t, h);
		c.capturing = false;
		const cudaError_t ce = cudaStreamEndCapture(c.stream, &graph);
		c.evals = ev0; c.pairs = pr0;                     // (counted per replay below)
		if (rc != SOL_OK || ce != cudaSuccess || graph == nullptr) {
			if (graph) cudaGraphDestroy(graph);
			if (rc == SOL_OK) c.err = std::string("graph capture: ") + cudaGetErrorString(ce);
			return SOL_ERR;
		}
		Ctx::GraphEntry e{};
		e.integrator = integrator; e.kind = kind; e.y0 = c.y0; e.epoch = c.cfg_epoch; e.launches = (int)(c.launches - l0);
		const cudaError_t ie = cudaGraphInstantiate(&e.exec, graph, 0);
		cudaGraphDestroy(graph);
		c.launches = l0;
		if (ie != cudaSuccess) { c.err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie); return SOL_ERR; }
		c.graphs.push_back(e);
		g = &c.graphs.back();
	}
	SOL_CUDA(cudaGraphLaunch(g->exec, c.stream));
	c.launches += g->launches;
	const int ne = integrator == SOL_RUNGE_KUTTA4 ? 4 : (kind == 0 ? 1 : (integrator == SOL_RUNGE_KUTTA_FEHLBERG78 ? 12 : 8));
	c.evals += ne; c.pairs += ne * pairs_per_eval(c);
	return SOL_OK;
}

static int driver_rk4(Ctx &c, double *time, double *hNext, double *hDid, double *info)
{
	const double t = *time, h = *hNext;
	if (use_small(c)) {
		SmallPlan P{};
		P.integrator = SOL_RUNGE_KUTTA4; P.h = h; P.first = 1; P.nevals = 4;
		P.ev[0] = small_eval(c, {}, 0, t, SOL_EVAL_ALL, false, 0.0);
		P.ev[1] = small_eval(c, {{0, 1.0 / 2.0}}, 1, t + (1.0 / 2.0) * h, SOL_EVAL_GAS_DRAG, false, 0.0);
		P.ev[2] = small_eval(c, {{1, 1.0 / 2.0}}, 2, t + (1.0 / 2.0) * h, SOL_EVAL_GAS_DRAG, false, 0.0);
		P.ev[3] = small_eval(c, {{2, 1.0}}, 3, t + 1.0 * h, SOL_EVAL_GAS_DRAG, true, 0.0);
		launch_attempt(c, P, h);
		small_account(c, 4);
	} else {
		if (graph_ok(c) && stage_scalars(c, SOL_RUNGE_KUTTA4, t, h) != SOL_OK) return SOL_ERR;
		if (run_segment(c, SOL_RUNGE_KUTTA4, 0, t, h) != SOL_OK) return SOL_ERR;
	}
	*hDid = h;
	*time += *hDid;
	*hNext = h;
	std::swap(c.y0, c.y);
	if (info) { info[0] = 1; info[1] = 0; }
	return SOL_OK;
}

static int driver_rkf78(Ctx &c, double *time, double *hNext, double *hDid, double *info)
{
	const double SAFETY = 0.9, PGROW = -0.2, PSHRNK = -0.25, ERRCON = 1.89e-4;
	const double epsilon = pow(10, -10.0);   // RungeKuttaFehlberg78.cpp:38-39 (XML <Accuracy> is ignored, Q10)
	const auto &T = rkf78_tableau();
	const double t = *time;
	double h = *hNext;
	const bool small = use_small(c);
	const unsigned flags = SOL_EVAL_GAS_DRAG;
	if (!small) {
		if (graph_ok(c) && stage_scalars(c, SOL_RUNGE_KUTTA_FEHLBERG78, t, h) != SOL_OK) return SOL_ERR;
		if (run_segment(c, SOL_RUNGE_KUTTA_FEHLBERG78, 0, t, h) != SOL_OK) return SOL_ERR;
	}
	double errorMax = 0.0;
	int attempts = 0;
	for (;;) {
		if (small) {
			SmallPlan P{};
			P.integrator = SOL_RUNGE_KUTTA_FEHLBERG78; P.h = h; P.first = attempts == 0 ? 1 : 0; P.nevals = 13;
			P.ev[0] = small_eval(c, {}, 0, t, SOL_EVAL_ALL, false, 0.0);
			for (int s = 1; s <= 12; s++) P.ev[s] = small_eval(c, T[s], s, t, flags, s == 12, 0.0);
			launch_attempt(c, P, *hNext);
			small_account(c, attempts == 0 ? 13 : 12);
		} else {
			if (attempts > 0) {
				// a repeated attempt starts from k0 again: its first trial state needs the new h
				launch_rk_stage(c, c.y0, h, make_stage(T[1], c.k), c.ytmp);
				if (graph_ok(c) && stage_scalars(c, SOL_RUNGE_KUTTA_FEHLBERG78, t, h) != SOL_OK) return SOL_ERR;
			}
			if (run_segment(c, SOL_RUNGE_KUTTA_FEHLBERG78, 1, t, h) != SOL_OK) return SOL_ERR;
		}
		attempts++;
		double emax;
		if (read_error_max(c, emax) != SOL_OK) return SOL_ERR;
		errorMax = emax / epsilon;
		if (errorMax < 1.0) { *hDid = h; break; }
		double hTemp = SAFETY * h * pow(errorMax, PSHRNK);
		h = fabs(hTemp) > fabs(0.1 * h) ? hTemp : 0.1 * h;
		double tNew = *time + h;
		if (tNew == *time) {
			c.err = "Stepsize-underflow occurred during Runge-Kutta-Fehlberg7(8) step!";
			if (info) { info[0] = attempts; info[1] = errorMax; }
			return SOL_ERR;
		}
	}
	*time += *hDid;
	*hNext = errorMax > ERRCON ? (SAFETY * h * pow(errorMax, PGROW)) : (5.0 * h);
	std::swap(c.y0, c.y);
	if (info) { info[0] = attempts; info[1] = errorMax; }
	return SOL_OK;
}

static int driver_rkn76(Ctx &c, double *time, double *hNext, double *hDid, double *info)
{
	const double epsilon = pow(10, -10.0);   // DormandPrince.cpp:31-32
	const int maxIter = 10;
	const RknTableau &T = rkn_tableau();
	const double t = *time;
	const bool small = use_small(c);
	if (!small) {
		if (graph_ok(c) && stage_scalars(c, SOL_DORMAND_PRINCE, t, *hNext) != SOL_OK) return SOL_ERR;
		if (run_segment(c, SOL_DORMAND_PRINCE, 0, t, *hNext) != SOL_OK) return SOL_ERR;
	}
	const unsigned flags = SOL_EVAL_GAS_DRAG;
	int iter = 0;
	double errorMax = 0.0;
	do {
		iter++;
		const double h = *hNext;
		if (small) {
			SmallPlan P{};
			P.integrator = SOL_DORMAND_PRINCE; P.h = h; P.first = iter == 1 ? 1 : 0; P.nevals = 9;
			for (int q = 0; q < 9; q++) { P.b[q] = T.b[q]; P.bd[q] = T.bd[q]; }
			P.ev[0] = small_eval(c, {}, 0, t, SOL_EVAL_ALL, false, 0.0);
			for (int k = 1; k <= 8; k++) P.ev[k] = small_eval(c, T.a[k], k, t + T.c[k] * h, flags, k == 8, T.c[k] * h);
			launch_attempt(c, P, h);
			small_account(c, iter == 1 ? 9 : 8);
		} else {
			if (iter > 1) {
				launch_rkn_stage(c, c.y0, h, T.c[1], make_stage(T.a[1], c.k), c.ytmp);   // repeated attempt: new h
				if (graph_ok(c) && stage_scalars(c, SOL_DORMAND_PRINCE, t, h) != SOL_OK) return SOL_ERR;
			}
			if (run_segment(c, SOL_DORMAND_PRINCE, 1, t, h) != SOL_OK) return SOL_ERR;
		}
		if (read_error_max(c, errorMax) != SOL_OK) return SOL_ERR;
		*hDid = h;
		*hNext = errorMax < 1.0e-20 ? 2.0 * h : 0.9 * h * pow(epsilon / errorMax, 1.0 / 7.0);
	} while (errorMax > epsilon && iter <= maxIter);
	if (info) { info[0] = iter; info[1] = errorMax; }
	if (iter > maxIter) {
		c.err = "An error occurred during Prince-Dormand driver: iteration number exceeded maxIter!";
		return SOL_ERR;
	}
	*time += *hDid;
	std::swap(c.y0, c.y);
	return SOL_OK;
}

// ---------------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

int sol_create(int device, sol_ctx **out)
{
	if (!out) return SOL_ERR;
	*out = nullptr;
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev <= 0) {
		g_create_error = std::string("no usable CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
		                 "); solaris_b200 has no CPU fallback";
		return SOL_ERR;
	}
	if (device < 0 || device >= ndev) { g_create_error = "device index out of range"; return SOL_ERR; }
	if ((e = cudaSetDevice(device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return SOL_ERR; }
	sol_ctx *h = new sol_ctx();
	Ctx &c = h->c;
	c.device = device;
	bool ok = cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking) == cudaSuccess;
	c.own_stream = ok;
	ok = ok && cudaMalloc((void **)&c.errBits, sizeof(unsigned long long)) == cudaSuccess;
	ok = ok && cudaMallocHost((void **)&c.errBitsHost, sizeof(unsigned long long)) == cudaSuccess;
	ok = ok && cudaMalloc((void **)&c.evCount, 8 * sizeof(int)) == cudaSuccess;
	ok = ok && cudaMallocHost((void **)&c.evCountHost, 8 * sizeof(int)) == cudaSuccess;
	ok = ok && cudaMalloc((void **)&c.ssDev, sizeof(StepScalars)) == cudaSuccess;
	ok = ok && cudaMalloc((void **)&c.fusedBar, sizeof(unsigned)) == cudaSuccess;
	ok = ok && cudaStreamCreateWithFlags(&c.side, cudaStreamNonBlocking) == cudaSuccess;
	if (const char *e = getenv("SOLARIS_B200_STAGE_IN_FINALIZE")) c.stage_in_finalize = atoi(e) != 0;
	ok = ok && cudaEventCreateWithFlags(&c.evFork, cudaEventDisableTiming) == cudaSuccess;
	ok = ok && cudaEventCreateWithFlags(&c.evJoin, cudaEventDisableTiming) == cudaSuccess;
	ok = ok && cudaMallocHost((void **)&c.ssHost, sizeof(StepScalars)) == cudaSuccess;
	ok = ok && cudaMalloc((void **)&c.runOut, sizeof(RunOut)) == cudaSuccess;
	ok = ok && cudaMallocHost((void **)&c.runOutHost, sizeof(RunOut)) == cudaSuccess;
	ok = ok && cudaMalloc((void **)&c.indPart, kIndirectBlocks * 6 * sizeof(double)) == cudaSuccess;
	ok = ok && cudaMalloc((void **)&c.indirect, 6 * sizeof(double)) == cudaSuccess;
	ok = ok && cudaMalloc((void **)&c.indCounter, sizeof(unsigned)) == cudaSuccess;
	ok = ok && cudaMalloc((void **)&c.integralsPart, kIndirectBlocks * 12 * sizeof(double)) == cudaSuccess;
	ok = ok && cudaMalloc((void **)&c.integralsDev, 12 * sizeof(double)) == cudaSuccess;
	ok = ok && cudaMallocHost((void **)&c.integralsHost, 12 * sizeof(double)) == cudaSuccess;
	ok = ok && cudaMalloc((void **)&c.stageSrc, 13 * kSmallMax * sizeof(double4)) == cudaSuccess;
	ok = ok && cudaMalloc((void **)&c.stageS6, 13 * 6 * sizeof(double)) == cudaSuccess;
	ok = ok && cudaEventCreate(&c.ev0) == cudaSuccess && cudaEventCreate(&c.ev1) == cudaSuccess;
	if (ok) {
		cudaMemset(c.indirect, 0, 6 * sizeof(double));
		cudaMemset(c.indCounter, 0, sizeof(unsigned));
		cudaMemset(c.fusedBar, 0, sizeof(unsigned));
		cudaMemset(c.evCount, 0, 8 * sizeof(int)); memset(c.evCountHost, 0, 8 * sizeof(int));
		cudaMemset(c.errBits, 0, sizeof(unsigned long long));
	}
	if (!ok) {
		g_create_error = std::string("context allocation failed: ") + cudaGetErrorString(cudaGetLastError());
		delete h;
		return SOL_ERR;
	}
	*out = h;
	return SOL_OK;
}

int sol_create_multi(int n_gpus, sol_ctx **out)
{
	if (!out) return SOL_ERR;
	*out = nullptr;
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev <= 0) {
		g_create_error = std::string("no usable CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
		                 "); solaris_b200 has no CPU fallback";
		return SOL_ERR;
	}
	if (n_gpus < 1 || n_gpus > ndev) { g_create_error = "sol_create_multi: " + std::to_string(n_gpus) + " GPUs requested, " + std::to_string(ndev) + " visible"; return SOL_ERR; }
	if (n_gpus == 1) return sol_create(0, out);
	if (!g_nccl.load(g_create_error)) return SOL_ERR;
	ncclUniqueId id;
	if (g_nccl.GetUniqueId(&id) != ncclSuccess) { g_create_error = "ncclGetUniqueId failed"; return SOL_ERR; }
	sol_ctx *front = new sol_ctx();
	MultiCtx *M = new MultiCtx();
	front->multi = M;
	M->ranks.resize(n_gpus, nullptr);
	M->rc.assign(n_gpus, SOL_OK);
	for (int r = 0; r < n_gpus; r++) {
		if (sol_create(r, &M->ranks[r]) != SOL_OK) {
			for (int q = 0; q < r; q++) sol_destroy(M->ranks[q]);
			delete M; delete front;
			return SOL_ERR;
		}
	}
	for (int r = 0; r < n_gpus; r++) M->threads.emplace_back(multi_worker, M, r);
	// every worker joins the communicator from its own thread (ncclCommInitRank blocks until all ranks arrive)
	if (fan_out(front, [&](sol_ctx *rk, int rank) { return sol_dist_init(rk, rank, n_gpus, &id); }) != SOL_OK) {
		g_create_error = front->c.err;
		sol_destroy(front);
		return SOL_ERR;
	}
	*out = front;
	return SOL_OK;
}

void sol_destroy(sol_ctx *h)
{
	if (!h) return;
	if (h->multi) {
		MultiCtx *M = h->multi;
		{
			std::lock_guard<std::mutex> lk(M->m);
			M->quit = true;
			M->cv_job.notify_all();
		}
		for (auto &t : M->threads) t.join();
		for (sol_ctx *r : M->ranks) sol_destroy(r);
		delete M;
		delete h;
		return;
	}
	Ctx &c = h->c;
	cudaSetDevice(c.device);
	cudaStreamSynchronize(c.stream);
	free_bodies(c);
	if (c.nccl && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)c.nccl);
	cudaFree(c.errBits); cudaFreeHost(c.errBitsHost); cudaFree(c.evCount); cudaFreeHost(c.evCountHost);
	cudaFree(c.indPart); cudaFree(c.indirect); cudaFree(c.indCounter); cudaFree(c.stageSrc); cudaFree(c.stageS6); cudaFree(c.integralsPart); cudaFree(c.integralsDev); cudaFreeHost(c.integralsHost);
	if (c.pin) cudaFreeHost(c.pin);
	cudaFree(c.runOut); cudaFreeHost(c.runOutHost); if (c.runRec) cudaFree(c.runRec);
	for (auto &g : c.graphs) cudaGraphExecDestroy(g.exec);
	for (auto &f : c.fused) if (f.program) cudaFree(f.program);
	cudaFree(c.fusedBar);
	if (c.side) cudaStreamDestroy(c.side);
	if (c.evFork) cudaEventDestroy(c.evFork);
	if (c.evJoin) cudaEventDestroy(c.evJoin);
	cudaFree(c.ssDev); cudaFreeHost(c.ssHost);
	cudaEventDestroy(c.ev0); cudaEventDestroy(c.ev1);
	for (auto e : c.ev_pool) cudaEventDestroy(e);
	if (c.own_stream) cudaStreamDestroy(c.stream);
	delete h;
}

const char *sol_last_error(const sol_ctx *h) { return h ? h->c.err.c_str() : g_create_error.c_str(); }

int sol_set_stream(sol_ctx *h, void *stream)
{
	if (!h) return SOL_ERR;
	if (h->multi) { h->c.err = "sol_set_stream: a multi-GPU handle runs one private stream per device"; return SOL_ERR; }
	Ctx &c = h->c;
	cudaStreamSynchronize(c.stream);
	if (c.own_stream) { cudaStreamDestroy(c.stream); c.own_stream = false; }
	c.stream = (cudaStream_t)stream;
	c.cfg_epoch++;
	return SOL_OK;
}

int sol_set_frame(sol_ctx *h, int barycentric)
{
	if (!h) return SOL_ERR;
	SOL_FANOUT(h, sol_set_frame(r, barycentric));
	h->c.barycentric = barycentric ? 1 : 0;
	h->c.cfg_epoch++;
	return SOL_OK;
}

int sol_set_nebula(sol_ctx *h, const sol_nebula_pod *neb)
{
	if (!h) return SOL_ERR;
	SOL_FANOUT(h, sol_set_nebula(r, neb));
	Ctx &c = h->c;
	c.has_nebula = neb != nullptr;
	if (neb) c.neb = *neb;
	refresh_gas(c);
	c.cfg_epoch++;
	return SOL_OK;
}

int sol_set_nn_tracking(sol_ctx *h, int mode)
{
	if (!h || mode < 0 || mode > 2) return SOL_ERR;
	SOL_FANOUT(h, sol_set_nn_tracking(r, mode));
	h->c.nn_mode = mode;
	h->c.cfg_epoch++;
	return SOL_OK;
}

int sol_body_count(const sol_ctx *h) { return h ? (h->multi ? sol_body_count(h->multi->ranks[0]) : h->c.cnt.n) : 0; }

int sol_set_bodies(sol_ctx *h, const int counts[7], const double *y0, const double *mass, const double *radius,
                   const double *density, const double *cD, const double *gS, const double *gE, const double *migStop,
                   const int *type, const int *migType, const int *id)
{
	if (!h) return SOL_ERR;
	SOL_FANOUT(h, sol_set_bodies(r, counts, y0, mass, radius, density, cD, gS, gE, migStop, type, migType, id));
	Ctx &c = h->c;
	SOL_CUDA(cudaSetDevice(c.device));
	Counts n{};
	n.c = counts[0]; n.g = counts[1]; n.r = counts[2]; n.p = counts[3]; n.s = counts[4]; n.l = counts[5]; n.t = counts[6];
	n.n = n.c + n.g + n.r + n.p + n.s + n.l + n.t;
	n.M = n.c + n.g + n.r + n.p;
	if (n.n <= 0) { c.err = "host memory allocation"; return SOL_ERR; }   // BodyData::Allocate, BodyData.cpp:68-72
	if (n.c != 1) { c.err = "exactly one central body is required (body 0)"; return SOL_ERR; }
	if (!y0 || !mass || !radius || !density || !cD || !gS || !gE || !migStop || !type || !migType || !id) {
		c.err = "sol_set_bodies: null array"; return SOL_ERR;
	}
	if (n.n > c.alloc_n || n.n < c.alloc_n / 2) {
		if (alloc_bodies(c, n.n) != SOL_OK) return SOL_ERR;
	}
	c.cnt = n;
	c.cfg_epoch++;
	if (c.nranks > 1) shard_of(n.n, c.nranks, c.rank, c.lo, c.hi);
	else { c.lo = 0; c.hi = n.n; }
	const size_t nb = (size_t)n.n;
	if (ensure_stage(c, 6 * nb) != SOL_OK) return SOL_ERR;
	SOL_CUDA(cudaMemcpyAsync(c.stage_aos, y0, 6 * nb * sizeof(double), cudaMemcpyHostToDevice, c.stream));
	launch_aos_to_planes(c, c.stage_aos, c.y0, n.n);
#define UP(dst, src, T) SOL_CUDA(cudaMemcpyAsync(dst, src, nb * sizeof(T), cudaMemcpyHostToDevice, c.stream));
	UP(c.mass, mass, double) UP(c.radius, radius, double) UP(c.density, density, double) UP(c.cD, cD, double)
	UP(c.gS, gS, double) UP(c.gE, gE, double) UP(c.migStop, migStop, double)
	UP(c.type, type, int) UP(c.migType, migType, int) UP(c.id, id, int)
#undef UP
	// Acceleration::rm3 starts zeroed (Acceleration.cpp:65-69); NN arrays start at -1 / 0
	SOL_CUDA(cudaMemsetAsync(c.rm3, 0, c.ld * sizeof(double), c.stream));
	SOL_CUDA(cudaMemsetAsync(c.nnDist, 0, c.ld * sizeof(double), c.stream));
	SOL_CUDA(cudaMemsetAsync(c.nnIdx, 0xff, c.ld * sizeof(int), c.stream));
	SOL_CUDA(cudaMemsetAsync(c.aGas, 0, 3 * (size_t)c.ld * sizeof(double), c.stream));
	SOL_CUDA(cudaMemsetAsync(c.aMig1, 0, 3 * (size_t)c.ld * sizeof(double), c.stream));
	SOL_CUDA(cudaMemsetAsync(c.aMig2, 0, 3 * (size_t)c.ld * sizeof(double), c.stream));
	SOL_CUDA(cudaMemsetAsync(c.indirect, 0, 6 * sizeof(double), c.stream));
	SOL_CUDA(cudaStreamSynchronize(c.stream));
	c.mass0 = mass[0];
	refresh_gas(c);
	return SOL_OK;
}

int sol_compute(sol_ctx *h, double t, const double *y_host, double *dydt_host, unsigned eval_flags)
{
	if (!h) return SOL_ERR;
	SOL_FANOUT(h, sol_compute(r, t, y_host, dydt_host, eval_flags));
	Ctx &c = h->c;
	if (c.cnt.n <= 0) { c.err = "sol_compute before sol_set_bodies"; return SOL_ERR; }
	if (!y_host || !dydt_host) { c.err = "sol_compute: null pointer"; return SOL_ERR; }
	SOL_CUDA(cudaSetDevice(c.device));
	const size_t nb = (size_t)c.cnt.n;
	if (ensure_stage(c, 6 * nb) != SOL_OK) return SOL_ERR;
	SOL_CUDA(cudaMemcpyAsync(c.stage_aos, y_host, 6 * nb * sizeof(double), cudaMemcpyHostToDevice, c.stream));
	launch_aos_to_planes(c, c.stage_aos, c.ytmp, c.cnt.n);
	if (eval_force(c, c.ytmp, c.k[1], t, eval_flags, true, true) != SOL_OK) return SOL_ERR;
	// a sharded context evaluates its own sinks [lo, hi) and fills exactly those rows of dydt_host (collective call: every
	// rank passes the full y; with a multi-GPU handle the ranks share the caller's array, which ends up complete)
	const int lo = c.nranks > 1 ? c.lo : 0, hi = c.nranks > 1 ? c.hi : c.cnt.n;
	if (hi > lo) {
		launch_planes_to_aos(c, c.k[1] + lo, c.stage_aos, hi - lo);
		SOL_CUDA(cudaMemcpyAsync(dydt_host + 6 * (size_t)lo, c.stage_aos, 6 * (size_t)(hi - lo) * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
	}
	SOL_CUDA(cudaStreamSynchronize(c.stream));
	return SOL_OK;
}

int sol_compute_device(sol_ctx *h, double t, unsigned eval_flags)
{
	if (!h) return SOL_ERR;
	SOL_FANOUT(h, sol_compute_device(r, t, eval_flags));
	Ctx &c = h->c;
	if (c.cnt.n <= 0) { c.err = "sol_compute_device before sol_set_bodies"; return SOL_ERR; }
	SOL_CUDA(cudaSetDevice(c.device));
	if (eval_force(c, c.y0, c.k[0], t, eval_flags, true, true) != SOL_OK) return SOL_ERR;
	SOL_CUDA(cudaStreamSynchronize(c.stream));
	return SOL_OK;
}

int sol_step(sol_ctx *h, int integrator, double *time, double *h_next, double *h_did, double *info)
{
	if (!h || !time || !h_next || !h_did) return SOL_ERR;
	if (h->multi) {
		// every rank runs the driver on its own copies of the scalars (identical on all ranks: the error norm is all-reduced)
		const int nr = (int)h->multi->ranks.size();
		std::vector<double> tt(nr, *time), hn(nr, *h_next), hd(nr, 0.0), inf(4 * (size_t)nr, 0.0);
		const int rc = fan_out(h, [&](sol_ctx *r, int rank) { return sol_step(r, integrator, &tt[rank], &hn[rank], &hd[rank], &inf[4 * (size_t)rank]); });
		*time = tt[0]; *h_next = hn[0]; *h_did = hd[0];
		if (info) for (int q = 0; q < 4; q++) info[q] = inf[q];
		return rc;
	}
	Ctx &c = h->c;
	if (c.cnt.n <= 0) { c.err = "sol_step before sol_set_bodies"; return SOL_ERR; }
	SOL_CUDA(cudaSetDevice(c.device));
	const double ev0 = c.evals, pr0 = c.pairs;
	int r;
	switch (integrator) {
	case SOL_RUNGE_KUTTA4:           r = driver_rk4(c, time, h_next, h_did, info); break;
	case SOL_RUNGE_KUTTA_FEHLBERG78: r = driver_rkf78(c, time, h_next, h_did, info); break;
	case SOL_DORMAND_PRINCE:         r = driver_rkn76(c, time, h_next, h_did, info); break;
	default: c.err = "Unknown integrator type!"; return SOL_ERR;   // Simulator.cpp:76-80
	}
	if (info) { info[2] = c.evals - ev0; info[3] = c.pairs - pr0; }
	if (r == SOL_OK) {
		cudaError_t e = cudaStreamSynchronize(c.stream);
		if (e != cudaSuccess) { c.err = cudaGetErrorString(e); return SOL_ERR; }
	}
	return r;
}

// ---- seam A, many steps per call ----
// Plan of one attempt for the one-warp kernel, as the single-step drivers build it (the per-attempt fields h, c_k h and
// the reduction factors are rewritten on the device between attempts).
static SmallPlan run_plan(Ctx &c, int integrator, double t, double h)
{
	SmallPlan P{};
	const unsigned flags = SOL_EVAL_GAS_DRAG;
	P.integrator = integrator; P.h = h; P.h_first = h; P.first = 1; P.n_active = c.cnt.n;
	if (integrator == SOL_RUNGE_KUTTA4) {
		P.nevals = 4;
		P.ev[0] = small_eval(c, {}, 0, t, SOL_EVAL_ALL, false, 0.0);
		P.ev[1] = small_eval(c, {{0, 1.0 / 2.0}}, 1, t, flags, false, 0.0);
		P.ev[2] = small_eval(c, {{1, 1.0 / 2.0}}, 2, t, flags, false, 0.0);
		P.ev[3] = small_eval(c, {{2, 1.0}}, 3, t, flags, true, 0.0);
	} else if (integrator == SOL_RUNGE_KUTTA_FEHLBERG78) {
		const auto &T = rkf78_tableau();
		P.nevals = 13;
		P.ev[0] = small_eval(c, {}, 0, t, SOL_EVAL_ALL, false, 0.0);
		for (int s = 1; s <= 12; s++) P.ev[s] = small_eval(c, T[s], s, t, flags, s == 12, 0.0);
	} else {
		const RknTableau &T = rkn_tableau();
		P.nevals = 9;
		for (int q = 0; q < 9; q++) { P.b[q] = T.b[q]; P.bd[q] = T.bd[q]; }
		P.ev[0] = small_eval(c, {}, 0, t, SOL_EVAL_ALL, false, 0.0);
		for (int k = 1; k <= 8; k++) P.ev[k] = small_eval(c, T.a[k], k, t, flags, k == 8, T.c[k] * h);
	}
	return P;
}

static int count_events(Ctx &c, double ejection, double hit_centrum, double collision_factor, int counts_out[3]);

int sol_run(sol_ctx *h, sol_run_args *A)
{
	if (!h || !A) return SOL_ERR;
	if (h->multi) {
		const int nr = (int)h->multi->ranks.size();
		std::vector<sol_run_args> args(nr, *A);
		for (int q = 1; q < nr; q++) args[q].records = nullptr;
		const int rc = fan_out(h, [&](sol_ctx *r, int rank) { return sol_run(r, &args[rank]); });
		*A = args[0];
		return rc;
	}
	Ctx &c = h->c;
	if (c.cnt.n <= 0) { c.err = "sol_run before sol_set_bodies"; return SOL_ERR; }
	if (A->max_steps < 1) { c.err = "sol_run: max_steps must be >= 1"; return SOL_ERR; }
	if (A->integrator != SOL_RUNGE_KUTTA4 && A->integrator != SOL_RUNGE_KUTTA_FEHLBERG78 && A->integrator != SOL_DORMAND_PRINCE) {
		c.err = "Unknown integrator type!"; return SOL_ERR;
	}
	SOL_CUDA(cudaSetDevice(c.device));
	A->steps = 0; A->stop_reason = SOL_RUN_MAX_STEPS; A->attempts = 0; A->err_max = 0.0; A->h_did = 0.0;
	A->event_counts[0] = A->event_counts[1] = A->event_counts[2] = 0;
	const double e3 = A->ejection > 0 ? 1.0 / (A->ejection * A->ejection * A->ejection) : 0.0;          // Simulator.cpp:626-629
	const double h3 = A->hit_centrum > 0 ? 1.0 / (A->hit_centrum * A->hit_centrum * A->hit_centrum) : 0.0;

	if (warp_run_eligible(c)) {
		// ---- one persistent launch (<= 32 bodies, all massive) ----
		RunCtl R{};
		R.max_steps = A->max_steps; R.time = A->time; R.h_next = A->h_next;
		R.millenium_days = A->millenium_days; R.length = A->length; R.output = A->output; R.last_save = A->last_save;
		R.e3 = e3; R.h3 = h3; R.ej_on = A->ejection > 0; R.hc_on = A->hit_centrum > 0; R.col_factor = A->collision_factor;
		R.step_counter = A->step_counter; R.flush_every = A->flush_every; R.tiny = A->flush_threshold;
		R.eps = pow(10, -10.0);                                   // RungeKuttaFehlberg78.cpp:38-39, DormandPrince.cpp:31-32
		if (A->integrator == SOL_RUNGE_KUTTA4) { R.cstage[1] = 1.0 / 2.0; R.cstage[2] = 1.0 / 2.0; R.cstage[3] = 1.0; }
		else if (A->integrator == SOL_DORMAND_PRINCE) { const RknTableau &T = rkn_tableau(); for (int k = 0; k < 9; k++) R.cstage[k] = T.c[k]; }
		R.time_dependent_factor = (c.has_nebula && c.neb.decrease_type == 1) ? 1 : 0;
		R.rec = nullptr;
		if (A->records != nullptr) {
			const size_t need = 4 * (size_t)A->max_steps;
			if (c.runRecCap < need) {
				if (c.runRec) cudaFree(c.runRec);
				c.runRec = nullptr; c.runRecCap = 0;
				SOL_CUDA(cudaMalloc((void **)&c.runRec, need * sizeof(double)));
				c.runRecCap = need;
			}
			R.rec = c.runRec;
		}
		const SmallPlan P = run_plan(c, A->integrator, A->time, A->h_next);
		launch_warp_run(c, P, R, c.runOut);
		SOL_CUDA(cudaMemcpyAsync(c.runOutHost, c.runOut, sizeof(RunOut), cudaMemcpyDeviceToHost, c.stream));
		SOL_CUDA(cudaStreamSynchronize(c.stream));
		const RunOut &o = *c.runOutHost;
		if (A->records != nullptr && o.steps > 0) {
			SOL_CUDA(cudaMemcpyAsync(A->records, c.runRec, 4 * (size_t)o.steps * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
			SOL_CUDA(cudaStreamSynchronize(c.stream));
		}
		A->time = o.time; A->h_next = o.h_next; A->h_did = o.h_did; A->last_save = o.last_save; A->err_max = o.err_max;
		A->step_counter = o.step_counter; A->attempts = o.attempts; A->steps = o.steps; A->stop_reason = o.stop_reason;
		A->event_counts[0] = o.ev[0]; A->event_counts[1] = o.ev[1]; A->event_counts[2] = o.ev[2];
		c.evals += (double)o.evals;
		c.pairs += (double)o.evals * pairs_per_eval(c);
		// keep sol_event_indices / sol_event_records usable after an event stop
		if (o.stop_reason == SOL_RUN_EVENT) {
			int cnt[3];
			if (count_events(c, A->ejection, A->hit_centrum, A->collision_factor, cnt) != SOL_OK) return SOL_ERR;
		}
		if (o.stop_reason == 4) {
			c.err = o.err_code == 1 ? "Stepsize-underflow occurred during Runge-Kutta-Fehlberg7(8) step!"
			                        : "An error occurred during Prince-Dormand driver: iteration number exceeded maxIter!";
			A->stop_reason = SOL_RUN_ERROR;
			return SOL_ERR;
		}
		return SOL_OK;
	}

	// ---- general systems: the same loop on the host, one Driver + one flag reduction per step ----
	double lastSave = A->last_save;
	long long counter = A->step_counter;
	while (A->steps < A->max_steps) {
		double info[4] = {0, 0, 0, 0}, hDid = 0.0;
		const double h_trial = A->h_next;
		int r;
		switch (A->integrator) {
		case SOL_RUNGE_KUTTA4:           r = driver_rk4(c, &A->time, &A->h_next, &hDid, info); break;
		case SOL_RUNGE_KUTTA_FEHLBERG78: r = driver_rkf78(c, &A->time, &A->h_next, &hDid, info); break;
		default:                         r = driver_rkn76(c, &A->time, &A->h_next, &hDid, info); break;
		}
		A->attempts += (long long)info[0]; A->err_max = info[1];
		if (r != SOL_OK) { A->stop_reason = SOL_RUN_ERROR; return SOL_ERR; }
		A->h_did = hDid;
		A->steps++; counter++;
		A->step_counter = counter;
		if (A->records != nullptr) {
			A->records[4 * (size_t)(A->steps - 1) + 0] = A->time;
			A->records[4 * (size_t)(A->steps - 1) + 1] = hDid;
			A->records[4 * (size_t)(A->steps - 1) + 2] = A->h_next;
			A->records[4 * (size_t)(A->steps - 1) + 3] = h_trial;
		}
		if (A->ejection > 0 || A->hit_centrum > 0 || A->collision_factor > 0) {
			if (count_events(c, A->ejection, A->hit_centrum, A->collision_factor, A->event_counts) != SOL_OK) return SOL_ERR;
			if (A->event_counts[0] > 0 || A->event_counts[1] > 0 || A->event_counts[2] > 0) { A->stop_reason = SOL_RUN_EVENT; break; }
		}
		const double ls = lastSave + hDid;
		const double actualTime = A->millenium_days + A->time;
		if (fabs(actualTime) >= fabs(A->length)) { A->stop_reason = SOL_RUN_END; break; }
		double hn = A->h_next;
		if (fabs(actualTime + hn) > fabs(A->length)) hn = A->length - actualTime;
		if (fabs(ls) >= fabs(A->output)) { A->stop_reason = SOL_RUN_SAVE; break; }
		if (fabs(ls + hn) > fabs(A->output)) hn = A->output - ls;
		lastSave = ls; A->h_next = hn; A->last_save = ls;
		if (A->flush_every > 0 && counter % A->flush_every == 0) {
			launch_flush_tiny(c, c.y, A->flush_threshold);
			launch_flush_tiny(c, c.y0, A->flush_threshold);
		}
	}
	SOL_CUDA(cudaStreamSynchronize(c.stream));
	return SOL_OK;
}

// Device flag reduction shared by sol_detect_events and sol_run: counts into evCountHost (see sol_detect_events).
static int count_events(Ctx &c, double ejection, double hit_centrum, double collision_factor, int counts_out[3])
{
	// thresholds exactly as Simulator.cpp:626-629
	const double e3 = ejection > 0 ? 1.0 / (ejection * ejection * ejection) : 0.0;
	const double h3 = hit_centrum > 0 ? 1.0 / (hit_centrum * hit_centrum * hit_centrum) : 0.0;
	SOL_CUDA(cudaMemsetAsync(c.evCount, 0, 4 * sizeof(int), c.stream));
	launch_detect_events(c, e3, h3, ejection > 0, hit_centrum > 0, collision_factor);
	// evCount[0..3] = this rank's counts (its indices stay with it), evCount[4..7] = global counts
	if (c.nranks > 1)
		SOL_NCCL(g_nccl.AllReduce(c.evCount, c.evCount + 4, 4, ncclInt, ncclSum, (ncclComm_t)c.nccl, c.stream));
	else
		SOL_CUDA(cudaMemcpyAsync(c.evCount + 4, c.evCount, 4 * sizeof(int), cudaMemcpyDeviceToDevice, c.stream));
	SOL_CUDA(cudaMemcpyAsync(c.evCountHost, c.evCount, 8 * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
	SOL_CUDA(cudaStreamSynchronize(c.stream));
	counts_out[0] = c.evCountHost[4]; counts_out[1] = c.evCountHost[5]; counts_out[2] = c.evCountHost[6];
	return SOL_OK;
}

int sol_detect_events(sol_ctx *h, double ejection, double hit_centrum, double collision_factor, int counts_out[3])
{
	if (!h || !counts_out) return SOL_ERR;
	if (h->multi) {
		const int nr = (int)h->multi->ranks.size();
		std::vector<int> cnt(3 * (size_t)nr, 0);
		const int rc = fan_out(h, [&](sol_ctx *r, int rank) { return sol_detect_events(r, ejection, hit_centrum, collision_factor, &cnt[3 * (size_t)rank]); });
		for (int q = 0; q < 3; q++) counts_out[q] = cnt[q];      // global counts, the same on every rank
		return rc;
	}
	Ctx &c = h->c;
	SOL_CUDA(cudaSetDevice(c.device));
	return count_events(c, ejection, hit_centrum, collision_factor, counts_out);
}

int sol_event_indices(sol_ctx *h, int kind, int *idx_out, int cap, int *n_out)
{
	if (!h || kind < 0 || kind > 2 || !n_out) return SOL_ERR;
	if (h->multi) {
		// every rank holds the candidates among its own sinks: collect and merge them into scan order
		const int nr = (int)h->multi->ranks.size();
		std::vector<std::vector<int>> part(nr);
		const int rc = fan_out(h, [&](sol_ctx *r, int rank) {
			int m = 0;
			if (sol_event_indices(r, kind, nullptr, 0, &m) != SOL_OK) return SOL_ERR;
			part[rank].resize(m);
			return m > 0 ? sol_event_indices(r, kind, part[rank].data(), m, &m) : SOL_OK;
		});
		if (rc != SOL_OK) return rc;
		std::vector<int> all;
		for (auto &v : part) all.insert(all.end(), v.begin(), v.end());
		std::sort(all.begin(), all.end());
		*n_out = (int)all.size();
		if (idx_out) memcpy(idx_out, all.data(), std::min<size_t>(all.size(), (size_t)std::max(cap, 0)) * sizeof(int));
		return SOL_OK;
	}
	Ctx &c = h->c;
	SOL_CUDA(cudaSetDevice(c.device));
	int n = c.evCountHost[kind];
	*n_out = n;
	int m = std::min(n, cap);
	if (m > 0 && idx_out) {
		std::vector<int> tmp(n);
		SOL_CUDA(cudaMemcpyAsync(tmp.data(), c.evIdx + (size_t)kind * c.ld, n * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
		SOL_CUDA(cudaStreamSynchronize(c.stream));
		std::sort(tmp.begin(), tmp.end());   // scan order of the reference's loops
		memcpy(idx_out, tmp.data(), m * sizeof(int));
	}
	return SOL_OK;
}

// records for the given (sorted) ejection / hit-centrum index lists, from the full accepted state this context holds
static int records_from_lists(Ctx &c, const std::vector<int> &ej, const std::vector<int> &hc, double time, int first_event_id,
                              void *records, int capacity, int *n_records)
{
	const int ne = (int)ej.size(), nh = (int)hc.size(), m = ne + nh;
	*n_records = m;
	if (m == 0 || records == nullptr) return SOL_OK;
	if (capacity < m) { c.err = "sol_event_records: buffer too small"; return SOL_ERR; }
	// TwoBodyAffair ids count up in the order the reference constructs the objects: one scan over the bodies, the
	// ejection test before the hit-centrum test (Simulator.cpp:631-646, TwoBodyAffair.cpp:11); the records are then
	// written list by list (ejections, hit centrums)
	std::vector<int> table(3 * (size_t)m);
	int a = 0, b = 0, id = first_event_id;
	while (a < ne || b < nh) {
		const bool take_ej = b >= nh || (a < ne && ej[a] <= hc[b]);
		const int k = take_ej ? a : ne + b;
		table[3 * (size_t)k + 0] = take_ej ? ej[a] : hc[b];
		table[3 * (size_t)k + 1] = id++;
		table[3 * (size_t)k + 2] = take_ej ? 0 : 1;                 // EventType: Ejection = 0, HitCentrum = 1 (TwoBodyAffair.h:6-13)
		if (take_ej) a++; else b++;
	}
	const size_t rec_bytes = 120 * (size_t)m, tab_bytes = 3 * (size_t)m * sizeof(int);
	if (ensure_stage(c, (rec_bytes + tab_bytes + 15) / 8) != SOL_OK) return SOL_ERR;
	int *d_table = reinterpret_cast<int *>(reinterpret_cast<unsigned char *>(c.stage_aos) + rec_bytes);
	SOL_CUDA(cudaMemcpyAsync(d_table, table.data(), tab_bytes, cudaMemcpyHostToDevice, c.stream));
	launch_event_records(c, d_table, c.stage_aos, m, time);
	SOL_CUDA(cudaMemcpyAsync(records, c.stage_aos, rec_bytes, cudaMemcpyDeviceToHost, c.stream));
	SOL_CUDA(cudaStreamSynchronize(c.stream));
	return SOL_OK;
}

int sol_event_records(sol_ctx *h, double time, int first_event_id, void *records, int capacity, int *n_records)
{
	if (!h || !n_records) return SOL_ERR;
	if (h->multi) {
		// the flagged bodies' phases live on their own ranks: merge the index lists, gather the accepted state, and let
		// rank 0 assemble the records
		int ne = 0, nh = 0;
		if (sol_event_indices(h, 0, nullptr, 0, &ne) != SOL_OK || sol_event_indices(h, 1, nullptr, 0, &nh) != SOL_OK) return SOL_ERR;
		std::vector<int> ej(ne), hc(nh);
		if (ne && sol_event_indices(h, 0, ej.data(), ne, &ne) != SOL_OK) return SOL_ERR;
		if (nh && sol_event_indices(h, 1, hc.data(), nh, &nh) != SOL_OK) return SOL_ERR;
		*n_records = ne + nh;
		if (ne + nh == 0 || records == nullptr) return SOL_OK;
		if (sol_gather_state(h) != SOL_OK) return SOL_ERR;
		sol_ctx *r0 = h->multi->ranks[0];
		return fan_out(h, [&](sol_ctx *r, int rank) {
			if (rank != 0) return SOL_OK;
			cudaSetDevice(r0->c.device);
			return records_from_lists(r->c, ej, hc, time, first_event_id, records, capacity, n_records);
		});
	}
	Ctx &c = h->c;
	if (c.nranks > 1) { c.err = "sol_event_records: one process per GPU keeps the flagged bodies' phases on their own ranks; use a sol_create_multi handle"; return SOL_ERR; }
	SOL_CUDA(cudaSetDevice(c.device));
	const int ne = c.evCountHost[0], nh = c.evCountHost[1];
	// indices of both lists (compacted in arbitrary order on the device) -> scan order, like sol_event_indices
	std::vector<int> ej(ne), hc(nh);
	if (ne) SOL_CUDA(cudaMemcpyAsync(ej.data(), c.evIdx, (size_t)ne * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
	if (nh) SOL_CUDA(cudaMemcpyAsync(hc.data(), c.evIdx + (size_t)c.ld, (size_t)nh * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
	SOL_CUDA(cudaStreamSynchronize(c.stream));
	std::sort(ej.begin(), ej.end());
	std::sort(hc.begin(), hc.end());
	return records_from_lists(c, ej, hc, time, first_event_id, records, capacity, n_records);
}

// rows [lo, hi) of a state array: host AoS6 <-> device planes (a sharded rank only ever holds its own rows up to date)
static int xfer_planes(Ctx &c, double *planes, void *host, bool down, int lo, int hi)
{
	if (hi <= lo) return SOL_OK;
	const size_t nb = (size_t)(hi - lo);
	double *hp = (double *)host + 6 * (size_t)lo;
	if (ensure_stage(c, 6 * nb) != SOL_OK) return SOL_ERR;
	if (down) {
		launch_planes_to_aos(c, planes + lo, c.stage_aos, hi - lo);
		SOL_CUDA(cudaMemcpyAsync(hp, c.stage_aos, 6 * nb * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
	} else {
		SOL_CUDA(cudaMemcpyAsync(c.stage_aos, hp, 6 * nb * sizeof(double), cudaMemcpyHostToDevice, c.stream));
		launch_aos_to_planes(c, c.stage_aos, planes + lo, hi - lo);
	}
	SOL_CUDA(cudaStreamSynchronize(c.stream));
	return SOL_OK;
}

// gas caches are 3 planes of stride ld on the device, AoS-3 on the host (Acceleration.h:46-48); entries [qlo, qhi)
static int xfer_cache(Ctx &c, double *planes, int qlo, int qhi, void *host, bool down)
{
	if (qhi <= qlo) return SOL_OK;
	std::vector<double> tmp(3 * (size_t)c.ld);
	double *hp = (double *)host;
	if (down) {
		SOL_CUDA(cudaMemcpyAsync(tmp.data(), planes, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
		SOL_CUDA(cudaStreamSynchronize(c.stream));
		for (int q = qlo; q < qhi; q++) for (int k = 0; k < 3; k++) hp[3 * q + k] = tmp[(size_t)k * c.ld + q];
	} else {
		SOL_CUDA(cudaMemcpyAsync(tmp.data(), planes, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
		SOL_CUDA(cudaStreamSynchronize(c.stream));
		for (int q = qlo; q < qhi; q++) for (int k = 0; k < 3; k++) tmp[(size_t)k * c.ld + q] = hp[3 * q + k];
		SOL_CUDA(cudaMemcpyAsync(planes, tmp.data(), tmp.size() * sizeof(double), cudaMemcpyHostToDevice, c.stream));
		SOL_CUDA(cudaStreamSynchronize(c.stream));
	}
	return SOL_OK;
}

// Transfers rows [lo, hi) of array `what` (whole array: lo = 0, hi = n).
static int xfer_rows(sol_ctx *h, int what, void *host, bool down, int lo, int hi)
{
	Ctx &c = h->c;
	if (c.cnt.n <= 0) { c.err = "transfer before sol_set_bodies"; return SOL_ERR; }
	SOL_CUDA(cudaSetDevice(c.device));
	const Counts &n = c.cnt;
	auto cache = [&](double *planes, int class_lo, int class_hi) {
		return xfer_cache(c, planes, std::max(lo, class_lo) - class_lo, std::min(hi, class_hi) - class_lo, host, down);
	};
	char *dev = nullptr; size_t elem = 0;
	switch (what) {
	case SOL_Y0: return xfer_planes(c, c.y0, host, down, lo, hi);
	case SOL_Y: return xfer_planes(c, c.y, host, down, lo, hi);
	case SOL_ACCEL: return xfer_planes(c, c.k[0], host, down, lo, hi);
	case SOL_YSCALE: return xfer_planes(c, c.yscale, host, down, lo, hi);
	case SOL_RM3: dev = (char *)c.rm3; elem = sizeof(double); break;
	case SOL_NN_INDEX: dev = (char *)c.nnIdx; elem = sizeof(int); break;
	case SOL_NN_DISTANCE: dev = (char *)c.nnDist; elem = sizeof(double); break;
	case SOL_MIGTYPE: dev = (char *)c.migType; elem = sizeof(int); break;
	case SOL_MASS: dev = (char *)c.mass; elem = sizeof(double); break;
	case SOL_RADIUS: dev = (char *)c.radius; elem = sizeof(double); break;
	case SOL_DENSITY: dev = (char *)c.density; elem = sizeof(double); break;
	case SOL_CD: dev = (char *)c.cD; elem = sizeof(double); break;
	case SOL_GAMMA_STOKES: dev = (char *)c.gS; elem = sizeof(double); break;
	case SOL_GAMMA_EPSTEIN: dev = (char *)c.gE; elem = sizeof(double); break;
	case SOL_MIGSTOPAT: dev = (char *)c.migStop; elem = sizeof(double); break;
	case SOL_TYPE: dev = (char *)c.type; elem = sizeof(int); break;
	case SOL_ID: dev = (char *)c.id; elem = sizeof(int); break;
	case SOL_ACCEL_GASDRAG: return cache(c.aGas, n.M, n.M + n.s + n.l);
	case SOL_ACCEL_MIGTYPE1: return cache(c.aMig1, n.c + n.g, n.M);
	case SOL_ACCEL_MIGTYPE2: return cache(c.aMig2, n.c, n.c + n.g);
	default: c.err = "unknown array id"; return SOL_ERR;
	}
	if (hi > lo) {
		char *hp = (char *)host + (size_t)lo * elem;
		const size_t bytes = (size_t)(hi - lo) * elem;
		if (down) SOL_CUDA(cudaMemcpyAsync(hp, dev + (size_t)lo * elem, bytes, cudaMemcpyDeviceToHost, c.stream));
		else SOL_CUDA(cudaMemcpyAsync(dev + (size_t)lo * elem, hp, bytes, cudaMemcpyHostToDevice, c.stream));
		SOL_CUDA(cudaStreamSynchronize(c.stream));
	}
	if (!down && what == SOL_MASS && lo == 0) { c.mass0 = ((const double *)host)[0]; refresh_gas(c); c.cfg_epoch++; }
	return SOL_OK;
}

static bool sharded_array(int what)
{   // arrays of which a rank keeps only its own sinks' rows current
	switch (what) {
	case SOL_Y0: case SOL_Y: case SOL_ACCEL: case SOL_YSCALE: case SOL_RM3: case SOL_NN_INDEX: case SOL_NN_DISTANCE: case SOL_MIGTYPE:
	case SOL_ACCEL_GASDRAG: case SOL_ACCEL_MIGTYPE1: case SOL_ACCEL_MIGTYPE2: return true;
	default: return false;
	}
}

static int xfer(sol_ctx *h, int what, void *host, bool down)
{
	if (!h || !host) return SOL_ERR;
	if (h->multi) {
		// uploads go to every rank in full; downloads of sharded arrays come rank by rank, each into its rows of the
		// caller's array; everything else is replicated and comes from rank 0
		if (!down) return fan_out(h, [&](sol_ctx *r, int) { return xfer_rows(r, what, host, false, 0, r->c.cnt.n); });
		if (sharded_array(what)) return fan_out(h, [&](sol_ctx *r, int) { return xfer_rows(r, what, host, true, r->c.lo, r->c.hi); });
		return fan_out(h, [&](sol_ctx *r, int rank) { return rank == 0 ? xfer_rows(r, what, host, true, 0, r->c.cnt.n) : SOL_OK; });
	}
	return xfer_rows(h, what, host, down, 0, h->c.cnt.n);
}

int sol_download(sol_ctx *h, int what, void *host) { return xfer(h, what, host, true); }
int sol_upload(sol_ctx *h, int what, const void *host) { return xfer(h, what, const_cast<void *>(host), false); }

int sol_integrals(sol_ctx *h, double out[16])
{
	if (!h || !out) return SOL_ERR;
	if (h->multi) {
		const int nr = (int)h->multi->ranks.size();
		std::vector<double> all(16 * (size_t)nr, 0.0);
		const int rc = fan_out(h, [&](sol_ctx *r, int rank) { return sol_integrals(r, &all[16 * (size_t)rank]); });
		for (int q = 0; q < 16; q++) out[q] = all[q];           // all-reduced: the same on every rank
		return rc;
	}
	Ctx &c = h->c;
	if (c.cnt.n <= 0) { c.err = "sol_integrals before sol_set_bodies"; return SOL_ERR; }
	SOL_CUDA(cudaSetDevice(c.device));
	if (c.nranks > 1 && sol_gather_state(h) != SOL_OK) return SOL_ERR;   // the potential needs every body's accepted position
	launch_integrals(c);
	if (c.nranks > 1)
		SOL_NCCL(g_nccl.AllReduce(c.integralsDev, c.integralsDev, 12, ncclDouble, ncclSum, (ncclComm_t)c.nccl, c.stream));
	SOL_CUDA(cudaMemcpyAsync(c.integralsHost, c.integralsDev, 12 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
	SOL_CUDA(cudaStreamSynchronize(c.stream));
	const double *s = c.integralsHost;
	const double M = s[0];                                   // Calculate::TotalMass: massive bodies only (Calculate.cpp:64-71)
	out[0] = M;
	for (int j = 0; j < 6; j++) out[1 + j] = s[1 + j] / M;   // PhaseOfBC, :73-92
	out[7] = sqrt(out[1] * out[1] + out[2] * out[2] + out[3] * out[3]);
	out[8] = sqrt(out[4] * out[4] + out[5] * out[5] + out[6] * out[6]);
	out[9] = s[7]; out[10] = s[8]; out[11] = s[9];            // AngularMomentum, :106-124
	out[12] = sqrt(s[7] * s[7] + s[8] * s[8] + s[9] * s[9]);
	out[13] = s[10];                                         // KineticEnergy, :161-172
	out[14] = 0.5 * kGauss2 * s[11];                         // PotentialEnergy, :139-159
	out[15] = out[13] - out[14];
	return SOL_OK;
}

int sol_flush_tiny(sol_ctx *h, double threshold)
{
	if (!h) return SOL_ERR;
	SOL_FANOUT(h, sol_flush_tiny(r, threshold));
	Ctx &c = h->c;
	if (c.cnt.n <= 0) return SOL_OK;
	SOL_CUDA(cudaSetDevice(c.device));
	launch_flush_tiny(c, c.y, threshold);
	launch_flush_tiny(c, c.y0, threshold);
	SOL_CUDA(cudaStreamSynchronize(c.stream));
	return SOL_OK;
}

// ---- (f) row 4: the loader's Kepler solves ----
int sol_elements_to_phases(sol_ctx *h, int n, const double *mu, const double *elements, double *phases, int *n_failed)
{
	if (!h || n < 0 || (n > 0 && (!mu || !elements || !phases))) return SOL_ERR;
	if (h->multi) {     // independent of the loaded system: rank 0's device does the batch
		const int rc = fan_out(h, [&](sol_ctx *r, int rank) { return rank == 0 ? sol_elements_to_phases(r, n, mu, elements, phases, n_failed) : SOL_OK; });
		if (rc != SOL_OK) h->c.err = h->multi->ranks[0]->c.err;
		return rc;
	}
	Ctx &c = h->c;
	if (n_failed) *n_failed = 0;
	if (n == 0) return SOL_OK;
	SOL_CUDA(cudaSetDevice(c.device));
	// scratch (independent of any loaded system): mu[n] | el[6n] | out[6n] | failed[n]
	const size_t nb = (size_t)n;
	double *buf = nullptr;
	SOL_CUDA(cudaMalloc((void **)&buf, (14 * nb) * sizeof(double)));
	double *d_mu = buf, *d_el = buf + nb, *d_out = buf + 7 * nb;
	int *d_failed = reinterpret_cast<int *>(buf + 13 * nb);
	cudaError_t e = cudaMemcpyAsync(d_mu, mu, nb * sizeof(double), cudaMemcpyHostToDevice, c.stream);
	if (e == cudaSuccess) e = cudaMemcpyAsync(d_el, elements, 6 * nb * sizeof(double), cudaMemcpyHostToDevice, c.stream);
	if (e == cudaSuccess) e = cudaMemcpyAsync(d_out, phases, 6 * nb * sizeof(double), cudaMemcpyHostToDevice, c.stream);   // failed rows stay as they are
	if (e == cudaSuccess) {
		launch_elements_to_phases(c, d_mu, d_el, d_out, d_failed, n);
		e = cudaMemcpyAsync(phases, d_out, 6 * nb * sizeof(double), cudaMemcpyDeviceToHost, c.stream);
	}
	std::vector<int> failed(nb);
	if (e == cudaSuccess) e = cudaMemcpyAsync(failed.data(), d_failed, nb * sizeof(int), cudaMemcpyDeviceToHost, c.stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
	cudaFree(buf);
	if (e != cudaSuccess) { c.err = std::string("sol_elements_to_phases: ") + cudaGetErrorString(e); return SOL_ERR; }
	int bad = 0, first = -1;
	for (int i = 0; i < n; i++) if (failed[i]) { if (first < 0) first = i; bad++; }
	if (n_failed) *n_failed = bad;
	if (bad > 0) {
		// Ephemeris.cpp:207-209 / Simulation.cpp:149-152
		c.err = "Could not compute the excentric anomaly E! The phase could not be computed for body with index: " + std::to_string(first) + "!";
		return SOL_ERR;
	}
	return SOL_OK;
}

// ---- (f) row 3: removal and patching of bodies on the device ----
int sol_remove_bodies(sol_ctx *h, const int *indices, int count)
{
	if (!h || (count > 0 && !indices) || count < 0) return SOL_ERR;
	SOL_FANOUT(h, sol_remove_bodies(r, indices, count));
	Ctx &c = h->c;
	if (count == 0) return SOL_OK;
	if (c.cnt.n <= 0) { c.err = "sol_remove_bodies before sol_set_bodies"; return SOL_ERR; }
	SOL_CUDA(cudaSetDevice(c.device));
	std::vector<int> r(indices, indices + count);
	std::sort(r.begin(), r.end());
	for (int m = 0; m < count; m++) {
		if (r[m] < 1 || r[m] >= c.cnt.n || (m > 0 && r[m] == r[m - 1])) {
			c.err = "sol_remove_bodies: indices must be distinct and in [1, n) (body 0 is the central body)";
			return SOL_ERR;
		}
	}
	// every rank compacts the full accepted state, so it has to hold it first
	if (c.nranks > 1 && sol_gather_state(h) != SOL_OK) return SOL_ERR;
	// NBodies::UpdateAfterRemove (NBodies.cpp:80-113): one count per removed body, by its type
	std::vector<int> types(c.cnt.n);
	SOL_CUDA(cudaMemcpyAsync(types.data(), c.type, (size_t)c.cnt.n * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
	SOL_CUDA(cudaStreamSynchronize(c.stream));
	Counts n = c.cnt;
	for (int m = 0; m < count; m++) {
		switch (types[r[m]]) {
		case 2: n.g--; break;
		case 3: n.r--; break;
		case 4: n.p--; break;
		case 5: n.s--; break;
		case 6: n.l--; break;
		case 7: n.t--; break;
		default: c.err = "Unknown or undefined Body Type!"; return SOL_ERR;
		}
	}
	n.n = c.cnt.n - count;
	n.M = n.c + n.g + n.r + n.p;
	std::vector<int> adj(count);
	for (int m = 0; m < count; m++) adj[m] = r[m] - m;
	int *adj_dev = c.evIdx;                                   // scratch: [3][ld] ints
	if ((size_t)count > 3 * (size_t)c.ld) { c.err = "sol_remove_bodies: too many indices"; return SOL_ERR; }
	SOL_CUDA(cudaMemcpyAsync(adj_dev, adj.data(), (size_t)count * sizeof(int), cudaMemcpyHostToDevice, c.stream));
	// y0 (6 planes) through ytmp; the per-body arrays the reference moves (Simulator.cpp:757-769) through one plane of ytmp.
	// cD and migStopAt keep their slots, exactly like the reference; y, rm3 and the NN arrays are not moved either.
	launch_compact(c, c.y0, c.ytmp, n.n, 6, adj_dev, count);
	SOL_CUDA(cudaMemcpyAsync(c.y0, c.ytmp, 6 * (size_t)c.ld * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
	double *dscratch = c.ytmp;
	double *darr[5] = {c.mass, c.radius, c.density, c.gS, c.gE};
	for (double *a : darr) {
		launch_compact(c, a, dscratch, n.n, 1, adj_dev, count);
		SOL_CUDA(cudaMemcpyAsync(a, dscratch, (size_t)n.n * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
	}
	int *iscratch = reinterpret_cast<int *>(c.ytmp + c.ld);
	int *iarr[3] = {c.id, c.type, c.migType};
	for (int *a : iarr) {
		launch_compact(c, a, iscratch, n.n, adj_dev, count);
		SOL_CUDA(cudaMemcpyAsync(a, iscratch, (size_t)n.n * sizeof(int), cudaMemcpyDeviceToDevice, c.stream));
	}
	SOL_CUDA(cudaStreamSynchronize(c.stream));
	c.cnt = n;
	c.cfg_epoch++;
	if (c.nranks > 1) shard_of(n.n, c.nranks, c.rank, c.lo, c.hi);
	else { c.lo = 0; c.hi = n.n; }
	return SOL_OK;
}

int sol_patch_body(sol_ctx *h, int index, const double y0[6], double mass, double radius, double density)
{
	if (!h || !y0) return SOL_ERR;
	SOL_FANOUT(h, sol_patch_body(r, index, y0, mass, radius, density));
	Ctx &c = h->c;
	if (index < 0 || index >= c.cnt.n) { c.err = "sol_patch_body: index out of range"; return SOL_ERR; }
	SOL_CUDA(cudaSetDevice(c.device));
	for (int k = 0; k < 6; k++)
		SOL_CUDA(cudaMemcpyAsync(c.y0 + (size_t)k * c.ld + index, &y0[k], sizeof(double), cudaMemcpyHostToDevice, c.stream));
	SOL_CUDA(cudaMemcpyAsync(c.mass + index, &mass, sizeof(double), cudaMemcpyHostToDevice, c.stream));
	SOL_CUDA(cudaMemcpyAsync(c.radius + index, &radius, sizeof(double), cudaMemcpyHostToDevice, c.stream));
	SOL_CUDA(cudaMemcpyAsync(c.density + index, &density, sizeof(double), cudaMemcpyHostToDevice, c.stream));
	SOL_CUDA(cudaStreamSynchronize(c.stream));
	if (index == 0) { c.mass0 = mass; refresh_gas(c); c.cfg_epoch++; }      // the gas constants depend on the star's mass
	return SOL_OK;
}

// ---- (f) row 2: snapshot records ----
static int pack_phases_to(Ctx &c, double time, void *host, size_t bytes)
{
	if (ensure_stage(c, (bytes + 15) / 16 * 2) != SOL_OK) return SOL_ERR;   // the kernel stores whole 16-byte vectors
	launch_pack_phases(c, c.y0, time, c.stage_aos);
	SOL_CUDA(cudaMemcpyAsync(host, c.stage_aos, bytes, cudaMemcpyDeviceToHost, c.stream));
	SOL_CUDA(cudaStreamSynchronize(c.stream));
	return SOL_OK;
}

static int write_phases_local(sol_ctx *h, const char *path, double time);

int sol_pack_phases(sol_ctx *h, double time, void *host, size_t capacity, size_t *nbytes)
{
	if (!h || !nbytes) return SOL_ERR;
	if (h->multi) {
		// all ranks gather the accepted state; rank 0 assembles the record
		sol_ctx *r0 = h->multi->ranks[0];
		const size_t bytes = 12 + 52 * (size_t)std::max(r0->c.cnt.n, 0);
		*nbytes = bytes;
		if (host == nullptr) return SOL_OK;
		if (capacity < bytes) { h->c.err = "sol_pack_phases: buffer too small"; return SOL_ERR; }
		if (sol_gather_state(h) != SOL_OK) return SOL_ERR;
		return fan_out(h, [&](sol_ctx *r, int rank) {
			if (rank != 0) return SOL_OK;
			cudaSetDevice(r->c.device);
			return pack_phases_to(r->c, time, host, bytes);
		});
	}
	Ctx &c = h->c;
	const size_t bytes = 12 + 52 * (size_t)std::max(c.cnt.n, 0);
	*nbytes = bytes;
	if (host == nullptr) return SOL_OK;                     // size query
	if (capacity < bytes) { c.err = "sol_pack_phases: buffer too small"; return SOL_ERR; }
	SOL_CUDA(cudaSetDevice(c.device));
	if (c.nranks > 1 && sol_gather_state(h) != SOL_OK) return SOL_ERR;
	return pack_phases_to(c, time, host, bytes);
}

int sol_write_phases(sol_ctx *h, const char *path, double time)
{
	if (!h || !path) return SOL_ERR;
	if (h->multi) {
		if (sol_gather_state(h) != SOL_OK) return SOL_ERR;
		return fan_out(h, [&](sol_ctx *r, int rank) { return rank == 0 ? write_phases_local(r, path, time) : SOL_OK; });
	}
	Ctx &c = h->c;
	SOL_CUDA(cudaSetDevice(c.device));
	if (c.nranks > 1 && sol_gather_state(h) != SOL_OK) return SOL_ERR;
	return write_phases_local(h, path, time);
}

// the record of the state this context holds (already gathered when sharded), appended to `path` with one write
static int write_phases_local(sol_ctx *h, const char *path, double time)
{
	Ctx &c = h->c;
	const size_t bytes = 12 + 52 * (size_t)std::max(c.cnt.n, 0);
	SOL_CUDA(cudaSetDevice(c.device));
	if (c.pin_cap < bytes) {
		if (c.pin) cudaFreeHost(c.pin);
		c.pin = nullptr; c.pin_cap = 0;
		SOL_CUDA(cudaHostAlloc(&c.pin, bytes, cudaHostAllocDefault));
		c.pin_cap = bytes;
	}
	if (pack_phases_to(c, time, c.pin, bytes) != SOL_OK) return SOL_ERR;
	// ios::out | ios::app | ios::binary (BinaryFileAdapter.cpp:114): append, create if missing - one write
	const int fd = open(path, O_WRONLY | O_CREAT | O_APPEND, 0644);
	if (fd < 0) { c.err = std::string("sol_write_phases: The file '") + path + "' could not opened!"; return SOL_ERR; }
	const char *p = (const char *)c.pin;
	size_t left = bytes;
	while (left > 0) {
		const ssize_t k = write(fd, p, left);
		if (k < 0) { close(fd); c.err = "sol_write_phases: An error occurred during writing the phase!"; return SOL_ERR; }
		p += k; left -= (size_t)k;
	}
	close(fd);
	return SOL_OK;
}

// ---- multi-GPU ----
int sol_nccl_unique_id(void *out128)
{
	std::string err;
	if (!out128 || !g_nccl.load(err)) { g_create_error = err; return SOL_ERR; }
	ncclUniqueId id;
	if (g_nccl.GetUniqueId(&id) != ncclSuccess) { g_create_error = "ncclGetUniqueId failed"; return SOL_ERR; }
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
	memcpy(out128, &id, 128);
	return SOL_OK;
}

int sol_dist_init(sol_ctx *h, int rank, int nranks, const void *unique_id128)
{
	if (!h || !unique_id128 || nranks < 1 || rank < 0 || rank >= nranks) return SOL_ERR;
	if (h->multi) { h->c.err = "sol_dist_init: a sol_create_multi handle owns its communicator"; return SOL_ERR; }
	Ctx &c = h->c;
	SOL_CUDA(cudaSetDevice(c.device));
	if (nranks == 1) { c.rank = 0; c.nranks = 1; return SOL_OK; }
	// the nearest-neighbour merge gathers one candidate row per rank into the kSymRounds-row slot buffers
	if (nranks > kSymRounds) { c.err = "sol_dist_init: at most " + std::to_string(kSymRounds) + " ranks are supported"; return SOL_ERR; }
	if (c.cnt.n > 0) { c.err = "sol_dist_init: join the communicator before sol_set_bodies (the device layout depends on the number of ranks)"; return SOL_ERR; }
	if (!g_nccl.load(c.err)) return SOL_ERR;
	ncclUniqueId id;
	memcpy(&id, unique_id128, 128);
	ncclComm_t comm;
	SOL_NCCL(g_nccl.CommInitRank(&comm, nranks, id, rank));
	c.nccl = comm; c.rank = rank; c.nranks = nranks;
	if (c.cnt.n > 0) shard_of(c.cnt.n, nranks, rank, c.lo, c.hi);
	c.alloc_n = 0;      // the plane stride depends on the number of ranks: the next sol_set_bodies re-allocates
	c.cfg_epoch++;
	return SOL_OK;
}

int sol_shard_of(int n, int nranks, int rank, int *lo, int *hi)
{
	if (!lo || !hi || n < 0 || nranks < 1 || rank < 0 || rank >= nranks) return SOL_ERR;
	shard_of(n, nranks, rank, *lo, *hi);
	return SOL_OK;
}

int sol_sym_round_pair(int nb, int round, int p, int *q)
{
	if (!q || nb < 1 || round < 0 || round > nb / 2 || p < 0 || p >= nb) return -1;
	if (2 * round == nb && p >= nb / 2) return 0;      // same test as sym_pair_kernel
	*q = (p + round) % nb;
	return 1;
}

int sol_sym_rounds_of_rank(int nb, int nranks, int rank, int *lo, int *hi)
{
	if (!lo || !hi || nb < 1 || nranks < 1 || rank < 0 || rank >= nranks) return SOL_ERR;
	const int rounds_total = nb / 2 + 1;               // same split as eval_force
	*lo = (int)((long long)rounds_total * rank / nranks);
	*hi = (int)((long long)rounds_total * (rank + 1) / nranks);
	return SOL_OK;
}

int sol_sym_work_of_rank(int nb, int nranks, int rank, int out4[4])
{
	if (!out4 || nb < 1 || nranks < 1 || rank < 0 || rank >= nranks) return SOL_ERR;
	sym_work_of_rank(nb, nranks, rank, out4);
	return SOL_OK;
}

int sol_plan_pairs(int sinks, int sources, int sinks_all, int out3[3])
{
	if (!out3 || sinks < 1 || sources < 1) return SOL_ERR;
	PairLaunch pl{};
	plan_pairs(sinks, sources, pl, kMaxSplit, sinks_all);
	out3[0] = pl.sinks_per_thread; out3[1] = pl.splits; out3[2] = pl.chunk;
	return SOL_OK;
}

int sol_shard_range(const sol_ctx *h, int *lo, int *hi)
{
	if (!h || !lo || !hi) return SOL_ERR;
	if (h->multi) { *lo = 0; *hi = h->multi->ranks[0]->c.cnt.n; return SOL_OK; }    // the handle as a whole integrates every sink
	*lo = h->c.lo; *hi = h->c.hi;
	return SOL_OK;
}

int sol_gather_state(sol_ctx *h)
{
	if (!h) return SOL_ERR;
	SOL_FANOUT(h, sol_gather_state(r));
	Ctx &c = h->c;
	if (c.nranks <= 1) return SOL_OK;
	SOL_CUDA(cudaSetDevice(c.device));
	SOL_NCCL(g_nccl.GroupStart());
	for (int r = 0; r < c.nranks; r++) {
		int lo, hi;
		shard_of(c.cnt.n, c.nranks, r, lo, hi);
		if (hi <= lo) continue;
		for (int p = 0; p < 6; p++) {
			double *ptr = c.y0 + (size_t)p * c.ld + lo;
			SOL_NCCL(g_nccl.Broadcast(ptr, ptr, (size_t)(hi - lo), ncclDouble, r, (ncclComm_t)c.nccl, c.stream));
		}
	}
	SOL_NCCL(g_nccl.GroupEnd());
	SOL_CUDA(cudaStreamSynchronize(c.stream));
	return SOL_OK;
}

// ---- measurement ----
int sol_time_gravity_kernel(sol_ctx *h, int reps, float *ms_out, double *pairs_out)
{
	if (!h || reps < 1 || !ms_out) return SOL_ERR;
	if (h->multi) return fan_out(h, [&](sol_ctx *r, int rank) { return rank == 0 ? sol_time_gravity_kernel(r, reps, ms_out, pairs_out) : SOL_OK; });
	Ctx &c = h->c;
	if (c.cnt.n <= 0) { c.err = "no bodies"; return SOL_ERR; }
	SOL_CUDA(cudaSetDevice(c.device));
	const Counts &n = c.cnt;
	const bool bary = c.barycentric != 0;
	const int jlo = bary ? 0 : 1;
	const int src_hi = bary ? n.M : n.M + n.s;
	launch_prep_sources(c, c.y0, 0, src_hi);
	const bool track = c.nn_mode == 1;
	const bool use_sym = c.sym_mode != 0 && c.nranks == 1 && (n.M - jlo) >= (c.sym_mode == 1 ? kSymMinBodies : kSymAutoBodies);   // kernel timing helper: single GPU
	PairLaunch pl{};
	SymLaunch L{};
	if (use_sym) {
		if (alloc_sym(c) != SOL_OK) return SOL_ERR;
		L.r0 = jlo; L.nR = n.M - jlo; L.nb = (L.nR + kSymB - 1) / kSymB; L.track_nn = track; L.tie_ge = bary;
		L.round_first = 0; L.p_first_lo = 0; L.round_last = L.nb / 2; L.p_last_hi = L.nb;
	} else {
		pl.track_nn = track; pl.tie_prefers_larger_j = bary;
		pl.i_lo = std::max(c.lo, jlo); pl.i_hi = c.hi; pl.j_lo = jlo; pl.j_hi = n.M;
		if (pl.i_hi <= pl.i_lo || pl.j_hi <= pl.j_lo) { c.err = "empty pair range"; return SOL_ERR; }
		plan_pairs(pl.i_hi - pl.i_lo, pl.j_hi - pl.j_lo, pl, kMaxSplit, n.n - jlo);
	}
	auto once = [&]() {
		if (use_sym) {
			const int rounds_total = L.nb / 2 + 1;
			for (int rb = 0; rb < rounds_total; rb += kSymRounds) {
				L.round_begin = rb; L.nrounds = std::min(kSymRounds, rounds_total - rb);
				launch_sym_phase(c, L, rb == 0);
			}
		} else {
			launch_pairs(c, c.y0, pl);
		}
	};
	once();   // warm-up
	cudaEvent_t a, b;
	SOL_CUDA(cudaEventCreate(&a)); SOL_CUDA(cudaEventCreate(&b));
	SOL_CUDA(cudaEventRecord(a, c.stream));
	for (int r = 0; r < reps; r++) once();
	SOL_CUDA(cudaEventRecord(b, c.stream));
	SOL_CUDA(cudaEventSynchronize(b));
	float ms = 0;
	SOL_CUDA(cudaEventElapsedTime(&ms, a, b));
	cudaEventDestroy(a); cudaEventDestroy(b);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { c.err = std::string("kernel launch: ") + cudaGetErrorString(e); return SOL_ERR; }
	*ms_out = ms / reps;
	if (pairs_out) *pairs_out = use_sym ? (double)L.nR * (double)(L.nR - 1) : (double)(pl.i_hi - pl.i_lo) * (double)(pl.j_hi - pl.j_lo);
	return SOL_OK;
}

int sol_set_small_system_kernel(sol_ctx *h, int on)
{
	if (!h) return SOL_ERR;
	if (on < 0 || on > 3) return SOL_ERR;
	SOL_FANOUT(h, sol_set_small_system_kernel(r, on));
	h->c.cfg_epoch++;
	h->c.small_mode = on == 0 ? 0 : (on == 1 ? 1 : 2);   // 1: automatic choice (up to kSmallAuto bodies), 2: whenever it fits
	h->c.warp_mode = (on == 1 || on == 3) ? 1 : 0;
	h->c.cp_mode = on == 1 ? 1 : 0;
	return SOL_OK;
}

int sol_set_graph_mode(sol_ctx *h, int on)
{
	if (!h) return SOL_ERR;
	SOL_FANOUT(h, sol_set_graph_mode(r, on));
	if (on < 0 || on > 2) { h->c.err = "sol_set_graph_mode: mode must be 0, 1 or 2"; return SOL_ERR; }
	h->c.graph_mode = on;
	return SOL_OK;
}

int sol_set_tracer_kernel(sol_ctx *h, int on)
{
	if (!h) return SOL_ERR;
	SOL_FANOUT(h, sol_set_tracer_kernel(r, on));
	h->c.tracer_mode = on ? 1 : 0;
	h->c.cfg_epoch++;
	return SOL_OK;
}

int sol_set_pair_algorithm(sol_ctx *h, int mode)
{
	if (!h || mode < 0 || mode > 2) return SOL_ERR;
	SOL_FANOUT(h, sol_set_pair_algorithm(r, mode));
	h->c.sym_mode = mode;
	h->c.cfg_epoch++;
	return SOL_OK;
}

int sol_selftest_fast_paths(sol_ctx *h, unsigned long long seed, long long samples, unsigned long long *mismatches_out)
{
	if (!h || !mismatches_out || samples <= 0) return SOL_ERR;
	if (h->multi) return fan_out(h, [&](sol_ctx *r, int rank) { return rank == 0 ? sol_selftest_fast_paths(r, seed, samples, mismatches_out) : SOL_OK; });
	Ctx &c = h->c;
	SOL_CUDA(cudaSetDevice(c.device));
	return selftest_fast_paths(c, seed, samples, mismatches_out);
}

int sol_measure_fp64_peak(sol_ctx *h, double *tflops_out)
{
	if (!h || !tflops_out) return SOL_ERR;
	if (h->multi) return fan_out(h, [&](sol_ctx *r, int rank) { return rank == 0 ? sol_measure_fp64_peak(r, tflops_out) : SOL_OK; });
	Ctx &c = h->c;
	SOL_CUDA(cudaSetDevice(c.device));
	double *out = nullptr;
	SOL_CUDA(cudaMalloc((void **)&out, sizeof(double)));
	cudaDeviceProp prop;
	SOL_CUDA(cudaGetDeviceProperties(&prop, c.device));
	const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
	launch_fp64_peak(c, out, 1024, blocks, threads);
	cudaEvent_t a, b;
	SOL_CUDA(cudaEventCreate(&a)); SOL_CUDA(cudaEventCreate(&b));
	float best = 1e30f;
	for (int rep = 0; rep < 3; rep++) {
		SOL_CUDA(cudaEventRecord(a, c.stream));
		launch_fp64_peak(c, out, iters, blocks, threads);
		SOL_CUDA(cudaEventRecord(b, c.stream));
		SOL_CUDA(cudaEventSynchronize(b));
		float ms = 0;
		SOL_CUDA(cudaEventElapsedTime(&ms, a, b));
		best = std::min(best, ms);
	}
	cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(out);
	double flops = (double)blocks * threads * 16.0 * iters * 2.0;
	*tflops_out = flops / (best * 1e-3) / 1e12;
	return SOL_OK;
}

long long sol_launch_count(const sol_ctx *h)
{
	if (!h) return 0;
	if (!h->multi) return h->c.launches;
	long long total = 0;
	for (const sol_ctx *r : h->multi->ranks) total += r->c.launches;
	return total;
}

int sol_profile_enable(sol_ctx *h, int on)
{
	if (!h) return SOL_ERR;
	SOL_FANOUT(h, sol_profile_enable(r, on));
	h->c.prof = on != 0;
	return SOL_OK;
}

int sol_profile_read(sol_ctx *h, double ms_out[6], long long launches_out[6], int reset)
{
	if (!h) return SOL_ERR;
	if (h->multi) return fan_out(h, [&](sol_ctx *r, int rank) {      // rank 0's figures (the ranks run the same kernels on equal shares)
		return sol_profile_read(r, rank == 0 ? ms_out : nullptr, rank == 0 ? launches_out : nullptr, reset); });
	Ctx &c = h->c;
	prof_resolve(c);
	for (int q = 0; q < 6; q++) {
		if (ms_out) ms_out[q] = c.prof_ms[q];
		if (launches_out) launches_out[q] = c.prof_n[q];
		if (reset) { c.prof_ms[q] = 0; c.prof_n[q] = 0; }
	}
	return SOL_OK;
}

}  // extern "C"

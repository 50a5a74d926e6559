// Device bodies of the ordered pair kernel and of the indirect-term reduction, shared by
//
//   gravity.cu      pair_kernel / indirect_kernel: one launch each, one CTA per (sink block, source chunk)
//   elementwise.cu  fused_attempt_kernel: the launches of a whole attempt run as PHASES of one cooperative kernel; its
//                   CTAs walk the same (sink block, source chunk) list, so every partial sum is formed by the same
//                   statements in the same order as in the one-launch-per-kernel path (bit-identical; tests assert it)
//
// Nothing here has an a*b+c the compiler could contract (every fused operation is an explicit fma(), the reductions use
// __dadd_rn / __dmul_rn), so the two translation units - one compiled with FMA contraction, one with -fmad=false - emit the
// same arithmetic.
//
// The pointers are deliberately NOT __restrict__: inside the fused kernel the arrays read here were written by other
// CTAs in an earlier phase of the same launch, which rules out the non-coherent load path.
#pragma once
#include "common.cuh"
#include "ilp_asm.cuh"

namespace sol {

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
	return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "WAIT_LOOP:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	    "@p bra WAIT_DONE;\n"
	    "bra WAIT_LOOP;\n"
	    "WAIT_DONE:\n"
	    "}\n" ::"r"(smem_u32(bar)),
	    "r"(parity)
	    : "memory");
}

// 1-D bulk async copy global -> shared, completion counted in bytes on `bar` (TMA engine, UBLKCP).
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
	                 smem_u32(dst_smem)),
	             "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}

// ---------------------------------------------------------------------------------------------
// astrocentric indirect term: S_M = sum_{1<=j<M} T_j,  S_Ms = sum_{1<=j<M+s} T_j,
// T_j = m_j * (r_j * rm3_j)   (the per-pair subtrahend of Acceleration.cpp:314-316, without k^2).
// Deterministic: fixed number of blocks of 256 (virtual) threads, fixed per-thread stride order, tree reduction, the
// block that finishes last sums the block partials in block order.
//
// PACK: the body also does the source staging (planes -> packed {x,y,z,m}, prep_sources_kernel) for the same bodies
// it reduces, reading the state planes instead of src4: one launch less per evaluation on an unsharded context
// (the sharded one exchanges the staged slices between the two steps).  Same j -> thread assignment, same sums.
//
// T = physical threads of the CTA: 256, or 128 with every thread doing the work of virtual threads tid and tid + 128 -
// its two accumulators are added exactly as the first level of the 256-wide tree would add them.
// `vb` / `nvb`: index of this block / number of blocks.  sh: [6][T] doubles, last: one flag, both in shared memory.
// ---------------------------------------------------------------------------------------------
template <bool PACK, int T>
__device__ __forceinline__ void indirect_body(const double4 *src4, int M, int Ms, double *partials, double *out, unsigned *counter,
                                              const double *state, int ld, const double *mass, double4 *src4_out, const int vb,
                                              const int nvb, double (*sh)[T], bool *last)
{
	static_assert(T == 256 || T == 128, "indirect_body: 256 virtual threads");
	constexpr int V = 256 / T;
	const int tid = threadIdx.x;
	double acc[V][6];
#pragma unroll
	for (int v = 0; v < V; v++)
#pragma unroll
		for (int q = 0; q < 6; q++) acc[v][q] = 0.0;
	if (PACK && vb == 0 && tid == 0) {
		double4 s0;
		s0.x = state[0]; s0.y = state[ld]; s0.z = state[2 * ld]; s0.w = mass[0];
		src4_out[0] = s0;                                      // body 0 is a source too, but has no indirect term
	}
#pragma unroll
	for (int v = 0; v < V; v++) {
		for (int j = 1 + vb * 256 + v * T + tid; j < Ms; j += nvb * 256) {
			double4 s;
			if (PACK) {
				s.x = state[0 * ld + j]; s.y = state[1 * ld + j]; s.z = state[2 * ld + j]; s.w = mass[j];
				src4_out[j] = s;
			} else {
				s = src4[j];
			}
			double r2 = __dadd_rn(__dadd_rn(__dmul_rn(s.x, s.x), __dmul_rn(s.y, s.y)), __dmul_rn(s.z, s.z));
			double r = __dsqrt_rn(r2);
			double rm3 = __ddiv_rn(1.0, __dmul_rn(r2, r));
			double tx = __dmul_rn(s.w, __dmul_rn(s.x, rm3));
			double ty = __dmul_rn(s.w, __dmul_rn(s.y, rm3));
			double tz = __dmul_rn(s.w, __dmul_rn(s.z, rm3));
			if (j < M) { acc[v][0] = __dadd_rn(acc[v][0], tx); acc[v][1] = __dadd_rn(acc[v][1], ty); acc[v][2] = __dadd_rn(acc[v][2], tz); }
			else       { acc[v][3] = __dadd_rn(acc[v][3], tx); acc[v][4] = __dadd_rn(acc[v][4], ty); acc[v][5] = __dadd_rn(acc[v][5], tz); }
		}
	}
#pragma unroll
	for (int q = 0; q < 6; q++) sh[q][tid] = V == 1 ? acc[0][q] : __dadd_rn(acc[0][q], acc[V - 1][q]);
	__syncthreads();
	for (int st = (V == 1 ? 128 : 64); st > 0; st >>= 1) {
		if (tid < st)
			for (int q = 0; q < 6; q++) sh[q][tid] = __dadd_rn(sh[q][tid], sh[q][tid + st]);
		__syncthreads();
	}
	if (tid < 6) partials[vb * 6 + tid] = sh[tid][0];
	__threadfence();
	__syncthreads();   // all six partials are stored and fenced before thread 0 publishes the ticket
	if (tid == 0) {
		unsigned done = atomicAdd(counter, 1u);
		*last = (done == (unsigned)nvb - 1u);
	}
	__syncthreads();
	const bool is_last = *last;
	if (is_last && tid < 6) {
		__threadfence();
		double s = 0.0;
		for (int b = 0; b < nvb; b++) s = __dadd_rn(s, ((volatile double *)partials)[b * 6 + tid]);
		sh[tid][0] = s;
	}
	__syncthreads();
	if (is_last && tid < 3) {
		out[tid] = sh[tid][0];                                    // S over j < M
		out[3 + tid] = __dadd_rn(sh[tid][0], sh[3 + tid][0]);     // S over j < M+s
		if (tid == 0) *counter = 0;
	}
	__syncthreads();   // sh / last may be reused by the caller's next block
}

// ---------------------------------------------------------------------------------------------
// the ordered pair kernel's inner loop and CTA body
// ---------------------------------------------------------------------------------------------
template <int I, bool NN, bool TIE_GE, bool CHECK_SELF>
__device__ __forceinline__ void tile_loop(const double4 *__restrict__ tile, int cnt, int j0, const int (&isink)[I],
                                          const double (&xi)[I], const double (&yi)[I], const double (&zi)[I],
                                          double (&ax)[I], double (&ay)[I], double (&az)[I], double (&r2min)[I],
                                          int (&jmin)[I])
{
	// Nearest neighbour: almost every candidate loses, so the loop only filters on the high word of d^2 against
	// the largest running minimum of this lane's I sinks and takes the exact update path (same candidates in
	// the same order, hence the same result) when any lane of the warp has a hit.
	int imax = 0;
	if (NN) {
		imax = __double2hiint(r2min[0]);
#pragma unroll
		for (int k = 1; k < I; k++) imax = max(imax, __double2hiint(r2min[k]));
	}
	int jj = 0;
	if constexpr (I == 1) {
		// One sink per thread is the mid-size regime: few warps per scheduler, so a warp has to bring its own independent
		// work.  The compiler emits unrolled iterations one after the other (a ~190-cycle dependent chain per source:
		// LDS, 3 DADD, d^2, MUFU, 7 refinement steps, 3 DFMA), so four sources are advanced in lock step here - the
		// same operations per pair, accumulated in the same order, hence the same bits.
		// Each stage of the four chains is ONE volatile asm block (ilp_asm.cuh): the front end otherwise re-serialises
		// the chains (depth first, to save registers) and the assembler keeps that order inside a large function.
		constexpr int U = 4;
		for (; jj + U <= cnt; jj += U) {
			using A = ilp::V<U>;
			double sx[U], sy[U], sz[U], sm[U], dx[U], dy[U], dz[U], r2[U], nr2[U], y0[U], c2[U], e[U], my[U], c3m[U], p[U], pe[U], w[U];
#pragma unroll
			for (int u = 0; u < U; u++) {
				const double4 s = tile[jj + u];
				sx[u] = s.x; sy[u] = s.y; sz[u] = s.z; sm[u] = s.w;
			}
			A::sub_vs(dx, sx, xi[0]); A::sub_vs(dy, sy, yi[0]); A::sub_vs(dz, sz, zi[0]);     // d = source - sink
			A::mul_vv(r2, dx, dx); A::fma_sq_acc(r2, dy); A::fma_sq_acc(r2, dz);
			A::rsqrt(y0, r2);
			// mass_over_r3, stage by stage:  c2 = y0^2, my = m y0;  e = 1 - r2 c2, c3m = c2 my;  p = 1.5 + 1.875 e;  w = c3m + c3m (p e)
			A::mul_vv(c2, y0, y0); A::mul_vv(my, sm, y0);
#pragma unroll
			for (int u = 0; u < U; u++) nr2[u] = -r2[u];
			A::fma_vvs(e, nr2, c2, 1.0); A::mul_vv(c3m, c2, my);
			A::fma_svs(p, 1.875, e, 1.5);
			A::mul_vv(pe, p, e);
			A::fma_vvv(w, c3m, pe, c3m);
#pragma unroll
			for (int u = 0; u < U; u++) {
				if (CHECK_SELF) w[u] = ((j0 + jj + u) == isink[0]) ? 0.0 : w[u];
				ax[0] = fma(w[u], dx[u], ax[0]);
				ay[0] = fma(w[u], dy[u], ay[0]);
				az[0] = fma(w[u], dz[u], az[0]);
			}
			if (NN) {
				bool hit = false;
#pragma unroll
				for (int u = 0; u < U; u++) hit |= __double2hiint(r2[u]) <= imax;
				if (__any_sync(0xffffffffu, hit)) {
#pragma unroll
					for (int u = 0; u < U; u++) {
						const bool closer = closer_than<TIE_GE>(r2[u], r2min[0]) && !(CHECK_SELF && (j0 + jj + u) == isink[0]);
						r2min[0] = closer ? r2[u] : r2min[0];
						jmin[0] = closer ? (j0 + jj + u) : jmin[0];
					}
					imax = __double2hiint(r2min[0]);
				}
			}
		}
	}
#pragma unroll 4
	for (; jj < cnt; jj++) {
		const double4 s = tile[jj];
		double r2[I];
#pragma unroll
		for (int k = 0; k < I; k++) {
			const double dx = s.x - xi[k];
			const double dy = s.y - yi[k];
			const double dz = s.z - zi[k];
			r2[k] = fma(dz, dz, fma(dy, dy, dx * dx));
			double w = mass_over_r3(r2[k], s.w);   // e uses y0^2 rounded: costs <= 1.5 ulp in w, saves one DMUL
			if (CHECK_SELF) w = ((j0 + jj) == isink[k]) ? 0.0 : w;
			ax[k] = fma(w, dx, ax[k]);
			ay[k] = fma(w, dy, ay[k]);
			az[k] = fma(w, dz, az[k]);
		}
		if (NN) {
			bool hit = false;
#pragma unroll
			for (int k = 0; k < I; k++) hit |= __double2hiint(r2[k]) <= imax;
			if (__any_sync(0xffffffffu, hit)) {
#pragma unroll
				for (int k = 0; k < I; k++) {
					const bool closer = closer_than<TIE_GE>(r2[k], r2min[k]) && !(CHECK_SELF && (j0 + jj) == isink[k]);
					r2min[k] = closer ? r2[k] : r2min[k];
					jmin[k] = closer ? (j0 + jj) : jmin[k];
				}
				imax = __double2hiint(r2min[0]);
#pragma unroll
				for (int k = 1; k < I; k++) imax = max(imax, __double2hiint(r2min[k]));
			}
		}
	}
}

// The mbarriers of the two tile buffers are initialised ONCE per CTA by the caller; `use0` / `use1` count how often each
// buffer has been filled so far (a kernel that runs a single block passes zeros), which gives the parity to wait for.
struct PairSmem {
	double4 (*tile)[kTileJ];       // [2][kTileJ], 128-byte aligned
	uint64_t *bar;                 // [2]
	double (*run)[kPairThreads];   // [3 * I][kPairThreads]
};

// One CTA of kPairThreads threads: sinks block `bx` (kPairThreads * I sinks) against source chunk `by`.
template <int I, bool NN, bool TIE_GE>
__device__ __forceinline__ void pair_body(const double *state, const int ld, const double4 *src4, const PairLaunch &pl, double *part,
                                          double *partR2, int *partIdx, const int bx, const int by, const PairSmem &sm,
                                          unsigned &use0, unsigned &use1, unsigned long long *tr = nullptr)
{
#define SOL_TR(k) do { if (tr != nullptr && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); tr[k] = t_; } } while (0)
	SOL_TR(0);
	// Two-level summation: the registers hold the sum over ONE tile (256 sources); the running sum over the tiles of
	// this CTA's chunk lives in sm.run.  A close neighbour's large term then perturbs the ~10^2 tile additions after it
	// instead of the ~3*10^4 pair additions a single running accumulator would make at N = 10^6 (measured against the
	// extended-precision oracle: 5e-13 -> 3e-14 of |a_i| on the worst-conditioned bodies).  One LDS + DADD + STS per
	// sink and tile; a chunk of a single tile gives 0.0 + tile sum, i.e. the same bits as before.
	const int tid = threadIdx.x;
	const int ibase = pl.i_lo + bx * (kPairThreads * I);
	const int split = by + pl.split_offset;
	const int jb = pl.j_lo + by * pl.chunk;
	const int je = min(jb + pl.chunk, pl.j_hi);
	const int ntiles = (je - jb + kTileJ - 1) / kTileJ;

	// first tile on its way before anything else (a mid-size launch is one short dependent chain per CTA: the copy's
	// latency then overlaps the loads of the sinks)
	if (tid == 0 && ntiles > 0) {
		unsigned cnt0 = (unsigned)min(kTileJ, je - jb);
		mbar_expect_tx(&sm.bar[0], cnt0 * 32u);
		bulk_g2s(&sm.tile[0][0], src4 + jb, cnt0 * 32u, &sm.bar[0]);
	}
	int isink[I];
	double xi[I], yi[I], zi[I], ax[I], ay[I], az[I], r2min[I];
	int jmin[I];
#pragma unroll
	for (int k = 0; k < I; k++) {
		int i = ibase + k * kPairThreads + tid;
		isink[k] = i;
		int ic = i < pl.i_hi ? i : pl.i_hi - 1;   // clamp: out-of-range lanes compute a duplicate, never store
		xi[k] = state[0 * ld + ic];
		yi[k] = state[1 * ld + ic];
		zi[k] = state[2 * ld + ic];
		ax[k] = ay[k] = az[k] = 0.0;
		sm.run[3 * k + 0][tid] = sm.run[3 * k + 1][tid] = sm.run[3 * k + 2][tid] = 0.0;
		r2min[k] = 1.0e20;   // (rMin = 1e10)^2, Acceleration.cpp:269 / :546
		jmin[k] = -1;
	}

	const int blk_lo = ibase, blk_hi = ibase + kPairThreads * I;
	for (int t = 0; t < ntiles; t++) {
		const int buf = t & 1;
		if (tid == 0 && t + 1 < ntiles) {
			const int jn = jb + (t + 1) * kTileJ;
			unsigned cntn = (unsigned)min(kTileJ, je - jn);
			mbar_expect_tx(&sm.bar[buf ^ 1], cntn * 32u);
			bulk_g2s(&sm.tile[buf ^ 1][0], src4 + jn, cntn * 32u, &sm.bar[buf ^ 1]);
		}
		mbar_wait(&sm.bar[buf], ((unsigned)(t >> 1) + (buf ? use1 : use0)) & 1u);
		if (t == 0) SOL_TR(1);
		const int j0 = jb + t * kTileJ;
		const int cnt = min(kTileJ, je - j0);
		const bool diag = (j0 < blk_hi) && (j0 + cnt > blk_lo);   // tile may contain one of this CTA's sinks
		if (diag)
			tile_loop<I, NN, TIE_GE, true>(sm.tile[buf], cnt, j0, isink, xi, yi, zi, ax, ay, az, r2min, jmin);
		else
			tile_loop<I, NN, TIE_GE, false>(sm.tile[buf], cnt, j0, isink, xi, yi, zi, ax, ay, az, r2min, jmin);
#pragma unroll
		for (int k = 0; k < I; k++) {
			sm.run[3 * k + 0][tid] += ax[k]; sm.run[3 * k + 1][tid] += ay[k]; sm.run[3 * k + 2][tid] += az[k];
			ax[k] = ay[k] = az[k] = 0.0;
		}
		__syncthreads();   // everyone is done with tile[buf] before it is refilled two iterations later
	}
	SOL_TR(2);
	use0 += (unsigned)((ntiles + 1) >> 1);
	use1 += (unsigned)(ntiles >> 1);

#pragma unroll
	for (int k = 0; k < I; k++) {
		int i = isink[k];
		if (i < pl.i_hi) {
			part[(size_t)(split * 3 + 0) * ld + i] = sm.run[3 * k + 0][tid];
			part[(size_t)(split * 3 + 1) * ld + i] = sm.run[3 * k + 1][tid];
			part[(size_t)(split * 3 + 2) * ld + i] = sm.run[3 * k + 2][tid];
			if (NN) {
				partR2[(size_t)split * ld + i] = r2min[k];
				partIdx[(size_t)split * ld + i] = jmin[k];
			}
		}
	}
	SOL_TR(3);
#undef SOL_TR
}

}  // namespace sol

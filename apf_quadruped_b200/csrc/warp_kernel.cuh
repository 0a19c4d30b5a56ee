// Warp-per-QP kernel of the batched interior-point solver (sm_100a, FP64): ONE WARP PER QP, the QP's whole
// working set -- value row, factor L, pivots, iterate, residuals (SliceLayout, program.h) -- resident in the
// warp's slice of SHARED memory from the coalesced load of its value row to the store of its solution (generated
// programs keep the value row itself in a per-warp scratch row in global memory / L2, in the order the residual
// products read it: see value_row_global, value_row_resid_order below).
//
// Launch shape: one persistent CTA per SM, as many warps as slices fit.  The warps of a CTA walk the
// interior-point iteration in lock step -- [residuals + stop test | factor | solve | step length, sigma |
// solve | update] -- so that they fetch the same code / instruction words at the same time.  A warp whose QP
// has left the loop stores it and pulls the next QP index from an atomic counter at the top of the cycle; the
// new QP's initial point (kkt_initialize) is just another factor + solve of the same cycle, so iteration
// counts that differ from QP to QP cost no idle cycles.
//
// The kernel is a template over where LDL_numeric and the triangular solves come from:
//   InterpretedWaves   the list-scheduled wave schedules of program.cpp, interpreted (any pattern);
//   a GENERATED program (apf_quadruped_b200/codegen_warp.py): straight-line warp code for one pattern,
//                      lane = row of L, L(j,i) read once per column step as a shared-memory broadcast.
// Reference: QP_SOLVE qpSWIFT.c:473-644; the elementwise steps are kernels.cuh's.
#pragma once

#include "kernels.cuh"
#include "warp_ipm.cuh"

namespace qpb {

struct ProgramView
{
	const uint2 *words;
	int f_first, f_waves, f_width, s_first, s_waves, s_width;
};

enum { MODE_IDLE = 0, MODE_INIT = 1, MODE_PRED = 2, MODE_CENT = 3 };
static const int MAX_WARPS_PER_CTA = 16; // interpreted schedules: 128 registers per thread
static const int GEN_WARPS_MAX = 16;     // generated programs: as many warps per CTA as their slices fit (trot 12, crawl 11, stance 10)

struct GeneratedWarpKernel
{
	unsigned long long key; // stream_program_key of the pattern the code was generated for
	const void *kernel;     // qp_solve_warp_kernel<PROG, max_warps, true>
	const void *kernel_free; // ... <PROG, max_warps, false>: the warps of a CTA never wait for one another
	const char *name;
	SliceLayout lay;        // the layout the addresses in the code assume (scratch cut down to what the code uses)
	int max_warps;          // warps per CTA the kernel was compiled for (its register budget follows: 65,536 / (32 max_warps))
	int value_row_doubles;  // > 0: the value row P | A | G lives in a scratch row of that many doubles per warp in GLOBAL memory
};

void register_generated_warp_kernel(unsigned long long key, const void *kernel, const void *kernel_free, const char *name, const SliceLayout &lay, int max_warps, int value_row_doubles);
const GeneratedWarpKernel *find_generated_warp_kernel(unsigned long long key, bool value_row_global);

#ifdef __CUDACC__

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

// One table-driven chain per lane (generated programs): acc -= S[ta[64 t]] * S[tb[64 t]], t ascending; ta / tb
// point at the lane's column of 16-bit byte offsets, a step's rows of 32 lie 64 entries apart.
__device__ __forceinline__ double lane_chain(double acc, const unsigned sb, const unsigned short *__restrict__ ta, const unsigned short *__restrict__ tb, const int steps)
{
	int t = 0;
	for (; t + 4 <= steps; t += 4) {
		unsigned oa[4], ob[4];
		double va[4], vb[4];
#pragma unroll
		for (int u = 0; u < 4; ++u) { oa[u] = __ldg(ta + 64 * (t + u)); ob[u] = __ldg(tb + 64 * (t + u)); }
#pragma unroll
		for (int u = 0; u < 4; ++u) { va[u] = lds64(sb + oa[u]); vb[u] = lds64(sb + ob[u]); }
#pragma unroll
		for (int u = 0; u < 4; ++u) acc = acc - va[u] * vb[u];
	}
	for (; t < steps; ++t) acc = acc - lds64(sb + __ldg(ta + 64 * t)) * lds64(sb + __ldg(tb + 64 * t));
	return acc;
}

// The warps of a CTA that walk the iteration together: the whole CTA (barrier 0), or -- QPB_WARP_GROUPS=G, an experiment --
// one of G groups (named barriers 1..G) that run 1/G of a cycle apart, so that the LSU-heavy and the FP64-heavy
// phases of the groups overlap instead of colliding (measured: erratic, -40 % to +2 % on trot; the default stays one group).
struct SyncGroup
{
	unsigned id, threads;
};
__device__ __forceinline__ unsigned group_count(const SyncGroup g, const bool pred)
{
	if (g.id == 0u) return (unsigned)__syncthreads_count(pred);
	unsigned r;
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\tbar.red.popc.u32 %0, %1, %2, p;\n\t}" : "=r"(r) : "r"(g.id), "r"(g.threads), "r"((unsigned)pred) : "memory");
	return r;
}
__device__ __forceinline__ bool group_or(const SyncGroup g, const bool pred) { return group_count(g, pred) != 0; }

// One QP's rows P | A | G | c | h | b for a program that knows its sizes at compile time: all loads are issued before the
// first store (a loop of load -> store pairs would pay one trip to HBM per pass, because the compiler must keep a store
// to the scratch row in global memory ahead of the next load), coalesced 8-byte requests, 256 bytes per warp each.
template <int COUNT, bool L2_ONLY>
struct RowInFlight
{
	static constexpr int K = COUNT > 0 ? (COUNT + 31) / 32 : 1;
	double v[K];
	__device__ __forceinline__ void load(const double *__restrict__ src, const int lane)
	{
#pragma unroll
		for (int k = 0; k < K; ++k) {
			const int i = lane + 32 * k;
			v[k] = 0.0;
			if (i < COUNT) v[k] = L2_ONLY ? __ldcg(src + i) : __ldg(src + i);
		}
	}
	template <bool PLUS_ZERO>
	__device__ __forceinline__ void store(double *dst, const int lane) const
	{
#pragma unroll
		for (int k = 0; k < K; ++k) {
			const int i = lane + 32 * k;
			if (i < COUNT) dst[i] = PLUS_ZERO ? v[k] + 0.0 : v[k];
		}
	}
	// entries [first, first + COUNT) of P | A | G to their slots of a value row in residual order (PROG::kVDst at most each), as v + 0.0
	template <class PROG>
	__device__ __forceinline__ void place(double *row, const int first, const int lane) const
	{
#pragma unroll
		for (int k = 0; k < K; ++k) {
			const int i = lane + 32 * k;
			unsigned d[PROG::kVDst > 0 ? PROG::kVDst : 1];
#pragma unroll
			for (int c = 0; c < PROG::kVDst; ++c) d[c] = i < COUNT ? PROG::vdst(c * PROG::kNV + first + i) : 0xffffu;
			const double t = v[k] + 0.0;
#pragma unroll
			for (int c = 0; c < PROG::kVDst; ++c)
				if (d[c] != 0xffffu) row[d[c]] = t;
		}
	}
};

template <class PROG, bool L2_ONLY>
__device__ __forceinline__ void load_qp_static(const WarpState &w, const BatchIO &io, const size_t qi)
{
	constexpr int nP = PROG::kOffA, nA = PROG::kOffG - PROG::kOffA, nG = PROG::kNV - PROG::kOffG;
	const int lane = lane_id();
	RowInFlight<nP, L2_ONLY> rP;
	RowInFlight<nA, L2_ONLY> rA;
	RowInFlight<nG, L2_ONLY> rG;
	RowInFlight<PROG::kN, L2_ONLY> rc;
	RowInFlight<PROG::kM, L2_ONLY> rh;
	RowInFlight<PROG::kP, L2_ONLY> rb;
	rP.load(io.Ppr + qi * nP, lane);
	if (nA) rA.load(io.Apr + qi * nA, lane);
	rG.load(io.Gpr + qi * nG, lane);
	rc.load(io.c + qi * PROG::kN, lane);
	rh.load(io.h + qi * PROG::kM, lane);
	if (PROG::kP) rb.load(io.b + qi * PROG::kP, lane);
	if (PROG::value_row_resid_order) {
		rP.template place<PROG>(w.V, 0, lane);
		if (nA) rA.template place<PROG>(w.V, nP, lane);
		rG.template place<PROG>(w.V, PROG::kOffG, lane);
	} else {
		rP.template store<true>(w.V, lane);
		if (nA) rA.template store<true>(w.V + nP, lane);
		rG.template store<true>(w.V + PROG::kOffG, lane);
	}
	rc.template store<false>(w.c, lane);
	rh.template store<false>(w.h, lane);
	if (PROG::kP) rb.template store<false>(w.b, lane);
}

struct InterpretedWaves
{
	static constexpr bool own_residuals = false;
	static constexpr bool realign = true;
	static constexpr bool value_row_global = false;
	static constexpr int value_row_doubles = 0;
	static constexpr bool value_row_resid_order = false;
	static constexpr int kVDst = 0;
	static __device__ __forceinline__ unsigned vdst(const int) { return 0u; }
	static __device__ __forceinline__ unsigned gslot(const int) { return 0u; }
	static constexpr bool static_shape = false; // sizes and layout come with the launch
	static constexpr int kN = 0, kM = 0, kP = 0, kOffA = 0, kOffG = 0, kNV = 0;
	static __device__ __forceinline__ SliceLayout layout() { return SliceLayout(); }
	static __device__ __forceinline__ void resid(const unsigned, const char *) {}
	static __device__ __forceinline__ void factor(const ProgramView &pg, const unsigned sb, unsigned, const char *) { run_segment(pg.words, pg.f_first, pg.f_waves, pg.f_width, sb); }
	static __device__ __forceinline__ void solve(const ProgramView &pg, const unsigned sb, unsigned, const char *) { run_segment(pg.words, pg.s_first, pg.s_waves, pg.s_width, sb); }
};

template <class PROG, int MAXW, bool CTA_SYNC>
__global__ void __launch_bounds__(MAXW * 32, 1) qp_solve_warp_kernel(const int *__restrict__ plan, ProgramView pg, SliceLayout lay_arg, BatchIO io, SolveParams prm, int *__restrict__ counter)
{
	extern __shared__ __align__(16) double smem[];
	// A generated program knows its pattern's sizes and slice layout at compile time: the elementwise steps below then
	// unroll, and every address in them is the slice base plus a constant.
	const SliceLayout lay = PROG::static_shape ? PROG::layout() : lay_arg;
	PlanView pv = view_of(plan);
	if (PROG::static_shape) { pv.n = PROG::kN; pv.m = PROG::kM; pv.p = PROG::kP; pv.N = PROG::kN + PROG::kM + PROG::kP; pv.offA = PROG::kOffA; pv.offG = PROG::kOffG; pv.nV = PROG::kNV; }
	const int lane = lane_id(), warp = threadIdx.x >> 5;
	double *S = smem + (size_t)warp * lay.total;
	const unsigned sb = (unsigned)__cvta_generic_to_shared(S);
	WarpState w = carve(S, lay);
	if (lane == 0) { S[lay.zero] = 0.0; S[lay.zero + 1] = -0.0; }
	// Generated programs for the DogBot patterns keep the value row V = P | A | G -- read-only during a solve, a third
	// of the working set -- in the warp's own scratch row in GLOBAL memory (L2-resident: a few MB for the whole grid)
	// instead of in the slice: 18.7 KB of shared memory per QP instead of 26.2 KB (trot), i.e. 12 warps per SM instead
	// of 8.  The programs address slice and scratch row as ONE space: a byte offset >= lay.V * 8 is in the scratch row.
	const char *vb = nullptr;
	if (PROG::value_row_global) {
		double *vrow = io.vscratch + ((size_t)blockIdx.x * MAXW + warp) * PROG::value_row_doubles;
		w.V = vrow;
		vb = reinterpret_cast<const char *>(vrow) - (size_t)lay.V * sizeof(double);
		if (PROG::value_row_resid_order) { // the slots without a term hold +0.0 for good: a QP's values only ever go to their own slots
			for (int i = lane; i < PROG::value_row_doubles; i += 32) vrow[i] = 0.0;
		} else if (lane == 0) { vrow[PROG::value_row_doubles - 2] = 0.0; vrow[PROG::value_row_doubles - 1] = 0.0; } // the zero behind the row
	}
	__syncwarp();
	const int n = pv.n, m = pv.m, p = pv.p;
	const double thr = prm.reltol / sqrt(3.0);

	SyncGroup grp = {0u, blockDim.x};
	if (CTA_SYNC && prm.groups >= 2 && prm.groups <= 8 && blockDim.x >= 32u * (unsigned)prm.groups) {
		// group g = warps [g nw / G, (g + 1) nw / G): named barrier g + 1, started g * delay behind the first group
		const unsigned nw = blockDim.x >> 5, G = (unsigned)prm.groups;
		unsigned g = 0;
		while ((g + 1) * nw / G <= (unsigned)warp) ++g;
		grp.id = g + 1u;
		grp.threads = 32u * ((g + 1) * nw / G - g * nw / G);
		for (unsigned t = 0; t < g; ++t) __nanosleep(prm.group_delay_ns);
	}
	int q = 0, iters = 0;
	bool active = false; // a QP is between its initial point and its exit
	double sigma = 0.0, mu = 0.0, ap = 0.0, ad = 0.0;
	Scalars sc;
	sc.n_rx = sc.n_ry = sc.n_rz = sc.n_mu = sc.fval = 0.0;
	const size_t trace_rows = (size_t)prm.maxit + 2;
	long long ldl_ns = 0;

	for (;;) {
		int mode = MODE_IDLE;
		if (active) { // top of QP_SOLVE's loop (qpSWIFT.c:502-534)
			int flag = -1;
			double mu_next = mu; // (mu is updated after the stop test, qpSWIFT.c:538: a QP that leaves keeps the last one)
			if (iters >= prm.maxit) flag = 2; // QP_MAXIT (qpSWIFT.c:604-607)
			else {
				if (PROG::own_residuals) PROG::resid(sb, vb);
				else ipm_residual_products(pv, w.V, w.c, w.h, w.b, w.x, w.y, w.z, w.s, w.r, S + lay.T /* scratch: -(P x) */);
				const IterSums sums = ipm_sums(n, m, p, sb, lay, lay.T);
				sc = sums.sc;
				mu_next = sums.mu;
				if (io.trace && lane == 0) {
					double *row = io.trace + ((size_t)q * trace_rows + 1 + iters) * 8;
					row[0] = sc.fval; row[1] = sc.n_rx; row[2] = sc.n_ry; row[3] = sc.n_rz; row[4] = sc.n_mu;
				}
				if (sc.n_rx < thr && sc.n_rz < thr && (p == 0 || sc.n_ry < thr) && sc.n_mu < prm.abstol) flag = 0;
			}
			if (flag >= 0) {
				if (io.trace && lane == 0) io.trace[(size_t)q * trace_rows * 8] = (double)ldl_ns * 1e-9;
				const size_t qi = (size_t)q;
				warp_copy_out(io.x ? io.x + qi * n : nullptr, w.x, n);
				if (p) warp_copy_out(io.y ? io.y + qi * p : nullptr, w.y, p);
				warp_copy_out(io.z ? io.z + qi * m : nullptr, w.z, m);
				warp_copy_out(io.s ? io.s + qi * m : nullptr, w.s, m);
				if (lane == 0) {
					if (io.flag) io.flag[q] = flag;
					if (io.iters) io.iters[q] = iters;
					if (io.fval) io.fval[q] = sc.fval;
					if (io.stat) {
						double *st = io.stat + qi * 8;
						st[0] = sc.n_rx; st[1] = sc.n_ry; st[2] = sc.n_rz; st[3] = sc.n_mu;
						st[4] = ap; st[5] = ad; st[6] = sigma; st[7] = mu;
					}
				}
				__syncwarp();
				active = false;
			} else {
				__syncwarp();
				mu = mu_next;
				const bool centering = !(sigma > prm.sigma_d); // predictor: ds = -lambda^2 (qpSWIFT.c:540-552); else the centering step only (572-579)
				if (centering) sigma = prm.sigma_d;
				mode = centering ? MODE_CENT : MODE_PRED;
				ipm_begin_step(n, m, p, w, centering, sigma * mu);
			}
		}
		bool may_leave = true;
		if (!active) { // next QP: load its value row, set up kkt_initialize's system
			if (lane == 0) {
				if (!io.server_done) {
					q = atomicAdd(counter, 1);
					const int limit = io.count_ptr ? *io.count_ptr : io.count;
					q = q < limit ? (io.index ? io.index[q] : q) : -1; // the QP behind the ticket (| iterations << 24)
				} else { // server: the list grows while the throughput kernel runs; one attempt per turn, never a wait
					const int done = *reinterpret_cast<const volatile int *>(io.server_done); // (read before the count)
					__threadfence();
					const int limit = *reinterpret_cast<const volatile int *>(io.count_ptr);
					const int cur = *reinterpret_cast<volatile int *>(counter);
					q = -1;
					if (cur < limit) {
						may_leave = false;
						const int listed = reinterpret_cast<const volatile int *>(io.index)[cur];
						if (listed >= 0 && atomicCAS(counter, cur, cur + 1) == cur) q = listed;
						__threadfence(); // the entry before what it points at
					} else if (!done) may_leave = false;
				}
			}
			q = __shfl_sync(0xffffffffu, q, 0);
			if (q >= 0) {
				// list entries carry the iteration count in the high byte only while QP indices fit 24 bits (stream.cuh)
				const bool packed = io.index && io.count < (1 << 24);
				const int resume_at = packed ? (q >> 24) : 0; // > 0: a straggler's iterate is waiting in the output arrays
				if (packed) q &= (1 << 24) - 1;
				const size_t qi = (size_t)q;
				// host-pointer entry: the value rows come up in chunks while the kernel runs (batch.cu); QPs [0, *ready) are there
				if (io.ready && !io.index) {
					if (lane == 0)
						while (*reinterpret_cast<const volatile int *>(io.ready) <= q) __nanosleep(500);
					__syncwarp();
				}
				// matrix values as v + 0.0 (-0.0 becomes +0.0): what "Y[i] += value" from Y[i] == 0 makes of them (ldl.c:284)
				if (PROG::static_shape) { // every load of the QP's six rows in flight before the first store: ONE round trip to HBM
					if (io.ready) load_qp_static<PROG, true>(w, io, qi);
					else load_qp_static<PROG, false>(w, io, qi);
				} else if (io.ready) { // (rows that arrive while the kernel runs: L2 only)
					warp_copy_in<true, true>(w.V, io.Ppr + qi * pv.offA, pv.offA);
					if (p) warp_copy_in<true, true>(w.V + pv.offA, io.Apr + qi * (pv.offG - pv.offA), pv.offG - pv.offA);
					warp_copy_in<true, true>(w.V + pv.offG, io.Gpr + qi * (pv.nV - pv.offG), pv.nV - pv.offG);
					warp_copy_in<false, true>(w.c, io.c + qi * n, n);
					warp_copy_in<false, true>(w.h, io.h + qi * m, m);
					if (p) warp_copy_in<false, true>(w.b, io.b + qi * p, p);
				} else {
					warp_copy_in<true>(w.V, io.Ppr + qi * pv.offA, pv.offA);
					if (p) warp_copy_in<true>(w.V + pv.offA, io.Apr + qi * (pv.offG - pv.offA), pv.offG - pv.offA);
					warp_copy_in<true>(w.V + pv.offG, io.Gpr + qi * (pv.nV - pv.offG), pv.nV - pv.offG);
					warp_copy_in(w.c, io.c + qi * n, n);
					warp_copy_in(w.h, io.h + qi * m, m);
					if (p) warp_copy_in(w.b, io.b + qi * p, p);
				}
				if (PROG::value_row_global) __threadfence_block(); // (the row was written through w.V by other lanes: visible before it is read)
				__syncwarp();
				if (resume_at > 0 && io.spill_sigma) {
					// the state of QP_SOLVE's loop between two iterations is (x, y, z, s, sigma, IterationCount);
					// everything else is recomputed at the top of the next one (qpSWIFT.c:502-552)
					warp_copy_in(w.x, io.x + qi * n, n);
					if (p) warp_copy_in(w.y, io.y + qi * p, p);
					warp_copy_in(w.z, io.z + qi * m, m);
					warp_copy_in(w.s, io.s + qi * m, m);
					__syncwarp();
					iters = resume_at;
					sigma = io.spill_sigma[q];
					mu = ap = ad = 0.0;
					active = true;
					may_leave = false;
				} else {
					ipm_init_rhs(pv, w);
					ldl_ns = 0;
					iters = 0;
					sigma = prm.sigma0;
					mu = ap = ad = 0.0;
					mode = MODE_INIT;
				}
			}
		}
		// how many threads of the CTA run factor + solve now: they meet at barriers inside the generated code
		const unsigned busy = CTA_SYNC ? group_count(grp, mode != MODE_IDLE) : (mode != MODE_IDLE ? 32u : 0u);
		if (!busy) {
			if (!(CTA_SYNC ? group_or(grp, !may_leave) : !may_leave)) break;
			__nanosleep(2000); // a server with nothing listed yet
			continue;
		}

		if (mode != MODE_IDLE) {
			long long t0 = 0;
			if (io.trace) t0 = (long long)globaltimer_ns();
			PROG::factor(pg, sb, busy, vb);
			if (io.trace) ldl_ns += (long long)globaltimer_ns() - t0;
			PROG::solve(pg, sb, busy, vb);
		}
		if (mode == MODE_INIT) {
			ipm_init_point(pv, w, [](const int t, const PlanView &v) { return PROG::value_row_resid_order ? (int)PROG::gslot(t) : __ldg(v.grow_val + t); });
			active = true;
		} else if (mode == MODE_PRED) { // affine step length, sigma, corrector right-hand side (qpSWIFT.c:564-570)
			ipm_recover_and_steplength(n, m, p, w, ap, ad);
			sigma = ipm_sigma_and_corrector(n, m, p, w, sb, lay.T, ap, ad, prm.sigma_d, mu);
		}
		// (re-aligns the warps where the program wants that: interpreted schedules stream their words together)
		const unsigned again = CTA_SYNC && PROG::realign ? group_count(grp, mode == MODE_PRED) : 32u;
		if (mode == MODE_PRED) PROG::solve(pg, sb, again, vb);
		if (mode == MODE_PRED || mode == MODE_CENT) { // step, damping, update (qpSWIFT.c:584-596)
			ipm_recover_and_steplength(n, m, p, w, ap, ad);
			ipm_update(n, m, p, w, ap, ad);
			if (io.trace && lane == 0) {
				double *row = io.trace + ((size_t)q * trace_rows + 1 + iters) * 8;
				row[5] = ap; row[6] = ad;
			}
			++iters;
		}
	}
}

#define QPB_REGISTER_GENERATED_WARP(tag, key, layout, PROG, MAXW)                                                                    \
	template __global__ void qp_solve_warp_kernel<PROG, MAXW, true>(const int *__restrict__, ProgramView, SliceLayout, BatchIO, SolveParams, int *__restrict__);  \
	template __global__ void qp_solve_warp_kernel<PROG, MAXW, false>(const int *__restrict__, ProgramView, SliceLayout, BatchIO, SolveParams, int *__restrict__); \
	namespace {                                                                                                                           \
	struct RegisterWarp_##tag                                                                                                             \
	{                                                                                                                                     \
		RegisterWarp_##tag()                                                                                                              \
		{                                                                                                                                 \
			register_generated_warp_kernel(key, reinterpret_cast<const void *>(&qp_solve_warp_kernel<PROG, MAXW, true>),                  \
			                               reinterpret_cast<const void *>(&qp_solve_warp_kernel<PROG, MAXW, false>), #tag, layout, MAXW,  \
			                               PROG::value_row_global ? PROG::value_row_doubles : 0);                                         \
		}                                                                                                                                 \
	} register_warp_##tag;                                                                                                                \
	}

#endif // __CUDACC__

} // namespace qpb

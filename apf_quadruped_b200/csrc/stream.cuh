// Throughput kernel of the batched interior-point solver (sm_100a, FP64): ONE THREAD PER QP.
//
// The 32 lanes of a warp run 32 QPs of the same sparsity pattern in lock step.  A QP's whole
// working set (value row, factor, iterate, residuals: StreamLayout) is a slice of HBM interleaved
// with the slices of the other 31 lanes -- element e of lane l lives at base + e*32 + l -- so every
// load and store of the warp is one fully used 256-byte line, whatever the sparsity pattern asks for.
// The factorisation, the triangular solves and the sparse residual products are micro-programs
// compiled per pattern on the host (stream_program.h) and interpreted here: the warp executes a
// GROUP of micro-ops at a time, first issuing every HBM load of the group (they are independent of
// the group's arithmetic by construction), then running the arithmetic in the reference's order with
// the two accumulators in registers and the row-local work vector in shared memory.  HBM latency is
// covered by the loads in flight per warp (up to 2 per micro-op) times the warps per SM, not by
// occupancy alone.
//
// Arithmetic contract: each thread performs exactly the reference's IEEE-754 binary64 operations in
// the reference's order (file compiled with -fmad=false), so x, y, z, s, the iteration count and the
// exit flag are bit-identical with reference qpSWIFT, and independent of the batch composition.
//
// Lanes are scheduled individually: a lane whose QP has left the interior-point loop writes it
// back and pulls the next QP of the batch from an atomic counter; the new QP's initial point
// (kkt_initialize) is one more factor + solve of the same programs, so lanes at different iteration
// counts never wait for one another.
//
// Reference routines restated (dogbot_controller/src/qpSWIFT/): see kernels.cuh for the list; the
// elementwise steps below follow Auxilary.c:205-393, 638-646, 879-892, 992-1089 and qpSWIFT.c:473-644.
#pragma once

#include "kernels.cuh"
#include "stream_program.h"

namespace qpb {

struct StreamView
{
	const uint2 *words;
	int first[SSEG_COUNT], groups[SSEG_COUNT];
};

struct StreamDims
{
	int n, m, p, N, nnzP, nnzA, nnzG;
};

#ifdef __CUDACC__

// Interpreter.  g: the lane's slice (element e at g[e << 5]); sb: the lane's window of shared memory
// (element e at sb + (e << 8)).
//
// Instruction supply: the words are the same for every lane, so the warp reads them a CHUNK (DEPTH
// groups of G micro-ops, at most 32) at a time with one coalesced load -- lane l takes word l, decodes it
// once into byte offsets and flags -- and every lane then gets the fields of a micro-op with a shuffle
// from its owner lane.  Software pipeline, per group i:
//            request the HBM operands of group i+DEPTH-1   (DEPTH operand buffers, rotated by unrolling)
//            execute group i                                (operands requested DEPTH-1 groups ago)
// so DEPTH-1 groups of operands (24 to 48 lines of 256 bytes per warp) are in flight while a group
// runs, and chunk c+2 is on its way from L2 while chunk c executes.
// A FAST group is executed without a single branch (flags become predicates and selects), a SLOW
// group micro-op by micro-op (the words are the same in all lanes, so no branch ever diverges).
// GS = false: factor / solve programs (results to the shared window); GS = true: residual programs
// (the window holds a copy of x | y | z, results to HBM).
// __noinline__: one copy serves all call sites (instruction cache).

// A load the compiler may neither delay to its first use nor move into a branch: the requests of a
// group have to be in flight together.
__device__ __forceinline__ double ldg_now(const char *p)
{
	double v;
	asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
	return v;
}

// A decoded micro-op as its owner lane holds it: byte offsets of f1, f2, c and (op | flags << 4).
#define LO(flag) ((unsigned)((flag) >> 48))
template <bool GS>
__device__ __forceinline__ uint4 decode_word(const uint2 w)
{
	return make_uint4((w.x & 0xffffu) << 8, (w.x >> 16) << 8, (w.y & 0xffffu) << 8, w.y >> 16);
}

// One micro-op of a SLOW group (quotients, pivots, copies, late operands).  A real call: slow groups
// are a small share of the work and the interpreter has to stay small enough for the instruction cache.
template <bool GS>
static __device__ __noinline__ double2 slow_op(const unsigned f1, const unsigned f2, const unsigned c, const unsigned fl, double P, double t, double d, char *g, const unsigned sb)
{
	const unsigned op = fl & 15u;
	if (fl & LO(SF_LATE1)) P = *reinterpret_cast<const double *>(g + f1);
	double S = 1.0;
	if (!(fl & LO(SF_ONE)) && (op == SOP_TERM || op == SOP_DIVACC || op == SOP_DIV || op == SOP_LOADS)) S = lds64(sb + f2);
	if (op == SOP_TERM) {
		const double base = (fl & LO(SF_FIRSTZ)) ? 0.0 : t;
		const double r = base - P * S;
		t = (fl & LO(SF_INIT)) ? P : r;
	} else if (op == SOP_DIVACC) {
		t = S / P;
		d = d - t * S;
	} else if (op == SOP_DIV) t = S / P;
	else if (op == SOP_INITD) d = P;
	else if (op == SOP_ENDD) t = regularise(d);
	else if (op == SOP_LOADS) t = S;
	if (fl & LO(SF_ST_HBM)) *reinterpret_cast<double *>(g + c) = t;
	if (!GS && (fl & LO(SF_ST_SHARED))) sts64(sb + c, t);
	return make_double2(t, d);
}

// slots [first, first + G) of `chunk`
template <int G>
__device__ __forceinline__ void request(const uint4 chunk, const int first, double (&p1)[G], const char *g)
{
	unsigned o1[G];
#pragma unroll
	for (int u = 0; u < G; ++u) o1[u] = __shfl_sync(0xffffffffu, chunk.x, first + u);
#pragma unroll
	for (int u = 0; u < G; ++u) p1[u] = ldg_now(g + o1[u]);
}

template <int G, bool GS>
__device__ __forceinline__ void execute(const uint4 chunk, const int first, const double (&p1)[G], double &t, double &d, char *g, const unsigned sb)
{
	unsigned f2[G], c[G], fl[G];
#pragma unroll
	for (int u = 0; u < G; ++u) {
		f2[u] = __shfl_sync(0xffffffffu, chunk.y, first + u);
		c[u] = __shfl_sync(0xffffffffu, chunk.z, first + u);
		fl[u] = __shfl_sync(0xffffffffu, chunk.w, first + u);
	}
	if (!(fl[0] & LO(SF_SLOW))) {
		// No micro-op of a fast group reads what an earlier one of the same group stores (the compiler
		// starts a new group instead): all second operands are read first, all results stored last, and
		// in between only the accumulator links the micro-ops.
		double s[G], out[G];
#pragma unroll
		for (int u = 0; u < G; ++u) s[u] = lds64(sb + f2[u]);
#pragma unroll
		for (int u = 0; u < G; ++u) {
			const double base = (fl[u] & LO(SF_FIRSTZ)) ? 0.0 : t;
			const double su = (GS && (fl[u] & LO(SF_ONE))) ? 1.0 : s[u];
			const double r = base - p1[u] * su;
			t = (fl[u] & LO(SF_INIT)) ? p1[u] : r;
			out[u] = t;
		}
#pragma unroll
		for (int u = 0; u < G; ++u) {
			if (GS) {
				if (fl[u] & LO(SF_ST_HBM)) *reinterpret_cast<double *>(g + c[u]) = out[u];
			} else {
				if (fl[u] & LO(SF_ST_SHARED)) sts64(sb + c[u], out[u]);
			}
		}
		return;
	}
#pragma unroll
	for (int u = 0; u < G; ++u) {
		const unsigned f1 = __shfl_sync(0xffffffffu, chunk.x, first + u);
		const double2 td = slow_op<GS>(f1, f2[u], c[u], fl[u], p1[u], t, d, g, sb);
		t = td.x;
		d = td.y;
	}
}

// ngroups is a multiple of DEPTH (the compiler pads); the program is followed by 4 chunks of padding.
// DEPTH = groups per chunk = operand buffers.
template <int G, bool GS, int DEPTH>
static __device__ __noinline__ void run_stream(const uint2 *__restrict__ prog, const int ngroups, double *slice, const unsigned sb)
{
	constexpr int CH = DEPTH * G; // micro-ops per chunk, one per lane (lanes beyond CH idle along)
	static_assert(CH <= 32, "a chunk is at most one micro-op per lane");
	char *g = reinterpret_cast<char *>(slice);
	const uint2 *mine = prog + (lane_id() < CH ? lane_id() : 0);
	uint4 here = decode_word<GS>(__ldg(mine)), next = decode_word<GS>(__ldg(mine + CH));
	uint2 far = __ldg(mine + 2 * CH);
	mine += 3 * CH;
	double p1[DEPTH][G];
#pragma unroll
	for (int k = 0; k + 1 < DEPTH; ++k) request<G>(here, k * G, p1[k], g);
	double t = 0.0, d = 0.0;
	for (int grp = 0; grp < ngroups; grp += DEPTH) {
#pragma unroll
		for (int k = 0; k < DEPTH; ++k) {
			const int ahead = (k + DEPTH - 1) % DEPTH; // buffer (and group within its chunk) of group i + DEPTH - 1
			request<G>(k == 0 ? here : next, ahead * G, p1[ahead], g);
			execute<G, GS>(here, k * G, p1[k], t, d, g, sb);
		}
		here = next;
		next = decode_word<GS>(far);
		far = __ldg(mine);
		mine += CH;
	}
}
#undef LO

// Sequential inner product sum_i a[i]*b[i] from 0.0 in ascending order (norm Auxilary.c:430-440,
// innerproduct 451-462); the loads and products of eight terms are issued together.
static __device__ __noinline__ double stream_dot(const double *a, const double *b, const int count)
{
	double acc = 0.0;
	for (int i0 = 0; i0 < count; i0 += 8) {
		double pa[8], pb[8];
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			const int i = min(i0 + j, count - 1);
			pa[j] = a[i << 5];
			pb[j] = b[i << 5];
		}
#pragma unroll
		for (int j = 0; j < 8; ++j)
			if (i0 + j < count) acc += pa[j] * pb[j];
	}
	return acc;
}

// One row of the API's row-per-QP arrays into a lane's slice (element e at to[e << 5]); the warp reads 256
// bytes at a time, eight such requests in flight.  PLUS_ZERO: v + 0.0 (-0.0 becomes +0.0).
template <bool PLUS_ZERO>
__device__ __forceinline__ void stream_take_row(double *to, const double *row, const int count, const int lane)
{
	for (int e0 = lane; e0 < count; e0 += 8 * 32) {
		double v[8];
#pragma unroll
		for (int j = 0; j < 8; ++j) v[j] = __ldcg(row + min(e0 + 32 * j, count - 1)); // (L2 only: rows may arrive while the kernel runs)
#pragma unroll
		for (int j = 0; j < 8; ++j)
			if (e0 + 32 * j < count) to[(size_t)(e0 + 32 * j) << 5] = PLUS_ZERO ? v[j] + 0.0 : v[j];
	}
}

// ... and a vector of a lane's slice out into a row of the API's arrays (up to 128 elements in flight).
__device__ __forceinline__ void stream_give_row(double *row, const double *from, const int count, const int lane)
{
	for (int e0 = lane; e0 < count; e0 += 4 * 32) {
		double v[4];
#pragma unroll
		for (int j = 0; j < 4; ++j) v[j] = from[(size_t)min(e0 + 32 * j, count - 1) << 5];
#pragma unroll
		for (int j = 0; j < 4; ++j)
			if (e0 + 32 * j < count) row[e0 + 32 * j] = v[j];
	}
}

// Out of line on purpose: an inlined FP64 division / square root is ~30 instructions, and the elementwise
// steps below are unrolled eight-fold; the kernel has to stay friendly to the instruction cache.
static __device__ __noinline__ double div_rn(const double a, const double b) { return a / b; }
static __device__ __noinline__ double sqrt_rn(const double a) { return sqrt(a); }

enum { SM_IDLE = 0, SM_INIT = 1, SM_PRED = 2, SM_CENT = 3 };
static const int STREAM_MAX_WARPS = 8;
static const int STREAM_U = 8; // elements of an elementwise step whose loads are issued together

#define SL(e) g[(unsigned)(e) << 5]
// for (i < count): LOADS; then BODY, eight elements at a time with all loads first
#define STREAM_FOR(count, LOADS, BODY)                                 \
	for (int i0__ = 0; i0__ < (count); i0__ += STREAM_U) {             \
		_Pragma("unroll") for (int j = 0; j < STREAM_U; ++j)           \
		{                                                              \
			const int i = min(i0__ + j, (count)-1);                    \
			LOADS                                                      \
		}                                                              \
		_Pragma("unroll") for (int j = 0; j < STREAM_U; ++j)           \
		{                                                              \
			const int i = i0__ + j;                                    \
			if (i < (count)) { BODY }                                  \
		}                                                              \
	}

// The last warp of the launch to leave tells the stragglers' server (batch.cu) that the list is complete.
__device__ __forceinline__ void stream_warp_leaves(const BatchIO &io)
{
	if (!io.all_done) return;
	__syncwarp();
	if (lane_id() == 0) {
		__threadfence(); // this warp's entries of the list before its count
		if (atomicAdd(io.done_ctas, 1) == (int)(gridDim.x * (blockDim.x >> 5)) - 1) {
			__threadfence();
			*reinterpret_cast<volatile int *>(io.all_done) = 1;
		}
	}
}

// Where the four micro-programs come from.  InterpretedProgram<G>: the words are fetched from memory
// and interpreted (run_stream above), whatever the pattern.  A GENERATED program (stream_gen.cuh,
// apf_quadruped_b200/codegen.py) is the same micro-op sequence of one particular pattern compiled into
// straight-line code; the kernel around the programs is the same.
template <int G>
struct InterpretedProgram
{
	static constexpr bool cta_sync = false; // the warps of a CTA never wait for one another
	static constexpr int ring_slots = 0;    // no operand ring in shared memory (see stream_gen.cuh)
	static __device__ __forceinline__ void factor(const StreamView &sv, double *slice, const unsigned sb, unsigned, bool) { run_stream<G, false, STREAM_FS_DEPTH(G)>(sv.words + sv.first[SSEG_FACTOR], sv.groups[SSEG_FACTOR], slice, sb); }
	static __device__ __forceinline__ void solve(const StreamView &sv, double *slice, const unsigned sb, unsigned, bool) { run_stream<G, false, STREAM_FS_DEPTH(G)>(sv.words + sv.first[SSEG_SOLVE], sv.groups[SSEG_SOLVE], slice, sb); }
	static __device__ __forceinline__ void resid(const StreamView &sv, double *slice, const unsigned sb, unsigned, bool) { run_stream<G, true, STREAM_FS_DEPTH(G)>(sv.words + sv.first[SSEG_RESID], sv.groups[SSEG_RESID], slice, sb); }
	static __device__ __forceinline__ void resz(const StreamView &sv, double *slice, const unsigned sb, unsigned, bool) { run_stream<G, true, STREAM_FS_DEPTH(G)>(sv.words + sv.first[SSEG_RESZ], sv.groups[SSEG_RESZ], slice, sb); }
};

// Does any lane want this step?  A program with cta_sync = true (generated straight-line code, which
// is far larger than the instruction caches) keeps the warps of its CTA walking the same code at the
// same time: the question is then put to the whole CTA, every warp enters the step (and meets the
// others at the barriers inside it), and a warp none of whose lanes wants it skips the work.
template <class PROG>
__device__ __forceinline__ bool any_wants(const bool mine)
{
	if (PROG::cta_sync) return __syncthreads_or(mine) != 0;
	return __any_sync(0xffffffffu, mine) != 0;
}

template <class PROG>
__global__ void __launch_bounds__(STREAM_MAX_WARPS * 32, 1) qp_solve_stream_kernel(const StreamView sv, const StreamLayout L, const StreamDims dm, const BatchIO io, const SolveParams prm, double *workspace, const int nslots, int *__restrict__ counter)
{
	extern __shared__ __align__(16) double smem[];
	const unsigned FULL = 0xffffffffu;
	const int lane = lane_id(), warp = threadIdx.x >> 5;
	const int slot = blockIdx.x * (blockDim.x >> 5) + warp;
	const bool live = slot < nslots;
	if (!live && !PROG::cta_sync) { // (with cta_sync a warp without a tile keeps the CTA's barriers company)
		stream_warp_leaves(io);
		return;
	}
	double *tile = workspace + (size_t)slot * L.total * 32; // the warp's 32 interleaved slices
	double *g = tile + lane;
	const unsigned sb = (unsigned)__cvta_generic_to_shared(smem) + (unsigned)(warp * L.scratch * 32 + lane) * 8u;
	// behind the windows of all warps: this warp's ring of HBM operands on their way in (generated programs only)
	const unsigned rb = (unsigned)__cvta_generic_to_shared(smem) + (unsigned)((blockDim.x >> 5) * L.scratch * 32 + warp * PROG::ring_slots * 32 + lane) * 8u;
	const int n = dm.n, m = dm.m, p = dm.p, np = dm.n + dm.p, N = dm.N;
	const double thr = prm.reltol / sqrt(3.0);
	if (live) {
		SL(L.zero) = 0.0;
		SL(L.one) = 1.0;
		SL(L.mone) = -1.0;
	}

	int q = -1, iters = 0;
	bool active = false, exhausted = !live;
	double sigma = 0.0, mu = 0.0, ap = 0.0, ad = 0.0;
	Scalars sc;
	sc.n_rx = sc.n_ry = sc.n_rz = sc.n_mu = sc.fval = 0.0;
	double a8[STREAM_U], b8[STREAM_U], c8[STREAM_U], d8[STREAM_U], e8[STREAM_U];

	for (;;) {
		// ---- top of QP_SOLVE's loop (qpSWIFT.c:502-534): residuals, stop test ---------------------
		int flag = -1;
		if (active && iters >= prm.maxit) flag = 2; // QP_MAXIT (qpSWIFT.c:604-607)
		else if (active && prm.spill_after && iters >= prm.spill_after) flag = 4; // a straggler: the latency kernel takes it (batch.cu)
		const bool need_res = active && flag < 0;
		if (any_wants<PROG>(need_res)) {
			const bool mine = __any_sync(FULL, need_res);
			if (need_res) { // s - h (= -(h - s), exactly), and x | y | z into the window for the residual program
				STREAM_FOR(m, a8[j] = SL(L.h + i); b8[j] = SL(L.s + i); c8[j] = SL(L.z + i);, SL(L.hs + i) = b8[j] - a8[j]; sts64(sb + ((unsigned)(np + i) << 8), c8[j]);)
				STREAM_FOR(n, a8[j] = SL(L.x + i);, sts64(sb + ((unsigned)i << 8), a8[j]);)
				STREAM_FOR(p, a8[j] = SL(L.y + i);, sts64(sb + ((unsigned)(n + i) << 8), a8[j]);)
			}
			PROG::resid(sv, g, sb, rb, mine);
			if (need_res) {
				sc.n_rx = sqrt(stream_dot(g + (L.r << 5), g + (L.r << 5), n));
				if (p) sc.n_ry = sqrt(stream_dot(g + ((L.r + n) << 5), g + ((L.r + n) << 5), p));
				sc.n_rz = sqrt(stream_dot(g + ((L.r + np) << 5), g + ((L.r + np) << 5), m));
				sc.n_mu = stream_dot(g + (L.s << 5), g + (L.z << 5), m) / m;
				const double q1 = stream_dot(g + (L.px << 5), g + (L.x << 5), n), q2 = stream_dot(g + (L.c << 5), g + (L.x << 5), n);
				sc.fval = -0.5 * q1 + q2;
				if (sc.n_rx < thr && sc.n_rz < thr && (p == 0 || sc.n_ry < thr) && sc.n_mu < prm.abstol) flag = 0;
			}
		}
		// ---- QPs that leave the loop: the warp writes them back, one finished lane at a time --------
		// (a straggler, flag 4, leaves its iterate in the output arrays too: the latency kernel resumes from it)
		unsigned done = __ballot_sync(FULL, active && flag >= 0);
		__syncwarp();
		while (done) {
			const int src = __ffs(done) - 1;
			done &= done - 1;
			const size_t qi = (size_t)__shfl_sync(FULL, q, src);
			const double *from = tile + src;
			if (io.x) stream_give_row(io.x + qi * n, from + ((size_t)L.x << 5), n, lane);
			if (io.y) stream_give_row(io.y + qi * p, from + ((size_t)L.y << 5), p, lane);
			if (io.z) stream_give_row(io.z + qi * m, from + ((size_t)L.z << 5), m, lane);
			if (io.s) stream_give_row(io.s + qi * m, from + ((size_t)L.s << 5), m, lane);
		}
		__syncwarp();
		if (active && flag == 4) {
			// list entry: QP | iterations << 24 when the iterate is handed over (all of x, z, s[, y] are written and
			// the batch is small enough for 24 bits), else the bare QP: the latency kernel starts it from scratch
			const bool hand_over = io.spill_sigma && io.x && io.z && io.s && (p == 0 || io.y) && io.count < (1 << 24) && iters < 128;
			if (hand_over) io.spill_sigma[q] = sigma;
			__threadfence(); // the iterate before the entry
			io.spill[atomicAdd(io.spill_count, 1)] = hand_over ? (q | (iters << 24)) : q;
			active = false;
		} else if (active && flag >= 0) {
			if (io.flag) io.flag[q] = flag;
			if (io.iters) io.iters[q] = iters;
			if (io.fval) io.fval[q] = sc.fval;
			if (io.stat) {
				double *st = io.stat + (size_t)q * 8;
				st[0] = sc.n_rx; st[1] = sc.n_ry; st[2] = sc.n_rz; st[3] = sc.n_mu;
				st[4] = ap; st[5] = ad; st[6] = sigma; st[7] = mu;
			}
			active = false;
		}
		// ---- free lanes pull the next QPs of the batch -----------------------------------------------
		int mode = SM_IDLE;
		const unsigned want = __ballot_sync(FULL, !active && !exhausted);
		if (want) {
			int base = 0;
			if (lane == 0) base = atomicAdd(counter, __popc(want));
			base = __shfl_sync(FULL, base, 0);
			bool fresh = false;
			if (!active && !exhausted) {
				q = base + __popc(want & ((1u << lane) - 1u));
				if (q < io.count) fresh = true;
				else { exhausted = true; q = -1; }
			}
			if (io.ready) { // host-pointer entry: wait until the rows of the QPs just taken have arrived (batch.cu)
				const int last = __reduce_max_sync(FULL, fresh ? q : -1);
				while (*reinterpret_cast<const volatile int *>(io.ready) <= last) __nanosleep(500);
				__threadfence(); // the rows were written by a copy engine before the mark
			}
			unsigned got = __ballot_sync(FULL, fresh);
			while (got) { // the warp transposes one value row at a time into the owner's slice
				const int dst = __ffs(got) - 1;
				got &= got - 1;
				const size_t qi = (size_t)__shfl_sync(FULL, q, dst);
				double *to = tile + dst;
				// "+ 0.0": the reference's Y[i] += value from Y[i] == 0 (ldl.c:284), see stream_program.h
				stream_take_row<true>(to + ((size_t)L.V << 5), io.Ppr + qi * dm.nnzP, dm.nnzP, lane);
				stream_take_row<true>(to + ((size_t)(L.V + dm.nnzP) << 5), io.Apr + qi * dm.nnzA, dm.nnzA, lane);
				stream_take_row<true>(to + ((size_t)(L.V + dm.nnzP + dm.nnzA) << 5), io.Gpr + qi * dm.nnzG, dm.nnzG, lane);
				stream_take_row<false>(to + ((size_t)L.c << 5), io.c + qi * n, n, lane);
				stream_take_row<false>(to + ((size_t)L.h << 5), io.h + qi * m, m, lane);
				for (int e = lane; e < p; e += 32) {
					const double be = __ldcg(io.b + qi * p + e);
					to[(L.b + e) << 5] = be;
					to[(L.nb + e) << 5] = -be;
				}
			}
			__syncwarp();
			if (fresh) { // kkt_initialize's system (Auxilary.c:1006-1030): rhs = [-c; b; h], (z,z) block = -I
				STREAM_FOR(n, a8[j] = SL(L.c + i);, SL(L.rhs + i) = -a8[j];)
				STREAM_FOR(p, a8[j] = SL(L.b + i);, SL(L.rhs + n + i) = a8[j];)
				STREAM_FOR(m, a8[j] = SL(L.h + i);, SL(L.rhs + np + i) = a8[j]; SL(L.zd + i) = -1.0;)
				iters = 0;
				sigma = prm.sigma0;
				mu = ap = ad = 0.0;
				mode = SM_INIT;
			}
		}
		// ---- an iteration of a running QP: lambda, mu, ds, W'W, right-hand side (qpSWIFT.c:537-552) ------
		if (active) {
			double acc = 0.0;
			STREAM_FOR(m, a8[j] = SL(L.s + i); b8[j] = SL(L.z + i);,
			           const double lam = sqrt_rn(a8[j] * b8[j]); SL(L.lam + i) = lam; acc += lam * lam;
			           SL(L.zd + i) = div_rn(-a8[j], b8[j]) + 0.0;)
			mu = acc / m;
			double sm = 0.0;
			if (sigma > prm.sigma_d) mode = SM_PRED;                         // predictor: ds = -lambda^2
			else { sigma = prm.sigma_d; sm = sigma * mu; mode = SM_CENT; }   // centering step only (qpSWIFT.c:572-579)
			const bool cent = mode == SM_CENT;
			STREAM_FOR(m, a8[j] = SL(L.lam + i); b8[j] = SL(L.z + i); c8[j] = SL(L.r + np + i);,
			           const double ds = cent ? -(a8[j] * a8[j]) + sm : -a8[j] * a8[j]; SL(L.ds + i) = ds;
			           SL(L.rhs + np + i) = c8[j] - div_rn(ds, b8[j]);)
			STREAM_FOR(np, a8[j] = SL(L.r + i);, SL(L.rhs + i) = a8[j];)
		}
		if (!any_wants<PROG>(mode != SM_IDLE)) break;
		const bool busy = __any_sync(FULL, mode != SM_IDLE);

		PROG::factor(sv, g, sb, rb, busy);
		PROG::solve(sv, g, sb, rb, busy);

		// ---- kkt_initialize, second half (Auxilary.c:1044-1083) ------------------------------------------
		if (any_wants<PROG>(mode == SM_INIT)) {
			const bool mine = __any_sync(FULL, mode == SM_INIT);
			if (mode == SM_INIT) {
				STREAM_FOR(n, a8[j] = SL(L.rhs + i);, SL(L.x + i) = a8[j]; sts64(sb + ((unsigned)i << 8), a8[j]);) // x also into the window
				STREAM_FOR(p, a8[j] = SL(L.rhs + n + i);, SL(L.y + i) = a8[j];)
				STREAM_FOR(m, a8[j] = SL(L.h + i);, SL(L.hs + i) = -a8[j];)
			}
			PROG::resz(sv, g, sb, rb, mine); // hs <- -(G x) - (-h) = h - G x
			if (mode == SM_INIT) {
				const double inf = __longlong_as_double(0x7ff0000000000000LL);
				double lo = inf, hi = -inf;
				STREAM_FOR(m, a8[j] = SL(L.hs + i);, if (a8[j] < lo) lo = a8[j]; if (a8[j] > hi) hi = a8[j];)
				const double alpha_p = -lo, alpha_d = hi;
				STREAM_FOR(m, a8[j] = SL(L.hs + i);,
				           SL(L.s + i) = (alpha_p < 0) ? a8[j] : a8[j] + (1 + alpha_p);
				           SL(L.z + i) = (alpha_d < 0) ? -a8[j] : -a8[j] + (1 + alpha_d);)
				active = true;
			}
		}
		// ---- delta_z, delta_s, step length, damping, update ------------------------------------------------
		// (Auxilary.c:508-511, 359-393, 246-254; qpSWIFT.c:584-596)
		auto step = [&](const bool take) {
			double tp = 1e10, td = 1e10;
			STREAM_FOR(m, a8[j] = SL(L.rhs + np + i); b8[j] = SL(L.ds + i); c8[j] = SL(L.s + i); d8[j] = SL(L.z + i);,
			           const double dzi = a8[j]; const double dsi = div_rn(b8[j] - (c8[j] * dzi), d8[j]);
			           SL(L.dz + i) = dzi; SL(L.dsl + i) = dsi;
			           if (dsi < 0) { const double ratio = div_rn(c8[j], dsi); if (-ratio < tp) tp = -ratio; }
			           if (dzi < 0) { const double ratio = div_rn(d8[j], dzi); if (-ratio < td) td = -ratio; })
			ap = tp < 1e10 ? tp : 1.0;
			ad = td < 1e10 ? td : 1.0;
			if (!take) return;
			ap = ((0.99 * ap) < (1.0)) ? (0.99 * ap) : (1.0);
			ad = ((0.99 * ad) < (1.0)) ? (0.99 * ad) : (1.0);
			STREAM_FOR(n, a8[j] = SL(L.x + i); b8[j] = SL(L.rhs + i);, SL(L.x + i) = a8[j] + b8[j] * ap;)
			STREAM_FOR(p, a8[j] = SL(L.y + i); b8[j] = SL(L.rhs + n + i);, SL(L.y + i) = a8[j] + b8[j] * ad;)
			STREAM_FOR(m, a8[j] = SL(L.s + i); b8[j] = SL(L.dsl + i); c8[j] = SL(L.z + i); d8[j] = SL(L.dz + i);,
			           SL(L.s + i) = a8[j] + b8[j] * ap; SL(L.z + i) = c8[j] + d8[j] * ad;)
			++iters;
		};
		if (mode == SM_CENT) step(true); // before the second solve overwrites this lane's solution
		if (mode == SM_PRED) { // affine step, rho, sigma (Auxilary.c:879-892, qpSWIFT.c:566-567), corrector right-hand side
			step(false);
			double num = 0.0, den = 0.0;
			STREAM_FOR(m, a8[j] = SL(L.s + i); b8[j] = SL(L.z + i); c8[j] = SL(L.dsl + i); d8[j] = SL(L.dz + i);,
			           num += (a8[j] + (ap * c8[j])) * (b8[j] + (ad * d8[j])); den += a8[j] * b8[j];)
			const double rho = num / den;
			const double mr = (1 < rho) ? 1 : rho;
			const double cube = mr * mr * mr;
			sigma = (prm.sigma_d > cube) ? prm.sigma_d : cube;
			const double sm = sigma * mu;
			STREAM_FOR(m, a8[j] = SL(L.lam + i); b8[j] = SL(L.dsl + i); c8[j] = SL(L.dz + i); d8[j] = SL(L.z + i); e8[j] = SL(L.r + np + i);,
			           const double ds = -(a8[j] * a8[j]) - (b8[j] * c8[j]) + sm; SL(L.ds + i) = ds;
			           SL(L.rhs + np + i) = e8[j] - div_rn(ds, d8[j]);)
			STREAM_FOR(np, a8[j] = SL(L.r + i);, SL(L.rhs + i) = a8[j];)
		}
		if (any_wants<PROG>(mode == SM_PRED)) {
			PROG::solve(sv, g, sb, rb, __any_sync(FULL, mode == SM_PRED));
			if (mode == SM_PRED) step(true);
		}
		(void)N;
	}
	stream_warp_leaves(io);
}

#undef SL
#undef STREAM_FOR

#endif // __CUDACC__

} // namespace qpb

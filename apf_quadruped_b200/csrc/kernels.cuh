// Device code of the batched interior-point solver (sm_100a, FP64, one QP per warp).
//
// Arithmetic contract: every floating-point result is produced by the same sequence of IEEE-754
// binary64 operations (same operands, same order, no fused multiply-add) as the reference's C
// code compiled for x86-64, so x, y, z, s, the iteration count and the exit flag are bit-identical
// with reference qpSWIFT.  The file must be compiled with -fmad=false.  Parallelism comes only
// from operations the reference performs on distinct targets (lanes own different targets) and
// from order-independent reductions (min / max); every sum the reference accumulates sequentially
// is accumulated sequentially here too.
//
// The factorisation and the triangular solves are not coded as loops: the host compiles them, per
// sparsity pattern, into wave schedules (program.h) that run_segment() below interprets.
//
// Reference routines restated (dogbot_controller/src/qpSWIFT/):
//   kkt_initialize Auxilary.c:992-1089      ipm_initial_point
//   computeresiduals 745-786, obj_value 1133-1141   ipm_residuals
//   formlambda 638-646, form_ds 315-345, updatekktmatrix 205-233, updatekktmatrix_b 274-295
//   kktsolve_1/2 471-564, findsteplength 359-393, formrho 879-892, updatevariables 246-254
//   LDL_numeric ldl.c:253-326               run_segment(factor schedule)
//   LDL_perm/lsolve/dsolve/ltsolve/permt ldl.c:495-597   run_segment(solve schedule)
//   QP_SOLVE qpSWIFT.c:473-644              qp_solve_cta_kernel (batch.cu)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "plan.h"
#include "program.h"

namespace qpb {

struct SolveParams
{
	int maxit;
	double reltol, abstol, sigma0, sigma_d;
	int spill_after; // throughput kernel: hand a QP that has not converged after this many iterations to the latency kernel (0: never)
	int groups;               // warp-per-QP kernel: 1 = the warps of a CTA walk the iteration together, 2 = two halves half a cycle apart
	unsigned group_delay_ns;
};

struct BatchIO
{
	const double *Ppr, *Apr, *Gpr, *c, *h, *b; // [count][...] value rows
	double *x, *y, *z, *s;                     // [count][...] outputs (may be null)
	int *flag, *iters;
	double *fval, *stat;                       // [count], [count][8] (may be null)
	int count;
	// second pass over the stragglers of the throughput kernel: solve QP index[i], i < *count_ptr
	const int *index;     // null: QP i is QP i
	const int *count_ptr; // null: count
	int *spill, *spill_count; // throughput kernel: where it lists the QPs it hands over (null: it never does)
	// hand-over of a straggler's iterate: x, y, z, s travel through the output arrays, the iteration count in the
	// list entry (QP | iterations << 24), sigma here (null: stragglers are restarted from scratch)
	double *spill_sigma;
	// throughput kernel behind QP_SOLVE_BATCH: the value rows are still arriving from the host while the kernel
	// runs; QPs [0, *ready) are there (null: all of them are)
	const int *ready;
	// latency kernel as the throughput kernel's straggler SERVER (launched next to it on an SM the throughput
	// kernel leaves free): QPs appear in index[] (-1 = not written yet) while *count_ptr grows; the server
	// leaves when *server_done is set and the list is empty (null: the list is complete at launch)
	const int *server_done;
	// throughput kernel: where its last warp to leave announces that nothing more will be listed
	int *done_ctas, *all_done; // (done_ctas counts warps)
	// warp-per-QP kernel, verbose single solves: per QP (maxit + 2) rows of 8 doubles -- row 0: seconds spent in
	// LDL_numeric (stats->ldl_numeric, Auxilary.c:476-484); row 1 + it: fval, n_rx, n_ry, n_rz, n_mu at the top of
	// iteration it (qpSWIFT.c:513-516) and the step lengths it ends with (598-601).  Null: no trace.
	double *trace;
	// warp-per-QP kernels whose programs keep the value row P | A | G outside the slice: one scratch row per warp of the
	// grid (GeneratedWarpKernel::value_row_doubles each) in global memory
	double *vscratch;
};

#ifdef __CUDACC__

struct PlanView
{
	int n, m, p, N, nV, lnz;
	const int *perm, *lp, *li;
	const int *scat_ptr, *scat_dst, *scat_src;
	const int *pat_ptr, *pat_col, *pat_slot;
	const int *prow_ptr, *prow_col, *prow_val;
	const int *arow_ptr, *arow_col, *arow_val;
	const int *grow_ptr, *grow_col, *grow_val;
	const int *acol_ptr, *acol_row;
	const int *gcol_ptr, *gcol_row;
	int offA, offG;
};

__device__ __forceinline__ PlanView view_of(const int *blob)
{
	const DevicePlanHeader *h = reinterpret_cast<const DevicePlanHeader *>(blob);
	PlanView v;
	v.n = h->n; v.m = h->m; v.p = h->p; v.N = h->N; v.nV = h->nV; v.lnz = h->lnz;
	v.perm = blob + h->perm; v.lp = blob + h->lp; v.li = blob + h->li;
	v.scat_ptr = blob + h->scat_ptr; v.scat_dst = blob + h->scat_dst; v.scat_src = blob + h->scat_src;
	v.pat_ptr = blob + h->pat_ptr; v.pat_col = blob + h->pat_col; v.pat_slot = blob + h->pat_slot;
	v.prow_ptr = blob + h->prow_ptr; v.prow_col = blob + h->prow_col; v.prow_val = blob + h->prow_val;
	v.arow_ptr = blob + h->arow_ptr; v.arow_col = blob + h->arow_col; v.arow_val = blob + h->arow_val;
	v.grow_ptr = blob + h->grow_ptr; v.grow_col = blob + h->grow_col; v.grow_val = blob + h->grow_val;
	v.acol_ptr = blob + h->acol_ptr; v.acol_row = blob + h->acol_row;
	v.gcol_ptr = blob + h->gcol_ptr; v.gcol_row = blob + h->gcol_row;
	v.offA = h->nnzP; v.offG = h->nnzP + h->nnzA;
	return v;
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// Coalesced copy of one QP's row, 128-bit when the row start is 16-byte aligned.  PLUS_ZERO: v + 0.0.
// L2_ONLY: ld.global.cg -- rows that a copy engine writes WHILE the kernel runs (host-pointer entry, batch.cu) must
// not be served from an L1 sector that an earlier read of the neighbouring row brought in before they arrived.
template <bool PLUS_ZERO = false, bool L2_ONLY = false>
__device__ __forceinline__ void warp_copy_in(double *dst, const double *__restrict__ src, int count)
{
	const int lane = lane_id();
	if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
		const int pairs = count >> 1;
		const double2 *s2 = reinterpret_cast<const double2 *>(src);
		double2 *d2 = reinterpret_cast<double2 *>(dst);
		for (int i = lane; i < pairs; i += 32) {
			double2 v = L2_ONLY ? __ldcg(s2 + i) : __ldg(s2 + i);
			if (PLUS_ZERO) { v.x += 0.0; v.y += 0.0; }
			d2[i] = v;
		}
		if ((count & 1) && lane == 0) {
			const double t = L2_ONLY ? __ldcg(src + count - 1) : __ldg(src + count - 1);
			dst[count - 1] = PLUS_ZERO ? t + 0.0 : t;
		}
	} else {
		for (int i = lane; i < count; i += 32) {
			const double t = L2_ONLY ? __ldcg(src + i) : __ldg(src + i);
			dst[i] = PLUS_ZERO ? t + 0.0 : t;
		}
	}
}

__device__ __forceinline__ void warp_copy_out(double *__restrict__ dst, const double *src, int count)
{
	if (!dst) return;
	const int lane = lane_id();
	if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
		const int pairs = count >> 1;
		const double2 *s2 = reinterpret_cast<const double2 *>(src);
		double2 *d2 = reinterpret_cast<double2 *>(dst);
		for (int i = lane; i < pairs; i += 32) d2[i] = s2[i];
		if ((count & 1) && lane == 0) dst[count - 1] = src[count - 1];
	} else {
		for (int i = lane; i < count; i += 32) dst[i] = src[i];
	}
}

// ------------------------------------------------------------------------------------------------
// Interpreter of the wave schedules (program.h).  `sb` is the 32-bit shared-window address of the
// QP's slice, `t` the lane's accumulator.  The wave kind is the same in all lanes (decided by a
// vote, so the branch is uniform); everything inside a wave is predicated, never divergent.
// Shared memory is addressed explicitly (ld/st.shared with 32-bit addresses): two integer
// instructions per operand address instead of a generic-pointer computation.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double regularise(double dk) // ldl.c:319-320
{
	const double sign = dk <= 0.0 ? -1.0 : 1.0;
	return (sign * dk <= 0.00000000000001) ? sign * 0.0000001 : dk;
}

__device__ __forceinline__ double lds64(unsigned addr)
{
	double v;
	asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
	return v;
}
__device__ __forceinline__ void sts64(unsigned addr, double v)
{
	asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}

// One slot of a quotient / ordered-sum super-wave.  Rare (a few dozen per factorisation) and long
// (an FP64 division is ~20 dependent instructions), so they are real calls, not inlined copies.
static __device__ __noinline__ void quotient_slot(const unsigned lo, const unsigned hi, const unsigned sb)
{
	const unsigned M = (1u << INS_ADDR_BITS) - 1u;
	// the divisor is a pivot: dynamic regularisation (ldl.c:319-320) applied on use
	if (hi & 0x8000u) sts64(sb + (((hi >> 16) & M) << 3), lds64(sb + ((lo & M) << 3)) / regularise(lds64(sb + ((hi & M) << 3))));
}

static __device__ __noinline__ void ordered_sum_slot(const unsigned lo, const unsigned hi, const unsigned sb)
{
	const unsigned M = (1u << INS_ADDR_BITS) - 1u;
	const int cnt = (hi & 0x8000u) ? (int)(hi & M) : 0;
	const int most = __reduce_max_sync(0xffffffffu, cnt);
	const unsigned a_at = sb + ((lo & M) << 3);
	double acc = lds64(sb + (((lo >> 16) & M) << 3)); // t = S[ci]; t -= S[a+i], i ascending
#pragma unroll 4
	for (int i = 0; i < most; ++i)
		if (i < cnt) acc = acc - lds64(a_at + 8u * (unsigned)i);
	if (cnt) sts64(sb + (((hi >> 16) & M) << 3), acc);
}

// One super-wave: U instruction words per lane, all of the same kind (a vote makes the branch
// uniform).  In a multiply-subtract super-wave the loads of all U terms are issued before any
// arithmetic, so their latencies overlap; everything is predicated, nothing diverges.
template <int U>
__device__ __forceinline__ void exec_super(const uint2 (&ins)[U], double (&t)[U], const unsigned sb)
{
	const unsigned M = (1u << INS_ADDR_BITS) - 1u;
	if (__any_sync(0xffffffffu, ins[0].y & 0xC0000000u)) {
		if (__any_sync(0xffffffffu, ins[0].y & 0x80000000u)) {
#pragma unroll
			for (int u = 0; u < U; ++u) quotient_slot(ins[u].x, ins[u].y, sb);
		} else {
#pragma unroll
			for (int u = 0; u < U; ++u) ordered_sum_slot(ins[u].x, ins[u].y, sb);
		}
		return;
	}
	double va[U], vb[U], vi[U];
#pragma unroll
	for (int u = 0; u < U; ++u) {
		va[u] = lds64(sb + ((ins[u].x & M) << 3));
		vb[u] = lds64(sb + ((ins[u].y & M) << 3));
	}
#pragma unroll
	for (int u = 0; u < U; ++u) { // FIRST: fetch the initial value
		vi[u] = 0.0;
		asm volatile("{\n\t.reg .pred p;\n\t"
		             "setp.ne.u32 p, %2, 0;\n\t"
		             "@p ld.shared.f64 %0, [%1];\n\t}"
		             : "+d"(vi[u])
		             : "r"(sb + (((ins[u].x >> 16) & M) << 3)), "r"(ins[u].x & 0x4000u));
	}
#pragma unroll
	for (int u = 0; u < U; ++u) {
		const unsigned lo = ins[u].x;
		// FIRST: t = S[ci] + (NEGZ ? -0.0 : +0.0);  then t -= (ADD ? -1 : 1) * va*vb
		if (lo & 0x4000u) t[u] = vi[u] + __hiloint2double((int)((lo << 1) & 0x80000000u), 0);
		double prod = va[u] * vb[u];
		prod = __hiloint2double(__double2hiint(prod) ^ (int)(lo & 0x80000000u), __double2loint(prod));
		t[u] = t[u] - prod;
	}
#pragma unroll
	for (int u = 0; u < U; ++u) // LAST: S[cs] = t
		asm volatile("{\n\t.reg .pred p;\n\t"
		             "setp.ne.u32 p, %2, 0;\n\t"
		             "@p st.shared.f64 [%0], %1;\n\t}" ::"r"(sb + (((ins[u].y >> 16) & M) << 3)), "d"(t[u]), "r"(ins[u].x & 0x8000u)
		             : "memory");
}

template <int U>
__device__ __forceinline__ void load_words(uint2 (&dst)[U], const uint2 *__restrict__ p)
{
	if (U == 1) dst[0] = __ldg(p);
	else {
#pragma unroll
		for (int u = 0; u < U; u += 2) {
			const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p + u));
			dst[u] = make_uint2(v.x, v.y);
			dst[u + 1] = make_uint2(v.z, v.w);
		}
	}
}

// Runs `waves` super-waves (a multiple of PROGRAM_WAVE_PAD) starting `first` * 32 words into `prog`.
// PROGRAM_WAVE_PAD super-waves' worth of instruction words are always in flight (read-only path): a warp
// that runs alone on its SM -- the stragglers of the throughput kernel, the single-QP QP_SOLVE -- gets
// every word from L2, and a super-wave is much shorter than that round trip; the program buffer is padded
// so the prefetch needs no bounds check.
template <int U>
__device__ __forceinline__ void run_segment_w(const uint2 *__restrict__ prog, int first, int waves, const unsigned sb)
{
	constexpr int D = PROGRAM_WAVE_PAD;
	const uint2 *p = prog + (size_t)first * 32 + lane_id() * U;
	uint2 W[D][U];
#pragma unroll
	for (int k = 0; k < D; ++k) load_words<U>(W[k], p + 32 * U * k);
	double t[U];
#pragma unroll
	for (int u = 0; u < U; ++u) t[u] = 0.0;
	for (int w = 0; w < waves; w += D) {
		p += 32 * U * D;
#pragma unroll
		for (int k = 0; k < D; ++k) {
			exec_super<U>(W[k], t, sb);
			__syncwarp();
			load_words<U>(W[k], p + 32 * U * k);
		}
	}
}

__device__ __forceinline__ void run_segment(const uint2 *__restrict__ prog, int first, int waves, int width, const unsigned sb)
{
	if (width == 4) run_segment_w<4>(prog, first, waves, sb);
	else if (width == 2) run_segment_w<2>(prog, first, waves, sb);
	else run_segment_w<1>(prog, first, waves, sb);
}

struct Scalars
{
	double n_rx, n_ry, n_rz, n_mu, fval;
};

// ------------------------------------------------------------------------------------------------
// computeresiduals (Auxilary.c:745-786) + obj_value (1133-1141).
// Each lane owns output rows and accumulates them in the order the reference's column-scatter /
// column-gather products deliver the terms: ascending column for P*x, A*x, G*x; ascending row
// inside a column for G'z then A'y.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ Scalars ipm_residuals(const PlanView &pv, const double *V, const double *c, const double *h, const double *b, const double *x, const double *y, const double *z, const double *s, double *r, double *tmp)
{
	const int lane = lane_id();
	const int n = pv.n, m = pv.m, p = pv.p;
	for (int i = lane; i < n; i += 32) {
		double acc = 0.0;
		for (int t = __ldg(pv.prow_ptr + i); t < __ldg(pv.prow_ptr + i + 1); ++t) acc -= x[__ldg(pv.prow_col + t)] * V[__ldg(pv.prow_val + t)];
		tmp[i] = acc;                                          // = -(P x)_i, reused by obj_value
		for (int t = __ldg(pv.gcol_ptr + i); t < __ldg(pv.gcol_ptr + i + 1); ++t) acc -= V[pv.offG + t] * z[__ldg(pv.gcol_row + t)];
		if (p)
			for (int t = __ldg(pv.acol_ptr + i); t < __ldg(pv.acol_ptr + i + 1); ++t) acc -= V[pv.offA + t] * y[__ldg(pv.acol_row + t)];
		acc += c[i] * -1.0;
		r[i] = acc;
	}
	for (int i = lane; i < p; i += 32) {
		double acc = 0.0;
		for (int t = __ldg(pv.arow_ptr + i); t < __ldg(pv.arow_ptr + i + 1); ++t) acc -= x[__ldg(pv.arow_col + t)] * V[__ldg(pv.arow_val + t)];
		acc += b[i] * 1.0;
		r[n + i] = acc;
	}
	for (int i = lane; i < m; i += 32) {
		double acc = 0.0;
		for (int t = __ldg(pv.grow_ptr + i); t < __ldg(pv.grow_ptr + i + 1); ++t) acc -= x[__ldg(pv.grow_col + t)] * V[__ldg(pv.grow_val + t)];
		acc += h[i] - s[i];
		r[n + p + i] = acc;
	}
	__syncwarp();
	// sequential sums, one per lane (norm 430-440, innerproduct 451-462)
	double v = 0.0;
	if (lane == 0) { for (int i = 0; i < n; ++i) v += r[i] * r[i]; v = sqrt(v); }
	else if (lane == 1) { for (int i = 0; i < p; ++i) v += r[n + i] * r[n + i]; v = sqrt(v); }
	else if (lane == 2) { for (int i = 0; i < m; ++i) v += r[n + p + i] * r[n + p + i]; v = sqrt(v); }
	else if (lane == 3) { for (int i = 0; i < m; ++i) v += s[i] * z[i]; v = v / m; }
	else if (lane == 4) { for (int i = 0; i < n; ++i) v += tmp[i] * x[i]; }
	else if (lane == 5) { for (int i = 0; i < n; ++i) v += c[i] * x[i]; }
	Scalars sc;
	sc.n_rx = __shfl_sync(0xffffffffu, v, 0);
	sc.n_ry = __shfl_sync(0xffffffffu, v, 1);
	sc.n_rz = __shfl_sync(0xffffffffu, v, 2);
	sc.n_mu = __shfl_sync(0xffffffffu, v, 3);
	const double q1 = __shfl_sync(0xffffffffu, v, 4), q2 = __shfl_sync(0xffffffffu, v, 5);
	sc.fval = -0.5 * q1 + q2;
	return sc;
}

__device__ __forceinline__ double warp_min(double v)
{
	for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}
__device__ __forceinline__ double warp_max(double v)
{
	for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}

// findsteplength (Auxilary.c:359-393): the smallest ratio below 1e10, else 1.  min() is exact and
// order independent, so a shuffle reduction gives the reference's value.
__device__ __forceinline__ void ipm_steplength(int m, const double *s, const double *dsl, const double *z, const double *dz, double &alpha_p, double &alpha_d)
{
	const int lane = lane_id();
	double ap = 1e10, ad = 1e10;
	for (int i = lane; i < m; i += 32) {
		if (dsl[i] < 0 && (-s[i] / dsl[i]) < ap) ap = -(s[i] / dsl[i]);
		if (dz[i] < 0 && (-z[i] / dz[i]) < ad) ad = -(z[i] / dz[i]);
	}
	ap = warp_min(ap);
	ad = warp_min(ad);
	alpha_p = ap < 1e10 ? ap : 1.0;
	alpha_d = ad < 1e10 ? ad : 1.0;
}

// rhs = [rx; ry; rz - ds/z]  (updatekktmatrix_b, Auxilary.c:274-295)
__device__ __forceinline__ void ipm_rhs(const PlanView &pv, const double *r, const double *ds, const double *z, double *rhs)
{
	const int lane = lane_id();
	const int np = pv.n + pv.p;
	for (int i = lane; i < np; i += 32) rhs[i] = r[i];
	for (int i = lane; i < pv.m; i += 32) rhs[np + i] = r[np + i] - (ds[i] / z[i]);
	__syncwarp();
}

struct WarpState
{
	double *S; // slice base
	double *V, *c, *h, *b, *zd, *rhs, *x, *y, *z, *s, *r, *lam, *ds, *dz, *dsl;
};

__device__ __forceinline__ WarpState carve(double *base, const SliceLayout &L)
{
	WarpState w;
	w.S = base;
	w.V = base + L.V; w.c = base + L.c; w.h = base + L.h; w.b = base + L.b; w.zd = base + L.zd;
	w.rhs = base + L.rhs;
	w.x = base + L.x; w.y = base + L.y; w.z = base + L.z; w.s = base + L.s;
	w.r = base + L.r; w.lam = base + L.lam; w.ds = base + L.ds; w.dz = base + L.dz; w.dsl = base + L.dsl;
	return w;
}

// kkt_initialize, first half (Auxilary.c:1006-1030): rhs = [-c; b; h], (z,z) block = -I
__device__ __forceinline__ void ipm_init_rhs(const PlanView &pv, WarpState &w)
{
	const int lane = lane_id();
	const int n = pv.n, m = pv.m, p = pv.p;
	for (int i = lane; i < n; i += 32) w.rhs[i] = -w.c[i];
	for (int i = lane; i < p; i += 32) w.rhs[n + i] = w.b[i];
	for (int i = lane; i < m; i += 32) { w.rhs[n + p + i] = w.h[i]; w.zd[i] = -1.0; }
	__syncwarp();
}

// kkt_initialize, second half (Auxilary.c:1044-1083): x, y from the solution, r = h - G x,
// s and z shifted into the cone.
// VSLOT(t, pv): where the t-th entry of G by rows (the plan's grow lists) sits in w.V: pv.grow_val[t], or its slot of a
// value row kept in residual order.
template <class VSLOT>
__device__ __forceinline__ void ipm_init_point(const PlanView &pv, WarpState &w, const VSLOT vslot)
{
	const int lane = lane_id();
	const int n = pv.n, m = pv.m, p = pv.p;
	for (int i = lane; i < n; i += 32) w.x[i] = w.rhs[i];
	for (int i = lane; i < p; i += 32) w.y[i] = w.rhs[n + i];
	__syncwarp();
	double lo = 0.0, hi = 0.0;
	bool first = true;
	double *zi = w.dz; // scratch
	for (int i = lane; i < m; i += 32) {
		double acc = 0.0;
		for (int t = __ldg(pv.grow_ptr + i); t < __ldg(pv.grow_ptr + i + 1); ++t) acc -= w.x[__ldg(pv.grow_col + t)] * w.V[vslot(t, pv)];
		acc += w.h[i] * 1.0;
		zi[i] = acc;
		if (first) { lo = hi = acc; first = false; }
		else { if (acc < lo) lo = acc; if (acc > hi) hi = acc; }
	}
	if (first) { lo = __longlong_as_double(0x7ff0000000000000LL); hi = -lo; }
	lo = warp_min(lo);
	hi = warp_max(hi);
	const double alpha_p = -lo, alpha_d = hi;
	for (int i = lane; i < m; i += 32) {
		const double v = zi[i];
		w.s[i] = (alpha_p < 0) ? v : v + (1 + alpha_p);
		w.z[i] = (alpha_d < 0) ? -v : -v + (1 + alpha_d);
	}
	__syncwarp();
}

// lambda = sqrt(s.z), mu = lambda'lambda / m   (qpSWIFT.c:537-538)
__device__ __forceinline__ double ipm_lambda_mu(int m, WarpState &w)
{
	const int lane = lane_id();
	for (int i = lane; i < m; i += 32) w.lam[i] = sqrt(w.s[i] * w.z[i]);
	__syncwarp();
	double mu = 0.0;
	if (lane == 0) { for (int i = 0; i < m; ++i) mu += w.lam[i] * w.lam[i]; mu = mu / m; }
	return __shfl_sync(0xffffffffu, mu, 0);
}

// (z,z) block of the KKT matrix: -s/z   (updatekktmatrix, Auxilary.c:211-215)
__device__ __forceinline__ void ipm_update_kkt(int m, WarpState &w)
{
	for (int i = lane_id(); i < m; i += 32) w.zd[i] = -w.s[i] / w.z[i];
}

// delta_z from the solution, delta_s = (ds - s.delta_z)/z   (Auxilary.c:508-511, 556-562)
__device__ __forceinline__ void ipm_recover_dz_ds(const PlanView &pv, WarpState &w)
{
	const int np = pv.n + pv.p;
	for (int i = lane_id(); i < pv.m; i += 32) {
		const double dzi = w.rhs[np + i];
		w.dz[i] = dzi;
		w.dsl[i] = (w.ds[i] - (w.s[i] * dzi)) / w.z[i];
	}
	__syncwarp();
}

// rho (Auxilary.c:879-892) and sigma = max(sigma_d, min(1,rho)^3) (qpSWIFT.c:566-567)
__device__ __forceinline__ double ipm_sigma(int m, const WarpState &w, double ap, double ad, double sigma_d)
{
	const int lane = lane_id();
	double a = 0.0;
	if (lane == 0) { for (int i = 0; i < m; ++i) a += (w.s[i] + (ap * w.dsl[i])) * (w.z[i] + (ad * w.dz[i])); }
	else if (lane == 1) { for (int i = 0; i < m; ++i) a += w.s[i] * w.z[i]; }
	const double num = __shfl_sync(0xffffffffu, a, 0), den = __shfl_sync(0xffffffffu, a, 1);
	const double rho = num / den;
	const double mr = (1 < rho) ? 1 : rho;
	const double cube = mr * mr * mr;
	return (sigma_d > cube) ? sigma_d : cube;
}

#endif // __CUDACC__

} // namespace qpb

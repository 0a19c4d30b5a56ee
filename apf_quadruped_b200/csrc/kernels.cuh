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
// Reference routines restated (dogbot_controller/src/qpSWIFT/):
//   kkt_initialize Auxilary.c:992-1089      ipm_initial_point
//   computeresiduals 745-786, obj_value 1133-1141   ipm_residuals
//   formlambda 638-646, form_ds 315-345, updatekktmatrix 205-233, updatekktmatrix_b 274-295
//   kktsolve_1/2 471-564, findsteplength 359-393, formrho 879-892, updatevariables 246-254
//   LDL_numeric ldl.c:253-326               ldl_factor
//   LDL_perm/lsolve/dsolve/ltsolve/permt ldl.c:495-597   ldl_solve
//   QP_SOLVE qpSWIFT.c:473-644              ipm_solve
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "plan.h"

namespace qpb {

struct SolveParams
{
	int maxit;
	double reltol, abstol, sigma0, sigma_d;
};

// Per-warp shared-memory layout, in doubles from the warp's base.
struct SmemLayout
{
	int V, c, h, b;      // value row: P|A|G values, then c, h, b
	int Lx, D, Y, rhs;   // factor, work vector, right-hand side / solution
	int x, y, z, s;      // iterate
	int r;               // residuals [rx; ry; rz]
	int lam, ds, dz, dsl;
	int total;           // doubles per warp
};

struct BatchIO
{
	const double *Ppr, *Apr, *Gpr, *c, *h, *b; // [count][...] value rows
	double *x, *y, *z, *s;                     // [count][...] outputs (may be null)
	int *flag, *iters;
	double *fval, *stat;                       // [count], [count][8] (may be null)
	int count;
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

// Coalesced copy of one QP's row, 128-bit when the row start is 16-byte aligned.
__device__ __forceinline__ void warp_copy_in(double *dst, const double *__restrict__ src, int count)
{
	const int lane = lane_id();
	if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
		const int pairs = count >> 1;
		const double2 *s2 = reinterpret_cast<const double2 *>(src);
		double2 *d2 = reinterpret_cast<double2 *>(dst);
		for (int i = lane; i < pairs; i += 32) d2[i] = __ldg(s2 + i);
		if ((count & 1) && lane == 0) dst[count - 1] = __ldg(src + count - 1);
	} else {
		for (int i = lane; i < count; i += 32) dst[i] = __ldg(src + i);
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
// LDL_numeric (ldl.c:253-326), up-looking, one pivot row at a time.
//   V      : the QP's value row (P|A|G) in shared memory
//   zd     : (z,z) diagonal source; zd_const != 0 -> every diagonal is zd_const (the -I of
//            kkt_initialize), else the diagonal of inequality i is -s[i]/z[i] (updatekktmatrix,
//            Auxilary.c:211-215)
// The column update  Y[Li[p]] -= Lx[p]*yi  touches distinct targets, so the lanes split it; the
// pivot recurrence  D[k] -= l_ki*yi  is carried redundantly by all lanes in the reference order.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldl_factor(const PlanView &pv, const double *V, const double *sv, const double *zv, double zd_const, double *Lx, double *D, double *Y)
{
	const int lane = lane_id();
	for (int i = lane; i < pv.N; i += 32) Y[i] = 0.0;
	__syncwarp();
	for (int k = 0; k < pv.N; ++k) {
		const int s0 = __ldg(pv.scat_ptr + k), s1 = __ldg(pv.scat_ptr + k + 1);
		for (int t = s0 + lane; t < s1; t += 32) {
			const int dst = __ldg(pv.scat_dst + t), src = __ldg(pv.scat_src + t);
			double val;
			if (src < pv.nV) val = V[src];
			else if (zd_const != 0.0) val = zd_const;
			else val = -sv[src - pv.nV] / zv[src - pv.nV];
			Y[dst] += val;
		}
		__syncwarp();
		double dk = Y[k];
		__syncwarp();
		if (lane == 0) Y[k] = 0.0;
		const int p0 = __ldg(pv.pat_ptr + k), p1 = __ldg(pv.pat_ptr + k + 1);
		for (int t = p0; t < p1; ++t) {
			const int i = __ldg(pv.pat_col + t), slot = __ldg(pv.pat_slot + t);
			const double yi = Y[i];
			__syncwarp();
			for (int q = __ldg(pv.lp + i) + lane; q < slot; q += 32) Y[__ldg(pv.li + q)] -= Lx[q] * yi;
			const double l = yi / D[i];
			dk -= l * yi;
			if (lane == 0) { Y[i] = 0.0; Lx[slot] = l; }
			__syncwarp();
		}
		const double sign = dk <= 0.0 ? -1.0 : 1.0;          // dynamic regularisation, ldl.c:319-320
		dk = (sign * dk <= 0.00000000000001) ? sign * 0.0000001 : dk;
		if (lane == 0) D[k] = dk;
		__syncwarp();
	}
}

// ------------------------------------------------------------------------------------------------
// x = P' L'^-1 D^-1 L^-1 P b (ldl.c:495-597), in place on `rhs`, `w` is an N-vector of scratch.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldl_solve(const PlanView &pv, const double *Lx, const double *D, double *rhs, double *w)
{
	const int lane = lane_id();
	const int N = pv.N;
	for (int j = lane; j < N; j += 32) w[j] = rhs[__ldg(pv.perm + j)];
	__syncwarp();
	for (int j = 0; j < N; ++j) {                              // forward: column scatter
		const int q0 = __ldg(pv.lp + j), q1 = __ldg(pv.lp + j + 1);
		if (q0 == q1) continue;
		const double xj = w[j];
		for (int q = q0 + lane; q < q1; q += 32) w[__ldg(pv.li + q)] -= Lx[q] * xj;
		__syncwarp();
	}
	for (int j = lane; j < N; j += 32) w[j] /= D[j];
	__syncwarp();
	for (int j = N - 1; j >= 0; --j) {                         // backward: column gather, sequential sum
		const int q0 = __ldg(pv.lp + j), q1 = __ldg(pv.lp + j + 1);
		if (q0 == q1) continue;
		// products in parallel, subtraction chain in the reference order by lane 0
		double acc = w[j];
		for (int base = q0; base < q1; base += 32) {
			const int q = base + lane;
			double t = 0.0;
			if (q < q1) t = Lx[q] * w[__ldg(pv.li + q)];
			const int cnt = min(32, q1 - base);
			for (int u = 0; u < cnt; ++u) acc -= __shfl_sync(0xffffffffu, t, u);
		}
		if (lane == 0) w[j] = acc;
		__syncwarp();
	}
	for (int j = lane; j < N; j += 32) rhs[__ldg(pv.perm + j)] = w[j];
	__syncwarp();
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
	double *V, *c, *h, *b, *Lx, *D, *Y, *rhs, *x, *y, *z, *s, *r, *lam, *ds, *dz, *dsl;
};

__device__ __forceinline__ WarpState carve(double *base, const SmemLayout &L)
{
	WarpState w;
	w.V = base + L.V; w.c = base + L.c; w.h = base + L.h; w.b = base + L.b;
	w.Lx = base + L.Lx; w.D = base + L.D; w.Y = base + L.Y; w.rhs = base + L.rhs;
	w.x = base + L.x; w.y = base + L.y; w.z = base + L.z; w.s = base + L.s;
	w.r = base + L.r; w.lam = base + L.lam; w.ds = base + L.ds; w.dz = base + L.dz; w.dsl = base + L.dsl;
	return w;
}

// kkt_initialize (Auxilary.c:992-1089)
__device__ __forceinline__ void ipm_initial_point(const PlanView &pv, WarpState &w)
{
	const int lane = lane_id();
	const int n = pv.n, m = pv.m, p = pv.p;
	for (int i = lane; i < n; i += 32) w.rhs[i] = -w.c[i];
	for (int i = lane; i < p; i += 32) w.rhs[n + i] = w.b[i];
	for (int i = lane; i < m; i += 32) w.rhs[n + p + i] = w.h[i];
	__syncwarp();
	ldl_factor(pv, w.V, nullptr, nullptr, -1.0, w.Lx, w.D, w.Y);
	ldl_solve(pv, w.Lx, w.D, w.rhs, w.Y);
	for (int i = lane; i < n; i += 32) w.x[i] = w.rhs[i];
	for (int i = lane; i < p; i += 32) w.y[i] = w.rhs[n + i];
	__syncwarp();
	// r = h - G x, its min and max
	double lo = 0.0, hi = 0.0;
	bool first = true;
	double *zi = w.dz; // scratch
	for (int i = lane; i < m; i += 32) {
		double acc = 0.0;
		for (int t = __ldg(pv.grow_ptr + i); t < __ldg(pv.grow_ptr + i + 1); ++t) acc -= w.x[__ldg(pv.grow_col + t)] * w.V[__ldg(pv.grow_val + t)];
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

struct SolveResult
{
	int flag, iters;
	Scalars sc;
	double alpha_p, alpha_d, sigma, mu;
};

// QP_SOLVE main loop (qpSWIFT.c:499-607)
__device__ __forceinline__ SolveResult ipm_solve(const PlanView &pv, WarpState &w, const SolveParams &prm)
{
	const int lane = lane_id();
	const int n = pv.n, m = pv.m, p = pv.p, np = n + p;
	SolveResult res;
	res.flag = 3; // QP_FATAL until proven otherwise (qpSWIFT.c:80)
	res.iters = 0;
	res.alpha_p = 0.0;
	res.alpha_d = 0.0;
	res.mu = 0.0;
	double sigma = prm.sigma0;
	const double thr = prm.reltol / sqrt(3.0);
	for (int it = 0; it < prm.maxit; ++it) {
		res.sc = ipm_residuals(pv, w.V, w.c, w.h, w.b, w.x, w.y, w.z, w.s, w.r, w.lam /* scratch: -(Px) */);
		const bool small = res.sc.n_rx < thr && res.sc.n_rz < thr && (p == 0 || res.sc.n_ry < thr) && res.sc.n_mu < prm.abstol;
		if (small) { res.flag = 0; break; }

		// lambda = sqrt(s.z), mu = lambda'lambda / m   (qpSWIFT.c:537-538)
		__syncwarp();
		for (int i = lane; i < m; i += 32) w.lam[i] = sqrt(w.s[i] * w.z[i]);
		__syncwarp();
		double mu = 0.0;
		if (lane == 0) { for (int i = 0; i < m; ++i) mu += w.lam[i] * w.lam[i]; mu = mu / m; }
		mu = __shfl_sync(0xffffffffu, mu, 0);
		res.mu = mu;

		double ap, ad;
		if (sigma > prm.sigma_d) {
			// predictor: ds = -lambda^2, refactor with K_zz = -s/z, solve
			for (int i = lane; i < m; i += 32) w.ds[i] = -w.lam[i] * w.lam[i];
			__syncwarp();
			ipm_rhs(pv, w.r, w.ds, w.z, w.rhs);
			ldl_factor(pv, w.V, w.s, w.z, 0.0, w.Lx, w.D, w.Y);
			ldl_solve(pv, w.Lx, w.D, w.rhs, w.Y);
			for (int i = lane; i < m; i += 32) {
				const double dzi = w.rhs[np + i];
				w.dz[i] = dzi;
				w.dsl[i] = (w.ds[i] - (w.s[i] * dzi)) / w.z[i];
			}
			__syncwarp();
			ipm_steplength(m, w.s, w.dsl, w.z, w.dz, ap, ad);
			// rho, sigma   (Auxilary.c:879-892, qpSWIFT.c:566-567)
			double a = 0.0;
			if (lane == 0) { for (int i = 0; i < m; ++i) a += (w.s[i] + (ap * w.dsl[i])) * (w.z[i] + (ad * w.dz[i])); }
			else if (lane == 1) { for (int i = 0; i < m; ++i) a += w.s[i] * w.z[i]; }
			const double num = __shfl_sync(0xffffffffu, a, 0), den = __shfl_sync(0xffffffffu, a, 1);
			const double rho = num / den;
			const double mr = (1 < rho) ? 1 : rho;
			const double cube = mr * mr * mr;
			sigma = (prm.sigma_d > cube) ? prm.sigma_d : cube;
			// corrector right-hand side
			const double sm = sigma * mu;
			for (int i = lane; i < m; i += 32) w.ds[i] = -(w.lam[i] * w.lam[i]) - (w.dsl[i] * w.dz[i]) + sm;
			__syncwarp();
			ipm_rhs(pv, w.r, w.ds, w.z, w.rhs);
		} else {
			sigma = prm.sigma_d;
			const double sm = sigma * mu;
			for (int i = lane; i < m; i += 32) w.ds[i] = -(w.lam[i] * w.lam[i]) + sm;
			__syncwarp();
			ipm_rhs(pv, w.r, w.ds, w.z, w.rhs);
			ldl_factor(pv, w.V, w.s, w.z, 0.0, w.Lx, w.D, w.Y);
		}
		ldl_solve(pv, w.Lx, w.D, w.rhs, w.Y);
		for (int i = lane; i < m; i += 32) {
			const double dzi = w.rhs[np + i];
			w.dz[i] = dzi;
			w.dsl[i] = (w.ds[i] - (w.s[i] * dzi)) / w.z[i];
		}
		__syncwarp();
		ipm_steplength(m, w.s, w.dsl, w.z, w.dz, ap, ad);
		ap = ((0.99 * ap) < (1.0)) ? (0.99 * ap) : (1.0);
		ad = ((0.99 * ad) < (1.0)) ? (0.99 * ad) : (1.0);
		res.alpha_p = ap;
		res.alpha_d = ad;
		for (int i = lane; i < n; i += 32) w.x[i] += w.rhs[i] * ap;
		for (int i = lane; i < p; i += 32) w.y[i] += w.rhs[n + i] * ad;
		for (int i = lane; i < m; i += 32) { w.s[i] += w.dsl[i] * ap; w.z[i] += w.dz[i] * ad; }
		__syncwarp();
		++res.iters;
	}
	if (res.iters == prm.maxit) res.flag = 2; // QP_MAXIT (qpSWIFT.c:604-607)
	res.sigma = sigma;
	return res;
}

#endif // __CUDACC__

} // namespace qpb

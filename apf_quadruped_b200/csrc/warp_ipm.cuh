// Elementwise steps of the interior-point iteration for the warp-per-QP kernels, fused: every vector of a
// QP is read once per step from the warp's slice of shared memory, and the reference's sequential sums
// (norms, inner products, rho) run in parallel LANES -- one sum per lane, each accumulated in the
// reference's own order from 0.0 -- instead of one after the other.
//
// Arithmetic contract as in kernels.cuh: every value is produced by the reference's IEEE-754 binary64
// operations in the reference's order (-fmad=false); only min/max are reordered.
// Reference: computeresiduals Auxilary.c:745-786 (sums 430-462), obj_value 1133-1141, formlambda 638-646,
// form_ds 315-345, updatekktmatrix 205-233, updatekktmatrix_b 274-295, kktsolve_1/2 508-511 / 556-562,
// findsteplength 359-393, formrho 879-892, updatevariables 246-254, qpSWIFT.c:537-596.
#pragma once

#include "kernels.cuh"

namespace qpb {

#ifdef __CUDACC__

// Up to 32 sequential inner products at once: lane k returns sum_{t < cnt} S[a + t] * S[b + t], accumulated
// from 0.0 in ascending t (norm / innerproduct, Auxilary.c:430-462).  a, b: shared-window byte addresses.
__device__ __forceinline__ double lane_inner_product(const unsigned a, const unsigned b, const int cnt, const int most)
{
	double acc = 0.0;
	for (int t0 = 0; t0 < most; t0 += 4) {
		double va[4], vb[4];
#pragma unroll
		for (int u = 0; u < 4; ++u) {
			const int t = min(t0 + u, cnt - 1);
			va[u] = cnt > 0 ? lds64(a + 8u * (unsigned)t) : 0.0;
			vb[u] = cnt > 0 ? lds64(b + 8u * (unsigned)t) : 0.0;
		}
#pragma unroll
		for (int u = 0; u < 4; ++u)
			if (t0 + u < cnt) acc += va[u] * vb[u];
	}
	return acc;
}

// ... and plain sequential sums sum_{t < cnt} S[a + t] from 0.0.
__device__ __forceinline__ double lane_sum(const unsigned a, const int cnt, const int most)
{
	double acc = 0.0;
	for (int t0 = 0; t0 < most; t0 += 4) {
		double va[4];
#pragma unroll
		for (int u = 0; u < 4; ++u) va[u] = cnt > 0 ? lds64(a + 8u * (unsigned)min(t0 + u, cnt - 1)) : 0.0;
#pragma unroll
		for (int u = 0; u < 4; ++u)
			if (t0 + u < cnt) acc += va[u];
	}
	return acc;
}

struct IterSums
{
	Scalars sc;
	double mu; // lambda'lambda / m  (qpSWIFT.c:538)
};

// After the residual products (r, and tmp = -(P x)) are in place: lambda = sqrt(s.z) and the seven sums of
// the top of an iteration, one per lane.  (lambda is needed only if the QP goes on; computing it here costs
// three square-root rounds on the last iteration and saves a pass over s, z on all the others.)
__device__ __forceinline__ IterSums ipm_sums(const int n, const int m, const int p, const unsigned sb, const SliceLayout &L, const int tmp)
{
	const int lane = lane_id();
	const unsigned B = 8u;
	for (int i = lane; i < m; i += 32) sts64(sb + B * (unsigned)(L.lam + i), sqrt(lds64(sb + B * (unsigned)(L.s + i)) * lds64(sb + B * (unsigned)(L.z + i))));
	__syncwarp();
	const int np = n + p;
	unsigned a = 0, b = 0;
	int cnt = 0;
	switch (lane) {
	case 0: a = b = L.r; cnt = n; break;                 // n_rx^2
	case 1: a = b = L.r + n; cnt = p; break;             // n_ry^2
	case 2: a = b = L.r + np; cnt = m; break;            // n_rz^2
	case 3: a = L.s; b = L.z; cnt = m; break;            // s'z
	case 4: a = tmp; b = L.x; cnt = n; break;            // -(P x)'x
	case 5: a = L.c; b = L.x; cnt = n; break;            // c'x
	case 6: a = b = L.lam; cnt = m; break;               // lambda'lambda
	default: break;
	}
	const int most = max(max(n, m), p);
	double v = lane_inner_product(sb + B * a, sb + B * b, cnt, most);
	if (lane < 3) v = sqrt(v);
	else if (lane == 3 || lane == 6) v = v / m;
	IterSums out;
	out.sc.n_rx = __shfl_sync(0xffffffffu, v, 0);
	out.sc.n_ry = __shfl_sync(0xffffffffu, v, 1);
	out.sc.n_rz = __shfl_sync(0xffffffffu, v, 2);
	out.sc.n_mu = __shfl_sync(0xffffffffu, v, 3);
	const double q1 = __shfl_sync(0xffffffffu, v, 4), q2 = __shfl_sync(0xffffffffu, v, 5);
	out.sc.fval = -0.5 * q1 + q2;
	out.mu = __shfl_sync(0xffffffffu, v, 6);
	return out;
}

// ds (predictor: -lambda^2; centering: -lambda^2 + sigma mu), the (z,z) block -s/z, and the right-hand side
// [rx; ry; rz - ds/z] in one pass (qpSWIFT.c:540-552, 572-579; Auxilary.c:211-215, 274-295).
__device__ __forceinline__ void ipm_begin_step(const int n, const int m, const int p, WarpState &w, const bool centering, const double sm)
{
	const int lane = lane_id(), np = n + p;
	for (int i = lane; i < m; i += 32) {
		const double lam = w.lam[i], s = w.s[i], z = w.z[i];
		const double ds = centering ? -(lam * lam) + sm : -lam * lam;
		w.ds[i] = ds;
		w.zd[i] = -s / z;
		w.rhs[np + i] = w.r[np + i] - (ds / z);
	}
	for (int i = lane; i < np; i += 32) w.rhs[i] = w.r[i];
	__syncwarp();
}

// delta_z from the solution, delta_s = (ds - s.delta_z)/z, and the step lengths of findsteplength
// (Auxilary.c:508-511, 556-562, 359-393) in one pass.
__device__ __forceinline__ void ipm_recover_and_steplength(const int n, const int m, const int p, WarpState &w, double &alpha_p, double &alpha_d)
{
	const int lane = lane_id(), np = n + p;
	double ap = 1e10, ad = 1e10;
	for (int i = lane; i < m; i += 32) {
		const double dzi = w.rhs[np + i], s = w.s[i], z = w.z[i];
		const double dsi = (w.ds[i] - (s * dzi)) / z;
		w.dz[i] = dzi;
		w.dsl[i] = dsi;
		if (dsi < 0 && (-s / dsi) < ap) ap = -(s / dsi);
		if (dzi < 0 && (-z / dzi) < ad) ad = -(z / dzi);
	}
	for (int o = 16; o > 0; o >>= 1) {
		ap = fmin(ap, __shfl_xor_sync(0xffffffffu, ap, o));
		ad = fmin(ad, __shfl_xor_sync(0xffffffffu, ad, o));
	}
	alpha_p = ap < 1e10 ? ap : 1.0;
	alpha_d = ad < 1e10 ? ad : 1.0;
	__syncwarp();
}

// rho (Auxilary.c:879-892), sigma = max(sigma_d, min(1,rho)^3) (qpSWIFT.c:566-567), then the corrector's
// ds = -lambda^2 - delta_s.delta_z + sigma mu and right-hand side.  `scratch`: 2m doubles of the slice.
__device__ __forceinline__ double ipm_sigma_and_corrector(const int n, const int m, const int p, WarpState &w, const unsigned sb, const int scratch, const double ap, const double ad, const double sigma_d, const double mu)
{
	const int lane = lane_id(), np = n + p;
	double *prod = w.S + scratch;
	for (int i = lane; i < m; i += 32) {
		const double s = w.s[i], z = w.z[i];
		prod[i] = (s + (ap * w.dsl[i])) * (z + (ad * w.dz[i]));
		prod[m + i] = s * z;
	}
	__syncwarp();
	const double a = lane_sum(sb + 8u * (unsigned)(scratch + (lane == 1 ? m : 0)), lane < 2 ? m : 0, m);
	const double num = __shfl_sync(0xffffffffu, a, 0), den = __shfl_sync(0xffffffffu, a, 1);
	const double rho = num / den;
	const double mr = (1 < rho) ? 1 : rho;
	const double cube = mr * mr * mr;
	const double sigma = (sigma_d > cube) ? sigma_d : cube;
	const double sm = sigma * mu;
	for (int i = lane; i < m; i += 32) {
		const double lam = w.lam[i];
		const double ds = -(lam * lam) - (w.dsl[i] * w.dz[i]) + sm;
		w.ds[i] = ds;
		w.rhs[np + i] = w.r[np + i] - (ds / w.z[i]);
	}
	for (int i = lane; i < np; i += 32) w.rhs[i] = w.r[i];
	__syncwarp();
	return sigma;
}

// damping and update (qpSWIFT.c:584-596, Auxilary.c:246-254)
__device__ __forceinline__ void ipm_update(const int n, const int m, const int p, WarpState &w, double &ap, double &ad)
{
	const int lane = lane_id();
	ap = ((0.99 * ap) < (1.0)) ? (0.99 * ap) : (1.0);
	ad = ((0.99 * ad) < (1.0)) ? (0.99 * ad) : (1.0);
	for (int i = lane; i < n; i += 32) w.x[i] += w.rhs[i] * ap;
	for (int i = lane; i < p; i += 32) w.y[i] += w.rhs[n + i] * ad;
	for (int i = lane; i < m; i += 32) { w.s[i] += w.dsl[i] * ap; w.z[i] += w.dz[i] * ad; }
	__syncwarp();
}

// The residual products of computeresiduals for any pattern, from the plan's row/column lists (the generated
// programs bring their own, table-driven): r = [rx; ry; rz] and tmp = -(P x).
__device__ __forceinline__ void ipm_residual_products(const PlanView &pv, const double *V, const double *c, const double *h, const double *b, const double *x, const double *y, const double *z, const double *s, double *r, double *tmp)
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
}

#endif // __CUDACC__

} // namespace qpb

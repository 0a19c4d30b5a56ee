// SURVEY 8 f2 + the second half of f1: the DogBot whole-body-control QP built ON THE DEVICE from the robot's
// dynamic quantities, and the joint torques computed from its solution.
//
// Reference: dogbot_controller/src/client/main.cpp -- stance 1471-1647 (QP) and 1667 (torques), trot 1734-1992 and
// 2023, crawl 2943-3230 and 3246.  The controller fills dense Q, A, D, c, C, b from iDynTree's mass matrix,
// Jacobians and bias forces and hands them to QP_SETUP_dense; here every robot of a batch brings ONE packed state
// row (468 doubles = 3.7 KB instead of the 8.4-9.9 KB of a finished value row)
//
//     Jc[12][6] | Jj[12][12] | Mcom[6][6] | Mjj[12][12] | bias[18] | Jdqd[12] | W[6] | q[12] | dq[12] | vdot_sw[12]
//
// (foot-row order br, bl, fl, fr as in JacCOM_lin; W = the desired CoM wrench Kcom dx + Dcom dv + m g, main.cpp:1572;
// vdot_sw = the desired swing-foot accelerations, main.cpp:1991) and the kernel writes the plan-ordered value rows
// Ppr | Apr | Gpr | c | h | b that QP_SOLVE_BATCH_DEVICE takes.  Which formula lands in which entry of a value row is
// a table of one 8-byte descriptor per entry, built on the host from the pattern's own row lists (so the output order
// is the plan's by construction).  Every number is produced by a fixed sequence of IEEE-754 operations (sums from 0.0
// in ascending index order, no FMA), restated in numpy by apf_quadruped_b200/wbc.py: bit-identical by construction.
//
// Kernel shape: HBM-bound (3.7 KB in, 8.4-9.9 KB out per robot, a handful of flops per entry).  One warp per robot at
// a time: the state row is pulled into shared memory with coalesced 16-byte loads, then the lanes walk the value row
// front to back, one entry each, so that every store of the warp is one contiguous 256-byte segment.
#pragma once

#include <cstdint>
#include <vector>

namespace qpb {

enum { WBC_STANCE = 0, WBC_TROT = 1, WBC_CRAWL = 2 };

// state row
enum {
	WS_JC = 0, WS_JJ = 72, WS_MCOM = 216, WS_MJJ = 252, WS_BIAS = 396, WS_JDQD = 414, WS_W = 426, WS_Q = 432, WS_DQ = 444, WS_VSW = 456,
	WBC_STATE = 468,
	// constants appended in shared memory
	WC_ZERO = 468, WC_ONE = 469, WC_MU = 470, WC_TAU = 471, WC_SLACK = 472, WC_FIFTY = 473, WC_K = 474, WC_DT = 475, WC_QMIN = 476, WC_QMAX = 488,
	WBC_SMEM = 500
};
enum { WOP_COPY = 0, WOP_QDOT = 1, WOP_CDOT = 2, WOP_SUB = 3, WOP_ADD = 4, WOP_DDQ = 5 };

struct WbcDesc
{
	uint32_t op_a; // op | flag << 4 | a << 8   (flag: negate the result; QDOT: the entry is on the diagonal)
	uint32_t b;
};

struct WbcTables
{
	int kind = -1;
	int nf = 0;            // force variables (3 per stance foot)
	int stance_rows[12];   // rows of Jc / Jj that belong to the stance feet, in variable order
	std::vector<WbcDesc> desc; // [nnzP + nnzA + nnzG + n + m + p]
	double consts[WBC_SMEM - WBC_STATE];
};

// the device's form: op [0,3) | flag [3] | a [4,14) | b [14,24)  (every index is below WBC_SMEM = 500 < 1024)
inline uint32_t wbc_pack(const WbcDesc &d) { return (d.op_a & 15u) | (((d.op_a >> 4) & 1u) << 3) | ((d.op_a >> 8) << 4) | (d.b << 14); }

inline WbcDesc wbc_desc(int op, int flag, int a, int b = 0) { return WbcDesc{(uint32_t)op | ((uint32_t)flag << 4) | ((uint32_t)a << 8), (uint32_t)b}; }

// The dense entry (r, c) of Q / A / G and the entries of c / h / b as formulas of the state row.  Returns false for an
// entry that is structurally zero in the controller's matrices (the pattern must not contain it).
struct WbcFormulas
{
	int kind, ns, nw, nf, nslack, p, m;
	int st[12], sw[12]; // Jacobian rows of the stance / swing feet
	explicit WbcFormulas(int kind_) : kind(kind_)
	{
		// (main.cpp: stance = all four feet; trot: stance legs bl, fr and swing legs br, fl [1734-1741]; crawl: one swing leg)
		const int stance_feet[3][4] = {{0, 1, 2, 3}, {1, 3, -1, -1}, {1, 2, 3, -1}};
		ns = kind == WBC_STANCE ? 4 : kind == WBC_TROT ? 2 : 3;
		nw = 4 - ns;
		nf = 3 * ns;
		nslack = 3 * nw;
		p = 6 + nf;
		m = 5 * ns + 24 + 2 * nslack + 24;
		int k = 0, l = 0;
		for (int f = 0; f < 4; ++f) {
			bool is_st = false;
			for (int i = 0; i < ns; ++i) is_st |= stance_feet[kind][i] == f;
			for (int d = 0; d < 3; ++d) (is_st ? st[k++] : sw[l++]) = 3 * f + d;
		}
	}
	bool Q(int r, int c, WbcDesc &d) const
	{
		const bool force = r >= 18 && r < 18 + nf && c >= 18 && c < 18 + nf;
		if (force) { d = wbc_desc(WOP_QDOT, r == c, WS_JC + 6 * st[r - 18], WS_JC + 6 * st[c - 18]); return true; } // I + 50 Ts'Ts (1476-1481)
		if (r != c) return false;
		d = wbc_desc(WOP_COPY, 0, r >= 18 + nf ? WC_SLACK : WC_ONE); // slack weight 1e8 (1751) / 1e4 (2976)
		return true;
	}
	bool A(int r, int c, WbcDesc &d) const // [Mcom 0 -Jst_c'; Jst_c Jst_j 0] (1579-1583)
	{
		if (r < 6 && c < 6) { d = wbc_desc(WOP_COPY, 0, WS_MCOM + 6 * r + c); return true; }
		if (r < 6 && c >= 18 && c < 18 + nf) { d = wbc_desc(WOP_COPY, 1, WS_JC + 6 * st[c - 18] + r); return true; }
		if (r >= 6 && c < 6) { d = wbc_desc(WOP_COPY, 0, WS_JC + 6 * st[r - 6] + c); return true; }
		if (r >= 6 && c >= 6 && c < 18) { d = wbc_desc(WOP_COPY, 0, WS_JJ + 12 * st[r - 6] + (c - 6)); return true; }
		return false;
	}
	bool G(int r, int c, WbcDesc &d) const
	{
		int r0 = 0;
		if (r < 5 * ns) { // friction pyramids (1612-1632): rows (t1 - mu n), (t2 - mu n), -(t1 + mu n), -(t2 + mu n), -n
			const int foot = r / 5, row = r % 5, col = c - 18 - 3 * foot;
			if (col < 0 || col > 2) return false;
			if (col == 2) { d = row == 4 ? wbc_desc(WOP_COPY, 1, WC_ONE) : wbc_desc(WOP_COPY, 1, WC_MU); return true; }
			if ((row == 0 && col == 0) || (row == 1 && col == 1)) { d = wbc_desc(WOP_COPY, 0, WC_ONE); return true; }
			if ((row == 2 && col == 0) || (row == 3 && col == 1)) { d = wbc_desc(WOP_COPY, 1, WC_ONE); return true; }
			return false;
		}
		r0 = 5 * ns;
		if (r < r0 + 24) { // torque limits: +-[0 Mjj -Jst_j'] (1594-1597)
			const int j = (r - r0) % 12, neg = (r - r0) / 12;
			if (c >= 6 && c < 18) { d = wbc_desc(WOP_COPY, neg, WS_MJJ + 12 * j + (c - 6)); return true; }
			if (c >= 18 && c < 18 + nf) { d = wbc_desc(WOP_COPY, !neg, WS_JJ + 12 * st[c - 18] + j); return true; }
			return false;
		}
		r0 += 24;
		if (r < r0 + 2 * nslack) { // swing-foot tracking with slack: +-Jsw [a; qdd] - gamma <= +-(vdot_des - Jdqd_sw) (1864-1875)
			const int k = (r - r0) % nslack, neg = (r - r0) / nslack;
			if (c < 6) { d = wbc_desc(WOP_COPY, neg, WS_JC + 6 * sw[k] + c); return true; }
			if (c < 18) { d = wbc_desc(WOP_COPY, neg, WS_JJ + 12 * sw[k] + (c - 6)); return true; }
			if (c == 18 + nf + k) { d = wbc_desc(WOP_COPY, 1, WC_ONE); return true; }
			return false;
		}
		r0 += 2 * nslack;
		const int j = (r - r0) % 12, neg = (r - r0) / 12; // joint acceleration limits +-I (1598-1599)
		if (c == 6 + j) { d = wbc_desc(WOP_COPY, neg, WC_ONE); return true; }
		return false;
	}
	WbcDesc c(int i) const // c = -Ts' 50 W (1573)
	{
		if (i >= 18 && i < 18 + nf) return wbc_desc(WOP_CDOT, 0, WS_JC + 6 * st[i - 18], WS_W);
		return wbc_desc(WOP_COPY, 0, WC_ZERO);
	}
	WbcDesc b(int i) const // [-bias_com; -Jdqd_st] (1586-1588); the trot branch never fills eigenb (1851)
	{
		if (kind == WBC_TROT) return wbc_desc(WOP_COPY, 0, WC_ZERO);
		return i < 6 ? wbc_desc(WOP_COPY, 1, WS_BIAS + i) : wbc_desc(WOP_COPY, 1, WS_JDQD + st[i - 6]);
	}
	WbcDesc h(int r) const
	{
		int r0 = 5 * ns;
		if (r < r0) return wbc_desc(WOP_COPY, 0, WC_ZERO);
		if (r < r0 + 12) return wbc_desc(WOP_SUB, 0, WC_TAU, WS_BIAS + 6 + (r - r0));      // tau_max - bias (1634)
		if (r < r0 + 24) return wbc_desc(WOP_ADD, 0, WC_TAU, WS_BIAS + 6 + (r - r0 - 12)); // -(tau_min - bias) (1635)
		r0 += 24;
		if (r < r0 + 2 * nslack) { const int k = (r - r0) % nslack; return wbc_desc(WOP_SUB, (r - r0) / nslack, WS_VSW + sw[k], WS_JDQD + sw[k]); } // (1991-1992)
		r0 += 2 * nslack;
		if (r < r0 + 12) return wbc_desc(WOP_DDQ, 0, WC_QMAX + (r - r0), r - r0);         // ddqmax (1642-1645)
		return wbc_desc(WOP_DDQ, 1, WC_QMIN + (r - r0 - 12), r - r0 - 12);                // -ddqmin (1646)
	}
};

// The descriptor table of `kind` in the order of the plan's value rows.  False: the pattern is not that of `kind`.
inline bool wbc_build_tables(WbcTables &T, int kind, const Plan &pl)
{
	const WbcFormulas F(kind);
	if (pl.n != 30 || pl.p != F.p || pl.m != F.m) return false;
	const DevicePlanHeader *hd = reinterpret_cast<const DevicePlanHeader *>(pl.blob.data());
	const int32_t *blob = pl.blob.data();
	T.kind = kind;
	T.nf = F.nf;
	for (int i = 0; i < 12; ++i) T.stance_rows[i] = i < F.nf ? F.st[i] : 0;
	const long nV = pl.nnzP + pl.nnzA + pl.nnzG;
	T.desc.assign((size_t)(nV + pl.n + pl.m + pl.p), wbc_desc(WOP_COPY, 0, WC_ZERO));
	std::vector<char> seen((size_t)nV, 0);
	struct View { long rows; int32_t ptr, col, val; int which; };
	const View views[3] = {{pl.n, hd->prow_ptr, hd->prow_col, hd->prow_val, 0}, {pl.p, hd->arow_ptr, hd->arow_col, hd->arow_val, 1}, {pl.m, hd->grow_ptr, hd->grow_col, hd->grow_val, 2}};
	for (const View &v : views)
		for (long r = 0; r < v.rows; ++r)
			for (int32_t t = blob[v.ptr + r]; t < blob[v.ptr + r + 1]; ++t) {
				const int c = blob[v.col + t], at = blob[v.val + t];
				WbcDesc d;
				const bool ok = v.which == 0 ? F.Q((int)r, c, d) : v.which == 1 ? F.A((int)r, c, d) : F.G((int)r, c, d);
				if (!ok || at < 0 || at >= nV || seen[(size_t)at]) return false;
				seen[(size_t)at] = 1;
				T.desc[(size_t)at] = d;
			}
	for (long i = 0; i < pl.n; ++i) T.desc[(size_t)(nV + i)] = F.c((int)i);
	for (long i = 0; i < pl.m; ++i) T.desc[(size_t)(nV + pl.n + i)] = F.h((int)i);
	for (long i = 0; i < pl.p; ++i) T.desc[(size_t)(nV + pl.n + pl.m + i)] = F.b((int)i);
	static const double qmin[12] = {-1.75, -1.75, -1.75, -1.75, -1.58, -2.62, -3.15, -0.02, -1.58, -2.62, -3.15, -0.02}; // main.cpp:493
	static const double qmax[12] = {1.75, 1.75, 1.75, 1.75, 3.15, 0.02, 1.58, 2.62, 3.15, 0.02, 1.58, 2.62};             // main.cpp:494
	const double dt = 0.025; // main.cpp:1638
	double *k = T.consts;
	k[WC_ZERO - WBC_STATE] = 0.0; k[WC_ONE - WBC_STATE] = 1.0; k[WC_MU - WBC_STATE] = 0.5; k[WC_TAU - WBC_STATE] = 60.0;
	k[WC_SLACK - WBC_STATE] = kind == WBC_TROT ? 1e8 : 1e4; k[WC_FIFTY - WBC_STATE] = 50.0; k[WC_K - WBC_STATE] = 2.0 / (dt * dt); k[WC_DT - WBC_STATE] = dt;
	for (int j = 0; j < 12; ++j) { k[WC_QMIN - WBC_STATE + j] = qmin[j]; k[WC_QMAX - WBC_STATE + j] = qmax[j]; }
	return true;
}

#ifdef __CUDACC__

struct WbcOut
{
	double *Ppr, *Apr, *Gpr, *c, *h, *b;
	int nnzP, nnzA, nnzG, n, m, p;
};

// The rare formulas (the 6-term dot products of the cost, the limits): out of the copy path's way.
__device__ __forceinline__ double wbc_eval_formula(const unsigned op, const unsigned flag, const unsigned a, const unsigned b, const double *S)
{
	double v;
	switch (op) {
	case WOP_QDOT: { // (diagonal ? 1 : 0) + 50 (Jc_i . Jc_j), the sum from 0.0 in ascending k
		double acc = 0.0;
#pragma unroll
		for (int k = 0; k < 6; ++k) acc = acc + S[a + k] * S[b + k];
		return (flag ? 1.0 : 0.0) + S[WC_FIFTY] * acc;
	}
	case WOP_CDOT: {
		double acc = 0.0;
#pragma unroll
		for (int k = 0; k < 6; ++k) acc = acc + S[a + k] * S[b + k];
		return -S[WC_FIFTY] * acc;
	}
	case WOP_SUB: v = S[a] - S[b]; return flag ? -v : v;
	case WOP_ADD: return S[a] + S[b];
	default: v = S[WC_K] * ((S[a] - S[WS_Q + b]) - S[WC_DT] * S[WS_DQ + b]); return flag ? -v : v; // WOP_DDQ
	}
}

// Nine entries in ten are (signed) copies of a state entry or a constant: one shared-memory load and a sign flip.
__device__ __forceinline__ double wbc_eval(const unsigned d, const double *S)
{
	const unsigned op = d & 7u, flag = (d >> 3) & 1u, a = (d >> 4) & 1023u;
	if (op != WOP_COPY) return wbc_eval_formula(op, flag, a, (d >> 14) & 1023u, S);
	const double v = S[a];
	return flag ? -v : v;
}

// The state row of robot r on its way into a warp's buffer: 16 bytes per lane and request, asynchronously
// (cp.async: global -> shared memory without a stop in registers).
__device__ __forceinline__ void wbc_fetch(double *buf, const double *__restrict__ state, const int r, const int lane)
{
	const double *src = state + (size_t)r * WBC_STATE;
	const unsigned dst = (unsigned)__cvta_generic_to_shared(buf);
	for (int i = lane; i < WBC_STATE / 2; i += 32)
		asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * (unsigned)i), "l"(src + 2 * i) : "memory");
	asm volatile("cp.async.commit_group;" ::: "memory");
}

__device__ __forceinline__ void wbc_emit(double *__restrict__ row, const unsigned *sdesc, const int len, const double *S, const int lane)
{
	for (int e0 = lane; e0 < len; e0 += 128) { // four entries per lane in flight
		double v[4];
#pragma unroll
		for (int j = 0; j < 4; ++j)
			if (e0 + 32 * j < len) v[j] = wbc_eval(sdesc[e0 + 32 * j], S);
#pragma unroll
		for (int j = 0; j < 4; ++j)
			if (e0 + 32 * j < len) __stcs(row + e0 + 32 * j, v[j]);
	}
}

// One warp per robot at a time, grid-stride over the robots (or over the list `index[0 .. *count_ptr)`).  Rows are
// written at the robot's own index.  The formula table sits in shared memory (one copy per CTA); every warp has TWO
// state buffers: while it turns one robot's state into its value rows, the next robot's state row is already arriving
// in the other (cp.async), so the HBM round trip of the load is off the critical path.
__global__ void __launch_bounds__(256, 3) wbc_formulate_kernel(const double *__restrict__ state, const unsigned *__restrict__ desc, const double *__restrict__ consts, const WbcOut out, const int *__restrict__ index, const int *__restrict__ count_ptr, const int count)
{
	extern __shared__ __align__(16) double wsm[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
	const int nV = out.nnzP + out.nnzA + out.nnzG, total = nV + out.n + out.m + out.p;
	unsigned *sdesc = reinterpret_cast<unsigned *>(wsm + 2 * warps * WBC_SMEM);
	for (int i = threadIdx.x; i < total; i += blockDim.x) sdesc[i] = __ldg(desc + i);
	double *buf[2] = {wsm + (2 * warp) * WBC_SMEM, wsm + (2 * warp + 1) * WBC_SMEM};
	for (int i = lane; i < WBC_SMEM - WBC_STATE; i += 32) { buf[0][WBC_STATE + i] = consts[i]; buf[1][WBC_STATE + i] = consts[i]; }
	__syncthreads();
	const int limit = count_ptr ? min(*count_ptr, count) : count;
	const int stride = gridDim.x * warps;
	int t = blockIdx.x * warps + warp, cur = 0;
	if (t < limit) wbc_fetch(buf[0], state, index ? index[t] : t, lane);
	for (; t < limit; t += stride, cur ^= 1) {
		const int r = index ? index[t] : t;
		if (t + stride < limit) {
			wbc_fetch(buf[cur ^ 1], state, index ? index[t + stride] : t + stride, lane);
			asm volatile("cp.async.wait_group 1;" ::: "memory"); // the current robot's row has landed; the next one may still be in flight
		} else asm volatile("cp.async.wait_group 0;" ::: "memory");
		__syncwarp();
		const double *S = buf[cur];
		wbc_emit(out.Ppr + (size_t)r * out.nnzP, sdesc, out.nnzP, S, lane);
		wbc_emit(out.Apr + (size_t)r * out.nnzA, sdesc + out.nnzP, out.nnzA, S, lane);
		wbc_emit(out.Gpr + (size_t)r * out.nnzG, sdesc + out.nnzP + out.nnzA, out.nnzG, S, lane);
		wbc_emit(out.c + (size_t)r * out.n, sdesc + nV, out.n, S, lane);
		wbc_emit(out.h + (size_t)r * out.m, sdesc + nV + out.n, out.m, S, lane);
		wbc_emit(out.b + (size_t)r * out.p, sdesc + nV + out.n + out.m, out.p, S, lane);
		__syncwarp(); // (every lane is done with this buffer before the fetch after next overwrites it)
	}
}

// tau = Mjj qdd + bias_j - Jst_j' f (main.cpp:1667 / 2023 / 3246): x = [a_com(6); qdd(12); f(nf); slack], each sum from
// 0.0 in ascending index order, then (sum1 + bias) - sum2.  One thread per (robot, joint).
__global__ void __launch_bounds__(256) wbc_torque_kernel(const double *__restrict__ state, const double *__restrict__ x, double *__restrict__ tau, const int nf, const int4 st0, const int4 st1, const int4 st2, const int *__restrict__ index, const int *__restrict__ count_ptr, const int count)
{
	const int limit = count_ptr ? min(*count_ptr, count) : count;
	const int st[12] = {st0.x, st0.y, st0.z, st0.w, st1.x, st1.y, st1.z, st1.w, st2.x, st2.y, st2.z, st2.w};
	for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < limit * 12; g += gridDim.x * blockDim.x) {
		const int t = g / 12, j = g - 12 * t;
		const int r = index ? index[t] : t;
		const double *S = state + (size_t)r * WBC_STATE, *xr = x + (size_t)r * 30;
		double a = 0.0, b = 0.0;
		for (int k = 0; k < 12; ++k) a = a + S[WS_MJJ + 12 * j + k] * xr[6 + k];
		for (int k = 0; k < nf; ++k) b = b + S[WS_JJ + 12 * st[k] + j] * xr[18 + k];
		tau[(size_t)r * 12 + j] = (a + S[WS_BIAS + 6 + j]) - b;
	}
}

#endif // __CUDACC__

} // namespace qpb

// Kernels and host-side launch code of the batched solver; the C ABI of include/qpswift_b200.h.
//
// Launch shape: one warp per CTA, one QP per warp at a time, persistent CTAs that pull QP indices
// from an atomic work counter (iteration counts differ per QP, so static striding would leave
// SMs idle behind a slow instance).  Each warp keeps its QP's whole working set -- value row,
// factor, iterate, residuals -- in its own slice of shared memory, so the interior-point loop
// never touches HBM between the coalesced load of the value row and the store of the solution.

#include "capi_internal.h"
#include "kernels.cuh"

#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>

namespace qpb {

static thread_local std::string g_last_error;

static int fail(int code, const char *what, cudaError_t e)
{
	char buf[256];
	snprintf(buf, sizeof buf, "%s: %s", what, e == cudaSuccess ? "" : cudaGetErrorString(e));
	g_last_error = buf;
	return code;
}

#define CUDA_TRY(expr)                                                        \
	do {                                                                      \
		cudaError_t e__ = (expr);                                             \
		if (e__ != cudaSuccess) return fail(QPB_E_CUDA, #expr, e__);          \
	} while (0)

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(32) qp_solve_fused_kernel(const int *__restrict__ plan, SmemLayout lay, BatchIO io, SolveParams prm, int *__restrict__ counter)
{
	extern __shared__ __align__(16) double smem[];
	const PlanView pv = view_of(plan);
	WarpState w = carve(smem, lay);
	const int lane = lane_id();
	for (;;) {
		int q = 0;
		if (lane == 0) q = atomicAdd(counter, 1);
		q = __shfl_sync(0xffffffffu, q, 0);
		if (q >= io.count) break;
		const size_t qi = (size_t)q;
		warp_copy_in(w.V, io.Ppr + qi * pv.offA, pv.offA);
		if (pv.p) warp_copy_in(w.V + pv.offA, io.Apr + qi * (pv.offG - pv.offA), pv.offG - pv.offA);
		warp_copy_in(w.V + pv.offG, io.Gpr + qi * (pv.nV - pv.offG), pv.nV - pv.offG);
		warp_copy_in(w.c, io.c + qi * pv.n, pv.n);
		warp_copy_in(w.h, io.h + qi * pv.m, pv.m);
		if (pv.p) warp_copy_in(w.b, io.b + qi * pv.p, pv.p);
		__syncwarp();

		ipm_initial_point(pv, w);
		const SolveResult res = ipm_solve(pv, w, prm);

		warp_copy_out(io.x ? io.x + qi * pv.n : nullptr, w.x, pv.n);
		if (pv.p) warp_copy_out(io.y ? io.y + qi * pv.p : nullptr, w.y, pv.p);
		warp_copy_out(io.z ? io.z + qi * pv.m : nullptr, w.z, pv.m);
		warp_copy_out(io.s ? io.s + qi * pv.m : nullptr, w.s, pv.m);
		if (lane == 0) {
			if (io.flag) io.flag[q] = res.flag;
			if (io.iters) io.iters[q] = res.iters;
			if (io.fval) io.fval[q] = res.sc.fval;
			if (io.stat) {
				double *st = io.stat + qi * 8;
				st[0] = res.sc.n_rx; st[1] = res.sc.n_ry; st[2] = res.sc.n_rz; st[3] = res.sc.n_mu;
				st[4] = res.alpha_p; st[5] = res.alpha_d; st[6] = res.sigma; st[7] = res.mu;
			}
		}
		__syncwarp();
	}
}

// LDL_numeric alone: value row + (z,z) diagonal from HBM, factor to HBM.
__global__ void __launch_bounds__(32) ldl_numeric_kernel(const int *__restrict__ plan, const double *__restrict__ vals, const double *__restrict__ zdiag, double *__restrict__ Lx, double *__restrict__ D, int count, int *__restrict__ counter)
{
	extern __shared__ __align__(16) double smem[];
	const PlanView pv = view_of(plan);
	const int lane = lane_id();
	// layout: V[nV+pad] | zd[m] (as -zd/1 trick is NOT used: we store s = -zdiag, z = 1) | Lx | D | Y
	const int nVp = (pv.nV + 1) & ~1;
	double *V = smem, *sv = V + nVp, *zv = sv + pv.m, *L = zv + pv.m + ((2 * pv.m) & 1), *Dd = L + pv.lnz + (pv.lnz & 1), *Y = Dd + pv.N;
	for (;;) {
		int q = 0;
		if (lane == 0) q = atomicAdd(counter, 1);
		q = __shfl_sync(0xffffffffu, q, 0);
		if (q >= count) break;
		const size_t qi = (size_t)q;
		warp_copy_in(V, vals + qi * pv.nV, pv.nV);
		// -s/z with z = 1 and s = -d reproduces d exactly: -(-d)/1 == d
		for (int i = lane; i < pv.m; i += 32) { sv[i] = -__ldg(zdiag + qi * pv.m + i); zv[i] = 1.0; }
		__syncwarp();
		ldl_factor(pv, V, sv, zv, 0.0, L, Dd, Y);
		warp_copy_out(Lx + qi * pv.lnz, L, pv.lnz);
		warp_copy_out(D + qi * pv.N, Dd, pv.N);
		__syncwarp();
	}
}

// LDL_perm + lsolve + dsolve + ltsolve + permt alone.
__global__ void __launch_bounds__(32) ldl_solve_kernel(const int *__restrict__ plan, const double *__restrict__ Lx, const double *__restrict__ D, const double *__restrict__ rhs, double *__restrict__ sol, int count, int *__restrict__ counter)
{
	extern __shared__ __align__(16) double smem[];
	const PlanView pv = view_of(plan);
	const int lane = lane_id();
	double *L = smem, *Dd = L + pv.lnz + (pv.lnz & 1), *r = Dd + pv.N + (pv.N & 1), *wk = r + pv.N + (pv.N & 1);
	for (;;) {
		int q = 0;
		if (lane == 0) q = atomicAdd(counter, 1);
		q = __shfl_sync(0xffffffffu, q, 0);
		if (q >= count) break;
		const size_t qi = (size_t)q;
		warp_copy_in(L, Lx + qi * pv.lnz, pv.lnz);
		warp_copy_in(Dd, D + qi * pv.N, pv.N);
		warp_copy_in(r, rhs + qi * pv.N, pv.N);
		__syncwarp();
		ldl_solve(pv, L, Dd, r, wk);
		warp_copy_out(sol + qi * pv.N, r, pv.N);
		__syncwarp();
	}
}

static SmemLayout make_layout(const Plan &pl)
{
	SmemLayout L;
	int at = 0;
	auto take = [&](long count) { int here = at; at += (int)((count + 1) & ~1L); return here; }; // keep 16-byte alignment
	const long nV = pl.nnzP + pl.nnzA + pl.nnzG, N = pl.N, lnz = pl.Lp[pl.N];
	L.V = take(nV); L.c = take(pl.n); L.h = take(pl.m); L.b = take(pl.p);
	L.Lx = take(lnz); L.D = take(N); L.Y = take(N); L.rhs = take(N);
	L.x = take(pl.n); L.y = take(pl.p); L.z = take(pl.m); L.s = take(pl.m);
	L.r = take(N);
	L.lam = take(pl.m > pl.n ? pl.m : pl.n); L.ds = take(pl.m); L.dz = take(pl.m); L.dsl = take(pl.m);
	L.total = at;
	return L;
}

} // namespace qpb

using namespace qpb;

struct QPBatch
{
	Plan plan;
	double sigma_d = 0.0;
	long capacity = 0;
	int device = 0;
	int sm_count = 0;
	int *d_plan = nullptr;
	int *d_counter = nullptr;
	SmemLayout layout;
	int ctas_per_sm = 1;
	// device staging for the host-pointer entry point
	double *d_in[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
	double *d_x = nullptr, *d_y = nullptr, *d_z = nullptr, *d_s = nullptr, *d_fval = nullptr, *d_stat = nullptr;
	int *d_flag = nullptr, *d_iters = nullptr;
	int *h_flag = nullptr, *h_iters = nullptr; // pinned
	cudaStream_t stream = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	double last_kernel_ms = 0.0;
	long last_launches = 0;
	std::mutex lock;
};

static int batch_alloc(QPBatch *B)
{
	const Plan &pl = B->plan;
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0) return fail(QPB_E_NODEVICE, "no CUDA device", e);
	if (B->device < 0) CUDA_TRY(cudaGetDevice(&B->device));
	CUDA_TRY(cudaSetDevice(B->device));
	cudaDeviceProp prop;
	CUDA_TRY(cudaGetDeviceProperties(&prop, B->device));
	B->sm_count = prop.multiProcessorCount;
	B->layout = make_layout(pl);
	const size_t smem = (size_t)B->layout.total * sizeof(double);
	if (smem > (size_t)prop.sharedMemPerBlockOptin) return fail(QPB_E_PATTERN, "pattern does not fit in shared memory", cudaSuccess);
	CUDA_TRY(cudaFuncSetAttribute(qp_solve_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	CUDA_TRY(cudaFuncSetAttribute(ldl_numeric_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin));
	CUDA_TRY(cudaFuncSetAttribute(ldl_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin));
	int occ = 0;
	CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, qp_solve_fused_kernel, 32, smem));
	B->ctas_per_sm = occ > 0 ? occ : 1;

	CUDA_TRY(cudaMalloc(&B->d_plan, pl.blob.size() * sizeof(int32_t)));
	CUDA_TRY(cudaMemcpy(B->d_plan, pl.blob.data(), pl.blob.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
	CUDA_TRY(cudaMalloc(&B->d_counter, 64 * sizeof(int)));
	CUDA_TRY(cudaStreamCreateWithFlags(&B->stream, cudaStreamNonBlocking));
	CUDA_TRY(cudaEventCreate(&B->ev0));
	CUDA_TRY(cudaEventCreate(&B->ev1));
	return 0;
}

static int batch_alloc_staging(QPBatch *B)
{
	if (B->d_x) return 0;
	const Plan &pl = B->plan;
	const size_t cap = (size_t)B->capacity;
	const size_t widths[6] = {(size_t)pl.nnzP, (size_t)pl.nnzA, (size_t)pl.nnzG, (size_t)pl.n, (size_t)pl.m, (size_t)pl.p};
	for (int i = 0; i < 6; ++i)
		if (widths[i]) CUDA_TRY(cudaMalloc(&B->d_in[i], cap * widths[i] * sizeof(double)));
	CUDA_TRY(cudaMalloc(&B->d_x, cap * pl.n * sizeof(double)));
	if (pl.p) CUDA_TRY(cudaMalloc(&B->d_y, cap * pl.p * sizeof(double)));
	CUDA_TRY(cudaMalloc(&B->d_z, cap * pl.m * sizeof(double)));
	CUDA_TRY(cudaMalloc(&B->d_s, cap * pl.m * sizeof(double)));
	CUDA_TRY(cudaMalloc(&B->d_fval, cap * sizeof(double)));
	CUDA_TRY(cudaMalloc(&B->d_stat, cap * 8 * sizeof(double)));
	CUDA_TRY(cudaMalloc(&B->d_flag, cap * sizeof(int)));
	CUDA_TRY(cudaMalloc(&B->d_iters, cap * sizeof(int)));
	CUDA_TRY(cudaMallocHost(&B->h_flag, cap * sizeof(int)));
	CUDA_TRY(cudaMallocHost(&B->h_iters, cap * sizeof(int)));
	return 0;
}

static int launch_fused(QPBatch *B, const BatchIO &io, const settings *opts, cudaStream_t st)
{
	SolveParams prm;
	prm.maxit = (int)opts->maxit;
	prm.reltol = opts->reltol;
	prm.abstol = opts->abstol;
	prm.sigma0 = opts->sigma;
	prm.sigma_d = B->sigma_d;
	CUDA_TRY(cudaMemsetAsync(B->d_counter, 0, sizeof(int), st));
	long grid = (long)B->sm_count * B->ctas_per_sm;
	if (grid > io.count) grid = io.count;
	if (grid < 1) grid = 1;
	const size_t smem = (size_t)B->layout.total * sizeof(double);
	qp_solve_fused_kernel<<<(unsigned)grid, 32, smem, st>>>(B->d_plan, B->layout, io, prm, B->d_counter);
	CUDA_TRY(cudaGetLastError());
	B->last_launches = 1;
	return 0;
}

extern "C" {

const char *QPB_LAST_ERROR(void) { return g_last_error.c_str(); }

int QPB_DEVICE_COUNT(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

QPBatch *QPB_BATCH_CREATE(const QPBPlan *plan, qp_real sigma_d, qp_int capacity, int device)
{
	if (!plan || capacity < 1) { g_last_error = "QPB_BATCH_CREATE: bad argument"; return nullptr; }
	if (!plan->plan.valid) { g_last_error = "QPB_BATCH_CREATE: unusable pattern (an inequality row is empty)"; return nullptr; }
	QPBatch *B = new QPBatch;
	B->plan = plan->plan;
	B->sigma_d = sigma_d;
	B->capacity = capacity;
	B->device = device;
	if (batch_alloc(B) != 0) { QP_CLEANUP_BATCH(B); return nullptr; }
	return B;
}

QPBatch *QP_SETUP_BATCH(const QP *proto, qp_int capacity, int device)
{
	if (!proto) { g_last_error = "QP_SETUP_BATCH: null prototype"; return nullptr; }
	QPBPlan tmp;
	const smat *A = proto->A;
	build_plan(tmp.plan, proto->n, proto->m, proto->p, proto->P->jc, proto->P->ir, A ? A->jc : nullptr, A ? A->ir : nullptr, proto->G->jc, proto->G->ir, proto->kkt ? proto->kkt->P : nullptr);
	return QPB_BATCH_CREATE(&tmp, proto->sigma_d, capacity, device);
}

void QP_CLEANUP_BATCH(QPBatch *B)
{
	if (!B) return;
	cudaSetDevice(B->device);
	cudaFree(B->d_plan);
	cudaFree(B->d_counter);
	for (int i = 0; i < 6; ++i) cudaFree(B->d_in[i]);
	cudaFree(B->d_x); cudaFree(B->d_y); cudaFree(B->d_z); cudaFree(B->d_s);
	cudaFree(B->d_fval); cudaFree(B->d_stat); cudaFree(B->d_flag); cudaFree(B->d_iters);
	if (B->h_flag) cudaFreeHost(B->h_flag);
	if (B->h_iters) cudaFreeHost(B->h_iters);
	if (B->stream) cudaStreamDestroy(B->stream);
	if (B->ev0) cudaEventDestroy(B->ev0);
	if (B->ev1) cudaEventDestroy(B->ev1);
	cudaGetLastError();
	delete B;
}

qp_int QP_SOLVE_BATCH_DEVICE(QPBatch *B, qp_int count, const settings *opts, const qp_real *dPpr, const qp_real *dApr, const qp_real *dGpr, const qp_real *dc, const qp_real *dh, const qp_real *db, qp_real *dx, qp_real *dy, qp_real *dz, qp_real *ds, int *dflag, int *diters, qp_real *dfval, qp_real *dstat, void *stream)
{
	if (!B || !opts || count < 0 || !dPpr || !dGpr || !dc || !dh || (B->plan.p && (!dApr || !db))) return fail(QPB_E_ARG, "QP_SOLVE_BATCH_DEVICE: bad argument", cudaSuccess);
	if (count == 0) return 0;
	CUDA_TRY(cudaSetDevice(B->device));
	BatchIO io;
	io.Ppr = dPpr; io.Apr = dApr; io.Gpr = dGpr; io.c = dc; io.h = dh; io.b = db;
	io.x = dx; io.y = dy; io.z = dz; io.s = ds; io.flag = dflag; io.iters = diters; io.fval = dfval; io.stat = dstat;
	io.count = (int)count;
	return launch_fused(B, io, opts, (cudaStream_t)stream);
}

qp_int QP_SOLVE_BATCH(QPBatch *B, qp_int count, const settings *opts, const qp_real *Ppr, const qp_real *Apr, const qp_real *Gpr, const qp_real *c, const qp_real *h, const qp_real *b, qp_real *x, qp_real *y, qp_real *z, qp_real *s, qp_int *flag, qp_int *iters)
{
	if (!B || !opts || count < 0 || count > B->capacity || !Ppr || !Gpr || !c || !h || (B->plan.p && (!Apr || !b))) return fail(QPB_E_ARG, "QP_SOLVE_BATCH: bad argument", cudaSuccess);
	if (count == 0) return 0;
	std::lock_guard<std::mutex> guard(B->lock);
	CUDA_TRY(cudaSetDevice(B->device));
	int rc = batch_alloc_staging(B);
	if (rc) return rc;
	const Plan &pl = B->plan;
	cudaStream_t st = B->stream;
	const size_t n = (size_t)count;
	const qp_real *src[6] = {Ppr, Apr, Gpr, c, h, b};
	const size_t widths[6] = {(size_t)pl.nnzP, (size_t)pl.nnzA, (size_t)pl.nnzG, (size_t)pl.n, (size_t)pl.m, (size_t)pl.p};
	for (int i = 0; i < 6; ++i)
		if (widths[i]) CUDA_TRY(cudaMemcpyAsync(B->d_in[i], src[i], n * widths[i] * sizeof(double), cudaMemcpyHostToDevice, st));
	BatchIO io;
	io.Ppr = B->d_in[0]; io.Apr = B->d_in[1]; io.Gpr = B->d_in[2]; io.c = B->d_in[3]; io.h = B->d_in[4]; io.b = B->d_in[5];
	io.x = B->d_x; io.y = B->d_y; io.z = B->d_z; io.s = B->d_s; io.flag = B->d_flag; io.iters = B->d_iters; io.fval = B->d_fval; io.stat = B->d_stat;
	io.count = (int)count;
	CUDA_TRY(cudaEventRecord(B->ev0, st));
	rc = launch_fused(B, io, opts, st);
	if (rc) return rc;
	CUDA_TRY(cudaEventRecord(B->ev1, st));
	if (x) CUDA_TRY(cudaMemcpyAsync(x, B->d_x, n * pl.n * sizeof(double), cudaMemcpyDeviceToHost, st));
	if (y && pl.p) CUDA_TRY(cudaMemcpyAsync(y, B->d_y, n * pl.p * sizeof(double), cudaMemcpyDeviceToHost, st));
	if (z) CUDA_TRY(cudaMemcpyAsync(z, B->d_z, n * pl.m * sizeof(double), cudaMemcpyDeviceToHost, st));
	if (s) CUDA_TRY(cudaMemcpyAsync(s, B->d_s, n * pl.m * sizeof(double), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaMemcpyAsync(B->h_flag, B->d_flag, n * sizeof(int), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaMemcpyAsync(B->h_iters, B->d_iters, n * sizeof(int), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	float ms = 0.f;
	CUDA_TRY(cudaEventElapsedTime(&ms, B->ev0, B->ev1));
	B->last_kernel_ms = ms;
	qp_int bad = 0;
	for (size_t i = 0; i < n; ++i) {
		if (flag) flag[i] = B->h_flag[i];
		if (iters) iters[i] = B->h_iters[i];
		if (B->h_flag[i] != QP_OPTIMAL) ++bad;
	}
	return bad;
}

qp_int QPB_BATCH_SET_SIGMA_D(QPBatch *B, qp_real sigma_d)
{
	if (!B) return fail(QPB_E_ARG, "QPB_BATCH_SET_SIGMA_D: null batch", cudaSuccess);
	B->sigma_d = sigma_d;
	return 0;
}

double QPB_BATCH_LAST_KERNEL_MS(const QPBatch *B) { return B ? B->last_kernel_ms : 0.0; }
qp_int QPB_BATCH_LAST_LAUNCHES(const QPBatch *B) { return B ? B->last_launches : 0; }

// Host access to fval / stat of the last host-pointer solve (used by QP_SOLVE to fill `stats`).
qp_int QPB_BATCH_READ_STATS(QPBatch *B, qp_int count, qp_real *fval, qp_real *stat)
{
	if (!B || !B->d_fval) return fail(QPB_E_ARG, "QPB_BATCH_READ_STATS: nothing solved yet", cudaSuccess);
	CUDA_TRY(cudaSetDevice(B->device));
	if (fval) CUDA_TRY(cudaMemcpy(fval, B->d_fval, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost));
	if (stat) CUDA_TRY(cudaMemcpy(stat, B->d_stat, (size_t)count * 8 * sizeof(double), cudaMemcpyDeviceToHost));
	return 0;
}

qp_int QPB_LDL_NUMERIC_BATCH(QPBatch *B, qp_int count, const qp_real *dvals, const qp_real *dzdiag, qp_real *dLx, qp_real *dD, void *stream)
{
	if (!B || count < 0 || !dvals || !dzdiag || !dLx || !dD) return fail(QPB_E_ARG, "QPB_LDL_NUMERIC_BATCH: bad argument", cudaSuccess);
	if (count == 0) return 0;
	CUDA_TRY(cudaSetDevice(B->device));
	const Plan &pl = B->plan;
	const long nV = pl.nnzP + pl.nnzA + pl.nnzG, lnz = pl.Lp[pl.N];
	const size_t smem = (size_t)(((nV + 1) & ~1L) + 2 * pl.m + ((2 * pl.m) & 1) + lnz + (lnz & 1) + 2 * pl.N + 2) * sizeof(double);
	cudaStream_t st = (cudaStream_t)stream;
	CUDA_TRY(cudaMemsetAsync(B->d_counter + 16, 0, sizeof(int), st));
	int occ = 0;
	CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ldl_numeric_kernel, 32, smem));
	long grid = (long)B->sm_count * (occ > 0 ? occ : 1);
	if (grid > count) grid = count;
	ldl_numeric_kernel<<<(unsigned)grid, 32, smem, st>>>(B->d_plan, dvals, dzdiag, dLx, dD, (int)count, B->d_counter + 16);
	CUDA_TRY(cudaGetLastError());
	return 0;
}

qp_int QPB_LDL_SOLVE_BATCH(QPBatch *B, qp_int count, const qp_real *dLx, const qp_real *dD, const qp_real *drhs, qp_real *dsol, void *stream)
{
	if (!B || count < 0 || !dLx || !dD || !drhs || !dsol) return fail(QPB_E_ARG, "QPB_LDL_SOLVE_BATCH: bad argument", cudaSuccess);
	if (count == 0) return 0;
	CUDA_TRY(cudaSetDevice(B->device));
	const Plan &pl = B->plan;
	const long lnz = pl.Lp[pl.N];
	const size_t smem = (size_t)(lnz + (lnz & 1) + 3 * (pl.N + (pl.N & 1))) * sizeof(double);
	cudaStream_t st = (cudaStream_t)stream;
	CUDA_TRY(cudaMemsetAsync(B->d_counter + 32, 0, sizeof(int), st));
	int occ = 0;
	CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ldl_solve_kernel, 32, smem));
	long grid = (long)B->sm_count * (occ > 0 ? occ : 1);
	if (grid > count) grid = count;
	ldl_solve_kernel<<<(unsigned)grid, 32, smem, st>>>(B->d_plan, dLx, dD, drhs, dsol, (int)count, B->d_counter + 32);
	CUDA_TRY(cudaGetLastError());
	return 0;
}

} // extern "C"

// Kernels and host-side launch code of the batched solver; the C ABI of include/qpswift_b200.h.
//
// Launch shape: ONE persistent CTA per SM, one QP per warp, as many warps as slices of shared
// memory fit (a slice = one QP's whole working set: value row, factor, scratch, iterate,
// residuals; SliceLayout in program.h).  Between the coalesced load of its value row and the store
// of its solution a QP never touches HBM.  The warps of a CTA walk the interior-point iteration in
// lock step -- [residuals + stop test | factor | solve | step length, sigma | solve | update] --
// so that they stream the same instruction words of the factor / solve schedules at the same time
// (one L2 read per CTA, L1 hits for the other warps).  A warp whose QP has converged stores it and
// pulls the next QP index from an atomic counter at the top of the cycle; the new QP's initial
// point (kkt_initialize) is just another factor + solve in the same cycle, so iteration counts that
// differ from QP to QP cost no idle cycles.

#include "capi_internal.h"
#include "kernels.cuh"
#include "stream.cuh"
#include "stream_gen.cuh"
#include "warp_kernel.cuh"
#include "wbc.cuh"
#include "rollout.cuh"

#include <algorithm>
#include <cstdio>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

namespace qpb {

static thread_local std::string g_last_error;

// Pattern-specialised kernels compiled into the library (stream_gen.cuh): filled by static
// initialisers of the generated translation units, read by batch_alloc.
static std::vector<GeneratedKernel> &generated_kernels()
{
	static std::vector<GeneratedKernel> v;
	return v;
}
void register_generated_kernel(unsigned long long key, const void *kernel, const char *name, int ring_slots)
{
	generated_kernels().push_back(GeneratedKernel{key, kernel, name, ring_slots});
}
const GeneratedKernel *find_generated_kernel(unsigned long long key)
{
	for (const GeneratedKernel &k : generated_kernels())
		if (k.key == key) return &k;
	return nullptr;
}
int list_generated_kernels(unsigned long long *keys, int capacity)
{
	const std::vector<GeneratedKernel> &v = generated_kernels();
	for (int i = 0; i < (int)v.size() && i < capacity; ++i) keys[i] = v[i].key;
	return (int)v.size();
}

// ... and the pattern-specialised warp-per-QP kernels (warp_kernel.cuh, codegen_warp.py)
static std::vector<GeneratedWarpKernel> &generated_warp_kernels()
{
	static std::vector<GeneratedWarpKernel> v;
	return v;
}
void register_generated_warp_kernel(unsigned long long key, const void *kernel, const void *kernel_free, const char *name, const SliceLayout &lay, int max_warps, int value_row_doubles)
{
	generated_warp_kernels().push_back(GeneratedWarpKernel{key, kernel, kernel_free, name, lay, max_warps, value_row_doubles});
}
// Two variants may be registered per pattern: the throughput one (value row in the global scratch row: more warps per
// SM) and the latency one (whole working set in shared memory: no L2 round trip on a lone warp's critical path).
const GeneratedWarpKernel *find_generated_warp_kernel(unsigned long long key, bool value_row_global)
{
	for (const GeneratedWarpKernel &k : generated_warp_kernels())
		if (k.key == key && (k.value_row_doubles > 0) == value_row_global) return &k;
	return nullptr;
}

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

// LDL_numeric alone: value row + (z,z) diagonal from HBM, factor to HBM.  One warp per CTA.
__global__ void __launch_bounds__(32) ldl_numeric_kernel(const int *__restrict__ plan, ProgramView pg, SliceLayout lay, const double *__restrict__ vals, const double *__restrict__ zdiag, double *__restrict__ Lx, double *__restrict__ D, int count, int *__restrict__ counter)
{
	extern __shared__ __align__(16) double smem[];
	const PlanView pv = view_of(plan);
	const int lane = lane_id();
	double *S = smem;
	if (lane == 0) { S[lay.zero] = 0.0; S[lay.zero + 1] = -0.0; }
	for (;;) {
		int q = 0;
		if (lane == 0) q = atomicAdd(counter, 1);
		q = __shfl_sync(0xffffffffu, q, 0);
		if (q >= count) break;
		const size_t qi = (size_t)q;
		warp_copy_in(S + lay.V, vals + qi * pv.nV, pv.nV);
		warp_copy_in(S + lay.zd, zdiag + qi * pv.m, pv.m);
		__syncwarp();
		run_segment(pg.words, pg.f_first, pg.f_waves, pg.f_width, (unsigned)__cvta_generic_to_shared(S));
		for (int i = lane; i < pv.N; i += 32) S[lay.D + i] = regularise(S[lay.D + i]); // chains store raw pivots (program.h)
		__syncwarp();
		warp_copy_out(Lx + qi * pv.lnz, S + lay.Lx, pv.lnz);
		warp_copy_out(D + qi * pv.N, S + lay.D, pv.N);
		__syncwarp();
	}
}

// LDL_perm + lsolve + dsolve + ltsolve + permt alone.
__global__ void __launch_bounds__(32) ldl_solve_kernel(const int *__restrict__ plan, ProgramView pg, SliceLayout lay, const double *__restrict__ Lx, const double *__restrict__ D, const double *__restrict__ rhs, double *__restrict__ sol, int count, int *__restrict__ counter)
{
	extern __shared__ __align__(16) double smem[];
	const PlanView pv = view_of(plan);
	const int lane = lane_id();
	double *S = smem;
	if (lane == 0) { S[lay.zero] = 0.0; S[lay.zero + 1] = -0.0; }
	for (;;) {
		int q = 0;
		if (lane == 0) q = atomicAdd(counter, 1);
		q = __shfl_sync(0xffffffffu, q, 0);
		if (q >= count) break;
		const size_t qi = (size_t)q;
		warp_copy_in(S + lay.Lx, Lx + qi * pv.lnz, pv.lnz);
		warp_copy_in(S + lay.D, D + qi * pv.N, pv.N);
		warp_copy_in(S + lay.rhs, rhs + qi * pv.N, pv.N);
		__syncwarp();
		run_segment(pg.words, pg.s_first, pg.s_waves, pg.s_width, (unsigned)__cvta_generic_to_shared(S));
		warp_copy_out(sol + qi * pv.N, S + lay.rhs, pv.N);
		__syncwarp();
	}
}

} // namespace qpb

using namespace qpb;

struct QPBatch
{
	Plan plan;
	Program program;
	double sigma_d = 0.0;
	long capacity = 0;
	int device = 0;
	int sm_count = 0;
	int *d_plan = nullptr;
	uint2 *d_prog = nullptr;
	int *d_counter = nullptr;
	int *d_spill = nullptr; // [capacity] QPs the throughput kernel handed over, d_counter[48] = how many
	double *d_spill_sigma = nullptr; // [capacity] sigma of a handed-over QP
	double *d_vscratch = nullptr;    // value rows of the QPs in flight (generated warp-per-QP kernels that keep V outside the slice)
	ProgramView pview;
	int max_warps = 1;      // slices of shared memory that fit next to each other in one CTA
	// throughput kernel (stream.cuh)
	StreamProgram sprog;
	uint2 *d_stream = nullptr;
	double *d_workspace = nullptr;
	StreamView sview;
	int stream_warps = 0;   // warps per CTA (0: the pattern cannot use the throughput kernel)
	long stream_slots = 0;  // warps in flight on the device = tiles of 32 slices in the workspace
	long stream_min = 0;    // smallest batch that goes to the throughput kernel
	const GeneratedKernel *generated = nullptr; // the pattern's specialised kernel, if the library carries one
	const GeneratedWarpKernel *gen_warp = nullptr; // the pattern's specialised warp-per-QP kernel, if the library carries one
	int gen_warps = 0;      // warps per CTA of that kernel
	const GeneratedWarpKernel *gen_warp_small = nullptr; // ... and its latency variant for batches that are in flight all at once
	int gen_warps_small = 0;
	size_t smem_limit = 0;
	// device staging for the host-pointer entry point
	double *d_in[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
	double *d_x = nullptr, *d_y = nullptr, *d_z = nullptr, *d_s = nullptr, *d_fval = nullptr, *d_stat = nullptr;
	int *d_flag = nullptr, *d_iters = nullptr;
	int *h_flag = nullptr, *h_iters = nullptr; // pinned
	cudaStream_t stream = nullptr;
	// dense front-end (QPB_DENSE_TO_VALUES_DEVICE): position of every dense entry in the value row, or -1
	int *d_dense_map[3] = {nullptr, nullptr, nullptr};
	int dense_map_ordering = 0;
	double *d_dense[3] = {nullptr, nullptr, nullptr}; // staging of the host-pointer entry
	int *d_violation = nullptr;
	cudaStream_t copy_stream = nullptr; // QP_SOLVE_BATCH: the value rows travel while the throughput kernel runs
	cudaEvent_t ev_go = nullptr, ev_go2 = nullptr, ev_server = nullptr;
	cudaStream_t aux_stream = nullptr;  // the stragglers' server (latency kernel next to the throughput kernel)
	int *h_marks = nullptr;             // pinned: the "rows [0, mark) have arrived" values written after each chunk
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	double *d_trace = nullptr; // per-iteration trace of the warp-per-QP kernel (QPB_BATCH_ENABLE_TRACE), [capacity][trace_rows][8]
	long trace_rows = 0;
	bool trace_on = false;
	double last_kernel_ms = 0.0;
	long last_launches = 0;
	int last_kernel = 0; // 0 warp-per-QP kernel (interpreted schedules), 1 throughput kernel (one QP per thread), 2 the same, pattern-specialised, 3 warp-per-QP kernel, pattern-specialised
	// batch of ONE (QP_SOLVE's path, the controller's 400 Hz call): inputs, outputs, statistics and the per-iteration trace
	// packed into one device block with a page-locked mirror, so that a solve is one upload, one launch, one download
	double *d_single = nullptr, *h_single = nullptr;
	size_t single_in = 0, single_total = 0; // doubles: inputs; inputs + outputs + fval + stat[8] + (flag, iters) + trace
	long single_trace_rows = -1;
	bool single_valid = false;              // h_single holds the results of the last solve
	// WBC formulation on the device (wbc.cuh): descriptor tables per contact pattern kind, built on first use
	unsigned *d_wbc_desc = nullptr;
	double *d_wbc_consts = nullptr;
	WbcTables wbc;
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
	const size_t slice = (size_t)B->program.lay.total * sizeof(double);
	B->smem_limit = (size_t)prop.sharedMemPerBlockOptin;
	if (slice > B->smem_limit) return fail(QPB_E_PATTERN, "pattern does not fit in shared memory", cudaSuccess);
	B->max_warps = (int)std::min<size_t>(MAX_WARPS_PER_CTA, B->smem_limit / slice);
	CUDA_TRY(cudaFuncSetAttribute(qp_solve_warp_kernel<InterpretedWaves, MAX_WARPS_PER_CTA, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B->smem_limit));
	if (B->sprog.ok && !getenv("QPB_WARP_INTERPRET")) {
		const unsigned long long key = stream_program_key(B->sprog);
		const GeneratedWarpKernel *g = find_generated_warp_kernel(key, true);
		const GeneratedWarpKernel *gs = find_generated_warp_kernel(key, false);
		if (!g) { g = gs; gs = nullptr; }
		if (g) {
			B->gen_warp = g;
			B->gen_warps = (int)std::min<size_t>((size_t)g->max_warps, B->smem_limit / ((size_t)g->lay.total * sizeof(double)));
			if (getenv("QPB_WARP_WARPS")) B->gen_warps = std::max(1, std::min(B->gen_warps, atoi(getenv("QPB_WARP_WARPS"))));
			CUDA_TRY(cudaFuncSetAttribute(g->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B->smem_limit));
			CUDA_TRY(cudaFuncSetAttribute(g->kernel_free, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B->smem_limit));
			if (g->value_row_doubles) // one scratch row per warp of the largest grid (+ a tail: lanes beyond a row's end may read past it)
				CUDA_TRY(cudaMalloc(&B->d_vscratch, ((size_t)B->sm_count * g->max_warps * g->value_row_doubles + 64) * sizeof(double)));
		}
		if (gs && !getenv("QPB_WARP_NO_SMALL")) {
			B->gen_warp_small = gs;
			B->gen_warps_small = (int)std::min<size_t>((size_t)gs->max_warps, B->smem_limit / ((size_t)gs->lay.total * sizeof(double)));
			CUDA_TRY(cudaFuncSetAttribute(gs->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B->smem_limit));
			CUDA_TRY(cudaFuncSetAttribute(gs->kernel_free, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B->smem_limit));
		}
	}
	CUDA_TRY(cudaFuncSetAttribute(ldl_numeric_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B->smem_limit));
	CUDA_TRY(cudaFuncSetAttribute(ldl_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B->smem_limit));

	CUDA_TRY(cudaMalloc(&B->d_plan, pl.blob.size() * sizeof(int32_t)));
	CUDA_TRY(cudaMemcpy(B->d_plan, pl.blob.data(), pl.blob.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
	const Program &pg = B->program;
	CUDA_TRY(cudaMalloc(&B->d_prog, pg.words.size() * sizeof(uint64_t)));
	CUDA_TRY(cudaMemcpy(B->d_prog, pg.words.data(), pg.words.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
	B->pview.words = B->d_prog;
	B->pview.f_first = pg.factor.first; B->pview.f_waves = pg.factor.waves; B->pview.f_width = pg.factor.width;
	B->pview.s_first = pg.solve.first; B->pview.s_waves = pg.solve.waves; B->pview.s_width = pg.solve.width;
	CUDA_TRY(cudaMalloc(&B->d_counter, 64 * sizeof(int)));
	// ---- throughput kernel: one thread per QP, slices interleaved in an HBM workspace ---------------
	if (B->sprog.ok) {
		if (!getenv("QPB_STREAM_INTERPRET")) B->generated = find_generated_kernel(stream_program_key(B->sprog));
		// shared memory per warp: the lanes' work vectors, plus the operand ring of a generated program
		const size_t window = (size_t)(B->sprog.lay.scratch + (B->generated ? B->generated->ring_slots : 0)) * 32 * sizeof(double);
		long cap = getenv("QPB_STREAM_WARPS") ? atol(getenv("QPB_STREAM_WARPS")) : STREAM_MAX_WARPS;
		if (cap > STREAM_MAX_WARPS) cap = STREAM_MAX_WARPS;
		B->stream_warps = (int)std::min<size_t>((size_t)std::max(cap, 1L), B->smem_limit / window);
	}
	if (B->stream_warps > 0) {
		const StreamProgram &sp = B->sprog;
		CUDA_TRY(cudaFuncSetAttribute(qp_solve_stream_kernel<InterpretedProgram<4>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B->smem_limit));
		CUDA_TRY(cudaFuncSetAttribute(qp_solve_stream_kernel<InterpretedProgram<8>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B->smem_limit));
		CUDA_TRY(cudaMalloc(&B->d_stream, sp.words.size() * sizeof(uint64_t)));
		CUDA_TRY(cudaMemcpy(B->d_stream, sp.words.data(), sp.words.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
		B->sview.words = B->d_stream;
		if (B->generated) CUDA_TRY(cudaFuncSetAttribute(B->generated->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B->smem_limit));
		for (int s = 0; s < SSEG_COUNT; ++s) { B->sview.first[s] = sp.seg[s].first; B->sview.groups[s] = sp.seg[s].groups; }
		const long full = (long)B->sm_count * B->stream_warps;
		B->stream_slots = std::min(full, (B->capacity + 31) / 32);
		// Which kernel takes a batch.  Without a pattern-specialised warp-per-QP kernel: the throughput kernel from
		// 2048 QPs on.  With one (DogBot trot / stance / crawl): the warp-per-QP kernel keeps the factor in shared
		// memory, has no tail (1,480-1,776 QPs in flight instead of 37,888) and no stragglers' hand-over, and since round 2
		// (value row outside the slice: 10-12 warps per SM) it is the faster one at every batch size measured
		// (trot 65,536: 2.05 M against 1.40 M solves/s; 262,144, at 10 warps: 1.71 M against 1.63 M).  QPB_STREAM_MIN overrides.
		B->stream_min = getenv("QPB_STREAM_MIN") ? atol(getenv("QPB_STREAM_MIN")) : (B->gen_warp ? LONG_MAX : 2048);
		// A pattern whose window leaves room for fewer than 4 warps per SM (e.g. the horizon-10 MPC QP:
		// N = 600, one warp) is faster on the latency kernel; QPB_STREAM_MIN_WARPS overrides.
		const long min_warps = getenv("QPB_STREAM_MIN_WARPS") ? atol(getenv("QPB_STREAM_MIN_WARPS")) : 4;
		if (B->stream_warps < min_warps) B->stream_min = LONG_MAX;
		if (B->capacity >= B->stream_min) CUDA_TRY(cudaMalloc(&B->d_spill, (size_t)B->capacity * sizeof(int)));
		if (B->capacity >= B->stream_min) CUDA_TRY(cudaMalloc(&B->d_spill_sigma, (size_t)B->capacity * sizeof(double)));
		if (B->capacity >= B->stream_min) CUDA_TRY(cudaMalloc(&B->d_workspace, (size_t)B->stream_slots * sp.lay.total * 32 * sizeof(double)));
	}
	CUDA_TRY(cudaStreamCreateWithFlags(&B->stream, cudaStreamNonBlocking));
	CUDA_TRY(cudaStreamCreateWithFlags(&B->copy_stream, cudaStreamNonBlocking));
	CUDA_TRY(cudaEventCreateWithFlags(&B->ev_go, cudaEventDisableTiming));
	CUDA_TRY(cudaEventCreateWithFlags(&B->ev_go2, cudaEventDisableTiming));
	CUDA_TRY(cudaEventCreateWithFlags(&B->ev_server, cudaEventDisableTiming));
	CUDA_TRY(cudaStreamCreateWithFlags(&B->aux_stream, cudaStreamNonBlocking));
	CUDA_TRY(cudaMallocHost(&B->h_marks, 256 * sizeof(int)));
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

static int launch_latency(QPBatch *B, const BatchIO &io, const SolveParams &prm, long count, int *counter, cudaStream_t st)
{
	// one persistent CTA per SM; small batches spread over the SMs with fewer warps per CTA.  A batch that is in flight
	// all at once (every QP has its warp from the start: control ticks, rollouts) is a matter of LATENCY: it goes to
	// the variant that keeps the whole working set in shared memory.
	const GeneratedWarpKernel *gk = B->gen_warp;
	int most = gk ? B->gen_warps : B->max_warps;
	if (B->gen_warp_small && !io.server_done && count <= (long)B->sm_count * B->gen_warps_small) { gk = B->gen_warp_small; most = B->gen_warps_small; }
	long warps = (count + B->sm_count - 1) / B->sm_count;
	if (warps > most) warps = most;
	if (warps < 1) warps = 1;
	long grid = (count + warps - 1) / warps;
	if (grid > B->sm_count) grid = B->sm_count;
	const SliceLayout &lay = gk ? gk->lay : B->program.lay;
	const size_t smem = (size_t)warps * lay.total * sizeof(double);
	if (gk) {
		const int *plan = B->d_plan;
		BatchIO gio = io;
		gio.vscratch = B->d_vscratch;
		void *args[] = {&plan, &B->pview, const_cast<SliceLayout *>(&lay), &gio, const_cast<SolveParams *>(&prm), &counter};
		CUDA_TRY(cudaLaunchKernel(getenv("QPB_WARP_FREE") && !io.server_done ? gk->kernel_free : gk->kernel, dim3((unsigned)grid), dim3((unsigned)warps * 32), args, smem, st));
	} else qp_solve_warp_kernel<InterpretedWaves, MAX_WARPS_PER_CTA, true><<<(unsigned)grid, (unsigned)warps * 32, smem, st>>>(B->d_plan, B->pview, B->program.lay, io, prm, counter);
	CUDA_TRY(cudaGetLastError());
	return 0;
}

// Large batches: one thread per QP (stream.cuh).  One CTA per SM, stream_warps warps each.  A QP that
// needs far more iterations than its neighbours (the rare trot QP that runs into maxit = 100) would keep
// its warp -- and the launch -- alive long after the batch is done, at one thread's pace; the throughput
// kernel therefore hands every QP that has not converged after QPB_STREAM_SPILL iterations (default 24)
// over to the latency kernel, a warp each, together with its iterate (x, y, z, s, sigma, iteration count:
// the complete state of QP_SOLVE's loop between two iterations), so that nothing is solved twice.  Results
// do not depend on which kernel produced them (both are bit-identical with the reference).
static int launch_stream(QPBatch *B, BatchIO io, SolveParams prm, cudaStream_t st)
{
	const StreamProgram &sp = B->sprog;
	long slots = std::min(B->stream_slots, ((long)io.count + 31) / 32);
	const long warps = std::min<long>(B->stream_warps, std::max(1L, (slots + B->sm_count - 1) / B->sm_count));
	const size_t smem = (size_t)warps * (sp.lay.scratch + (B->generated ? B->generated->ring_slots : 0)) * 32 * sizeof(double);
	StreamDims dm;
	dm.n = (int)B->plan.n; dm.m = (int)B->plan.m; dm.p = (int)B->plan.p; dm.N = (int)B->plan.N;
	dm.nnzP = (int)B->plan.nnzP; dm.nnzA = (int)B->plan.nnzA; dm.nnzG = (int)B->plan.nnzG;
	const long spill = getenv("QPB_STREAM_SPILL") ? atol(getenv("QPB_STREAM_SPILL")) : 24;
	prm.spill_after = spill > 0 && spill < prm.maxit ? (int)spill : 0;
	int *spill_count = B->d_counter + 48, *second_counter = B->d_counter + 56, *done_ctas = B->d_counter + 57, *all_done = B->d_counter + 58;
	io.done_ctas = io.all_done = nullptr;
	// The stragglers' server runs NEXT TO the throughput kernel, on an SM that kernel leaves free (it needs a
	// whole SM's shared memory); QPB_STREAM_SERVER=0: one launch after the other on the same stream instead.
	const bool server = prm.spill_after && !(getenv("QPB_STREAM_SERVER") && atoi(getenv("QPB_STREAM_SERVER")) == 0);
	long grid = (slots + warps - 1) / warps;
	if (prm.spill_after) {
		CUDA_TRY(cudaMemsetAsync(spill_count, 0, 16 * sizeof(int), st));
		io.spill = B->d_spill;
		io.spill_count = spill_count;
		io.spill_sigma = getenv("QPB_STREAM_RESTART") ? nullptr : B->d_spill_sigma; // QPB_STREAM_RESTART=1: stragglers start from scratch
		if (server) {
			if (grid >= B->sm_count) { grid = B->sm_count - 1; slots = grid * warps; }
			CUDA_TRY(cudaMemsetAsync(B->d_spill, 0xff, (size_t)std::min<long>(B->capacity, io.count) * sizeof(int), st));
			io.done_ctas = done_ctas;
			io.all_done = all_done;
			CUDA_TRY(cudaEventRecord(B->ev_go2, st));
		}
	}
	if (B->generated) {
		int nslots = (int)slots;
		double *ws = B->d_workspace;
		int *ctr = B->d_counter;
		void *args[] = {&B->sview, const_cast<StreamLayout *>(&sp.lay), &dm, &io, &prm, &ws, &nslots, &ctr};
		CUDA_TRY(cudaLaunchKernel(B->generated->kernel, dim3((unsigned)grid), dim3((unsigned)warps * 32), args, smem, st));
	} else if (sp.group == 4) qp_solve_stream_kernel<InterpretedProgram<4>><<<(unsigned)grid, (unsigned)warps * 32, smem, st>>>(B->sview, sp.lay, dm, io, prm, B->d_workspace, (int)slots, B->d_counter);
	else if (sp.group == 8) qp_solve_stream_kernel<InterpretedProgram<8>><<<(unsigned)grid, (unsigned)warps * 32, smem, st>>>(B->sview, sp.lay, dm, io, prm, B->d_workspace, (int)slots, B->d_counter);
	else return fail(QPB_E_ARG, "stream program: unsupported group size", cudaSuccess);
	CUDA_TRY(cudaGetLastError());
	B->last_launches = 1;
	B->last_kernel = B->generated ? 2 : 1;
	if (prm.spill_after) { // the stragglers: solved from scratch by the latency kernel, a warp each
		io.index = B->d_spill;
		io.count_ptr = spill_count;
		io.spill = nullptr;
		io.spill_count = nullptr;
		io.ready = nullptr;
		io.trace = nullptr;
		io.done_ctas = nullptr;
		io.server_done = server ? all_done : nullptr;
		io.all_done = nullptr;
		prm.spill_after = 0;
		cudaStream_t second = st;
		if (server) { // launched after the throughput kernel (a tool that serialises launches then simply runs it afterwards)
			second = B->aux_stream;
			CUDA_TRY(cudaStreamWaitEvent(second, B->ev_go2, 0));
		}
		const long lat_warps = B->gen_warp ? B->gen_warps : B->max_warps;
		const int rc = launch_latency(B, io, prm, server ? lat_warps : (long)B->sm_count * lat_warps, second_counter, second);
		if (rc) return rc;
		if (server) {
			CUDA_TRY(cudaEventRecord(B->ev_server, second));
			CUDA_TRY(cudaStreamWaitEvent(st, B->ev_server, 0));
		}
		B->last_launches = 2;
	}
	return 0;
}

static int launch_fused(QPBatch *B, BatchIO io, const settings *opts, cudaStream_t st)
{
	SolveParams prm;
	prm.maxit = (int)opts->maxit;
	prm.reltol = opts->reltol;
	prm.abstol = opts->abstol;
	prm.sigma0 = opts->sigma;
	prm.sigma_d = B->sigma_d;
	prm.spill_after = 0;
	prm.groups = getenv("QPB_WARP_GROUPS") ? atoi(getenv("QPB_WARP_GROUPS")) : 1;
	prm.group_delay_ns = getenv("QPB_WARP_GROUP_DELAY") ? (unsigned)atol(getenv("QPB_WARP_GROUP_DELAY")) : 25000u;
	io.index = nullptr;
	io.count_ptr = nullptr;
	io.spill = nullptr;
	io.spill_count = nullptr;
	io.spill_sigma = nullptr;
	io.server_done = nullptr;
	io.done_ctas = io.all_done = nullptr;
	CUDA_TRY(cudaMemsetAsync(B->d_counter, 0, sizeof(int), st));
	if (B->d_workspace && io.count >= B->stream_min) return launch_stream(B, io, prm, st);
	B->last_kernel = B->gen_warp ? 3 : 0;
	B->last_launches = 1;
	return launch_latency(B, io, prm, io.count, B->d_counter, st);
}

static qp_int qp_solve_batch_host(QPBatch *B, qp_int count, const settings *opts, const qp_real *Ppr, const qp_real *Apr, const qp_real *Gpr, const qp_real *c, const qp_real *h, const qp_real *b, qp_real *x, qp_real *y, qp_real *z, qp_real *s, qp_int *flag, qp_int *iters);

// ---- dense front-end ----------------------------------------------------------------
// densetosparse / densetosparse_ROWMAJOR (Auxilary.c:1154-1273) for a whole batch against ONE fixed
// pattern: the value of every entry of the plan's pattern goes to its place in the value row; what
// the reference's data-dependent pattern would have done differently is reported per QP
// (bit 0: a nonzero outside the pattern, bit 1: an exact zero inside it).  The dense blocks are read
// front to back (coalesced), the 4-byte map entries stay in L1/L2, a row's values land next to each other.
// Every thread walks the cells of ALL QPs of one matrix with a grid stride (no idle lanes at the end of a
// 900- or 360-cell block), four requests in flight; the violation bits are rare and go out by atomicOr.
__device__ __forceinline__ void dense_gather_all(const double *__restrict__ dense, const int *__restrict__ map, const unsigned cells, const unsigned nnz, double *__restrict__ rows, int *__restrict__ violation, const unsigned first_qp, const unsigned total)
{
	// (32-bit cell numbers: the host splits a batch whose matrix has 2^32 cells or more into several launches)
	const unsigned stride = gridDim.x * blockDim.x;
	dense += (size_t)first_qp * cells;
	for (unsigned g0 = blockIdx.x * blockDim.x + threadIdx.x; g0 < total; g0 += 4 * stride) {
		double v[4];
		int k[4];
		unsigned q[4];
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const unsigned g = min(g0 + j * stride, total - 1);
			q[j] = g / cells;
			v[j] = dense[g];
			k[j] = __ldg(map + (g - q[j] * cells));
		}
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			if (g0 + j * stride >= total || g0 + j * stride < g0) continue;
			int bits = 0;
			if (k[j] >= 0) {
				rows[(size_t)(first_qp + q[j]) * nnz + k[j]] = v[j];
				if (v[j] == 0.0) bits = 2;
			} else if (v[j] != 0.0) bits = 1;
			if (bits && violation) atomicOr(violation + first_qp + q[j], bits);
		}
	}
}

__global__ void __launch_bounds__(256) dense_gather_kernel(const double *__restrict__ dP, const double *__restrict__ dA, const double *__restrict__ dG, const int *__restrict__ mapP, const int *__restrict__ mapA, const int *__restrict__ mapG, const int n, const int m, const int p, const int nnzP, const int nnzA, const int nnzG, double *__restrict__ oP, double *__restrict__ oA, double *__restrict__ oG, int *__restrict__ violation, const int first_qp, const int count)
{
	dense_gather_all(dG, mapG, (unsigned)(m * n), (unsigned)nnzG, oG, violation, (unsigned)first_qp, (unsigned)count * (unsigned)(m * n));
	dense_gather_all(dP, mapP, (unsigned)(n * n), (unsigned)nnzP, oP, violation, (unsigned)first_qp, (unsigned)count * (unsigned)(n * n));
	if (p) dense_gather_all(dA, mapA, (unsigned)(p * n), (unsigned)nnzA, oA, violation, (unsigned)first_qp, (unsigned)count * (unsigned)(p * n));
}

template <bool FUSED>
__global__ void __launch_bounds__(1024) fp64_peak_kernel(double *sink, const int steps, const double b, const double c)
{
	double a[8];
#pragma unroll
	for (int k = 0; k < 8; ++k) a[k] = 1.0 + 1e-3 * (double)(threadIdx.x + k);
	for (int i = 0; i < steps; ++i) {
#pragma unroll
		for (int k = 0; k < 8; ++k) a[k] = FUSED ? __fma_rn(a[k], b, c) : __dadd_rn(__dmul_rn(a[k], b), c);
	}
	double t = 0.0;
#pragma unroll
	for (int k = 0; k < 8; ++k) t += a[k];
	sink[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = t;
}

static int ensure_dense_maps(QPBatch *B, int ordering)
{
	if (B->d_dense_map[0] && B->dense_map_ordering == ordering) return 0;
	const Plan &pl = B->plan;
	const DevicePlanHeader *hd = reinterpret_cast<const DevicePlanHeader *>(pl.blob.data());
	const int32_t *blob = pl.blob.data();
	const long n = pl.n;
	struct View { long rows; int32_t ptr, col, val; long off; };
	const View views[3] = {{pl.n, hd->prow_ptr, hd->prow_col, hd->prow_val, 0}, {pl.p, hd->arow_ptr, hd->arow_col, hd->arow_val, pl.nnzP}, {pl.m, hd->grow_ptr, hd->grow_col, hd->grow_val, pl.nnzP + pl.nnzA}};
	for (int k = 0; k < 3; ++k) {
		const View &v = views[k];
		if (v.rows == 0) continue;
		std::vector<int32_t> map((size_t)v.rows * n, -1);
		for (long r = 0; r < v.rows; ++r)
			for (int32_t t = blob[v.ptr + r]; t < blob[v.ptr + r + 1]; ++t) {
				const long c = blob[v.col + t];
				const size_t cell = ordering == ROW_MAJOR_ORDERING ? (size_t)r * n + c : (size_t)c * v.rows + r;
				map[cell] = (int32_t)(blob[v.val + t] - v.off);
			}
		if (!B->d_dense_map[k]) CUDA_TRY(cudaMalloc(&B->d_dense_map[k], map.size() * sizeof(int32_t)));
		CUDA_TRY(cudaMemcpy(B->d_dense_map[k], map.data(), map.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
	}
	B->dense_map_ordering = ordering;
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
	if (!plan->program.ok) { g_last_error = "QPB_BATCH_CREATE: " + plan->program.error; return nullptr; }
	QPBatch *B = new QPBatch;
	B->plan = plan->plan;
	B->program = plan->program;
	B->sprog = plan->stream;
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
	qpb_build(tmp, proto->n, proto->m, proto->p, proto->P->jc, proto->P->ir, A ? A->jc : nullptr, A ? A->ir : nullptr, proto->G->jc, proto->G->ir, proto->kkt ? proto->kkt->P : nullptr);
	return QPB_BATCH_CREATE(&tmp, proto->sigma_d, capacity, device);
}

void QP_CLEANUP_BATCH(QPBatch *B)
{
	if (!B) return;
	cudaSetDevice(B->device);
	cudaFree(B->d_plan);
	cudaFree(B->d_prog);
	cudaFree(B->d_stream);
	cudaFree(B->d_spill_sigma);
	cudaFree(B->d_vscratch);
	for (int k = 0; k < 3; ++k) { cudaFree(B->d_dense_map[k]); cudaFree(B->d_dense[k]); }
	cudaFree(B->d_violation);
	cudaFree(B->d_single);
	if (B->h_single) cudaFreeHost(B->h_single);
	cudaFree(B->d_wbc_desc);
	cudaFree(B->d_wbc_consts);
	cudaFree(B->d_workspace);
	cudaFree(B->d_trace);
	cudaFree(B->d_spill);
	cudaFree(B->d_counter);
	for (int i = 0; i < 6; ++i) cudaFree(B->d_in[i]);
	cudaFree(B->d_x); cudaFree(B->d_y); cudaFree(B->d_z); cudaFree(B->d_s);
	cudaFree(B->d_fval); cudaFree(B->d_stat); cudaFree(B->d_flag); cudaFree(B->d_iters);
	if (B->h_flag) cudaFreeHost(B->h_flag);
	if (B->h_iters) cudaFreeHost(B->h_iters);
	if (B->stream) cudaStreamDestroy(B->stream);
	if (B->copy_stream) cudaStreamDestroy(B->copy_stream);
	if (B->ev_go) cudaEventDestroy(B->ev_go);
	if (B->ev_go2) cudaEventDestroy(B->ev_go2);
	if (B->ev_server) cudaEventDestroy(B->ev_server);
	if (B->aux_stream) cudaStreamDestroy(B->aux_stream);
	if (B->h_marks) cudaFreeHost(B->h_marks);
	if (B->ev0) cudaEventDestroy(B->ev0);
	if (B->ev1) cudaEventDestroy(B->ev1);
	cudaGetLastError();
	delete B;
}

qp_int QP_SOLVE_BATCH_DEVICE(QPBatch *B, qp_int count, const settings *opts, const qp_real *dPpr, const qp_real *dApr, const qp_real *dGpr, const qp_real *dc, const qp_real *dh, const qp_real *db, qp_real *dx, qp_real *dy, qp_real *dz, qp_real *ds, int *dflag, int *diters, qp_real *dfval, qp_real *dstat, void *stream)
{
	// (the spill list and sigma array of the throughput kernel are sized by the capacity)
	if (!B || !opts || count < 0 || count > B->capacity || !dPpr || !dGpr || !dc || !dh || (B->plan.p && (!dApr || !db))) return fail(QPB_E_ARG, "QP_SOLVE_BATCH_DEVICE: bad argument (null pointer, count > capacity)", cudaSuccess);
	if (count == 0) return 0;
	CUDA_TRY(cudaSetDevice(B->device));
	BatchIO io;
	io.Ppr = dPpr; io.Apr = dApr; io.Gpr = dGpr; io.c = dc; io.h = dh; io.b = db;
	io.x = dx; io.y = dy; io.z = dz; io.s = ds; io.flag = dflag; io.iters = diters; io.fval = dfval; io.stat = dstat;
	io.count = (int)count;
	io.ready = nullptr;
	io.trace = nullptr;
	return launch_fused(B, io, opts, (cudaStream_t)stream);
}

// The QPs whose rows are listed in dindex[0 .. *dcount_ptr) (device memory; the count is read by the kernel, so a
// list compacted on the device needs no host round trip and the launch can sit in a CUDA graph); rows = the
// largest index + 1 the arrays hold.  Always the warp-per-QP kernel.
qp_int QP_SOLVE_BATCH_DEVICE_INDEXED(QPBatch *B, qp_int rows, const settings *opts, const int *dindex, const int *dcount_ptr, const qp_real *dPpr, const qp_real *dApr, const qp_real *dGpr, const qp_real *dc, const qp_real *dh, const qp_real *db, qp_real *dx, qp_real *dy, qp_real *dz, qp_real *ds, int *dflag, int *diters, void *stream)
{
	if (!B || !opts || rows < 1 || rows > B->capacity || rows >= (1 << 24) || !dindex || !dcount_ptr || !dPpr || !dGpr || !dc || !dh || (B->plan.p && (!dApr || !db))) return fail(QPB_E_ARG, "QP_SOLVE_BATCH_DEVICE_INDEXED: bad argument (null pointer, rows > capacity)", cudaSuccess);
	CUDA_TRY(cudaSetDevice(B->device));
	BatchIO io;
	io.Ppr = dPpr; io.Apr = dApr; io.Gpr = dGpr; io.c = dc; io.h = dh; io.b = db;
	io.x = dx; io.y = dy; io.z = dz; io.s = ds; io.flag = dflag; io.iters = diters; io.fval = nullptr; io.stat = nullptr;
	io.count = (int)rows;
	io.index = dindex; io.count_ptr = dcount_ptr;
	io.spill = nullptr; io.spill_count = nullptr; io.spill_sigma = nullptr; io.ready = nullptr; io.server_done = nullptr; io.done_ctas = io.all_done = nullptr; io.trace = nullptr;
	SolveParams prm;
	prm.maxit = (int)opts->maxit; prm.reltol = opts->reltol; prm.abstol = opts->abstol; prm.sigma0 = opts->sigma; prm.sigma_d = B->sigma_d; prm.spill_after = 0; prm.groups = 1; prm.group_delay_ns = 0;
	cudaStream_t st = (cudaStream_t)stream;
	CUDA_TRY(cudaMemsetAsync(B->d_counter, 0, sizeof(int), st));
	B->last_kernel = B->gen_warp ? 3 : 0;
	B->last_launches = 1;
	return launch_latency(B, io, prm, rows, B->d_counter, st);
}

qp_int QP_SOLVE_BATCH(QPBatch *B, qp_int count, const settings *opts, const qp_real *Ppr, const qp_real *Apr, const qp_real *Gpr, const qp_real *c, const qp_real *h, const qp_real *b, qp_real *x, qp_real *y, qp_real *z, qp_real *s, qp_int *flag, qp_int *iters)
{
	if (!B || !opts || count < 0 || count > B->capacity || !Ppr || !Gpr || !c || !h || (B->plan.p && (!Apr || !b))) return fail(QPB_E_ARG, "QP_SOLVE_BATCH: bad argument", cudaSuccess);
	return qp_solve_batch_host(B, count, opts, Ppr, Apr, Gpr, c, h, b, x, y, z, s, flag, iters);
}

} // extern "C"

// Host-pointer solve.  Ppr == nullptr: the value rows P|A|G are in the staging buffers already (dense front-end).
// One QP through a batch of capacity one (QP_SOLVE; the controller calls it every 2.5 ms, main.cpp:1649-1663): the
// value row goes up in ONE copy from a page-locked block, and x, y, z, s, the objective, the statistics, flag,
// iteration count and the per-iteration trace come back in ONE copy -- eight driver calls per solve instead of two dozen.
static qp_int qp_solve_single(QPBatch *B, const settings *opts, const qp_real *Ppr, const qp_real *Apr, const qp_real *Gpr, const qp_real *c, const qp_real *h, const qp_real *b, qp_real *x, qp_real *y, qp_real *z, qp_real *s, qp_int *flag, qp_int *iters)
{
	std::lock_guard<std::mutex> guard(B->lock);
	CUDA_TRY(cudaSetDevice(B->device));
	const Plan &pl = B->plan;
	const size_t w[6] = {(size_t)pl.nnzP, (size_t)pl.nnzA, (size_t)pl.nnzG, (size_t)pl.n, (size_t)pl.m, (size_t)pl.p};
	const long rows = B->trace_on ? (long)opts->maxit + 2 : 0;
	if (!B->d_single || rows != B->single_trace_rows) {
		cudaFree(B->d_single);
		if (B->h_single) cudaFreeHost(B->h_single);
		B->d_single = B->h_single = nullptr;
		size_t in = 0;
		for (int i = 0; i < 6; ++i) in += (w[i] + 1) & ~(size_t)1; // (every array starts on 16 bytes: warp_copy_in's fast path)
		B->single_in = in;
		B->single_total = in + (((size_t)pl.n + 1) & ~(size_t)1) + (((size_t)pl.p + 1) & ~(size_t)1) + 2 * (((size_t)pl.m + 1) & ~(size_t)1) + 2 + 8 + 2 + (size_t)rows * 8;
		CUDA_TRY(cudaMalloc(&B->d_single, B->single_total * sizeof(double)));
		CUDA_TRY(cudaMallocHost(&B->h_single, B->single_total * sizeof(double)));
		B->single_trace_rows = rows;
	}
	B->single_valid = false;
	const qp_real *src[6] = {Ppr, Apr, Gpr, c, h, b};
	double *din[6];
	size_t at = 0;
	for (int i = 0; i < 6; ++i) {
		din[i] = B->d_single + at;
		if (w[i]) memcpy(B->h_single + at, src[i], w[i] * sizeof(double));
		at += (w[i] + 1) & ~(size_t)1;
	}
	BatchIO io;
	io.Ppr = din[0]; io.Apr = din[1]; io.Gpr = din[2]; io.c = din[3]; io.h = din[4]; io.b = din[5];
	const size_t ox = at, oy = ox + (((size_t)pl.n + 1) & ~(size_t)1), oz = oy + (((size_t)pl.p + 1) & ~(size_t)1), os = oz + (((size_t)pl.m + 1) & ~(size_t)1);
	const size_t ofval = os + (((size_t)pl.m + 1) & ~(size_t)1), ostat = ofval + 2, oint = ostat + 8, otrace = oint + 2;
	io.x = B->d_single + ox; io.y = B->d_single + oy; io.z = B->d_single + oz; io.s = B->d_single + os;
	io.fval = B->d_single + ofval; io.stat = B->d_single + ostat;
	io.flag = reinterpret_cast<int *>(B->d_single + oint); io.iters = io.flag + 2;
	io.count = 1;
	io.ready = nullptr;
	io.trace = rows ? B->d_single + otrace : nullptr; // (only the rows of iterations that ran are read back, and those are written whole)
	cudaStream_t st = B->stream;
	CUDA_TRY(cudaMemcpyAsync(B->d_single, B->h_single, B->single_in * sizeof(double), cudaMemcpyHostToDevice, st));
	CUDA_TRY(cudaEventRecord(B->ev0, st));
	const int rc = launch_fused(B, io, opts, st);
	if (rc) return rc;
	CUDA_TRY(cudaEventRecord(B->ev1, st));
	CUDA_TRY(cudaMemcpyAsync(B->h_single + B->single_in, B->d_single + B->single_in, (B->single_total - B->single_in) * sizeof(double), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	float ms = 0.f;
	CUDA_TRY(cudaEventElapsedTime(&ms, B->ev0, B->ev1));
	B->last_kernel_ms = ms;
	if (x) memcpy(x, B->h_single + ox, pl.n * sizeof(double));
	if (y && pl.p) memcpy(y, B->h_single + oy, pl.p * sizeof(double));
	if (z) memcpy(z, B->h_single + oz, pl.m * sizeof(double));
	if (s) memcpy(s, B->h_single + os, pl.m * sizeof(double));
	const int *ints = reinterpret_cast<const int *>(B->h_single + oint);
	if (flag) flag[0] = ints[0];
	if (iters) iters[0] = ints[2];
	B->single_valid = true;
	return ints[0] != QP_OPTIMAL ? 1 : 0;
}

static qp_int qp_solve_batch_host(QPBatch *B, qp_int count, const settings *opts, const qp_real *Ppr, const qp_real *Apr, const qp_real *Gpr, const qp_real *c, const qp_real *h, const qp_real *b, qp_real *x, qp_real *y, qp_real *z, qp_real *s, qp_int *flag, qp_int *iters)
{
	if (count == 0) return 0;
	if (count == 1 && B->capacity == 1 && Ppr && !getenv("QPB_NO_SINGLE_PATH")) return qp_solve_single(B, opts, Ppr, Apr, Gpr, c, h, b, x, y, z, s, flag, iters);
	B->single_valid = false;
	std::lock_guard<std::mutex> guard(B->lock);
	CUDA_TRY(cudaSetDevice(B->device));
	int rc = batch_alloc_staging(B);
	if (rc) return rc;
	const Plan &pl = B->plan;
	cudaStream_t st = B->stream;
	const size_t n = (size_t)count;
	const qp_real *src[6] = {Ppr, Apr, Gpr, c, h, b};
	size_t widths[6] = {(size_t)pl.nnzP, (size_t)pl.nnzA, (size_t)pl.nnzG, (size_t)pl.n, (size_t)pl.m, (size_t)pl.p};
	if (!Ppr) widths[0] = widths[1] = widths[2] = 0; // already on the device
	BatchIO io;
	io.Ppr = B->d_in[0]; io.Apr = B->d_in[1]; io.Gpr = B->d_in[2]; io.c = B->d_in[3]; io.h = B->d_in[4]; io.b = B->d_in[5];
	io.x = B->d_x; io.y = B->d_y; io.z = B->d_z; io.s = B->d_s; io.flag = B->d_flag; io.iters = B->d_iters; io.fval = B->d_fval; io.stat = B->d_stat;
	io.count = (int)count;
	io.ready = nullptr;
	io.trace = nullptr;
	if (B->trace_on) { // verbose single solves (QP_SOLVE): per-iteration lines and the time in LDL_numeric
		const long rows = (long)opts->maxit + 2;
		if (rows > B->trace_rows) {
			cudaFree(B->d_trace);
			B->d_trace = nullptr;
			CUDA_TRY(cudaMalloc(&B->d_trace, (size_t)B->capacity * rows * 8 * sizeof(double)));
			B->trace_rows = rows;
		}
		CUDA_TRY(cudaMemsetAsync(B->d_trace, 0, (size_t)B->capacity * B->trace_rows * 8 * sizeof(double), st));
		io.trace = B->d_trace;
	}
	// Results straight into the caller's buffers where those are page-locked (and therefore mapped into the
	// device's address space): the throughput kernel writes a finished QP's rows 256 bytes at a time over the
	// bus while the others are still being solved, instead of one copy of everything after the kernel.
	bool direct[4] = {false, false, false, false};
	// (the throughput kernel and the warp-per-QP kernels, generated or interpreted, take QPs in index order from a queue)
	const bool overlap = ((B->d_workspace && count >= B->stream_min) || count >= 4096) && !getenv("QPB_NO_OVERLAP");
	if (overlap) {
		qp_real *host_out[4] = {x, y, z, s};
		double **dev_out[4] = {&io.x, &io.y, &io.z, &io.s};
		for (int k = 0; k < 4; ++k) {
			if (!host_out[k] || (k == 1 && !pl.p)) continue;
			cudaPointerAttributes attr;
			if (cudaPointerGetAttributes(&attr, host_out[k]) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer) {
				*dev_out[k] = static_cast<double *>(attr.devicePointer);
				direct[k] = true;
			} else cudaGetLastError(); // (an unregistered pointer is not an error here)
		}
	}
	if (!overlap) {
		for (int i = 0; i < 6; ++i)
			if (widths[i]) CUDA_TRY(cudaMemcpyAsync(B->d_in[i], src[i], n * widths[i] * sizeof(double), cudaMemcpyHostToDevice, st));
		CUDA_TRY(cudaEventRecord(B->ev0, st));
		rc = launch_fused(B, io, opts, st);
		if (rc) return rc;
	} else {
		// The throughput kernel takes QPs in index order from a queue, so it can start on the first rows while
		// the rest is still on its way: the rows go up in chunks on a second stream, each followed by the mark
		// "rows [0, m) have arrived" (d_counter[40]); a warp that has taken QPs beyond the mark waits for it.
		int *ready = B->d_counter + 40;
		io.ready = ready;
		CUDA_TRY(cudaMemsetAsync(ready, 0, sizeof(int), st));
		CUDA_TRY(cudaEventRecord(B->ev_go, st));
		CUDA_TRY(cudaStreamWaitEvent(B->copy_stream, B->ev_go, 0));
		// (the copies are queued BEFORE the launch: a tool that makes launches synchronous must not wait for
		// a kernel whose input is not even on its way)
		// (small first chunks: the kernel starts on the first 256 rows while the rest is on its way)
		const size_t chunk = std::max<size_t>(2048, (n + 63) / 64);
		int k = 0;
		size_t step = 256;
		for (size_t lo = 0; lo < n; lo += step, step = std::min(chunk, step * 2), ++k) {
			const size_t cnt = std::min(step, n - lo);
			for (int i = 0; i < 6; ++i)
				if (widths[i]) CUDA_TRY(cudaMemcpyAsync(B->d_in[i] + lo * widths[i], src[i] + lo * widths[i], cnt * widths[i] * sizeof(double), cudaMemcpyHostToDevice, B->copy_stream));
			B->h_marks[k] = (int)(lo + cnt);
			CUDA_TRY(cudaMemcpyAsync(ready, B->h_marks + k, sizeof(int), cudaMemcpyHostToDevice, B->copy_stream));
		}
		CUDA_TRY(cudaEventRecord(B->ev0, st));
		rc = launch_fused(B, io, opts, st);
		if (rc) return rc;
	}
	CUDA_TRY(cudaEventRecord(B->ev1, st));
	if (x && !direct[0]) CUDA_TRY(cudaMemcpyAsync(x, B->d_x, n * pl.n * sizeof(double), cudaMemcpyDeviceToHost, st));
	if (y && pl.p && !direct[1]) CUDA_TRY(cudaMemcpyAsync(y, B->d_y, n * pl.p * sizeof(double), cudaMemcpyDeviceToHost, st));
	if (z && !direct[2]) CUDA_TRY(cudaMemcpyAsync(z, B->d_z, n * pl.m * sizeof(double), cudaMemcpyDeviceToHost, st));
	if (s && !direct[3]) CUDA_TRY(cudaMemcpyAsync(s, B->d_s, n * pl.m * sizeof(double), cudaMemcpyDeviceToHost, st));
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

extern "C" {

qp_int QPB_DENSE_TO_VALUES_DEVICE(QPBatch *B, qp_int count, int ordering, const qp_real *dP, const qp_real *dA, const qp_real *dG, qp_real *dPpr, qp_real *dApr, qp_real *dGpr, int *dviolation, void *stream)
{
	if (!B || count < 0 || (ordering != ROW_MAJOR_ORDERING && ordering != COLUMN_MAJOR_ORDERING) || !dP || !dG || !dPpr || !dGpr || (B->plan.p && (!dA || !dApr))) return fail(QPB_E_ARG, "QPB_DENSE_TO_VALUES_DEVICE: bad argument", cudaSuccess);
	if (count == 0) return 0;
	CUDA_TRY(cudaSetDevice(B->device));
	int rc = ensure_dense_maps(B, ordering);
	if (rc) return rc;
	const Plan &pl = B->plan;
	const long grid = (long)B->sm_count * 8;
	if (dviolation) CUDA_TRY(cudaMemsetAsync(dviolation, 0, (size_t)count * sizeof(int), (cudaStream_t)stream));
	const long largest = std::max(std::max(pl.n * pl.n, pl.p * pl.n), pl.m * pl.n);
	const long per_launch = std::max(1L, (1L << 31) / std::max(1L, largest)); // cell numbers stay below 2^31
	for (long lo = 0; lo < (long)count; lo += per_launch) {
		const long cnt = std::min(per_launch, (long)count - lo);
		dense_gather_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(dP, dA, dG, B->d_dense_map[0], B->d_dense_map[1], B->d_dense_map[2], (int)pl.n, (int)pl.m, (int)pl.p, (int)pl.nnzP, (int)pl.nnzA, (int)pl.nnzG, dPpr, dApr, dGpr, dviolation, (int)lo, (int)cnt);
		CUDA_TRY(cudaGetLastError());
	}
	return 0;
}

qp_int QP_SOLVE_BATCH_DENSE(QPBatch *B, qp_int count, const settings *opts, int ordering, const qp_real *P, const qp_real *A, const qp_real *G, const qp_real *c, const qp_real *h, const qp_real *b, qp_real *x, qp_real *y, qp_real *z, qp_real *s, qp_int *flag, qp_int *iters, qp_int *violation)
{
	if (!B || !opts || count < 0 || count > B->capacity || !P || !G || !c || !h || (B->plan.p && (!A || !b))) return fail(QPB_E_ARG, "QP_SOLVE_BATCH_DENSE: bad argument", cudaSuccess);
	if (count == 0) return 0;
	std::vector<int> viol((size_t)count, 0);
	{
		std::lock_guard<std::mutex> guard(B->lock);
		CUDA_TRY(cudaSetDevice(B->device));
		int rc = batch_alloc_staging(B);
		if (rc) return rc;
		const Plan &pl = B->plan;
		const size_t cap = (size_t)B->capacity, n = (size_t)count;
		const size_t cells[3] = {(size_t)(pl.n * pl.n), (size_t)(pl.p * pl.n), (size_t)(pl.m * pl.n)};
		const qp_real *src[3] = {P, A, G};
		cudaStream_t st = B->stream;
		for (int k = 0; k < 3; ++k) {
			if (!cells[k]) continue;
			if (!B->d_dense[k]) CUDA_TRY(cudaMalloc(&B->d_dense[k], cap * cells[k] * sizeof(double)));
			CUDA_TRY(cudaMemcpyAsync(B->d_dense[k], src[k], n * cells[k] * sizeof(double), cudaMemcpyHostToDevice, st));
		}
		if (!B->d_violation) CUDA_TRY(cudaMalloc(&B->d_violation, cap * sizeof(int)));
		rc = (int)QPB_DENSE_TO_VALUES_DEVICE(B, count, ordering, B->d_dense[0], B->d_dense[1], B->d_dense[2], B->d_in[0], B->d_in[1], B->d_in[2], B->d_violation, st);
		if (rc) return rc;
		CUDA_TRY(cudaMemcpyAsync(viol.data(), B->d_violation, n * sizeof(int), cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaStreamSynchronize(st));
	}
	if (violation)
		for (qp_int i = 0; i < count; ++i) violation[i] = viol[(size_t)i];
	// the value rows are on the device already: the rest is QP_SOLVE_BATCH without their upload
	return qp_solve_batch_host(B, count, opts, nullptr, nullptr, nullptr, c, h, b, x, y, z, s, flag, iters);
}

qp_int QPB_BATCH_SET_SIGMA_D(QPBatch *B, qp_real sigma_d)
{
	if (!B) return fail(QPB_E_ARG, "QPB_BATCH_SET_SIGMA_D: null batch", cudaSuccess);
	B->sigma_d = sigma_d;
	return 0;
}

// Per-iteration trace of the warp-per-QP kernel for the next host-pointer solves (small batches only: the throughput
// kernel does not trace).  Layout per QP: (maxit + 2) rows of 8 doubles, see BatchIO::trace (kernels.cuh).
qp_int QPB_BATCH_ENABLE_TRACE(QPBatch *B, int on)
{
	if (!B) return fail(QPB_E_ARG, "QPB_BATCH_ENABLE_TRACE: null batch", cudaSuccess);
	if (on && B->capacity >= B->stream_min && B->d_workspace) return fail(QPB_E_ARG, "QPB_BATCH_ENABLE_TRACE: batches of this capacity run the throughput kernel, which keeps no trace", cudaSuccess);
	B->trace_on = on != 0;
	return 0;
}

qp_int QPB_BATCH_READ_TRACE(QPBatch *B, qp_int count, qp_int maxit, qp_real *trace)
{
	if (B && trace && B->single_valid && count == 1 && maxit + 2 == B->single_trace_rows) { // (the single-QP path brought it back with the results)
		memcpy(trace, B->h_single + (B->single_total - (size_t)B->single_trace_rows * 8), (size_t)B->single_trace_rows * 8 * sizeof(double));
		return 0;
	}
	if (!B || !trace || !B->d_trace || count < 0 || count > B->capacity || maxit + 2 != B->trace_rows) return fail(QPB_E_ARG, "QPB_BATCH_READ_TRACE: no trace of that shape", cudaSuccess);
	CUDA_TRY(cudaSetDevice(B->device));
	CUDA_TRY(cudaMemcpy(trace, B->d_trace, (size_t)count * B->trace_rows * 8 * sizeof(double), cudaMemcpyDeviceToHost));
	return 0;
}

// One control tick of a rollout, device side (rollout.cuh): integrate the previous tick's solutions x [robots][30]
// (skipped while *dtick == 0), advance the contact schedule, move the swing feet, rebuild the state rows, then list the
// robots of each contact pattern: dlists [3][robots] (stance, trot, crawl; ascending), dcounts [3]; *dtick += 1.
// The model's constants are fixed here (and mirrored by apf_quadruped_b200/rollout.py, which steps the same model in numpy).
qp_int QPB_ROLLOUT_TICK_DEVICE(int device, qp_int robots, qp_int first_robot, unsigned long long seed, qp_real *dstate, qp_real *drobot, const qp_real *dx, int *dkind, int *dlists, int *dcounts, long long *dtick, void *stream)
{
	if (robots < 1 || !dstate || !drobot || !dx || !dkind || !dlists || !dcounts || !dtick) return fail(QPB_E_ARG, "QPB_ROLLOUT_TICK_DEVICE: bad argument", cudaSuccess);
	if (device >= 0) CUDA_TRY(cudaSetDevice(device));
	RolloutParams P;
	P.dt = 0.0025;            // 400 Hz control loop (main.cpp:1107)
	P.kp_com = 3000.0; P.kd_com = 50.0; // Kcom, Dcom (main.cpp:1498-1499)
	P.g = 9.81; P.mass_g = 21.27 * 9.81; // robot_mass * g_acc (main.cpp:1572; dogbot.urdf)
	P.goal_spread = 0.03; P.sat = 0.05; P.step = 0.025; P.half = 0.5; P.kpa = 1.5; // APF step planner (main.cpp:1331-1390)
	P.kp_sw = 300.0;          // Kp of the swing-foot tracking (main.cpp:1985)
	P.qdd_sat = 40.0; P.kq = 400.0; P.kdq = 40.0; P.a_sat = 30.0; P.replan = 0.25; // the rollout model's joint posture spring-damper and saturations (rollout.cuh)
	P.phase_ticks = 25; P.robot_stride = 7; P.first_robot = (int)first_robot;
	P.seed = seed;
	cudaStream_t st = (cudaStream_t)stream;
	rollout_tick_kernel<<<(unsigned)((robots + 127) / 128), 128, 0, st>>>(dstate, drobot, dx, dkind, dtick, (int)robots, P);
	rollout_lists_kernel<<<1, 1024, 0, st>>>(dkind, dlists, dcounts, dtick, (int)robots);
	CUDA_TRY(cudaGetLastError());
	return 0;
}

// The descriptor table of `kind` for a plan's pattern (host only): desc[i] = op_a | b << 32 for every entry of the value
// rows Ppr | Apr | Gpr | c | h | b, and the 32 constants behind the state row.  Returns the number of descriptors, or
// QPB_E_PATTERN when the pattern is not that kind's.  Lets the tables be checked on the host (tests/test_wbc.py).
qp_int QPB_WBC_DESCRIPTORS(const QPBPlan *plan, int kind, unsigned long long *desc, qp_int capacity, qp_real *consts)
{
	WbcTables T;
	if (!plan || kind < 0 || kind > 2 || !wbc_build_tables(T, kind, plan->plan)) return fail(QPB_E_PATTERN, "QPB_WBC_DESCRIPTORS: not the DogBot pattern of that kind", cudaSuccess);
	for (size_t i = 0; i < T.desc.size() && (qp_int)i < capacity; ++i) desc[i] = (unsigned long long)T.desc[i].op_a | ((unsigned long long)T.desc[i].b << 32);
	if (consts)
		for (int i = 0; i < WBC_SMEM - WBC_STATE; ++i) consts[i] = T.consts[i];
	return (qp_int)T.desc.size();
}

static int ensure_wbc(QPBatch *B, int kind)
{
	if (B->d_wbc_desc && B->wbc.kind == kind) return 0;
	if (B->d_wbc_desc) return fail(QPB_E_ARG, "WBC: this batch's pattern was bound to another contact pattern kind", cudaSuccess);
	if (kind < 0 || kind > 2 || !wbc_build_tables(B->wbc, kind, B->plan)) { B->wbc.kind = -1; return fail(QPB_E_PATTERN, "WBC: the batch's sparsity pattern is not the DogBot pattern of that kind (0 stance, 1 trot, 2 crawl)", cudaSuccess); }
	CUDA_TRY(cudaFuncSetAttribute(wbc_formulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * 8 * WBC_SMEM * sizeof(double) + 8192)));
	std::vector<uint32_t> packed(B->wbc.desc.size());
	for (size_t i = 0; i < packed.size(); ++i) packed[i] = wbc_pack(B->wbc.desc[i]);
	CUDA_TRY(cudaMalloc(&B->d_wbc_desc, packed.size() * sizeof(uint32_t)));
	CUDA_TRY(cudaMemcpy(B->d_wbc_desc, packed.data(), packed.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
	CUDA_TRY(cudaMalloc(&B->d_wbc_consts, sizeof(B->wbc.consts)));
	CUDA_TRY(cudaMemcpy(B->d_wbc_consts, B->wbc.consts, sizeof(B->wbc.consts), cudaMemcpyHostToDevice));
	return 0;
}

qp_int QPB_WBC_FORMULATE_DEVICE(QPBatch *B, int kind, qp_int count, const qp_real *dstate, const int *dindex, const int *dcount_ptr, qp_real *dPpr, qp_real *dApr, qp_real *dGpr, qp_real *dc, qp_real *dh, qp_real *db, void *stream)
{
	if (!B || count < 0 || !dstate || !dPpr || !dApr || !dGpr || !dc || !dh || !db) return fail(QPB_E_ARG, "QPB_WBC_FORMULATE_DEVICE: bad argument", cudaSuccess);
	if (count == 0) return 0;
	CUDA_TRY(cudaSetDevice(B->device));
	const int rc = ensure_wbc(B, kind);
	if (rc) return rc;
	const Plan &pl = B->plan;
	WbcOut out;
	out.Ppr = dPpr; out.Apr = dApr; out.Gpr = dGpr; out.c = dc; out.h = dh; out.b = db;
	out.nnzP = (int)pl.nnzP; out.nnzA = (int)pl.nnzA; out.nnzG = (int)pl.nnzG; out.n = (int)pl.n; out.m = (int)pl.m; out.p = (int)pl.p;
	const int warps = 8;
	const long grid = std::min<long>((count + warps - 1) / warps, (long)B->sm_count * 3);
	const size_t smem = 2 * warps * WBC_SMEM * sizeof(double) + B->wbc.desc.size() * sizeof(unsigned);
	wbc_formulate_kernel<<<(unsigned)grid, warps * 32, smem, (cudaStream_t)stream>>>(dstate, B->d_wbc_desc, B->d_wbc_consts, out, dindex, dcount_ptr, (int)count);
	CUDA_TRY(cudaGetLastError());
	return 0;
}

qp_int QPB_WBC_TORQUES_DEVICE(QPBatch *B, int kind, qp_int count, const qp_real *dstate, const qp_real *dx, qp_real *dtau, const int *dindex, const int *dcount_ptr, void *stream)
{
	if (!B || count < 0 || !dstate || !dx || !dtau) return fail(QPB_E_ARG, "QPB_WBC_TORQUES_DEVICE: bad argument", cudaSuccess);
	if (count == 0) return 0;
	CUDA_TRY(cudaSetDevice(B->device));
	const int rc = ensure_wbc(B, kind);
	if (rc) return rc;
	const int *st = B->wbc.stance_rows;
	const long grid = std::min<long>((count * 12 + 255) / 256, (long)B->sm_count * 8);
	wbc_torque_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(dstate, dx, dtau, B->wbc.nf, make_int4(st[0], st[1], st[2], st[3]), make_int4(st[4], st[5], st[6], st[7]), make_int4(st[8], st[9], st[10], st[11]), dindex, dcount_ptr, (int)count);
	CUDA_TRY(cudaGetLastError());
	return 0;
}

// FP64 peak of the device, measured: every thread runs 8 independent chains of a <- a * b + c, either as one
// DFMA (fused != 0) or as DMUL + DADD, which is what this library's kernels execute (-fmad=false is part of the
// arithmetic contract).  2 flops per step either way.  One CTA of 1024 threads per SM x 2, long enough to dwarf
// the launch.  Returns TFLOP/s, or a negative QPB_E_* code.
double QPB_FP64_PEAK(int device, int fused)
{
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return (double)fail(QPB_E_NODEVICE, "no CUDA device", cudaSuccess);
	if (device >= 0 && cudaSetDevice(device) != cudaSuccess) return (double)fail(QPB_E_CUDA, "cudaSetDevice", cudaGetLastError());
	cudaDeviceProp prop;
	int dev = 0;
	cudaGetDevice(&dev);
	if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return (double)fail(QPB_E_CUDA, "cudaGetDeviceProperties", cudaGetLastError());
	const int grid = prop.multiProcessorCount * 2, block = 1024, steps = 1 << 15;
	double *sink = nullptr;
	if (cudaMalloc(&sink, sizeof(double) * (size_t)grid * block) != cudaSuccess) return (double)fail(QPB_E_CUDA, "cudaMalloc", cudaGetLastError());
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	float best = 1e30f;
	for (int rep = 0; rep < 4; ++rep) {
		cudaEventRecord(e0, 0);
		if (fused) fp64_peak_kernel<true><<<grid, block>>>(sink, steps, 1.0000001, 1e-9);
		else fp64_peak_kernel<false><<<grid, block>>>(sink, steps, 1.0000001, 1e-9);
		cudaEventRecord(e1, 0);
		cudaEventSynchronize(e1);
		float ms = 0.f;
		cudaEventElapsedTime(&ms, e0, e1);
		if (rep > 0 && ms < best) best = ms;
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(sink);
	if (cudaGetLastError() != cudaSuccess) return (double)fail(QPB_E_CUDA, "fp64_peak_kernel", cudaGetLastError());
	return 2.0 * 8.0 * (double)steps * (double)grid * block / (best * 1e-3) / 1e12;
}

void QPB_BATCH_ROW_WIDTHS(const QPBatch *B, qp_int *nnz)
{
	if (!B || !nnz) return;
	nnz[0] = B->plan.nnzP; nnz[1] = B->plan.nnzA; nnz[2] = B->plan.nnzG;
}

double QPB_BATCH_LAST_KERNEL_MS(const QPBatch *B) { return B ? B->last_kernel_ms : 0.0; }
qp_int QPB_BATCH_LAST_LAUNCHES(const QPBatch *B) { return B ? B->last_launches : 0; }
int QPB_BATCH_LAST_KERNEL(const QPBatch *B) { return B ? B->last_kernel : -1; }
qp_int QPB_GENERATED_KERNELS(unsigned long long *keys, int capacity) { return list_generated_kernels(keys, keys ? capacity : 0); }

// Host access to fval / stat of the last host-pointer solve (used by QP_SOLVE to fill `stats`).
qp_int QPB_BATCH_READ_STATS(QPBatch *B, qp_int count, qp_real *fval, qp_real *stat)
{
	if (B && B->single_valid && count == 1) {
		const double *st = B->h_single + (B->single_total - (size_t)B->single_trace_rows * 8 - 12); // fval(2) | stat[8] | flag, iters
		if (fval) fval[0] = st[0];
		if (stat) memcpy(stat, st + 2, 8 * sizeof(double));
		return 0;
	}
	if (!B || !B->d_fval) return fail(QPB_E_ARG, "QPB_BATCH_READ_STATS: nothing solved yet", cudaSuccess);
	if (count < 0 || count > B->capacity) return fail(QPB_E_ARG, "QPB_BATCH_READ_STATS: count > capacity", cudaSuccess);
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
	const size_t smem = (size_t)B->program.lay.total * sizeof(double);
	cudaStream_t st = (cudaStream_t)stream;
	CUDA_TRY(cudaMemsetAsync(B->d_counter + 16, 0, sizeof(int), st));
	int occ = 0;
	CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ldl_numeric_kernel, 32, smem));
	long grid = (long)B->sm_count * (occ > 0 ? occ : 1);
	if (grid > count) grid = count;
	ldl_numeric_kernel<<<(unsigned)grid, 32, smem, st>>>(B->d_plan, B->pview, B->program.lay, dvals, dzdiag, dLx, dD, (int)count, B->d_counter + 16);
	CUDA_TRY(cudaGetLastError());
	return 0;
}

qp_int QPB_LDL_SOLVE_BATCH(QPBatch *B, qp_int count, const qp_real *dLx, const qp_real *dD, const qp_real *drhs, qp_real *dsol, void *stream)
{
	if (!B || count < 0 || !dLx || !dD || !drhs || !dsol) return fail(QPB_E_ARG, "QPB_LDL_SOLVE_BATCH: bad argument", cudaSuccess);
	if (count == 0) return 0;
	CUDA_TRY(cudaSetDevice(B->device));
	const size_t smem = (size_t)B->program.lay.total * sizeof(double);
	cudaStream_t st = (cudaStream_t)stream;
	CUDA_TRY(cudaMemsetAsync(B->d_counter + 32, 0, sizeof(int), st));
	int occ = 0;
	CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ldl_solve_kernel, 32, smem));
	long grid = (long)B->sm_count * (occ > 0 ? occ : 1);
	if (grid > count) grid = count;
	ldl_solve_kernel<<<(unsigned)grid, 32, smem, st>>>(B->d_plan, B->pview, B->program.lay, dLx, dD, drhs, dsol, (int)count, B->d_counter + 32);
	CUDA_TRY(cudaGetLastError());
	return 0;
}

} // extern "C"

// Multi-GPU split of a batch behind the C ABI (SURVEY 8 e, 8 b): QP_SETUP_BATCH_MULTI / QP_SOLVE_BATCH_MULTI.
//
// QPs are independent (the reference solves one per control tick, qpSWIFT.c:473-644), so the batch is cut
// into contiguous shards, one per device, and nothing crosses devices while they are solved:
//   * host-pointer entry: every device gets its shard of the caller's arrays (pointer + offset) and writes
//     its results into its shard of the caller's result arrays -- no collective at all;
//   * device-pointer entry: every device solves the shard that already lies in its memory, then the
//     solutions, exit flags and iteration counts are all-gathered with ncclAllGather over one communicator
//     per device (ncclCommInitAll: one process drives all GPUs) -- the only collective north_star names.
// One host thread per device drives its stream; per-QP arithmetic does not depend on the shard boundaries,
// so results are bit-identical for 1, 2, 4 or 8 devices.  NCCL is loaded on first use (dlopen libnccl.so.2):
// a process that never gathers on the device never needs it.
#include <dlfcn.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "qpswift_b200.h"

namespace {

// ---- the few NCCL entry points used, bound at run time (types as in nccl.h) -------------------------
typedef struct ncclComm *ncclComm_t;
typedef int ncclResult_t;
enum { NCCL_INT32 = 2, NCCL_FLOAT64 = 8 };
struct Nccl
{
	void *lib = nullptr;
	ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
	bool load(std::string &err)
	{
		if (lib) return true;
		for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
			lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
			if (lib) break;
		}
		if (!lib) { err = std::string("NCCL not found: ") + dlerror(); return false; }
		CommInitAll = reinterpret_cast<decltype(CommInitAll)>(dlsym(lib, "ncclCommInitAll"));
		CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
		AllGather = reinterpret_cast<decltype(AllGather)>(dlsym(lib, "ncclAllGather"));
		GroupStart = reinterpret_cast<decltype(GroupStart)>(dlsym(lib, "ncclGroupStart"));
		GroupEnd = reinterpret_cast<decltype(GroupEnd)>(dlsym(lib, "ncclGroupEnd"));
		GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
		if (!CommInitAll || !CommDestroy || !AllGather || !GroupStart || !GroupEnd || !GetErrorString) { err = "NCCL: missing symbols"; return false; }
		return true;
	}
};
Nccl g_nccl;
thread_local std::string g_multi_error;

} // namespace

struct QPBatchMulti
{
	std::vector<int> devices;
	std::vector<QPBatch *> shards;
	std::vector<cudaStream_t> streams;
	std::vector<ncclComm_t> comms; // created by the first device-pointer gather
	qp_int capacity = 0, per_shard = 0;
	qp_int n = 0, m = 0, p = 0;
};

extern "C" {

const char *QPB_MULTI_LAST_ERROR(void) { return g_multi_error.c_str(); }

void QP_CLEANUP_BATCH_MULTI(QPBatchMulti *M)
{
	if (!M) return;
	for (size_t d = 0; d < M->comms.size(); ++d)
		if (M->comms[d]) g_nccl.CommDestroy(M->comms[d]);
	for (size_t d = 0; d < M->shards.size(); ++d) {
		if (d < M->streams.size() && M->streams[d]) {
			cudaSetDevice(M->devices[d]);
			cudaStreamDestroy(M->streams[d]);
		}
		QP_CLEANUP_BATCH(M->shards[d]);
	}
	delete M;
}

static QPBatchMulti *multi_create(const QPBPlan *plan, const QP *proto, qp_real sigma_d, qp_int capacity, const int *devices, int ndev)
{
	if ((!plan && !proto) || capacity < 1 || ndev < 1 || !devices) { g_multi_error = "QP_SETUP_BATCH_MULTI: bad argument"; return nullptr; }
	const int have = QPB_DEVICE_COUNT();
	if (have < 1) { g_multi_error = "QP_SETUP_BATCH_MULTI: no CUDA device (there is no CPU fallback)"; return nullptr; }
	for (int d = 0; d < ndev; ++d)
		if (devices[d] < 0 || devices[d] >= have) { g_multi_error = "QP_SETUP_BATCH_MULTI: no such device"; return nullptr; }
	QPBatchMulti *M = new QPBatchMulti;
	M->devices.assign(devices, devices + ndev);
	M->capacity = capacity;
	M->per_shard = (capacity + ndev - 1) / ndev;
	M->streams.assign(ndev, nullptr);
	for (int d = 0; d < ndev; ++d) {
		QPBatch *B = plan ? QPB_BATCH_CREATE(plan, sigma_d, M->per_shard, devices[d]) : QP_SETUP_BATCH(proto, M->per_shard, devices[d]);
		if (!B) { g_multi_error = std::string("QP_SETUP_BATCH_MULTI: ") + QPB_LAST_ERROR(); QP_CLEANUP_BATCH_MULTI(M); return nullptr; }
		M->shards.push_back(B);
		cudaSetDevice(devices[d]);
		if (cudaStreamCreateWithFlags(&M->streams[d], cudaStreamNonBlocking) != cudaSuccess) { g_multi_error = "QP_SETUP_BATCH_MULTI: cudaStreamCreate failed"; QP_CLEANUP_BATCH_MULTI(M); return nullptr; }
	}
	return M;
}

QPBatchMulti *QP_SETUP_BATCH_MULTI(const QP *proto, qp_int capacity, const int *devices, int ndev)
{
	if (!proto) { g_multi_error = "QP_SETUP_BATCH_MULTI: null prototype"; return nullptr; }
	QPBatchMulti *M = multi_create(nullptr, proto, proto->sigma_d, capacity, devices, ndev);
	if (M) { M->n = proto->n; M->m = proto->m; M->p = proto->p; }
	return M;
}

QPBatchMulti *QPB_BATCH_CREATE_MULTI(const QPBPlan *plan, qp_int n, qp_int m, qp_int p, qp_real sigma_d, qp_int capacity, const int *devices, int ndev)
{
	if (!plan) { g_multi_error = "QPB_BATCH_CREATE_MULTI: null plan"; return nullptr; }
	QPBatchMulti *M = multi_create(plan, nullptr, sigma_d, capacity, devices, ndev);
	if (M) { M->n = n; M->m = m; M->p = p; }
	return M;
}

int QPB_MULTI_DEVICES(const QPBatchMulti *M) { return M ? (int)M->devices.size() : 0; }

// Shard d of a batch of `count`: QPs [lo, hi).  Contiguous; the first count % ndev shards hold one more.
void QPB_MULTI_SHARD(const QPBatchMulti *M, qp_int count, int shard, qp_int *lo, qp_int *hi)
{
	const qp_int nd = M ? (qp_int)M->devices.size() : 1;
	const qp_int base = count / nd, extra = count % nd;
	const qp_int a = shard * base + std::min<qp_int>(shard, extra);
	if (lo) *lo = a;
	if (hi) *hi = a + base + (shard < extra ? 1 : 0);
}

qp_int QP_SOLVE_BATCH_MULTI(QPBatchMulti *M, qp_int count, const settings *opts, const qp_real *Ppr, const qp_real *Apr, const qp_real *Gpr, const qp_real *c, const qp_real *h, const qp_real *b, qp_real *x, qp_real *y, qp_real *z, qp_real *s, qp_int *flag, qp_int *iters)
{
	if (!M || !opts || count < 0 || count > M->capacity) { g_multi_error = "QP_SOLVE_BATCH_MULTI: bad argument (count > capacity)"; return QPB_E_ARG; }
	const int nd = (int)M->devices.size();
	std::vector<qp_int> rc(nd, 0);
	std::vector<std::string> err(nd);
	std::vector<std::thread> workers;
	// (row widths of the value arrays: every shard starts at its first QP's rows)
	qp_int nnz[3];
	QPB_BATCH_ROW_WIDTHS(M->shards[0], nnz);
	for (int d = 0; d < nd; ++d) {
		workers.emplace_back([&, d]() {
			qp_int lo, hi;
			QPB_MULTI_SHARD(M, count, d, &lo, &hi);
			if (hi <= lo) return;
			auto at = [lo](const qp_real *ptr, qp_int w) { return ptr ? ptr + (size_t)lo * w : nullptr; };
			auto out = [lo](qp_real *ptr, qp_int w) { return ptr ? ptr + (size_t)lo * w : nullptr; };
			rc[d] = QP_SOLVE_BATCH(M->shards[d], hi - lo, opts, at(Ppr, nnz[0]), at(Apr, nnz[1]), at(Gpr, nnz[2]), at(c, M->n), at(h, M->m), at(b, M->p),
			                       out(x, M->n), out(y, M->p), out(z, M->m), out(s, M->m), flag ? flag + lo : nullptr, iters ? iters + lo : nullptr);
			if (rc[d] < 0) err[d] = QPB_LAST_ERROR();
		});
	}
	for (auto &w : workers) w.join();
	qp_int bad = 0;
	for (int d = 0; d < nd; ++d) {
		if (rc[d] < 0) { g_multi_error = "QP_SOLVE_BATCH_MULTI: device " + std::to_string(M->devices[d]) + ": " + err[d]; return rc[d]; }
		bad += rc[d];
	}
	return bad;
}

// Device-pointer entry.  Arrays of ndev pointers: entry d points at shard d's rows in device d's memory (count_d =
// hi - lo of QPB_MULTI_SHARD).  gx / gflag / giters (each may be NULL; entries in device d's memory) receive the
// all-gather of x [ndev][most][n] and of the int32 flags / iteration counts [ndev][most], most = the largest shard
// (= QPB_MULTI_SHARD of shard 0); rows beyond a shard's count are padding.  Returns 0 or a negative QPB_E_* code;
// synchronises every device before it returns.
qp_int QP_SOLVE_BATCH_MULTI_DEVICE(QPBatchMulti *M, qp_int count, const settings *opts, const qp_real *const *dPpr, const qp_real *const *dApr, const qp_real *const *dGpr, const qp_real *const *dc, const qp_real *const *dh, const qp_real *const *db, qp_real *const *dx, qp_real *const *dy, qp_real *const *dz, qp_real *const *ds, int *const *dflag, int *const *diters, qp_real *const *gx, int *const *gflag, int *const *giters)
{
	if (!M || !opts || count < 0 || count > M->capacity || !dPpr || !dGpr || !dc || !dh || !dx || !dz || !ds || !dflag || !diters) { g_multi_error = "QP_SOLVE_BATCH_MULTI_DEVICE: bad argument"; return QPB_E_ARG; }
	const int nd = (int)M->devices.size();
	const bool gather = gx || gflag || giters;
	if (gather && M->comms.empty()) {
		std::vector<int> sorted(M->devices);
		std::sort(sorted.begin(), sorted.end());
		if (std::adjacent_find(sorted.begin(), sorted.end()) != sorted.end()) { g_multi_error = "QP_SOLVE_BATCH_MULTI_DEVICE: NCCL cannot gather between shards on the same device"; return QPB_E_ARG; }
		if (!g_nccl.load(g_multi_error)) return QPB_E_CUDA;
		M->comms.assign(nd, nullptr);
		const ncclResult_t r = g_nccl.CommInitAll(M->comms.data(), nd, M->devices.data());
		if (r != 0) { g_multi_error = std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(r); M->comms.clear(); return QPB_E_CUDA; }
	}
	for (int d = 0; d < nd; ++d) {
		qp_int lo, hi;
		QPB_MULTI_SHARD(M, count, d, &lo, &hi);
		if (hi <= lo) continue;
		const qp_int rc = QP_SOLVE_BATCH_DEVICE(M->shards[d], hi - lo, opts, dPpr[d], dApr ? dApr[d] : nullptr, dGpr[d], dc[d], dh[d], db ? db[d] : nullptr,
		                                        dx[d], dy ? dy[d] : nullptr, dz[d], ds[d], dflag[d], diters[d], nullptr, nullptr, M->streams[d]);
		if (rc < 0) { g_multi_error = std::string("QP_SOLVE_BATCH_MULTI_DEVICE: ") + QPB_LAST_ERROR(); return rc; }
	}
	if (gather) {
		qp_int lo0, hi0;
		QPB_MULTI_SHARD(M, count, 0, &lo0, &hi0);
		const size_t most = (size_t)(hi0 - lo0);
		g_nccl.GroupStart();
		for (int d = 0; d < nd; ++d) {
			cudaSetDevice(M->devices[d]);
			ncclResult_t r = 0;
			if (gx) r = g_nccl.AllGather(dx[d], gx[d], most * (size_t)M->n, NCCL_FLOAT64, M->comms[d], M->streams[d]);
			if (!r && gflag) r = g_nccl.AllGather(dflag[d], gflag[d], most, NCCL_INT32, M->comms[d], M->streams[d]);
			if (!r && giters) r = g_nccl.AllGather(diters[d], giters[d], most, NCCL_INT32, M->comms[d], M->streams[d]);
			if (r) { g_nccl.GroupEnd(); g_multi_error = std::string("ncclAllGather: ") + g_nccl.GetErrorString(r); return QPB_E_CUDA; }
		}
		const ncclResult_t r = g_nccl.GroupEnd();
		if (r) { g_multi_error = std::string("ncclGroupEnd: ") + g_nccl.GetErrorString(r); return QPB_E_CUDA; }
	}
	for (int d = 0; d < nd; ++d) {
		cudaSetDevice(M->devices[d]);
		const cudaError_t e = cudaStreamSynchronize(M->streams[d]);
		if (e != cudaSuccess) { g_multi_error = std::string("QP_SOLVE_BATCH_MULTI_DEVICE: ") + cudaGetErrorString(e); return QPB_E_CUDA; }
	}
	return 0;
}

} // extern "C"

// qpSWIFT's single-problem C API (QP_SETUP / QP_SETUP_dense / QP_SOLVE / QP_CLEANUP[_dense]) on top
// of the batched B200 solver: the host does the integer analysis, the GPU does all arithmetic.
//
// Behaviour mirrored from dogbot_controller/src/qpSWIFT/qpSWIFT.c:
//   QP_SETUP 60-234, QP_SETUP_dense 260-456, QP_SOLVE 473-644, QP_CLEANUP 661-737,
//   QP_CLEANUP_dense 754-839; densetosparse Auxilary.c:1154-1273; SparseMatrixTranspose 901-951.
// Differences, all deliberate (DESIGN.md "boundary"):
//   * the initial point (kkt_initialize) is computed by the solve kernel, not at setup;
//   * QP_SOLVE always restarts from that initial point (the reference's second call on the same
//     QP is undefined: IterationCount is never reset, qpSWIFT.c:596);
//   * plans are cached per sparsity pattern (at most PLAN_CACHE_SIZE of them, least recently used
//     out first), so a controller that calls QP_SETUP_dense every tick (main.cpp:1649) pays for AMD +
//     symbolic analysis once per pattern; a pattern change (an entry that becomes exactly 0.0) costs
//     one analysis + program compilation + cudaMalloc, i.e. a latency spike of a few milliseconds;
//   * what QP_SOLVE needs beyond the public structs hangs off the QP itself (a private tail behind
//     its kkt struct): there is no table of live QPs, and a controller that never calls QP_CLEANUP
//     (the reference controller does not) leaks exactly what it leaks with the reference -- its own
//     QP -- and nothing inside the library grows;
//   * without a usable CUDA device QP_SOLVE returns QP_FATAL; there is no CPU solver here.

#include "capi_internal.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace {

using qpb::Plan;

struct CachedPlan
{
	QPBPlan handle;
	std::vector<long> key; // the exact pattern this plan was built for (collision check)
	uint64_t hash = 0;
	QPBatch *single = nullptr;
	bool single_failed = false;
	std::mutex solve_lock; // one QP_SOLVE at a time on the shared single-QP batch (sigma_d, solve, statistics)
	unsigned long long last_use = 0;
	~CachedPlan() { if (single) QP_CLEANUP_BATCH(single); }
};

const size_t PLAN_CACHE_SIZE = 16;
std::mutex g_lock;
std::vector<std::shared_ptr<CachedPlan>> g_plans; // bounded; a QP keeps its plan alive on its own
unsigned long long g_clock = 0;

// What QP_SOLVE needs to find again, allocated in one block with the public kkt struct (which comes first, so
// that qp->kkt points at it): no global table of live QPs.
struct KktBlock
{
	kkt pub;
	uint64_t magic;
	std::shared_ptr<CachedPlan> plan;
	bool dense;
};
const uint64_t KKT_MAGIC = 0x51505342323030ull; // "QPSB200"

KktBlock *block_of(const QP *qp)
{
	if (!qp || !qp->kkt) return nullptr;
	KktBlock *b = reinterpret_cast<KktBlock *>(qp->kkt);
	return b->magic == KKT_MAGIC ? b : nullptr;
}

template <class T>
T *alloc(size_t count) { return (T *)malloc((count ? count : 1) * sizeof(T)); }

void push(std::vector<long> &key, const qp_int *a, long count) { key.insert(key.end(), a, a + count); }

std::shared_ptr<CachedPlan> plan_for(qp_int n, qp_int m, qp_int p, const qp_int *Pjc, const qp_int *Pir, const qp_int *Ajc, const qp_int *Air, const qp_int *Gjc, const qp_int *Gir, const qp_int *Permut)
{
	std::vector<long> key;
	key.push_back(n); key.push_back(m); key.push_back(p); key.push_back(Permut ? 1 : 0);
	push(key, Pjc, n + 1); push(key, Pir, Pjc[n]);
	if (p) { push(key, Ajc, n + 1); push(key, Air, Ajc[n]); }
	push(key, Gjc, n + 1); push(key, Gir, Gjc[n]);
	if (Permut) push(key, Permut, n + m + p);
	const uint64_t h = qpb::fnv1a_words(key.data(), (long)key.size());
	std::lock_guard<std::mutex> guard(g_lock);
	for (auto &cand : g_plans)
		if (cand->hash == h && cand->key == key) { cand->last_use = ++g_clock; return cand; }
	auto made = std::make_shared<CachedPlan>();
	qpb_build(made->handle, n, m, p, Pjc, Pir, p ? Ajc : nullptr, p ? Air : nullptr, Gjc, Gir, Permut);
	made->key.swap(key);
	made->hash = h;
	made->last_use = ++g_clock;
	if (g_plans.size() >= PLAN_CACHE_SIZE) { // least recently used out; QPs that still hold it keep it alive
		size_t oldest = 0;
		for (size_t i = 1; i < g_plans.size(); ++i)
			if (g_plans[i]->last_use < g_plans[oldest]->last_use) oldest = i;
		g_plans.erase(g_plans.begin() + (long)oldest);
	}
	g_plans.push_back(made);
	return made;
}

// dense -> CCS dropping exact zeros (Auxilary.c:1154-1206 column-major, 1219-1273 row-major)
smat *dense_to_ccs(qp_int rows, qp_int cols, const qp_real *a, int ordering)
{
	smat *M = alloc<smat>(1);
	const bool colmajor = ordering == COLUMN_MAJOR_ORDERING;
	auto at = [&](qp_int i, qp_int j) { return colmajor ? a[j * rows + i] : a[i * cols + j]; };
	qp_int nz = 0;
	for (qp_int j = 0; j < cols; ++j)
		for (qp_int i = 0; i < rows; ++i)
			if (at(i, j) != 0.0) ++nz;
	M->jc = alloc<qp_int>(cols + 1);
	M->ir = alloc<qp_int>(nz);
	M->pr = alloc<qp_real>(nz);
	qp_int k = 0;
	M->jc[0] = 0;
	for (qp_int j = 0; j < cols; ++j) {
		for (qp_int i = 0; i < rows; ++i) {
			const qp_real v = at(i, j);
			if (v != 0.0) { M->ir[k] = i; M->pr[k] = v; ++k; }
		}
		M->jc[j + 1] = k;
	}
	M->m = rows; M->n = cols; M->nnz = nz;
	return M;
}

smat *wrap_ccs(qp_int rows, qp_int cols, qp_int *jc, qp_int *ir, qp_real *pr)
{
	smat *M = alloc<smat>(1);
	M->jc = jc; M->ir = ir; M->pr = pr; M->m = rows; M->n = cols; M->nnz = jc[cols];
	return M;
}

smat *transpose_of(const smat *A)
{
	smat *T = alloc<smat>(1);
	T->m = A->n; T->n = A->m; T->nnz = A->nnz;
	T->jc = alloc<qp_int>(A->m + 1);
	T->ir = alloc<qp_int>(A->nnz);
	T->pr = alloc<qp_real>(A->nnz);
	std::vector<qp_int> cnt(A->m, 0);
	for (qp_int q = 0; q < A->nnz; ++q) ++cnt[A->ir[q]];
	T->jc[0] = 0;
	for (qp_int r = 0; r < A->m; ++r) { T->jc[r + 1] = T->jc[r] + cnt[r]; cnt[r] = 0; }
	for (qp_int c = 0; c < A->n; ++c)
		for (qp_int q = A->jc[c]; q < A->jc[c + 1]; ++q) {
			const qp_int r = A->ir[q], at = T->jc[r] + cnt[r]++;
			T->ir[at] = c;
			T->pr[at] = A->pr[q];
		}
	return T;
}

qp_int *copy_longs(const std::vector<long> &v)
{
	qp_int *a = alloc<qp_int>(v.size());
	if (!v.empty()) memcpy(a, v.data(), v.size() * sizeof(qp_int));
	return a;
}

QP *setup_common(qp_int n, qp_int m, qp_int p, smat *P, smat *A, smat *G, qp_real *c, qp_real *h, qp_real *b, qp_real sigma_d, qp_int *Permut, bool dense)
{
	const auto t0 = std::chrono::steady_clock::now();
	QP *qp = (QP *)calloc(1, sizeof(QP));
	qp->options = alloc<settings>(1);
	qp->options->abstol = ABSTOL;
	qp->options->reltol = RELTOL;
	qp->options->maxit = MAXIT;
	qp->options->sigma = SIGMA;
	qp->options->verbose = VERBOSE;
	qp->stats = (stats *)calloc(1, sizeof(stats));
	qp->stats->Flag = QP_FATAL;
	qp->n = n; qp->m = m; qp->p = A ? p : 0;
	qp->P = P; qp->G = G; qp->A = A;
	qp->c = c; qp->h = h; qp->b = A ? b : nullptr;
	qp->sigma_d = sigma_d;
	p = qp->p;
	const qp_int N = n + m + p;
	qp->At = A ? transpose_of(A) : nullptr;
	qp->Gt = transpose_of(G);
	qp->x = alloc<qp_real>(n); qp->z = alloc<qp_real>(m); qp->s = alloc<qp_real>(m);
	qp->rx = alloc<qp_real>(n); qp->rz = alloc<qp_real>(m);
	qp->delta = alloc<qp_real>(N); qp->delta_x = alloc<qp_real>(n); qp->delta_z = alloc<qp_real>(m); qp->delta_s = alloc<qp_real>(m);
	qp->ds = alloc<qp_real>(m); qp->lambda = alloc<qp_real>(m); qp->temp = alloc<qp_real>(n);
	if (A) { qp->y = alloc<qp_real>(p); qp->ry = alloc<qp_real>(p); qp->delta_y = alloc<qp_real>(p); }

	auto cached = plan_for(n, m, p, P->jc, P->ir, A ? A->jc : nullptr, A ? A->ir : nullptr, G->jc, G->ir, Permut);
	const Plan &pl = cached->handle.plan;

	KktBlock *blk = new KktBlock();
	memset(&blk->pub, 0, sizeof(kkt));
	blk->magic = KKT_MAGIC;
	blk->plan = cached;
	blk->dense = dense;
	kkt *k = &blk->pub;
	qp->kkt = k;
	k->kktmatrix = alloc<smat>(1);
	k->kktmatrix->m = k->kktmatrix->n = N;
	k->kktmatrix->nnz = (qp_int)pl.Kir.size();
	k->kktmatrix->jc = copy_longs(pl.Kjc);
	k->kktmatrix->ir = copy_longs(pl.Kir);
	k->kktmatrix->pr = alloc<qp_real>(pl.Kir.size());
	{
		const long nV = pl.nnzP + pl.nnzA + pl.nnzG;
		for (size_t q = 0; q < pl.Ksrc.size(); ++q) {
			const long src = pl.Ksrc[q];
			qp_real v;
			if (src >= nV) v = -1.0;
			else if (src < pl.nnzP) v = P->pr[src];
			else if (src < pl.nnzP + pl.nnzA) v = A->pr[src - pl.nnzP];
			else v = G->pr[src - pl.nnzP - pl.nnzA];
			k->kktmatrix->pr[q] = v;
		}
	}
	k->b = alloc<qp_real>(N);
	if (Permut) { k->P = Permut; qp->stats->AMD_RESULT = -3; }
	else { k->P = copy_longs(pl.P); qp->stats->AMD_RESULT = pl.amd_status; }
	k->Pinv = copy_longs(pl.Pinv);
	k->UPattern = alloc<qp_int>(N);
	k->Parent = copy_longs(pl.Parent);
	k->Lnz = copy_longs(pl.Lnz);
	k->Lp = copy_longs(pl.Lp);
	k->Li = copy_longs(pl.Li);
	k->Ltp = copy_longs(pl.Ltp);
	k->Lti = copy_longs(pl.Lti);
	k->Flag = alloc<qp_int>(N);
	k->Pattern = alloc<qp_int>(N);
	k->Y = alloc<qp_real>(N);
	k->D = alloc<qp_real>(N);
	k->Lx = alloc<qp_real>(pl.Li.size() + 1);

	qp->stats->ldl_numeric = 0.0;
	qp->stats->tsetup = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	return qp;
}

void free_smat_owned(smat *M)
{
	if (!M) return;
	free(M->jc); free(M->ir); free(M->pr); free(M);
}

void cleanup_common(QP *qp, bool dense)
{
	if (!qp) return;
	free(qp->x); free(qp->y); free(qp->z); free(qp->s);
	free(qp->rx); free(qp->ry); free(qp->rz);
	free(qp->delta); free(qp->delta_x); free(qp->delta_y); free(qp->delta_z); free(qp->delta_s);
	free(qp->ds); free(qp->lambda); free(qp->temp);
	if (dense) { free_smat_owned(qp->P); free_smat_owned(qp->G); free_smat_owned(qp->A); }
	else { free(qp->P); free(qp->G); free(qp->A); }
	free_smat_owned(qp->At);
	free_smat_owned(qp->Gt);
	kkt *k = qp->kkt;
	free_smat_owned(k->kktmatrix);
	free(k->b); free(k->Parent); free(k->Flag); free(k->Lnz); free(k->Li); free(k->Lp); free(k->Lti); free(k->Ltp);
	free(k->Pattern); free(k->UPattern); free(k->Y); free(k->Lx); free(k->D); free(k->Pinv);
	if (qp->stats->AMD_RESULT != -3) free(k->P);
	if (KktBlock *blk = block_of(qp)) delete blk; // (drops the QP's reference to its plan)
	else free(k);
	free(qp->options);
	free(qp->stats);
	free(qp);
}

} // namespace

extern "C" {

QP *QP_SETUP(qp_int n, qp_int m, qp_int p, qp_int *Pjc, qp_int *Pir, qp_real *Ppr, qp_int *Ajc, qp_int *Air, qp_real *Apr, qp_int *Gjc, qp_int *Gir, qp_real *Gpr, qp_real *c, qp_real *h, qp_real *b, qp_real sigma_d, qp_int *Permut)
{
	smat *A = (Apr && Ajc && Air && b && p != 0) ? wrap_ccs(p, n, Ajc, Air, Apr) : nullptr;
	return setup_common(n, m, p, wrap_ccs(n, n, Pjc, Pir, Ppr), A, wrap_ccs(m, n, Gjc, Gir, Gpr), c, h, b, sigma_d, Permut, false);
}

QP *QP_SETUP_dense(qp_int n, qp_int m, qp_int p, qp_real *Ppr, qp_real *Apr, qp_real *Gpr, qp_real *c, qp_real *h, qp_real *b, qp_int *Permut, int ordering)
{
	smat *A = (Apr && b && p != 0) ? dense_to_ccs(p, n, Apr, ordering) : nullptr;
	return setup_common(n, m, p, dense_to_ccs(n, n, Ppr, ordering), A, dense_to_ccs(m, n, Gpr, ordering), c, h, b, 0.0, Permut, true);
}

qp_int QP_SOLVE(QP *qp)
{
	const auto t0 = std::chrono::steady_clock::now();
	if (qp->options->verbose > 0) {
		printf("****qpSWIFT : Sparse Quadratic Programming Solver****\n\n");
		printf("================Data Statistics======================\n");
	}
	KktBlock *blk = block_of(qp);
	std::shared_ptr<CachedPlan> cached = blk ? blk->plan : nullptr;
	qp->stats->Flag = QP_FATAL;
	qp_int rc = QPB_E_ARG;
	if (cached && cached->handle.plan.valid) {
		// One lock around sigma_d, the solve and its statistics: QPs of one pattern share the single-QP batch, and
		// the reference allows independent QPs to be solved from different threads.
		std::lock_guard<std::mutex> solving(cached->solve_lock);
		if (!cached->single && !cached->single_failed) {
			cached->single = QPB_BATCH_CREATE(&cached->handle, qp->sigma_d, 1, -1);
			if (!cached->single) cached->single_failed = true;
		}
		if (cached->single) {
			qp_int flag = QP_FATAL, iters = 0;
			const bool verbose = qp->options->verbose > 0;
			// sigma_d belongs to the QP, not to the cached plan
			rc = QPB_BATCH_SET_SIGMA_D(cached->single, qp->sigma_d);
			QPB_BATCH_ENABLE_TRACE(cached->single, 1); // per-iteration lines (verbose) and stats->ldl_numeric
			rc = QP_SOLVE_BATCH(cached->single, 1, qp->options, qp->P->pr, qp->A ? qp->A->pr : nullptr, qp->G->pr, qp->c, qp->h, qp->b, qp->x, qp->y, qp->z, qp->s, &flag, &iters);
			if (rc >= 0) {
				qp_real fval = 0.0, st[8];
				if (QPB_BATCH_READ_STATS(cached->single, 1, &fval, st) == 0) {
					qp->stats->fval = fval;
					qp->stats->n_rx = st[0]; qp->stats->n_ry = st[1]; qp->stats->n_rz = st[2]; qp->stats->n_mu = st[3];
					qp->stats->alpha_p = st[4]; qp->stats->alpha_d = st[5];
					qp->options->sigma = st[6];
					qp->mu = st[7];
				}
				qp->stats->Flag = flag;
				qp->stats->IterationCount = iters;
				qp->stats->kkt_time = QPB_BATCH_LAST_KERNEL_MS(cached->single) * 1e-3;
				qp->stats->resolve_kkt = 0;
				std::vector<qp_real> trace((size_t)(qp->options->maxit + 2) * 8, 0.0);
				if (QPB_BATCH_READ_TRACE(cached->single, 1, qp->options->maxit, trace.data()) == 0) {
					qp->stats->ldl_numeric += trace[0]; // time inside LDL_numeric (Auxilary.c:476-484), device clock
					if (verbose) {
						// the reference's per-iteration lines (qpSWIFT.c:513-516, 598-601), printed after the fact
						const qp_int shown = flag == QP_OPTIMAL ? iters + 1 : iters;
						for (qp_int it = 0; it < shown; ++it) {
							const qp_real *row = &trace[(size_t)(1 + it) * 8];
							printf("It: %ld || pcost : %e || rx:%e   ||  ry:%e ||  rz:%e || mu:%e\n", it, row[0], row[1], row[2], row[3], row[4]);
							if (it < iters) printf("      || Primal Step Size : %f || Dual Step Size   : %f\n", row[5], row[6]);
						}
					}
				}
			}
		}
	}
	if (rc < 0 && qp->options->verbose > 0) printf("qpSWIFT-B200: %s\n", QPB_LAST_ERROR());
	qp->stats->tsolve = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	if (qp->options->verbose > 0) {
		if (qp->stats->Flag == QP_OPTIMAL) {
			printf("\nOptimal Solution Found\n");
			printf("Solve Time     : %f ms\n", (qp->stats->tsolve + qp->stats->tsetup) * 1000.0);
			printf("KKT_Solve Time : %f ms\n", qp->stats->kkt_time * 1000.0);
			printf("LDL Time       : %f ms\n", qp->stats->ldl_numeric * 1000.0);
			printf("Iterations     : %ld\n\n", qp->stats->IterationCount);
		}
		if (qp->stats->Flag == QP_MAXIT) {
			printf("\nMaximum Iterations reached\n");
			printf("Solve Time     : %f ms\n", qp->stats->tsolve * 1000.0);
			printf("KKT_Solve Time : %f ms\n", qp->stats->kkt_time * 1000.0);
			printf("LDL Time       : %f ms\n", qp->stats->ldl_numeric * 1000.0);
			printf("Iterations     : %ld\n\n", qp->stats->IterationCount);
		}
		if (qp->stats->Flag == QP_FATAL) printf("\nUnknown Error Detected\n\n");
		if (qp->stats->Flag == QP_KKTFAIL) printf("\nLDL Factorization fail\n\n");
	}
	return qp->stats->Flag;
}

// Introspection for the tests: 0 = plans in the (bounded) cache, 1 = its capacity.  There is nothing else to
// count: the library keeps no table of QPs.
qp_int QPB_PLAN_CACHE_INFO(int what)
{
	std::lock_guard<std::mutex> guard(g_lock);
	return what == 0 ? (qp_int)g_plans.size() : (qp_int)PLAN_CACHE_SIZE;
}

void QP_CLEANUP(QP *qp) { cleanup_common(qp, false); }
void QP_CLEANUP_dense(QP *qp) { cleanup_common(qp, true); }

} // extern "C"

/*
 * oracle/ref_driver.c -- multi-threaded batch driver for the CPU oracles.  TEST INFRASTRUCTURE:
 * used by tests (fast oracle batches), bench.py's cpu_baseline leg and `bench.py --impl reference`.
 *
 * dlopen()s an oracle library (oracle/_ref/libqpswift_ref.so = the unmodified reference, or
 * oracle/liboracle_port.so) and runs, one instance per host thread at a time,
 *     QP_SETUP(CCS pattern shared by the batch, this instance's values, Permut) -> QP_SOLVE -> QP_CLEANUP
 * which is exactly the parity procedure of SURVEY.md section 8c and the "plan-reuse" CPU baseline of
 * BASELINE.md section 4 (the permutation is computed once and supplied, so AMD is not re-run per QP).
 * With permut == NULL every instance runs AMD itself (controller style).
 * The reference is re-entrant across distinct QP handles (no mutable globals on this path).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../include/qpSWIFT/qpSWIFT.h"

typedef QP *(*setup_fn)(qp_int, qp_int, qp_int, qp_int *, qp_int *, qp_real *, qp_int *, qp_int *, qp_real *, qp_int *, qp_int *, qp_real *, qp_real *, qp_real *, qp_real *, qp_real, qp_int *);
typedef qp_int (*solve_fn)(QP *);
typedef void (*cleanup_fn)(QP *);

typedef struct
{
	setup_fn setup; solve_fn solve; cleanup_fn cleanup;
	qp_int n, m, p, nnzP, nnzA, nnzG;
	qp_int *Pjc, *Pir, *Ajc, *Air, *Gjc, *Gir, *permut;
	const qp_real *Ppr, *Apr, *Gpr, *c, *h, *b;
	qp_real *x, *y, *z, *s, *fval;
	qp_int *flag, *iters;
	qp_real sigma_d, reltol, abstol;
	qp_int maxit;
	long lo, hi;
	int pin;
} job_t;

/* Thread t of the batch runs on the t-th CPU this process is allowed on (stable timings on a busy box). */
static void pin_to_cpu(int t)
{
	cpu_set_t allowed, one;
	if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return;
	const int have = CPU_COUNT(&allowed);
	if (have < 1) return;
	int want = t % have;
	for (int c = 0; c < CPU_SETSIZE; ++c) {
		if (!CPU_ISSET(c, &allowed)) continue;
		if (want-- == 0) {
			CPU_ZERO(&one);
			CPU_SET(c, &one);
			pthread_setaffinity_np(pthread_self(), sizeof one, &one);
			return;
		}
	}
}

static void *worker(void *arg)
{
	job_t *J = (job_t *)arg;
	if (J->pin >= 0) pin_to_cpu(J->pin);
	const qp_int n = J->n, m = J->m, p = J->p, N = n + m + p;
	qp_real *vals = malloc(sizeof(qp_real) * (size_t)(J->nnzP + J->nnzA + J->nnzG + N + 4));
	qp_int *perm = J->permut ? malloc(sizeof(qp_int) * (size_t)N) : NULL;
	for (long i = J->lo; i < J->hi; ++i) {
		qp_real *Pv = vals, *Av = Pv + J->nnzP, *Gv = Av + J->nnzA, *cv = Gv + J->nnzG, *hv = cv + n, *bv = hv + m;
		memcpy(Pv, J->Ppr + i * J->nnzP, sizeof(qp_real) * (size_t)J->nnzP);
		if (p) memcpy(Av, J->Apr + i * J->nnzA, sizeof(qp_real) * (size_t)J->nnzA);
		memcpy(Gv, J->Gpr + i * J->nnzG, sizeof(qp_real) * (size_t)J->nnzG);
		memcpy(cv, J->c + i * n, sizeof(qp_real) * (size_t)n);
		memcpy(hv, J->h + i * m, sizeof(qp_real) * (size_t)m);
		if (p) memcpy(bv, J->b + i * p, sizeof(qp_real) * (size_t)p);
		if (perm) memcpy(perm, J->permut, sizeof(qp_int) * (size_t)N);
		QP *q = J->setup(n, m, p, J->Pjc, J->Pir, Pv, p ? J->Ajc : NULL, p ? J->Air : NULL, p ? Av : NULL, J->Gjc, J->Gir, Gv, cv, hv, p ? bv : NULL, J->sigma_d, perm);
		if (J->maxit >= 0) q->options->maxit = J->maxit;
		if (J->reltol > 0) q->options->reltol = J->reltol;
		if (J->abstol > 0) q->options->abstol = J->abstol;
		const qp_int flag = J->solve(q);
		if (J->x) memcpy(J->x + i * n, q->x, sizeof(qp_real) * (size_t)n);
		if (J->y && p) memcpy(J->y + i * p, q->y, sizeof(qp_real) * (size_t)p);
		if (J->z) memcpy(J->z + i * m, q->z, sizeof(qp_real) * (size_t)m);
		if (J->s) memcpy(J->s + i * m, q->s, sizeof(qp_real) * (size_t)m);
		if (J->fval) J->fval[i] = q->stats->fval;
		if (J->flag) J->flag[i] = flag;
		if (J->iters) J->iters[i] = q->stats->IterationCount;
		J->cleanup(q);
	}
	free(vals);
	free(perm);
	return NULL;
}

/* Returns wall seconds for the whole batch (CLOCK_MONOTONIC), or -1 if the library cannot be loaded. */
double oracle_run_batch(const char *lib_path, int nthreads, long count, qp_int n, qp_int m, qp_int p,
                        qp_int *Pjc, qp_int *Pir, qp_int *Ajc, qp_int *Air, qp_int *Gjc, qp_int *Gir, qp_int *permut,
                        const qp_real *Ppr, const qp_real *Apr, const qp_real *Gpr, const qp_real *c, const qp_real *h, const qp_real *b,
                        qp_real sigma_d, qp_int maxit, qp_real reltol, qp_real abstol,
                        qp_real *x, qp_real *y, qp_real *z, qp_real *s, qp_int *flag, qp_int *iters, qp_real *fval)
{
	void *lib = dlopen(lib_path, RTLD_NOW | RTLD_LOCAL);
	if (!lib) return -1.0;
	job_t proto;
	memset(&proto, 0, sizeof proto);
	proto.setup = (setup_fn)dlsym(lib, "QP_SETUP");
	proto.solve = (solve_fn)dlsym(lib, "QP_SOLVE");
	proto.cleanup = (cleanup_fn)dlsym(lib, "QP_CLEANUP");
	if (!proto.setup || !proto.solve || !proto.cleanup) return -1.0;
	proto.n = n; proto.m = m; proto.p = p; proto.nnzP = Pjc[n]; proto.nnzA = p ? Ajc[n] : 0; proto.nnzG = Gjc[n];
	proto.Pjc = Pjc; proto.Pir = Pir; proto.Ajc = Ajc; proto.Air = Air; proto.Gjc = Gjc; proto.Gir = Gir; proto.permut = permut;
	proto.Ppr = Ppr; proto.Apr = Apr; proto.Gpr = Gpr; proto.c = c; proto.h = h; proto.b = b;
	proto.x = x; proto.y = y; proto.z = z; proto.s = s; proto.flag = flag; proto.iters = iters; proto.fval = fval;
	proto.sigma_d = sigma_d; proto.maxit = maxit; proto.reltol = reltol; proto.abstol = abstol;
	if (nthreads < 1) nthreads = 1;
	if (nthreads > count) nthreads = count > 0 ? (int)count : 1;
	job_t *jobs = malloc(sizeof(job_t) * (size_t)nthreads);
	pthread_t *tid = malloc(sizeof(pthread_t) * (size_t)nthreads);
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	for (int t = 0; t < nthreads; ++t) {
		jobs[t] = proto;
		jobs[t].lo = count * t / nthreads;
		jobs[t].hi = count * (t + 1) / nthreads;
		jobs[t].pin = t;
		pthread_create(&tid[t], NULL, worker, &jobs[t]);
	}
	for (int t = 0; t < nthreads; ++t) pthread_join(tid[t], NULL);
	clock_gettime(CLOCK_MONOTONIC, &t1);
	free(jobs);
	free(tid);
	return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ---- controller style: what dogbot_controller does every tick (main.cpp:1649-1663) ----
 * QP_SETUP_dense (dense column-major P, A, G; densetosparse + AMD + symbolic analysis per QP) -> tolerances ->
 * QP_SOLVE -> read x -> QP_CLEANUP_dense.  Works on any library that exports qpSWIFT's C API: the reference and
 * libqpswift_b200.so alike (the drop-in check). */
typedef QP *(*setup_dense_fn)(qp_int, qp_int, qp_int, qp_real *, qp_real *, qp_real *, qp_real *, qp_real *, qp_real *, qp_int *, int);

typedef struct
{
	setup_dense_fn setup; solve_fn solve; cleanup_fn cleanup;
	qp_int n, m, p;
	const qp_real *P, *A, *G, *c, *h, *b;
	qp_real *x;
	qp_int *flag, *iters;
	qp_real reltol, abstol;
	long lo, hi;
	int pin;
} dense_job_t;

static void *dense_worker(void *arg)
{
	dense_job_t *J = (dense_job_t *)arg;
	if (J->pin >= 0) pin_to_cpu(J->pin);
	const qp_int n = J->n, m = J->m, p = J->p;
	const size_t sz[6] = {(size_t)(n * n), (size_t)(p * n), (size_t)(m * n), (size_t)n, (size_t)m, (size_t)p};
	const qp_real *src[6] = {J->P, J->A, J->G, J->c, J->h, J->b};
	qp_real *buf[6];
	for (int k = 0; k < 6; ++k) buf[k] = malloc(sizeof(qp_real) * (sz[k] ? sz[k] : 1));
	for (long i = J->lo; i < J->hi; ++i) {
		for (int k = 0; k < 6; ++k)
			if (sz[k]) memcpy(buf[k], src[k] + (size_t)i * sz[k], sizeof(qp_real) * sz[k]);
		QP *q = J->setup(n, m, p, buf[0], p ? buf[1] : NULL, buf[2], buf[3], buf[4], p ? buf[5] : NULL, NULL, COLUMN_MAJOR_ORDERING);
		if (J->reltol > 0) q->options->reltol = J->reltol;
		if (J->abstol > 0) q->options->abstol = J->abstol;
		const qp_int flag = J->solve(q);
		if (J->x) memcpy(J->x + i * n, q->x, sizeof(qp_real) * (size_t)n);
		if (J->flag) J->flag[i] = flag;
		if (J->iters) J->iters[i] = q->stats->IterationCount;
		J->cleanup(q);
	}
	for (int k = 0; k < 6; ++k) free(buf[k]);
	return NULL;
}

double oracle_run_batch_dense(const char *lib_path, int nthreads, long count, qp_int n, qp_int m, qp_int p,
                              const qp_real *P, const qp_real *A, const qp_real *G, const qp_real *c, const qp_real *h, const qp_real *b,
                              qp_real reltol, qp_real abstol, qp_real *x, qp_int *flag, qp_int *iters)
{
	void *lib = dlopen(lib_path, RTLD_NOW | RTLD_LOCAL);
	if (!lib) return -1.0;
	dense_job_t proto;
	memset(&proto, 0, sizeof proto);
	proto.setup = (setup_dense_fn)dlsym(lib, "QP_SETUP_dense");
	proto.solve = (solve_fn)dlsym(lib, "QP_SOLVE");
	proto.cleanup = (cleanup_fn)dlsym(lib, "QP_CLEANUP_dense");
	if (!proto.setup || !proto.solve || !proto.cleanup) return -1.0;
	proto.n = n; proto.m = m; proto.p = p; proto.P = P; proto.A = A; proto.G = G; proto.c = c; proto.h = h; proto.b = b;
	proto.x = x; proto.flag = flag; proto.iters = iters; proto.reltol = reltol; proto.abstol = abstol;
	if (nthreads < 1) nthreads = 1;
	if (nthreads > count) nthreads = count > 0 ? (int)count : 1;
	dense_job_t *jobs = malloc(sizeof(dense_job_t) * (size_t)nthreads);
	pthread_t *tid = malloc(sizeof(pthread_t) * (size_t)nthreads);
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	for (int t = 0; t < nthreads; ++t) {
		jobs[t] = proto;
		jobs[t].lo = count * t / nthreads;
		jobs[t].hi = count * (t + 1) / nthreads;
		jobs[t].pin = t;
		pthread_create(&tid[t], NULL, dense_worker, &jobs[t]);
	}
	for (int t = 0; t < nthreads; ++t) pthread_join(tid[t], NULL);
	clock_gettime(CLOCK_MONOTONIC, &t1);
	free(jobs);
	free(tid);
	return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* One control loop: `ticks` QPs one after the other on the calling thread, each through QP_SETUP_dense -> QP_SOLVE ->
 * x -> QP_CLEANUP_dense (cleanup != 0; the controller itself never cleans up, main.cpp:1649-1667), the wall time of
 * every tick in us[ticks] (microseconds; setup_us[ticks] the QP_SETUP_dense part, may be NULL).  The inputs cycle
 * through `count` dense instances.  Returns 0, or -1 if the library lacks the API. */
int oracle_control_loop(const char *lib_path, long ticks, long count, qp_int n, qp_int m, qp_int p,
                        const qp_real *P, const qp_real *A, const qp_real *G, const qp_real *c, const qp_real *h, const qp_real *b,
                        qp_real reltol, qp_real abstol, int cleanup, double *us, double *setup_us, qp_real *x, qp_int *flag, qp_int *iters)
{
	void *lib = dlopen(lib_path, RTLD_NOW | RTLD_LOCAL);
	if (!lib) return -1;
	setup_dense_fn setup = (setup_dense_fn)dlsym(lib, "QP_SETUP_dense");
	solve_fn solve = (solve_fn)dlsym(lib, "QP_SOLVE");
	cleanup_fn clean = (cleanup_fn)dlsym(lib, "QP_CLEANUP_dense");
	if (!setup || !solve || !clean) return -1;
	const size_t sz[6] = {(size_t)(n * n), (size_t)(p * n), (size_t)(m * n), (size_t)n, (size_t)m, (size_t)p};
	const qp_real *src[6] = {P, A, G, c, h, b};
	qp_real *buf[6];
	for (int k = 0; k < 6; ++k) buf[k] = malloc(sizeof(qp_real) * (sz[k] ? sz[k] : 1));
	for (long t = 0; t < ticks; ++t) {
		const long i = t % count;
		for (int k = 0; k < 6; ++k)
			if (sz[k]) memcpy(buf[k], src[k] + (size_t)i * sz[k], sizeof(qp_real) * sz[k]);
		struct timespec t0, t1, t2;
		clock_gettime(CLOCK_MONOTONIC, &t0);
		QP *q = setup(n, m, p, buf[0], p ? buf[1] : NULL, buf[2], buf[3], buf[4], p ? buf[5] : NULL, NULL, COLUMN_MAJOR_ORDERING);
		clock_gettime(CLOCK_MONOTONIC, &t1);
		if (reltol > 0) q->options->reltol = reltol;
		if (abstol > 0) q->options->abstol = abstol;
		const qp_int fl = solve(q);
		qp_real first = 0;
		for (qp_int k = 0; k < n; ++k) first += q->x[k]; /* the controller copies x out (main.cpp:1662-1664) */
		clock_gettime(CLOCK_MONOTONIC, &t2);
		if (x) memcpy(x + (size_t)t * n, q->x, sizeof(qp_real) * (size_t)n);
		if (flag) flag[t] = fl;
		if (iters) iters[t] = q->stats->IterationCount;
		if (us) us[t] = 1e6 * (double)(t2.tv_sec - t0.tv_sec) + 1e-3 * (double)(t2.tv_nsec - t0.tv_nsec) + 0.0 * first;
		if (setup_us) setup_us[t] = 1e6 * (double)(t1.tv_sec - t0.tv_sec) + 1e-3 * (double)(t1.tv_nsec - t0.tv_nsec);
		if (cleanup) clean(q);
	}
	for (int k = 0; k < 6; ++k) free(buf[k]);
	return 0;
}

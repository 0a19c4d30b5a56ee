/*
 * oracle/qpswift_port.c -- CPU restatement of the reference's QP hot path.  TEST INFRASTRUCTURE.
 *
 * This file is NOT part of the product: nothing under apf_quadruped_b200/ links, loads or calls it.
 * It exists so that the parity tests have a checker even where the real reference binary
 * (oracle/_ref/libqpswift_ref.so, built by oracle/Makefile from /root/reference) is absent, and it
 * is itself pinned against that binary by tests/test_oracle.py (bit-identical x, y, z, s, flags,
 * iteration counts, factor values and plan arrays on seeded DogBot batches) and against the
 * committed golden vectors under tests/golden/.  Parity status: PINNED (reference executed here).
 *
 * It is a restatement, in our own words, of the algorithm in
 * /root/reference/dogbot_controller/src/qpSWIFT/{qpSWIFT.c, Auxilary.c, ldl.c}; each function
 * below names the reference lines it follows.  The fill-reducing ordering (SuiteSparse AMD,
 * amd_*.c) is NOT restated here: QP_SETUP must be given the permutation (`Permut`), which the
 * Python side (oracle/reference.py) takes from the reference binary or from tests/golden/plans.npz.
 * With Permut == NULL the identity ordering is used and stats->AMD_RESULT is set to -1, which is
 * the reference's own fallback when AMD fails (qpSWIFT.c:207-216).
 *
 * The structs mirror include/qpSWIFT/Auxilary.h:18-151 field for field so that the same ctypes
 * binding drives both oracles.  Arithmetic: plain IEEE doubles, sequential sums, no FMA
 * (build with -O2 -ffp-contract=off, see oracle/Makefile).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef double real;
typedef long idx;

typedef struct { idx *jc, *ir; real *pr; idx n, m, nnz; } smat; /* n = columns, m = rows */

typedef struct
{
	smat *kktmatrix; real *b; idx *Parent, *Flag, *Lnz, *Li, *Lp, *Lti, *Ltp, *Pattern, *UPattern;
	real *Y, *Lx, *D; idx *P, *Pinv;
} kkt_t;

typedef struct
{
	real tsetup, tsolve, kkt_time, ldl_numeric; idx IterationCount; real n_rx, n_ry, n_rz, n_mu, alpha_p, alpha_d, fval;
	idx Flag, AMD_RESULT, resolve_kkt;
} stats_t;

typedef struct { idx maxit; real reltol, abstol, sigma; idx verbose; } settings_t;

typedef struct
{
	idx n, m, p; real sigma_d, mu, rho;
	real *x, *y, *z, *s, *rx, *ry, *rz, *delta, *delta_x, *delta_y, *delta_z, *delta_s, *ds, *lambda, *temp;
	smat *P; real *c; smat *G; real *h; smat *A; real *b; smat *At, *Gt;
	kkt_t *kkt; settings_t *options; stats_t *stats;
} QP;

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec; }
static void *grab(size_t count, size_t size) { return malloc((count ? count : 1) * size); }

static smat *view(idx rows, idx cols, idx *jc, idx *ir, real *pr)
{
	smat *M = grab(1, sizeof *M);
	M->jc = jc; M->ir = ir; M->pr = pr; M->m = rows; M->n = cols; M->nnz = jc[cols];
	return M;
}

/* Auxilary.c:901-951 -- transpose by counting; inside a column of the result the entries keep the
 * ascending column order of the source. */
static smat *transposed(const smat *A)
{
	idx *fill = calloc(A->m ? A->m : 1, sizeof(idx)), q, c;
	smat *T = view(A->n, A->m, calloc(A->m + 1, sizeof(idx)), grab(A->nnz, sizeof(idx)), grab(A->nnz, sizeof(real)));
	for (q = 0; q < A->nnz; ++q) ++T->jc[A->ir[q] + 1];
	for (c = 0; c < A->m; ++c) T->jc[c + 1] += T->jc[c];
	for (c = 0; c < A->n; ++c)
		for (q = A->jc[c]; q < A->jc[c + 1]; ++q) {
			const idx at = T->jc[A->ir[q]] + fill[A->ir[q]]++;
			T->ir[at] = c;
			T->pr[at] = A->pr[q];
		}
	T->nnz = A->nnz;
	free(fill);
	return T;
}

/* Auxilary.c:853-859 (y -= A x, column scatter) and 816-822 (y -= A' x, column gather). */
static void sub_Ax(const smat *A, const real *x, real *y, int clear)
{
	idx c, q;
	if (clear) for (c = 0; c < A->m; ++c) y[c] = 0;
	for (c = 0; c < A->n; ++c)
		for (q = A->jc[c]; q < A->jc[c + 1]; ++q) y[A->ir[q]] -= x[c] * A->pr[q];
}
static void sub_Atx(const smat *A, const real *x, real *y)
{
	idx c, q;
	for (c = 0; c < A->n; ++c)
		for (q = A->jc[c]; q < A->jc[c + 1]; ++q) y[c] -= A->pr[q] * x[A->ir[q]];
}
static real dot(const real *a, const real *b, idx k) { real s = 0; idx i; for (i = 0; i < k; ++i) s += a[i] * b[i]; return s; }
static void axpy(real *x, const real *d, real a, idx k) { idx i; for (i = 0; i < k; ++i) x[i] += d[i] * a; }

/* Auxilary.c:71-181 -- K = [P A' G'; A 0 0; G 0 -I], column by column; the -1 of a z column is
 * written right after the last entry of that row of G. */
static void assemble_kkt(QP *q)
{
	const idx n = q->n, p = q->p, m = q->m, N = n + p + m;
	const idx nz = q->P->nnz + 2 * q->G->nnz + m + (p ? 2 * q->A->nnz : 0);
	smat *K = view(N, N, calloc(N + 1, sizeof(idx)), grab(nz, sizeof(idx)), grab(nz, sizeof(real)));
	idx at = 0, c, t;
	for (c = 0; c < n; ++c) {
		for (t = q->P->jc[c]; t < q->P->jc[c + 1]; ++t) { K->ir[at] = q->P->ir[t]; K->pr[at++] = q->P->pr[t]; }
		if (p) for (t = q->A->jc[c]; t < q->A->jc[c + 1]; ++t) { K->ir[at] = n + q->A->ir[t]; K->pr[at++] = q->A->pr[t]; }
		for (t = q->G->jc[c]; t < q->G->jc[c + 1]; ++t) { K->ir[at] = n + p + q->G->ir[t]; K->pr[at++] = q->G->pr[t]; }
		K->jc[c + 1] = at;
	}
	for (c = 0; c < p; ++c) {
		for (t = q->At->jc[c]; t < q->At->jc[c + 1]; ++t) { K->ir[at] = q->At->ir[t]; K->pr[at++] = q->At->pr[t]; }
		K->jc[n + c + 1] = at;
	}
	for (c = 0; c < m; ++c) {
		for (t = q->Gt->jc[c]; t < q->Gt->jc[c + 1]; ++t) { K->ir[at] = q->Gt->ir[t]; K->pr[at++] = q->Gt->pr[t]; }
		if (q->Gt->jc[c + 1] > q->Gt->jc[c]) { K->ir[at] = n + p + c; K->pr[at++] = -1.0; }
		K->jc[n + p + c + 1] = at;
	}
	K->nnz = at;
	q->kkt->kktmatrix = K;
}

/* ldl.c:187-240 -- inverse permutation, elimination tree, column counts, column pointers. */
static void symbolic(kkt_t *k)
{
	const smat *K = k->kktmatrix;
	const idx N = K->n;
	idx r, t;
	for (r = 0; r < N; ++r) k->Pinv[k->P[r]] = r;
	for (r = 0; r < N; ++r) {
		k->Parent[r] = -1; k->Flag[r] = r; k->Lnz[r] = 0;
		for (t = K->jc[k->P[r]]; t < K->jc[k->P[r] + 1]; ++t) {
			idx i = k->Pinv[K->ir[t]];
			if (i >= r) continue;
			for (; k->Flag[i] != r; i = k->Parent[i]) {
				if (k->Parent[i] == -1) k->Parent[i] = r;
				++k->Lnz[i];
				k->Flag[i] = r;
			}
		}
	}
	k->Lp[0] = 0;
	for (r = 0; r < N; ++r) k->Lp[r + 1] = k->Lp[r] + k->Lnz[r];
}

/* ldl.c:253-326 -- up-looking LDL' with the reference's dynamic regularisation (319-320). */
static void numeric(kkt_t *k)
{
	const smat *K = k->kktmatrix;
	const idx N = K->n;
	idx r, t, top, len, i;
	for (r = 0; r < N; ++r) {
		k->Y[r] = 0.0; top = N; k->Flag[r] = r; k->Lnz[r] = 0;
		for (t = K->jc[k->P[r]]; t < K->jc[k->P[r] + 1]; ++t) {
			i = k->Pinv[K->ir[t]];
			if (i > r) continue;
			k->Y[i] += K->pr[t];
			for (len = 0; k->Flag[i] != r; i = k->Parent[i]) { k->Pattern[len++] = i; k->Flag[i] = r; }
			while (len > 0) k->Pattern[--top] = k->Pattern[--len];
		}
		k->D[r] = k->Y[r];
		k->Y[r] = 0.0;
		for (; top < N; ++top) {
			real yi, l;
			idx q, stop;
			i = k->Pattern[top];
			yi = k->Y[i];
			k->Y[i] = 0.0;
			stop = k->Lp[i] + k->Lnz[i];
			for (q = k->Lp[i]; q < stop; ++q) k->Y[k->Li[q]] -= k->Lx[q] * yi;
			l = yi / k->D[i];
			k->D[r] -= l * yi;
			k->Li[stop] = r;
			k->Lx[stop] = l;
			++k->Lnz[i];
		}
		{
			const real sign = k->D[r] <= 0 ? -1.0 : 1.0;
			if (sign * k->D[r] <= 0.00000000000001) k->D[r] = sign * 0.0000001;
		}
	}
}

/* ldl.c:495-597 -- b <- P' L'^-1 D^-1 L^-1 P b, using `w` as the permuted work vector. */
static void solve_kkt(kkt_t *k, real *w)
{
	const idx N = k->kktmatrix->n;
	idx j, q;
	for (j = 0; j < N; ++j) w[j] = k->b[k->P[j]];
	for (j = 0; j < N; ++j)
		for (q = k->Lp[j]; q < k->Lp[j + 1]; ++q) w[k->Li[q]] -= k->Lx[q] * w[j];
	for (j = 0; j < N; ++j) w[j] /= k->D[j];
	for (j = N - 1; j >= 0; --j)
		for (q = k->Lp[j]; q < k->Lp[j + 1]; ++q) w[j] -= k->Lx[q] * w[k->Li[q]];
	for (j = 0; j < N; ++j) k->b[k->P[j]] = w[j];
}

/* Auxilary.c:992-1089 with ldlinitialsolve 578-625. */
static void initial_point(QP *q)
{
	const idx n = q->n, p = q->p, m = q->m, N = n + p + m;
	kkt_t *k = q->kkt;
	real *r = grab(m, sizeof(real)), lo, hi;
	idx i;
	for (i = 0; i < n; ++i) k->b[i] = -q->c[i];
	for (i = 0; i < p; ++i) k->b[n + i] = q->b[i];
	for (i = 0; i < m; ++i) k->b[n + p + i] = q->h[i];
	k->Y = grab(N, sizeof(real)); k->D = grab(N, sizeof(real));
	k->Lp = grab(N + 1, sizeof(idx)); k->Parent = grab(N, sizeof(idx)); k->Pattern = grab(N, sizeof(idx));
	k->Flag = grab(N, sizeof(idx)); k->Lnz = grab(N, sizeof(idx));
	symbolic(k);
	k->Li = grab(k->Lp[N] + 1, sizeof(idx)); k->Lx = grab(k->Lp[N] + 1, sizeof(real));
	k->Lti = grab(k->Lp[N], sizeof(idx)); k->Ltp = grab(N + 1, sizeof(idx));
	numeric(k);
	solve_kkt(k, q->delta);
	for (i = 0; i < n; ++i) q->x[i] = k->b[i];
	for (i = 0; i < p; ++i) q->y[i] = k->b[n + i];
	sub_Ax(q->G, q->x, r, 1);
	axpy(r, q->h, 1, m);
	lo = hi = r[0];
	for (i = 1; i < m; ++i) { if (r[i] < lo) lo = r[i]; if (r[i] > hi) hi = r[i]; }
	lo = -lo;
	for (i = 0; i < m; ++i) q->s[i] = lo < 0 ? r[i] : r[i] + (1 + lo);
	for (i = 0; i < m; ++i) q->z[i] = hi < 0 ? -r[i] : -r[i] + (1 + hi);
	free(r);
}

/* qpSWIFT.c:60-234 */
QP *QP_SETUP(idx n, idx m, idx p, idx *Pjc, idx *Pir, real *Ppr, idx *Ajc, idx *Air, real *Apr, idx *Gjc, idx *Gir, real *Gpr, real *c, real *h, real *b, real sigma_d, idx *Permut)
{
	const double t0 = now();
	QP *q = calloc(1, sizeof *q);
	idx N, i;
	q->options = grab(1, sizeof(settings_t));
	q->options->maxit = 100; q->options->reltol = 1e-6; q->options->abstol = 1e-6; q->options->sigma = 100; q->options->verbose = 0;
	q->stats = calloc(1, sizeof(stats_t));
	q->stats->Flag = 3;
	q->n = n; q->m = m;
	if (Apr && Ajc && Air && b && p != 0) { q->p = p; q->A = view(p, n, Ajc, Air, Apr); q->b = b; q->At = transposed(q->A); }
	q->P = view(n, n, Pjc, Pir, Ppr);
	q->G = view(m, n, Gjc, Gir, Gpr);
	q->Gt = transposed(q->G);
	q->c = c; q->h = h; q->sigma_d = sigma_d;
	N = n + m + q->p;
	q->x = grab(n, sizeof(real)); q->y = grab(q->p, sizeof(real)); q->z = grab(m, sizeof(real)); q->s = grab(m, sizeof(real));
	q->rx = grab(n, sizeof(real)); q->ry = grab(q->p, sizeof(real)); q->rz = grab(m, sizeof(real));
	q->delta = grab(N, sizeof(real)); q->delta_x = grab(n, sizeof(real)); q->delta_y = grab(q->p, sizeof(real));
	q->delta_z = grab(m, sizeof(real)); q->delta_s = grab(m, sizeof(real));
	q->ds = grab(m, sizeof(real)); q->lambda = grab(m, sizeof(real)); q->temp = grab(n, sizeof(real));
	q->kkt = calloc(1, sizeof(kkt_t));
	assemble_kkt(q);
	q->kkt->b = grab(N, sizeof(real));
	q->kkt->P = grab(N, sizeof(idx));
	if (Permut) { memcpy(q->kkt->P, Permut, N * sizeof(idx)); q->stats->AMD_RESULT = -3; }
	else { for (i = 0; i < N; ++i) q->kkt->P[i] = i; q->stats->AMD_RESULT = -1; }
	q->kkt->Pinv = grab(N, sizeof(idx));
	q->kkt->UPattern = grab(N, sizeof(idx));
	initial_point(q);
	q->stats->tsetup = now() - t0;
	return q;
}

/* Auxilary.c:745-786 + 1133-1141 */
static void residuals(QP *q)
{
	idx i;
	sub_Ax(q->P, q->x, q->rx, 1);
	sub_Atx(q->G, q->z, q->rx);
	if (q->p) sub_Atx(q->A, q->y, q->rx);
	axpy(q->rx, q->c, -1, q->n);
	q->stats->n_rx = sqrt(dot(q->rx, q->rx, q->n));
	if (q->p) {
		sub_Ax(q->A, q->x, q->ry, 1);
		axpy(q->ry, q->b, 1, q->p);
		q->stats->n_ry = sqrt(dot(q->ry, q->ry, q->p));
	}
	sub_Ax(q->G, q->x, q->rz, 1);
	for (i = 0; i < q->m; ++i) q->rz[i] += q->h[i] - q->s[i];
	q->stats->n_rz = sqrt(dot(q->rz, q->rz, q->m));
	q->stats->n_mu = dot(q->s, q->z, q->m) / q->m;
	sub_Ax(q->P, q->x, q->temp, 1);
	q->stats->fval = -0.5 * dot(q->temp, q->x, q->n) + dot(q->c, q->x, q->n);
}

/* Auxilary.c:205-216 (indicator 0) and 274-295 */
static void load_system(QP *q, int refresh_diagonal)
{
	const idx n = q->n, p = q->p, m = q->m;
	smat *K = q->kkt->kktmatrix;
	idx i;
	if (refresh_diagonal)
		for (i = 0; i < m; ++i) K->pr[K->jc[n + p + i + 1] - 1] = -q->s[i] / q->z[i];
	for (i = 0; i < n; ++i) q->kkt->b[i] = q->rx[i];
	for (i = 0; i < p; ++i) q->kkt->b[n + i] = q->ry[i];
	for (i = 0; i < m; ++i) q->kkt->b[n + p + i] = q->rz[i] - (q->ds[i] / q->z[i]);
}

/* Auxilary.c:359-393 */
static void step_lengths(const QP *q, real *ap, real *ad)
{
	int hit_p = 0, hit_d = 0;
	idx i;
	*ap = 1e10; *ad = 1e10;
	for (i = 0; i < q->m; ++i) {
		if (q->delta_s[i] < 0 && (-q->s[i] / q->delta_s[i]) < *ap) { *ap = -(q->s[i] / q->delta_s[i]); hit_p = 1; }
		if (q->delta_z[i] < 0 && (-q->z[i] / q->delta_z[i]) < *ad) { *ad = -(q->z[i] / q->delta_z[i]); hit_d = 1; }
	}
	if (!hit_p) *ap = 1;
	if (!hit_d) *ad = 1;
}

static void recover_slack_direction(QP *q)
{
	const idx np = q->n + q->p;
	idx i;
	for (i = 0; i < q->m; ++i) q->delta_z[i] = q->kkt->b[np + i];
	for (i = 0; i < q->m; ++i) q->delta_s[i] = (q->ds[i] - (q->s[i] * q->delta_z[i])) / q->z[i];
}

/* qpSWIFT.c:473-644 */
idx QP_SOLVE(QP *q)
{
	const double t0 = now();
	const idx n = q->n, p = q->p, m = q->m;
	stats_t *st = q->stats;
	settings_t *o = q->options;
	idx it, i;
	if (st->Flag != 1) {
		for (it = 0; it < o->maxit; ++it) {
			const real thr = o->reltol / sqrt(3.0);
			residuals(q);
			if (st->n_rx < thr && st->n_rz < thr && (!p || st->n_ry < thr) && st->n_mu < o->abstol) { st->Flag = 0; break; }
			for (i = 0; i < m; ++i) q->lambda[i] = sqrt(q->s[i] * q->z[i]);
			q->mu = dot(q->lambda, q->lambda, m) / m;
			if (o->sigma > q->sigma_d) {
				real num = 0.0, mr;
				for (i = 0; i < m; ++i) q->ds[i] = -q->lambda[i] * q->lambda[i];
				load_system(q, 1);
				numeric(q->kkt);
				solve_kkt(q->kkt, q->delta);
				recover_slack_direction(q);
				step_lengths(q, &st->alpha_p, &st->alpha_d);
				for (i = 0; i < m; ++i) num += (q->s[i] + (st->alpha_p * q->delta_s[i])) * (q->z[i] + (st->alpha_d * q->delta_z[i]));
				q->rho = num / dot(q->s, q->z, m);
				mr = (1 < q->rho) ? 1 : q->rho;
				o->sigma = (q->sigma_d > mr * mr * mr) ? q->sigma_d : mr * mr * mr;
				for (i = 0; i < m; ++i) q->ds[i] = -(q->lambda[i] * q->lambda[i]) - (q->delta_s[i] * q->delta_z[i]) + (o->sigma * q->mu);
				load_system(q, 0);
				st->resolve_kkt = 0;
			} else {
				o->sigma = q->sigma_d;
				for (i = 0; i < m; ++i) q->ds[i] = -(q->lambda[i] * q->lambda[i]) + (o->sigma * q->mu);
				load_system(q, 1);
				st->resolve_kkt = 1;
				numeric(q->kkt);
			}
			solve_kkt(q->kkt, q->delta);
			for (i = 0; i < n; ++i) q->delta_x[i] = q->kkt->b[i];
			for (i = 0; i < p; ++i) q->delta_y[i] = q->kkt->b[n + i];
			recover_slack_direction(q);
			step_lengths(q, &st->alpha_p, &st->alpha_d);
			st->alpha_p = (0.99 * st->alpha_p) < 1.0 ? (0.99 * st->alpha_p) : 1.0;
			st->alpha_d = (0.99 * st->alpha_d) < 1.0 ? (0.99 * st->alpha_d) : 1.0;
			axpy(q->x, q->delta_x, st->alpha_p, n);
			axpy(q->y, q->delta_y, st->alpha_d, p);
			axpy(q->s, q->delta_s, st->alpha_p, m);
			axpy(q->z, q->delta_z, st->alpha_d, m);
			++st->IterationCount;
		}
		if (st->IterationCount == o->maxit) st->Flag = 2;
	}
	st->tsolve = now() - t0;
	return st->Flag;
}

static void drop(smat *M, int owns_arrays)
{
	if (!M) return;
	if (owns_arrays) { free(M->jc); free(M->ir); free(M->pr); }
	free(M);
}

/* qpSWIFT.c:661-737 */
void QP_CLEANUP(QP *q)
{
	kkt_t *k = q->kkt;
	free(q->x); free(q->y); free(q->z); free(q->s); free(q->rx); free(q->ry); free(q->rz);
	free(q->delta); free(q->delta_x); free(q->delta_y); free(q->delta_z); free(q->delta_s);
	free(q->ds); free(q->lambda); free(q->temp);
	drop(q->P, 0); drop(q->G, 0); drop(q->A, 0); drop(q->At, 1); drop(q->Gt, 1);
	drop(k->kktmatrix, 1);
	free(k->b); free(k->Parent); free(k->Flag); free(k->Lnz); free(k->Li); free(k->Lp); free(k->Lti); free(k->Ltp);
	free(k->Pattern); free(k->UPattern); free(k->Y); free(k->Lx); free(k->D); free(k->P); free(k->Pinv);
	free(k); free(q->options); free(q->stats); free(q);
}
void QP_CLEANUP_dense(QP *q) { QP_CLEANUP(q); }
QP *QP_SETUP_dense(idx n, idx m, idx p, real *Ppr, real *Apr, real *Gpr, real *c, real *h, real *b, idx *Permut, int ordering)
{
	(void)n; (void)m; (void)p; (void)Ppr; (void)Apr; (void)Gpr; (void)c; (void)h; (void)b; (void)Permut; (void)ordering;
	return NULL; /* the port only restates the sparse entry point (the parity procedure of SURVEY.md 8c) */
}

/*
 * qpSWIFT-compatible C API of the B200-native batched QP solver (drop-in boundary).
 *
 * A translation unit that today does  #include "qpSWIFT/qpSWIFT.h"  and links the
 * reference's `liblopt` (dogbot_controller/CMakeLists.txt:220) can be pointed at this
 * header + libqpswift_b200.so without source changes: the type names, the field order of
 * every struct a caller can reach through a `QP *`, the macro constants and the five entry
 * points are those of the reference.  Citations are relative to
 * /root/reference/dogbot_controller/ (INC = include/qpSWIFT, SRC = src/qpSWIFT).
 *
 *   replaces  INC/qpSWIFT.h:14-26      (entry points)
 *             INC/Auxilary.h:18-151    (smat, kkt, stats, settings, QP)
 *             INC/GlobalOptions.h:36-66 (qp_real, qp_int, defaults, exit flags, orderings)
 *
 * What is different behind the boundary: there is no CPU solver in this library.  QP_SETUP
 * does the integer analysis on the host (KKT pattern, AMD ordering, elimination tree, factor
 * pattern -- bit-exact with the reference) and QP_SOLVE runs the interior-point loop as a
 * batch of one on the current CUDA device.  If no CUDA device is usable QP_SOLVE reports
 * QP_FATAL; it never falls back to host arithmetic.  The batched entry points live in
 * "qpswift_b200.h".
 */
#ifndef QPSWIFT_B200_COMPAT_H
#define QPSWIFT_B200_COMPAT_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- scalar types (INC/GlobalOptions.h:36-43) -------------------------------------- */
#ifndef qp_real
#define qp_real double
#endif
#ifndef qp_int
#define qp_int long
#endif

/* ---- solver defaults (INC/GlobalOptions.h:46-50) ------------------------------------ */
#define MAXIT (100)
#define RELTOL (1e-6)
#define ABSTOL (1e-6)
#define SIGMA (100)
#define VERBOSE (0)

/* ---- exit flags (INC/GlobalOptions.h:54-57) ------------------------------------------ */
#define QP_OPTIMAL (0) /* converged                                   */
#define QP_KKTFAIL (1) /* factorisation failed (kept; unreachable, the LDL regularises) */
#define QP_MAXIT (2)   /* iteration count reached options->maxit      */
#define QP_FATAL (3)   /* not solved (initial value; also: no usable CUDA device)        */

/* ---- dense input orderings for QP_SETUP_dense (INC/GlobalOptions.h:59-60) ------------ */
#define ROW_MAJOR_ORDERING (20)
#define COLUMN_MAJOR_ORDERING (30)

/* ---- compressed-column matrix (INC/Auxilary.h:18-26).
 * NOTE the reference's field comments are inverted; in the code `m` is the ROW count and
 * `n` the COLUMN count (SRC/Auxilary.c:660-669).  We keep the code's meaning. ----------- */
typedef struct smat
{
	qp_int *jc;  /* column pointers, n+1 entries            */
	qp_int *ir;  /* row index of each stored entry          */
	qp_real *pr; /* value of each stored entry              */
	qp_int n;    /* number of columns                       */
	qp_int m;    /* number of rows                          */
	qp_int nnz;  /* stored entries (= jc[n])                */
} smat;

/* ---- KKT system and its symbolic/numeric factor (INC/Auxilary.h:28-50).
 * The integer members are the "plan": they are filled by QP_SETUP on the host and are
 * bit-identical with what the reference computes.  Lx/D/Y are host mirrors that are only
 * valid after QP_SOLVE when options->verbose >= 2 (they live on the device otherwise). --- */
typedef struct kkt
{
	smat *kktmatrix;  /* [P A' G'; A 0 0; G 0 -I], full symmetric CCS               */
	qp_real *b;       /* right-hand side workspace, n+p+m                            */
	qp_int *Parent;   /* elimination tree of P*K*P'                                  */
	qp_int *Flag;     /* workspace (kept for layout compatibility)                   */
	qp_int *Lnz;      /* column counts of L                                          */
	qp_int *Li;       /* row indices of L (value independent, computed at setup)     */
	qp_int *Lp;       /* column pointers of L                                        */
	qp_int *Lti;      /* row structure of L (Transpose_Row_Count), unused by solve   */
	qp_int *Ltp;
	qp_int *Pattern;  /* workspace (kept for layout compatibility)                   */
	qp_int *UPattern; /* allocated, unused (as in the reference)                     */
	qp_real *Y;       /* host mirror, see above                                      */
	qp_real *Lx;      /* host mirror, see above                                      */
	qp_real *D;       /* host mirror, see above                                      */
	qp_int *P;        /* fill-reducing permutation: P[k] = original index of pivot k */
	qp_int *Pinv;     /* inverse permutation                                         */
} kkt;

/* ---- statistics (INC/Auxilary.h:55-84) ------------------------------------------------ */
typedef struct stats
{
	qp_real tsetup;      /* seconds spent in QP_SETUP                       */
	qp_real tsolve;      /* seconds spent in QP_SOLVE (host wall clock)     */
	qp_real kkt_time;    /* device time of the solve kernel (CUDA events)   */
	qp_real ldl_numeric; /* not separable in the fused kernel: 0            */

	qp_int IterationCount;
	qp_real n_rx;
	qp_real n_ry;
	qp_real n_rz;
	qp_real n_mu;

	qp_real alpha_p;
	qp_real alpha_d;

	qp_real fval;
	qp_int Flag;
	qp_int AMD_RESULT; /* >=0 AMD ran; -3 caller supplied Permut; <0 identity fallback */

	qp_int resolve_kkt;
} stats;

/* ---- settings (INC/Auxilary.h:89-100); read at QP_SOLVE time, sigma is written back ---- */
typedef struct settings
{
	qp_int maxit;
	qp_real reltol;
	qp_real abstol;
	qp_real sigma;
	qp_int verbose;
} settings;

/* ---- the problem handle (INC/Auxilary.h:106-151) --------------------------------------- */
typedef struct QP
{
	qp_int n; /* variables      */
	qp_int m; /* inequalities   */
	qp_int p; /* equalities     */

	qp_real sigma_d;
	qp_real mu;
	qp_real rho;

	qp_real *x; /* [n] primal solution          */
	qp_real *y; /* [p] equality multipliers     */
	qp_real *z; /* [m] inequality multipliers   */
	qp_real *s; /* [m] slacks                   */

	qp_real *rx;
	qp_real *ry;
	qp_real *rz;

	qp_real *delta;
	qp_real *delta_x;
	qp_real *delta_y;
	qp_real *delta_z;
	qp_real *delta_s;

	qp_real *ds;
	qp_real *lambda;

	qp_real *temp;

	smat *P;
	qp_real *c;
	smat *G;
	qp_real *h;
	smat *A;
	qp_real *b;

	smat *At;
	smat *Gt;

	kkt *kkt;
	settings *options;
	stats *stats;
} QP;

/* ---- entry points (INC/qpSWIFT.h:14-26; behaviour SRC/qpSWIFT.c:60-839) ---------------- */

/* CCS inputs.  Aliases the caller's arrays (they must outlive the QP).  A/b may be NULL with
 * p == 0.  Permut == NULL -> AMD ordering; otherwise Permut is used and AMD_RESULT = -3.
 * Never returns NULL.  (SRC/qpSWIFT.c:60-234) */
QP *QP_SETUP(qp_int n, qp_int m, qp_int p, qp_int *Pjc, qp_int *Pir, qp_real *Ppr, qp_int *Ajc, qp_int *Air, qp_real *Apr, qp_int *Gjc, qp_int *Gir, qp_real *Gpr, qp_real *c, qp_real *h, qp_real *b, qp_real sigma_d, qp_int *Permut);

/* Dense inputs (column- or row-major); exact zeros are dropped, copies are made, sigma_d is
 * forced to 0.  Still aliases c, h, b.  (SRC/qpSWIFT.c:260-456) */
QP *QP_SETUP_dense(qp_int n, qp_int m, qp_int p, qp_real *Ppr, qp_real *Apr, qp_real *Gpr, qp_real *c, qp_real *h, qp_real *b, qp_int *Permut, int ordering);

/* Runs the interior-point loop on the GPU; returns and stores stats->Flag.  (SRC/qpSWIFT.c:473-644) */
qp_int QP_SOLVE(QP *myQP);

/* Release everything QP_SETUP / QP_SETUP_dense allocated.  (SRC/qpSWIFT.c:661-839) */
void QP_CLEANUP(QP *myQP);
void QP_CLEANUP_dense(QP *myQP);

#ifdef __cplusplus
}
#endif

#endif /* QPSWIFT_B200_COMPAT_H */

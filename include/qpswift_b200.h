/*
 * Batched / B200-specific C ABI of libqpswift_b200.so (plain pointers and sizes only).
 *
 * The reference has no batched entry point; BASELINE.json's north_star names QP_SOLVE_BATCH as
 * the new export next to the unchanged QP_SETUP / QP_SOLVE / QP_CLEANUP (qpSWIFT/qpSWIFT.h).
 * Everything here is what a cgo / ctypes / C++ caller binds; see INTEGRATION.md.
 *
 * Conventions
 *   - "value row" of one QP = the nonzero values of P, A, G in the CCS order of the prototype's
 *     pattern (exactly what QP_SETUP receives as Ppr/Apr/Gpr), plus c[n], h[m], b[p].
 *   - batch arrays are row-per-QP, C contiguous: Ppr[batch][nnzP], ..., x[batch][n], ...
 *   - return values < 0 are QPB_E_* errors (CUDA failures are never silent, and there is no
 *     CPU fallback: without a usable device every compute call returns QPB_E_NODEVICE).
 */
#ifndef QPSWIFT_B200_H
#define QPSWIFT_B200_H

#include "qpSWIFT/qpSWIFT.h"

#ifdef __cplusplus
extern "C" {
#endif

#define QPB_E_NODEVICE (-101) /* no CUDA device / driver                        */
#define QPB_E_CUDA (-102)     /* a CUDA call failed; see QPB_LAST_ERROR()       */
#define QPB_E_ARG (-103)      /* bad argument (null pointer, count > capacity)  */
#define QPB_E_PATTERN (-104)  /* unusable sparsity pattern (empty row of G)     */

/* ---------------------------------------------------------------------------------------------
 * Symbolic plan (host only, no GPU needed).  Replaces, once per sparsity pattern, what the
 * reference recomputes per QP: formkktmatrix_full (src/qpSWIFT/Auxilary.c:71-181), amd_l_order
 * (amd_order.c:21-199), LDL_symbolic (ldl.c:187-240) and the integer half of LDL_numeric
 * (ldl.c:276-298,315).
 * ------------------------------------------------------------------------------------------- */
typedef struct QPBPlan QPBPlan;

/* A* may be NULL with p == 0; Permut may be NULL (AMD ordering). Never returns NULL. */
QPBPlan *QPB_PLAN_CREATE(qp_int n, qp_int m, qp_int p, const qp_int *Pjc, const qp_int *Pir, const qp_int *Ajc, const qp_int *Air, const qp_int *Gjc, const qp_int *Gir, const qp_int *Permut);
void QPB_PLAN_DESTROY(QPBPlan *plan);

enum QPBPlanArray
{
	QPB_PLAN_KJC = 0, /* KKT column pointers [N+1]           */
	QPB_PLAN_KIR = 1, /* KKT row indices [nnzK]              */
	QPB_PLAN_P = 2,   /* permutation [N]                     */
	QPB_PLAN_PINV = 3,
	QPB_PLAN_PARENT = 4, /* elimination tree [N]             */
	QPB_PLAN_LNZ = 5,    /* column counts of L [N]           */
	QPB_PLAN_LP = 6,     /* column pointers of L [N+1]       */
	QPB_PLAN_LI = 7      /* row indices of L [Lp[N]]         */
};
/* Pointer to a plan array (owned by the plan) and its length. */
const qp_int *QPB_PLAN_ARRAY(const QPBPlan *plan, int which, qp_int *count);

enum QPBPlanInfo
{
	QPB_INFO_N = 0, /* KKT dimension n+p+m */
	QPB_INFO_NNZK = 1,
	QPB_INFO_LNZ = 2,
	QPB_INFO_AMD_STATUS = 3,
	QPB_INFO_FACTOR_FLOPS = 4, /* sum Lnz_k (Lnz_k + 2), ldl.c:182-184 */
	QPB_INFO_NNZU = 5,         /* upper-triangular KKT entries the factorisation reads */
	QPB_INFO_VALID = 6,
	QPB_INFO_PLAN_WORDS = 7, /* int32 words of the device plan */
	QPB_INFO_PROGRAM_OK = 8,    /* 1 when the factor/solve schedules could be compiled */
	QPB_INFO_SLICE_DOUBLES = 9, /* doubles of shared memory one QP occupies */
	QPB_INFO_PROGRAM_WORDS = 10, /* 64-bit instruction words of the schedules */
	/* per schedule, + {0 waves, 1 multiply-subtract waves, 2 quotient waves, 3 multiply-subtracts,
	 * 4 quotients, 5 chains, 6 occupied lane slots, 7 scratch doubles, 8 ordered-sum waves,
	 * 9 ordered sums, 10 suspensions, 11 first wave} */
	QPB_INFO_FACTOR_STATS = 32,
	QPB_INFO_SOLVE_STATS = 48
};
qp_int QPB_PLAN_INFO(const QPBPlan *plan, int what);

/* The compiled factor / solve schedules (csrc/program.h): 64-bit instruction words, 32 per wave,
 * and the layout of one QP's shared-memory slice as int32 offsets in doubles, in the order
 * {zero, V, c, h, b, zd, Lx, D, T, rhs, x, y, z, s, r, lam, ds, dz, dsl, total, tsize}.
 * Exposed so that the schedules can be checked on the host (tests/test_program.py). */
const unsigned long long *QPB_PLAN_PROGRAM(const QPBPlan *plan, qp_int *words);
qp_int QPB_PLAN_LAYOUT(const QPBPlan *plan, int *out, int capacity);

/* The micro-programs of the throughput kernel (csrc/stream_program.h): one thread per QP, working
 * sets interleaved in HBM.  QPB_PLAN_STREAM: the 64-bit micro-op words of all segments.
 * QPB_PLAN_STREAM_INFO: the slice layout {zero, one, mone, V, c, h, b, zd, Lx, D, rhs, x, y, z, s, r, lam, ds,
 * dz, dsl, px, hs, nb, total, scratch}, then {ok, micro-ops per group}, then for each segment (0 factor,
 * 1 solve, 2 residuals, 3 h - G x) {first word, groups, micro-ops, multiply-subtracts, quotients, late
 * loads, prefetched HBM loads, HBM stores, padding, slow groups}.  Returns the number of values available.
 * Exposed so that the programs can be checked on the host (tests/test_stream_program.py). */
const unsigned long long *QPB_PLAN_STREAM(const QPBPlan *plan, qp_int *words);
qp_int QPB_PLAN_STREAM_INFO(const QPBPlan *plan, long long *out, int capacity);
/* FNV-1a of the micro-programs and their layout: the name under which a kernel specialised for this
 * pattern (generated from these words by apf_quadruped_b200/codegen.py, compiled into the library) is
 * registered.  QPB_GENERATED_KERNELS: how many such kernels the library carries; keys[i] for i < capacity. */
unsigned long long QPB_PLAN_STREAM_KEY(const QPBPlan *plan);
qp_int QPB_GENERATED_KERNELS(unsigned long long *keys, int capacity);

/* The ordering alone (test hook for amd_l_order, amd_order.c:21-199): returns the AMD status
 * (0 ok, 1 ok-but-jumbled, -2 invalid) and writes Perm[n]. */
qp_int QPB_AMD_ORDER(qp_int n, const qp_int *Ap, const qp_int *Ai, qp_int *Perm);

/* ---------------------------------------------------------------------------------------------
 * Batched solve.  One plan, `capacity` QPs resident on ONE device (the caller shards a larger
 * batch across GPUs/processes: QPs are independent, see DESIGN.md "multi-GPU").
 * ------------------------------------------------------------------------------------------- */
typedef struct QPBatch QPBatch;

/* proto supplies n, m, p, the CCS patterns of P, A, G, the permutation and sigma_d.  Settings
 * are NOT captured: they are read from `proto->options` by every QP_SOLVE_BATCH call, like
 * QP_SOLVE reads them at call time (qpSWIFT.c:502,521).  device < 0 -> current device. */
QPBatch *QP_SETUP_BATCH(const QP *proto, qp_int capacity, int device);

/* Same, from a bare plan plus settings (no QP struct needed; used by bindings). */
QPBatch *QPB_BATCH_CREATE(const QPBPlan *plan, qp_real sigma_d, qp_int capacity, int device);

/* Host-pointer solve (pageable or pinned).  Copies the value rows H2D, solves, copies results
 * D2H.  Apr/b may be NULL when p == 0; any output pointer may be NULL.  Returns the number of
 * QPs whose flag is not QP_OPTIMAL, or a negative QPB_E_* code.
 * Per-QP semantics = QP_SETUP (initial point, Auxilary.c:992-1089) + QP_SOLVE (qpSWIFT.c:473-644). */
qp_int QP_SOLVE_BATCH(QPBatch *batch, qp_int count, const settings *opts, const qp_real *Ppr, const qp_real *Apr, const qp_real *Gpr, const qp_real *c, const qp_real *h, const qp_real *b, qp_real *x, qp_real *y, qp_real *z, qp_real *s, qp_int *flag, qp_int *iters);

/* Device-pointer solve on `stream` (a cudaStream_t passed as void*; NULL = default stream).
 * Asynchronous: nothing is synchronised, the return value is 0 or a negative QPB_E_* code.
 * flag/iters are int32 on the device; fval/stat (may be NULL) receive the objective value and
 * [n_rx, n_ry, n_rz, n_mu, alpha_p, alpha_d, sigma, mu] per QP. */
qp_int QP_SOLVE_BATCH_DEVICE(QPBatch *batch, qp_int count, const settings *opts, const qp_real *dPpr, const qp_real *dApr, const qp_real *dGpr, const qp_real *dc, const qp_real *dh, const qp_real *db, qp_real *dx, qp_real *dy, qp_real *dz, qp_real *ds, int *dflag, int *diters, qp_real *dfval, qp_real *dstat, void *stream);

/* Batched dense front-end.  The controller hands qpSWIFT DENSE matrices (QP_SETUP_dense,
 * dogbot_controller/src/client/main.cpp:1649-1650; qpSWIFT.c:260-456), which densetosparse /
 * densetosparse_ROWMAJOR (Auxilary.c:1154-1273) turn into CCS by dropping exact zeros -- a data-dependent
 * pattern.  A batch has ONE pattern (the plan's); these entries gather, on the device, the value of every
 * entry of that pattern from dense P [count][n*n], A [count][p*n], G [count][m*n] (ordering =
 * COLUMN_MAJOR_ORDERING or ROW_MAJOR_ORDERING, GlobalOptions.h:59-60) into the value rows QP_SOLVE_BATCH
 * takes, and report per QP what the reference's own pattern would have been different in:
 *   violation bit 0: a nonzero entry lies OUTSIDE the plan's pattern -- the plan does not describe this QP,
 *                    its solution must not be used;
 *   violation bit 1: an entry of the pattern is exactly 0.0 -- QP_SETUP_dense would have dropped it and
 *                    ordered a smaller KKT system (the solution here is that of the explicit pattern).
 * QPB_DENSE_TO_VALUES_DEVICE: device pointers, asynchronous on `stream`, dviolation (int32) may be NULL.
 * QP_SOLVE_BATCH_DENSE: host pointers; gather + QP_SOLVE_BATCH; violation[count] may be NULL. */
qp_int QPB_DENSE_TO_VALUES_DEVICE(QPBatch *batch, qp_int count, int ordering, const qp_real *dP, const qp_real *dA, const qp_real *dG, qp_real *dPpr, qp_real *dApr, qp_real *dGpr, int *dviolation, void *stream);
qp_int QP_SOLVE_BATCH_DENSE(QPBatch *batch, qp_int count, const settings *opts, int ordering, const qp_real *P, const qp_real *A, const qp_real *G, const qp_real *c, const qp_real *h, const qp_real *b, qp_real *x, qp_real *y, qp_real *z, qp_real *s, qp_int *flag, qp_int *iters, qp_int *violation);

/* The QPs whose rows are listed in dindex[0 .. *dcount_ptr) -- list and count in device memory, read by the kernel,
 * so a list compacted on the device (the robots that are in this contact pattern this tick) needs no host round trip
 * and the launch can be captured in a CUDA graph.  `rows` = rows the arrays hold (<= capacity, < 2^24).  Rows that
 * are not listed are not touched.  Asynchronous on `stream`. */
qp_int QP_SOLVE_BATCH_DEVICE_INDEXED(QPBatch *batch, qp_int rows, const settings *opts, const int *dindex, const int *dcount_ptr, const qp_real *dPpr, const qp_real *dApr, const qp_real *dGpr, const qp_real *dc, const qp_real *dh, const qp_real *db, qp_real *dx, qp_real *dy, qp_real *dz, qp_real *ds, int *dflag, int *diters, void *stream);

/* ---------------------------------------------------------------------------------------------
 * The DogBot whole-body-control QP built on the device, and the joint torques from its solution
 * (8 f1).  What dogbot_controller formulates every tick from iDynTree's quantities (main.cpp:1471-1647 stance,
 * 1734-1992 trot, 2943-3230 crawl) comes as ONE packed state row of QPB_WBC_STATE_DOUBLES doubles per robot:
 *   Jc[12][6] | Jj[12][12] | Mcom[6][6] | Mjj[12][12] | bias[18] | Jdqd[12] | W[6] | q[12] | dq[12] | vdot_sw[12]
 * (feet in the order br, bl, fl, fr; row-major blocks; W = the desired CoM wrench, main.cpp:1572; vdot_sw = the
 * desired swing-foot accelerations, main.cpp:1991).  kind: 0 stance (4 feet), 1 trot (bl, fr), 2 crawl (bl, fl, fr);
 * the batch's pattern must be that kind's (checked once; QPB_E_PATTERN otherwise).
 *   QPB_WBC_FORMULATE_DEVICE  writes the plan-ordered value rows Ppr | Apr | Gpr | c | h | b of robots 0 .. count-1, or
 *                             of the robots listed in dindex[0 .. *dcount_ptr) (both may be NULL), at the robot's row.
 *   QPB_WBC_TORQUES_DEVICE    tau[robot][12] = Mjj qdd + bias_j - Jst_j' f from x[robot][30] (main.cpp:1667, 2023, 3246).
 * Every number is a fixed sequence of IEEE-754 operations (apf_quadruped_b200/wbc.py restates them in numpy).
 * ------------------------------------------------------------------------------------------- */
#define QPB_WBC_STATE_DOUBLES 468
qp_int QPB_WBC_FORMULATE_DEVICE(QPBatch *batch, int kind, qp_int count, const qp_real *dstate, const int *dindex, const int *dcount_ptr, qp_real *dPpr, qp_real *dApr, qp_real *dGpr, qp_real *dc, qp_real *dh, qp_real *db, void *stream);
/* (host only) the formula table the formulation kernel walks: one descriptor per entry of Ppr | Apr | Gpr | c | h | b
 * (op | flag << 4 | a << 8 | b << 32, csrc/wbc.cuh) and the 32 constants that follow the state row. */
qp_int QPB_WBC_DESCRIPTORS(const QPBPlan *plan, int kind, unsigned long long *desc, qp_int capacity, qp_real *consts);
qp_int QPB_WBC_TORQUES_DEVICE(QPBatch *batch, int kind, qp_int count, const qp_real *dstate, const qp_real *dx, qp_real *dtau, const int *dindex, const int *dcount_ptr, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Rollout (BASELINE.json configs[3]): the device side of one control tick for `robots` robots whose
 * state rows (QPB_WBC_STATE_DOUBLES each) and rollout rows (QPB_ROLLOUT_ROW_DOUBLES each: r[4][3] | goal[4][2] | dx[6] |
 * dv[6] | noise[72]) live in device memory.  Integrates the previous tick's solutions dx [robots][30] (not on tick 0),
 * advances the contact schedule (stance / trot / stance / crawl, 25 ticks each, staggered over the robots), draws
 * perturbed footholds for the swing feet (counter-based, reproducible), moves the swing feet along the attractive
 * potential field (the model of main.cpp:1261-1421), rebuilds Jc, W, q, dq, vdot_sw of every state row, and writes
 * dkind [robots], the ascending robot lists dlists [3][robots] with dcounts [3] for the three contact patterns, and
 * *dtick += 1.  Follow with QPB_WBC_FORMULATE_DEVICE / QP_SOLVE_BATCH_DEVICE_INDEXED / QPB_WBC_TORQUES_DEVICE per pattern on
 * its list.  No host round trip: a tick can be captured in a CUDA graph and replayed.  first_robot: the index of robot 0
 * in a rollout that is split over several GPUs (schedule and random draws follow the global index). */
#define QPB_ROLLOUT_ROW_DOUBLES 104
qp_int QPB_ROLLOUT_TICK_DEVICE(int device, qp_int robots, qp_int first_robot, unsigned long long seed, qp_real *dstate, qp_real *drobot, const qp_real *dx, int *dkind, int *dlists, int *dcounts, long long *dtick, void *stream);

void QP_CLEANUP_BATCH(QPBatch *batch);

/* nnz[3] = entries of P, A, G per QP (the widths of the value rows Ppr / Apr / Gpr). */
void QPB_BATCH_ROW_WIDTHS(const QPBatch *batch, qp_int *nnz);

/* ---------------------------------------------------------------------------------------------
 * The same batch split across several GPUs of one box, behind the C ABI (SURVEY.md 8b proposed
 * QP_SETUP_BATCH(proto, batch, devices, ndev); the reference's caller is a C++ node, main.cpp:1649).
 * QPs are independent, so the batch is cut into contiguous shards, one per entry of `devices` (the first
 * count % ndev shards hold one QP more: QPB_MULTI_SHARD), every device solves its shard with its own copy
 * of the plan, and nothing crosses devices during the solve.  Results are bit-identical for any ndev.
 *   QP_SOLVE_BATCH_MULTI         host pointers (as QP_SOLVE_BATCH): every device reads its shard of the caller's
 *                                arrays and writes its shard of the results -- no collective.  Returns the
 *                                number of QPs whose flag is not QP_OPTIMAL, or a negative QPB_E_* code.
 *   QP_SOLVE_BATCH_MULTI_DEVICE  arrays of ndev device pointers, entry d = shard d's rows in device d's memory
 *                                (flags / iterations int32).  gx / gflag / giters (each may be NULL) receive, on
 *                                EVERY device, the ncclAllGather of x [ndev][most][n], flags and iteration counts
 *                                [ndev][most] (most = the largest shard; dx / dflag / diters must hold `most`
 *                                rows; rows beyond a shard's count are padding) -- the only collective on the
 *                                path, over one communicator per device (ncclCommInitAll).  Synchronous.
 * A device may appear more than once in `devices` (shards then share it; no NCCL gather between them).
 * ------------------------------------------------------------------------------------------- */
typedef struct QPBatchMulti QPBatchMulti;
QPBatchMulti *QP_SETUP_BATCH_MULTI(const QP *proto, qp_int capacity, const int *devices, int ndev);
QPBatchMulti *QPB_BATCH_CREATE_MULTI(const QPBPlan *plan, qp_int n, qp_int m, qp_int p, qp_real sigma_d, qp_int capacity, const int *devices, int ndev);
int QPB_MULTI_DEVICES(const QPBatchMulti *multi);
void QPB_MULTI_SHARD(const QPBatchMulti *multi, qp_int count, int shard, qp_int *lo, qp_int *hi);
qp_int QP_SOLVE_BATCH_MULTI(QPBatchMulti *multi, qp_int count, const settings *opts, const qp_real *Ppr, const qp_real *Apr, const qp_real *Gpr, const qp_real *c, const qp_real *h, const qp_real *b, qp_real *x, qp_real *y, qp_real *z, qp_real *s, qp_int *flag, qp_int *iters);
qp_int QP_SOLVE_BATCH_MULTI_DEVICE(QPBatchMulti *multi, qp_int count, const settings *opts, const qp_real *const *dPpr, const qp_real *const *dApr, const qp_real *const *dGpr, const qp_real *const *dc, const qp_real *const *dh, const qp_real *const *db, qp_real *const *dx, qp_real *const *dy, qp_real *const *dz, qp_real *const *ds, int *const *dflag, int *const *diters, qp_real *const *gx, int *const *gflag, int *const *giters);
void QP_CLEANUP_BATCH_MULTI(QPBatchMulti *multi);
const char *QPB_MULTI_LAST_ERROR(void);

/* Device time (ms, CUDA events) of the kernels of the last QP_SOLVE_BATCH call and the number
 * of kernel launches it made. */
/* fval[count] and stat[count][8] = {n_rx, n_ry, n_rz, n_mu, alpha_p, alpha_d, sigma, mu} of the last
 * host-pointer QP_SOLVE_BATCH call (what QP_SOLVE stores into `stats`, qpSWIFT.c:507-596). */
qp_int QPB_BATCH_READ_STATS(QPBatch *batch, qp_int count, qp_real *fval, qp_real *stat);

/* sigma_d belongs to a QP, not to its pattern: change it between solves. */
qp_int QPB_BATCH_SET_SIGMA_D(QPBatch *batch, qp_real sigma_d);

/* Per-iteration trace of the next host-pointer solves on this batch (small capacities only: the warp-per-QP
 * kernel keeps it, the throughput kernel does not).  QPB_BATCH_READ_TRACE copies trace[count][maxit + 2][8]:
 * row 0 = {seconds in LDL_numeric (stats->ldl_numeric, Auxilary.c:476-484), 0...}; row 1 + it =
 * {fval, n_rx, n_ry, n_rz, n_mu at the top of iteration it (qpSWIFT.c:513-516), alpha_p, alpha_d of its step
 * (qpSWIFT.c:598-601), 0}.  QP_SOLVE uses it for verbose > 0. */
qp_int QPB_BATCH_ENABLE_TRACE(QPBatch *batch, int on);
qp_int QPB_BATCH_READ_TRACE(QPBatch *batch, qp_int count, qp_int maxit, qp_real *trace);

double QPB_BATCH_LAST_KERNEL_MS(const QPBatch *batch);
qp_int QPB_BATCH_LAST_LAUNCHES(const QPBatch *batch);
/* Which kernel the last solve used: 0 = warp-per-QP kernel (one QP per warp, working set in shared memory)
 * interpreting the pattern's wave schedules, 1 = throughput kernel (one QP per thread, working sets
 * interleaved in HBM) interpreting the pattern's micro-programs, 2 = the throughput kernel specialised for the
 * pattern (straight-line code, see QPB_PLAN_STREAM_KEY), 3 = the warp-per-QP kernel specialised for the pattern
 * (lane = row of L, factor resident in shared memory).  Dispatch: where the library carries a specialised warp-per-QP
 * kernel for the pattern (DogBot stance / trot / crawl) it takes every batch; otherwise the throughput kernel takes
 * batches of at least QPB_STREAM_MIN QPs [default 2048] and the interpreted warp-per-QP kernel everything smaller. */
int QPB_BATCH_LAST_KERNEL(const QPBatch *batch);

/* ---------------------------------------------------------------------------------------------
 * Kernel-level entry points (device pointers), used by the parity tests to check each
 * reference routine on its own and by bench.py to time the factor / solve kernels alone.
 * ------------------------------------------------------------------------------------------- */

/* LDL_numeric (ldl.c:253-326) for `count` QPs: vals[count][nV] value rows, zdiag[count][m]
 * the (z,z) diagonal (e.g. -1 or -s/z); writes Lx[count][lnz], D[count][N]. */
qp_int QPB_LDL_NUMERIC_BATCH(QPBatch *batch, qp_int count, const qp_real *dvals, const qp_real *dzdiag, qp_real *dLx, qp_real *dD, void *stream);

/* LDL_perm, lsolve, dsolve, ltsolve, permt (ldl.c:495-597) for `count` QPs:
 * rhs[count][N] -> sol[count][N] (may alias). */
qp_int QPB_LDL_SOLVE_BATCH(QPBatch *batch, qp_int count, const qp_real *dLx, const qp_real *dD, const qp_real *drhs, qp_real *dsol, void *stream);

/* QP_SETUP caches the symbolic plan (and the single-QP batch) per sparsity pattern; the cache is bounded and
 * evicts the least recently used plan.  what = 0: plans cached now, 1: the cache's capacity. */
qp_int QPB_PLAN_CACHE_INFO(int what);

/* Measured FP64 peak of `device` (< 0: the current one) in TFLOP/s: 8 independent chains a <- a * b + c per thread as
 * one DFMA (fused != 0) or as DMUL + DADD -- the form this library executes, since unfused arithmetic is part of
 * the bit-exactness contract.  The denominator of the FP64-pipe roofline. */
double QPB_FP64_PEAK(int device, int fused);

const char *QPB_LAST_ERROR(void);
int QPB_DEVICE_COUNT(void);

#ifdef __cplusplus
}
#endif
#endif /* QPSWIFT_B200_H */

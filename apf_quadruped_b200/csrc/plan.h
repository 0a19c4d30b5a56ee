// Host-side symbolic plan of one QP sparsity pattern and its device image.
//
// Everything integer that the reference recomputes on every QP_SETUP / every LDL_numeric call
// (KKT pattern, AMD permutation, elimination tree, column counts, row indices of L, the
// topological row patterns) is value independent, so it is computed ONCE per pattern here and
// shared by every QP of a batch.  Reference functions covered (dogbot_controller/src/qpSWIFT):
//   formkktmatrix_full  Auxilary.c:71-181      -> Plan::Kjc/Kir (+ Ksrc: where each value comes from)
//   amd_l_order         amd_order.c:21-199     -> Plan::P        (amd_order.cpp)
//   LDL_symbolic        ldl.c:187-240          -> Plan::Pinv/Parent/Lnz/Lp
//   LDL_numeric (integer half) ldl.c:276-298,315 -> Plan::Li, row patterns, store slots
//   SparseMatrixTranspose Auxilary.c:901-951   -> row-wise views used by the residual kernels
#ifndef QPB_PLAN_H
#define QPB_PLAN_H

#include <cstdint>
#include <vector>

namespace qpb {

enum { AMD_STATUS_OK = 0, AMD_STATUS_JUMBLED = 1, AMD_STATUS_NOMEM = -1, AMD_STATUS_INVALID = -2 };

// amd_order.cpp -- returns AMD_STATUS_*; perm[k] = original index of the k-th pivot.
int amd_order(long n, const long *colptr, const long *rowidx, long *perm, double dense_ratio = 10.0, int aggressive = 1);

// Offsets (in int32 words) of the sections of the device plan blob.  The blob is one
// contiguous int32 array so that a CTA can pull it into shared memory with one bulk copy.
struct DevicePlanHeader
{
	int32_t n, m, p, N;
	int32_t nnzP, nnzA, nnzG, nV; // nV = nnzP + nnzA + nnzG (length of a QP's value row)
	int32_t lnz;                  // entries of L
	int32_t nscat;                // upper-triangular KKT entries (what the factorisation reads)
	int32_t words;                // total blob length
	// sections ------------------------------------------------------------------------------
	int32_t perm;      // [N]   P[k]
	int32_t lp;        // [N+1] column pointers of L
	int32_t li;        // [lnz] row indices of L
	int32_t scat_ptr;  // [N+1] per pivot row k: range in scat_* (entries of column P[k] with Pinv[row] <= k)
	int32_t scat_dst;  // [nscat] Y slot (permuted row index)
	int32_t scat_src;  // [nscat] value source: < nV -> value row index; >= nV -> z-diagonal of inequality (src - nV)
	int32_t pat_ptr;   // [N+1] per pivot row k: range in pat_* (topological row pattern)
	int32_t pat_col;   // [lnz] column i of L touched
	int32_t pat_slot;  // [lnz] where L(k,i) is stored (= Lp[i] + number of earlier rows in column i)
	// row/column views of P, A, G used by the residual and initial-point code -------------------
	int32_t prow_ptr;  // [n+1]  row r of P: entries in ascending column order
	int32_t prow_col;  // [nnzP]
	int32_t prow_val;  // [nnzP] value row index
	int32_t arow_ptr;  // [p+1]
	int32_t arow_col;  // [nnzA]
	int32_t arow_val;  // [nnzA]
	int32_t grow_ptr;  // [m+1]
	int32_t grow_col;  // [nnzG]
	int32_t grow_val;  // [nnzG]
	int32_t acol_ptr;  // [n+1]  column c of A: entries in stored order (value index = nnzP + position)
	int32_t acol_row;  // [nnzA]
	int32_t gcol_ptr;  // [n+1]
	int32_t gcol_row;  // [nnzG]
	int32_t reserved[4];
};

struct Plan
{
	long n = 0, m = 0, p = 0, N = 0;
	long nnzP = 0, nnzA = 0, nnzG = 0;
	int amd_status = 0;      // what AMD returned (>= 0), -3 when the caller supplied the permutation
	bool valid = false;      // false: an inequality row is empty (the reference misbehaves there too)

	// full symmetric KKT pattern exactly as the reference assembles it
	std::vector<long> Kjc, Kir;
	std::vector<int32_t> Ksrc; // per stored entry: value row index, or nV + i for the (z_i, z_i) diagonal

	std::vector<long> P, Pinv, Parent, Lnz, Lp, Li;
	std::vector<long> Ltp, Lti; // Transpose_Row_Count (Auxilary.c:683-732); API compatibility only

	// row patterns of LDL_numeric: for pivot row k, pat_col[pat_ptr[k] .. pat_ptr[k+1]) in the
	// order the reference visits them, and the slot each L(k,i) lands in
	std::vector<int32_t> pat_ptr, pat_col, pat_slot;
	std::vector<int32_t> scat_ptr, scat_dst, scat_src;

	std::vector<int32_t> blob; // DevicePlanHeader followed by the sections
	uint64_t pattern_hash = 0; // FNV-1a over (n,m,p,jc/ir of P,A,G, permutation source)

	long factor_flops() const; // sum Lnz_k (Lnz_k + 2), ldl.c:182-184
};

// Build the plan from CCS patterns.  A* may be null with p == 0.  user_perm may be null (AMD).
// Values are NOT needed.  Returns false (plan.valid == false) if the pattern is unusable.
bool build_plan(Plan &plan, long n, long m, long p, const long *Pjc, const long *Pir, const long *Ajc, const long *Air, const long *Gjc, const long *Gir, const long *user_perm);

uint64_t fnv1a_words(const long *a, long count, uint64_t h = 1469598103934665603ULL);

} // namespace qpb

#endif

// Static dataflow programs: the numeric factorisation and the triangular solves of one sparsity
// pattern, compiled once on the host into a schedule that a warp executes for every QP of a batch.
//
// Why: in LDL_numeric (dogbot_controller/src/qpSWIFT/ldl.c:253-326) and the LDL_lsolve / dsolve /
// ltsolve sweeps (ldl.c:495-557) every index, every loop bound and the order of every
// floating-point operation depend on the sparsity pattern only.  Each number the reference
// produces is the end of a CHAIN
//        t = init;  t = t - a1*b1;  t = t - a2*b2;  ...          (terms in the reference's order)
// (an entry of the row-k triangular solve, a pivot D[k], a row of the forward sweep, a column of
// the backward sweep) or a QUOTIENT a/b (L(k,i) = y_i / D[i], w[j] / D[j]).  Chains that do not
// depend on each other may run at the same time without changing a single rounding, so the host
// list-schedules all chains of a routine onto the 32 lanes of a warp: one WAVE = one term of up
// to 32 different chains (or up to 32 quotients), the accumulator of a chain stays in its lane's
// register from the first term to the last, and a __syncwarp() separates waves.  The result is
// bit-identical with the sequential reference and needs N_waves ~ max(critical path, work/32)
// steps instead of one step per entry of L.
//
// The program addresses one QP's "slice" of shared memory by double index (14 bits).
#ifndef QPB_PROGRAM_H
#define QPB_PROGRAM_H

#include <cstdint>
#include <string>
#include <vector>

#include "plan.h"

namespace qpb {

// Layout of one QP's slice of shared memory, in doubles from the slice base.  Every offset is even
// (16-byte aligned).  Shared by the host compiler (addresses inside instruction words) and the kernels.
struct SliceLayout
{
	int32_t zero;           // [2]  constants +0.0 (operand of idle lanes) and -0.0 (exact additive identity)
	int32_t V, c, h, b;     // value row: P|A|G values [nV], c[n], h[m], b[p]
	int32_t zd;             // [m]  (z,z) diagonal of the KKT matrix: -1 (initial point) or -s/z
	int32_t Lx, D;          // factor [lnz], pivots [N]
	int32_t T;              // scratch: temporaries of the factorisation / work vectors of the solves
	int32_t rhs;            // [N]  right-hand side in, solution out (original ordering)
	int32_t x, y, z, s;     // iterate
	int32_t r;              // [N]  residuals rx|ry|rz
	int32_t lam, ds, dz, dsl;
	int32_t total;          // doubles per slice
	int32_t tsize;          // doubles in T
};

// Instruction word, one per lane per wave (64 bit = four 16-bit fields, index in the low 14 bits):
//   field 0: a   | FIRST << 14 | LAST << 15          (low half of the low word)
//   field 1: ci  | NEGZ  << 14 | ADD  << 15
//   field 2: b   |      0      | ACTIVE << 15        (low half of the high word)
//   field 3: cs  | SUMWAVE << 14 | DIVWAVE << 15
// multiply-subtract wave:  if FIRST: t = S[ci] + (NEGZ ? -0.0 : +0.0);   t = t -/+ S[a]*S[b];
//                          if LAST:  S[cs] = t
// quotient wave (DIVWAVE set in every lane): if ACTIVE: S[cs] = S[a] / regularise(S[b])
//   -- every divisor of the factorisation and of the solves is a pivot D[i]; the dynamic
//   regularisation of ldl.c:319-320 is idempotent, so it is applied where D is USED and the
//   chains store the raw pivot (the stand-alone factor kernel regularises D on the way out).
// ordered-sum wave (SUMWAVE set in every lane): if ACTIVE: t = S[ci]; for i < b: t = t - S[a+i]; S[cs] = t
//   (field 2 holds the COUNT b, not an address)
enum : uint64_t
{
	INS_FIRST = 1ull << 14, INS_LAST = 1ull << 15,
	INS_NEGZ = 1ull << 30, INS_ADD = 1ull << 31,
	INS_ACTIVE = 1ull << 47,
	INS_SUMWAVE = 1ull << 62, INS_DIVWAVE = 1ull << 63
};
static const int INS_ADDR_BITS = 14;
static const int PROGRAM_WAVE_PAD = 6; // segments are padded to a multiple of this many super-waves (kernel prefetch depth)

struct SegmentStats
{
	int32_t waves = 0, mul_waves = 0, div_waves = 0, sum_waves = 0;
	int64_t terms = 0, divs = 0, sums = 0, chains = 0, suspensions = 0;
	int64_t lane_slots = 0;   // issued term slots incl. stalls of running chains
	int32_t temps = 0;        // scratch doubles used
};

struct Segment
{
	int32_t first = 0; // offset of the segment in Program::words, in units of 32 words
	int32_t waves = 0; // padded super-wave count
	int32_t width = 1; // instruction words per lane per super-wave (1, 2 or 4)
	SegmentStats stats;
};

struct Program
{
	SliceLayout lay;
	std::vector<uint64_t> words; // per segment [super-wave][lane][slot]
	Segment factor, solve;
	bool ok = false;
	std::string error;
};

// Compile the factor and solve programs of `plan`.  Returns false (and sets error) when the pattern
// cannot be handled: slice larger than the 14-bit address space, duplicate KKT entries.
bool build_program(const Plan &plan, Program &prog);

} // namespace qpb

#endif

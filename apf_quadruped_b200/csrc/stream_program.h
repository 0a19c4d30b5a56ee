// Stream programs: the numeric factorisation, the triangular solves and the residual products of
// one sparsity pattern as straight-line micro-programs for the THROUGHPUT kernel (stream.cuh), in
// which one THREAD owns one QP and the 32 lanes of a warp execute the same micro-op on 32 different
// QPs whose working sets are interleaved in HBM (element e of lane l at base + e*32 + l: every
// access of a warp is one 256-byte line).
//
// Each thread literally performs the reference's floating-point operations in the reference's order
// (dogbot_controller/src/qpSWIFT/ldl.c:253-326 LDL_numeric, 495-597 the solves, Auxilary.c:745-786
// computeresiduals), so there is nothing to schedule for parity; what the host compiles in is
// (1) every index (the kernel never touches Lp/Li/Pattern), (2) where a value lives -- the row-local
// work vector Y of LDL_numeric and the work vector of the solves sit in a small per-lane window of
// SHARED memory, everything else in the lane's slice of HBM -- and (3) which HBM operands can be
// fetched ahead.  Micro-ops are executed in GROUPS; the kernel requests all HBM operands of group
// g+DEPTH-1 before it executes the arithmetic of group g, so an operand that is stored by one of the
// DEPTH-1 previous groups or by an earlier micro-op of its own group is marked LATE (loaded at
// execution time instead).  A group is either
// FAST -- only multiply-subtracts and chain starts, executed branch-free -- or SLOW (quotients, pivots,
// copies, late operands: interpreted micro-op by micro-op); the compiler pads with no-ops so that
// the long runs of multiply-subtracts stay in fast groups.
#ifndef QPB_STREAM_PROGRAM_H
#define QPB_STREAM_PROGRAM_H

#include <cstdint>
#include <string>
#include <vector>

#include "plan.h"

namespace qpb {

// One lane's slice of HBM, offsets in doubles (element e lives at slice_base + e * 32 + lane).
struct StreamLayout
{
	int32_t zero, one, mone; // constants +0.0, 1.0 and -1.0
	int32_t V, c, h, b;     // value row: P|A|G values [nV] (normalised v + 0.0), c[n], h[m], b[p]
	int32_t zd;             // [m]  (z,z) diagonal of the KKT matrix: -1 (initial point) or -s/z
	int32_t Lx, D;          // factor [lnz], pivots [N] (regularised, ldl.c:319-320)
	int32_t rhs;            // [N]  right-hand side in, solution out (original ordering)
	int32_t x, y, z, s;     // iterate
	int32_t r;              // [N]  residuals rx|ry|rz
	int32_t lam, ds, dz, dsl;
	int32_t px;             // [n]  -(P x), kept for the objective value
	int32_t hs;             // [m]  s - h going into the residual program (-h for the initial point); h - G x out of RESZ
	int32_t nb;             // [p]  -b
	int32_t total;          // doubles per slice
	int32_t scratch;        // doubles of shared memory per lane
};

// Micro-op word (64 bit):  f1 (16) | f2 << 16 | c << 32 | op << 48 (4 bits) | flags << 52
//   f1  index of the PREFETCHED operand P in the lane's HBM slice (always fetched with its group; slots
//       without one point at a constant)
//   f2  second operand: an index S into the lane's window of shared memory, read at execution time (the
//       row-local Y / the work vector in the factor and solve programs, "FS"; a copy of x | y | z in the
//       residual programs, "GS", whose results go to HBM)
//   c   where the result goes (flags say whether it does, and into which space)
// Every lane keeps two registers: t (chain accumulator) and d (the pivot being accumulated).
// FAST groups hold only SOP_TERM words; their whole behaviour is in four flags:
//       r = (FIRSTZ ? +0.0 : t) - P*(ONE ? 1.0 : S);   t = INIT ? P : r;   if ST_SHARED (FS) / ST_HBM (GS): [c] = t
// (padding in a fast group is a TERM on the constant +0.0: it may clobber t, which is dead there).
enum StreamOp
{
	SOP_TERM = 0,   // see above
	SOP_DIVACC = 2, // t = S / P;  d = d - t*S          (l = y/D[i];  D[k] -= l*y,  ldl.c:305-309)
	SOP_DIV = 3,    // t = S / P
	SOP_INITD = 4,  // d = P
	SOP_ENDD = 5,   // t = regularise(d)                 (ldl.c:319-320)
	SOP_LOADS = 6,  // t = S
	SOP_NOP = 8     // t unchanged (padding of a slow group, or a pure store of t)
};
enum : uint64_t
{
	SF_FIRSTZ = 1ull << 52,
	SF_INIT = 1ull << 53,
	SF_ST_SHARED = 1ull << 54, // shared window[c] = t
	SF_ST_HBM = 1ull << 55,    // HBM slice[c] = t
	SF_LATE1 = 1ull << 56,     // P cannot be fetched ahead (stored too recently): fetch it at execution time
	SF_ONE = 1ull << 57,       // the second operand is the constant 1.0 (t = t - P: "+= c*(-1)", "+= b", "+= h - s")
	SF_SLOW = 1ull << 58       // set in EVERY word of a slow group
};
// Pipeline depth of the interpreter = groups per instruction chunk = operand buffers: the operands
// of group g + DEPTH - 1 are requested before group g executes.  Factor / solve programs carry one
// HBM operand per micro-op; 32 micro-ops are in the pipeline (4 groups of 8).
#define STREAM_FS_DEPTH(group) (32 / (group))
enum { SSEG_FACTOR = 0, SSEG_SOLVE = 1, SSEG_RESID = 2, SSEG_RESZ = 3, SSEG_COUNT = 4 };

struct StreamSegment
{
	int32_t first = 0;  // offset in StreamProgram::words
	int32_t groups = 0; // groups of `group` micro-ops
	int64_t ops = 0, terms = 0, divs = 0, late = 0, hbm_loads = 0, hbm_stores = 0, nops = 0, slow_groups = 0;
};

struct StreamProgram
{
	StreamLayout lay;
	int32_t group = 8; // micro-ops per group
	std::vector<uint64_t> words; // segments, each padded to a multiple of its pipeline depth in groups; 16 groups of padding at the end
	StreamSegment seg[SSEG_COUNT];
	bool ok = false;
	std::string error;
};

bool build_stream_program(const Plan &plan, StreamProgram &prog, int group = 8);

// FNV-1a over the layout and the words: the name under which a pattern-specialised kernel generated
// from this very program (apf_quadruped_b200/codegen.py) registers itself (stream_gen.cuh).
uint64_t stream_program_key(const StreamProgram &prog);

} // namespace qpb

#endif

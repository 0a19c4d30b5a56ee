// Host compiler of the stream programs described in stream_program.h.
//
// Reference order restated (dogbot_controller/src/qpSWIFT):
//   LDL_numeric ldl.c:276-322.  Row k scatters column P[k] of K into Y ("Y[i] += value" from
//     Y[i] == 0), then for every column i of Pattern[top..] in that order:  yi = Y[i];
//     Y[Li[q]] -= Lx[q]*yi for the entries q column i already holds;  l = yi/D[i];  D[k] -= l*yi.
//     Hence  y(k,j) = (0 + K(k,j)) - sum_{i before j in Pattern, L(j,i) != 0} L(j,i)*y(k,i), terms in
//     Pattern order -- one accumulator per j, finished before column j itself is visited.
//   LDL_lsolve 495-513   w[r] = b[P[r]] - sum_{j<r} L(r,j) w[j], j ascending (column sweeps deliver the
//     terms of row r in ascending column order);  LDL_dsolve 520-532  w[j] /= D[j];
//   LDL_ltsolve 539-557  w[j] -= Lx[q]*w[Li[q]], q ascending inside column j, j descending.
//   computeresiduals Auxilary.c:745-786 (column scatter for P x, A x, G x; column gather for G'z, A'y);
//     "+= c*(-1)", "+= b*1", "+= h - s" are exact as  t - c,  t - (-b),  t - (s-h).

#include "stream_program.h"

#include <algorithm>
#include <cstring>

namespace qpb {
namespace {

enum { F2_NONE = 0, F2_SHARED = 1, F2_ONE = 2 };
enum { ST_NONE = 0, ST_SHARED = 1, ST_HBM = 2 };

struct Emitter
{
	StreamProgram &prog;
	StreamSegment *seg = nullptr;
	const int group;
	bool gs = false;        // residual program: window = copy of x | y | z, results to HBM
	int depth = 4; // pipeline depth of the interpreter for this kind of program
	size_t seg_start = 0;
	bool cur_slow = false;  // class of the group being filled
	std::vector<int64_t> last_store; // per slice double: position (within the segment) of the last HBM store, -1 none
	std::vector<int64_t> last_window; // the same for the doubles of the shared window
	bool overflow = false;

	Emitter(StreamProgram &p) : prog(p), group(p.group) {}

	int64_t pos() const { return (int64_t)(prog.words.size() - seg_start); }

	void begin(int which, bool residual)
	{
		seg = &prog.seg[which];
		*seg = StreamSegment();
		seg->first = (int32_t)prog.words.size();
		seg_start = prog.words.size();
		gs = residual;
		depth = STREAM_FS_DEPTH(group);
		cur_slow = false;
		last_store.assign((size_t)prog.lay.total, -1);
		last_window.assign((size_t)prog.lay.total, -1);
	}

	// The kernel requests the HBM operands of group g+depth-1 before it executes group g: an operand
	// stored by the depth-1 previous groups or earlier in this one is not there yet.
	bool is_late(int idx, int64_t at) const
	{
		const int64_t window_start = at - at % group - (int64_t)(depth - 1) * group;
		return last_store[idx] >= 0 && last_store[idx] >= window_start;
	}

	void raw(uint64_t w, bool slow)
	{
		if (pos() % group == 0) cur_slow = slow;
		prog.words.push_back(w);
	}

	void pad()
	{
		while (pos() % group != 0) {
			if (cur_slow) raw(((uint64_t)SOP_NOP << 48) | SF_SLOW | (uint64_t)prog.lay.zero, true);
			else raw((uint64_t)prog.lay.zero, false); // TERM on +0.0: t is dead at every padding point
			++seg->nops;
			++seg->ops;
		}
	}

	void mark_group_slow()
	{
		const size_t start = prog.words.size() - (size_t)(pos() % group);
		for (size_t i = start; i < prog.words.size(); ++i) prog.words[i] |= SF_SLOW;
		cur_slow = true;
	}

	// f1 < 0: no prefetched operand (the word points at the constant +0.0)
	void op(int code, int f1, int f2kind, int f2, int store, int c, bool firstz = false, bool init = false)
	{
		if (f1 < 0) f1 = prog.lay.zero;
		if (f2kind != F2_SHARED) f2 = 0;
		if (store == ST_NONE) c = 0;
		if (f1 > 0xffff || f2 < 0 || f2 > 0xffff || c < 0 || c > 0xffff) overflow = true;
		for (int attempt = 0;; ++attempt) {
			const int64_t at = pos();
			const bool late1 = is_late(f1, at);
			const bool fast = code == SOP_TERM && !late1 && store != (gs ? ST_SHARED : ST_HBM);
			const int filled = (int)(at % group);
			if (filled != 0 && attempt == 0) {
				if (fast && cur_slow) { pad(); continue; } // leave a slow group
				if (!fast && !cur_slow) {                  // leave a fast group
					if (filled <= 2) mark_group_slow();
					else { pad(); continue; }
				}
			}
			uint64_t w = ((uint64_t)code << 48) | (uint64_t)f1 | ((uint64_t)f2 << 16) | ((uint64_t)c << 32);
			if (firstz) w |= SF_FIRSTZ;
			if (init) w |= SF_INIT;
			if (store == ST_SHARED) w |= SF_ST_SHARED;
			if (store == ST_HBM) w |= SF_ST_HBM;
			if (f2kind == F2_ONE) w |= SF_ONE;
			if (late1) { w |= SF_LATE1; ++seg->late; } else ++seg->hbm_loads;
			const bool slow = filled == 0 ? !fast : cur_slow;
			if (slow) w |= SF_SLOW;
			raw(w, slow);
			++seg->ops;
			if (code == SOP_TERM && !init) ++seg->terms;
			if (code == SOP_DIV || code == SOP_DIVACC) ++seg->divs;
			if (store == ST_HBM) { last_store[c] = at; ++seg->hbm_stores; }
			if (store == ST_SHARED) last_window[c] = at;
			return;
		}
	}

	// A fast group reads all its window operands before it stores any result, so a chain must not
	// read, in the group it starts in, what an earlier chain stored in that group.  Called where a
	// chain starts (the accumulator is dead there, padding may clobber it).
	void fresh_group_if_reads(const std::vector<int> &window_slots)
	{
		const int64_t at = pos();
		const int filled = (int)(at % group);
		if (filled == 0 || cur_slow) return; // a slow group executes micro-op by micro-op; a fast op leaves it anyway
		for (int s : window_slots)
			if (last_window[s] >= at - filled) { pad(); return; }
	}

	void end()
	{
		pad();
		while ((pos() / group) % depth != 0) { // the kernel runs `depth` groups per turn of its pipeline
			raw((uint64_t)prog.lay.zero, false);
			++seg->nops;
			++seg->ops;
			pad();
		}
		seg->groups = (int32_t)(pos() / group);
		for (int64_t g = 0; g < seg->groups; ++g)
			if (prog.words[seg_start + (size_t)g * group] & SF_SLOW) ++seg->slow_groups;
	}
};

struct Term
{
	int a, b; // HBM index of the matrix entry, index of the vector entry (shared window or HBM, per program)
};

// One accumulator: t = init (HBM index, or -1 for +0.0), minus the terms in order; the last micro-op
// stores the result.  All of it fast-path words.
void emit_chain(Emitter &E, int init, const std::vector<Term> &terms, int f2kind, int store, int c)
{
	if (f2kind == F2_SHARED && !E.gs) {
		std::vector<int> reads;
		for (const Term &t : terms) reads.push_back(t.b);
		E.fresh_group_if_reads(reads);
	}
	bool started = false;
	if (init >= 0 || terms.empty()) { // "+0.0 with nothing to subtract" starts from the constant
		E.op(SOP_TERM, init, F2_NONE, 0, terms.empty() ? store : ST_NONE, c, false, true);
		started = true;
	}
	for (size_t i = 0; i < terms.size(); ++i) {
		const bool last = i + 1 == terms.size();
		E.op(SOP_TERM, terms[i].a, f2kind, terms[i].b, last ? store : ST_NONE, c, !started);
		started = true;
	}
}

int even(long v) { return (int)((v + 1) & ~1L); }

} // namespace

bool build_stream_program(const Plan &pl, StreamProgram &prog, int group)
{
	prog = StreamProgram();
	prog.group = group;
	if (!pl.valid || pl.blob.empty()) { prog.error = "unusable plan"; return false; }
	const int n = (int)pl.n, m = (int)pl.m, p = (int)pl.p, N = (int)pl.N;
	const int nV = (int)(pl.nnzP + pl.nnzA + pl.nnzG);
	const int offA = (int)pl.nnzP, offG = (int)(pl.nnzP + pl.nnzA);
	const int lnz = (int)pl.Lp[N];
	StreamLayout &L = prog.lay;
	int at = 0;
	auto take = [&](long count) { const int here = at; at += even(count); return here; };
	L.zero = take(4); L.one = L.zero + 1; L.mone = L.zero + 2;
	L.V = take(nV); L.c = take(n); L.h = take(m); L.b = take(p);
	L.zd = take(m);
	L.Lx = take(lnz); L.D = take(N);
	L.rhs = take(N);
	L.x = take(n); L.y = take(p); L.z = take(m); L.s = take(m);
	L.r = take(N);
	L.lam = take(m); L.ds = take(m); L.dz = take(m); L.dsl = take(m);
	L.px = take(n);
	L.hs = take(m);
	L.nb = take(p);
	L.total = at;
	if (L.total > 0xffff) { prog.error = "pattern too large: a QP's working set exceeds 65535 doubles"; return false; }

	const DevicePlanHeader *hd = reinterpret_cast<const DevicePlanHeader *>(pl.blob.data());
	const int32_t *blob = pl.blob.data();
	const int32_t *prow_ptr = blob + hd->prow_ptr, *prow_col = blob + hd->prow_col, *prow_val = blob + hd->prow_val;
	const int32_t *arow_ptr = blob + hd->arow_ptr, *arow_col = blob + hd->arow_col, *arow_val = blob + hd->arow_val;
	const int32_t *grow_ptr = blob + hd->grow_ptr, *grow_col = blob + hd->grow_col, *grow_val = blob + hd->grow_val;
	const int32_t *acol_ptr = blob + hd->acol_ptr, *acol_row = blob + hd->acol_row;
	const int32_t *gcol_ptr = blob + hd->gcol_ptr, *gcol_row = blob + hd->gcol_row;

	Emitter E(prog);
	int maxpat = 0;

	// ---- LDL_numeric ---------------------------------------------------------------------------------
	E.begin(SSEG_FACTOR, false);
	{
		std::vector<int> row_init(N, -1), where(N, -1);
		std::vector<std::vector<Term>> terms(N);
		for (int k = 0; k < N; ++k) {
			int diag_init = -1;
			for (int s = pl.scat_ptr[k]; s < pl.scat_ptr[k + 1]; ++s) {
				const int dst = pl.scat_dst[s], src = pl.scat_src[s];
				const int addr = src < nV ? L.V + src : L.zd + (src - nV);
				int &slot = dst == k ? diag_init : row_init[dst];
				if (slot >= 0) { prog.error = "duplicate entries in the KKT pattern"; return false; }
				slot = addr;
			}
			const int p0 = pl.pat_ptr[k], p1 = pl.pat_ptr[k + 1];
			maxpat = std::max(maxpat, p1 - p0);
			for (int t = p0; t < p1; ++t) { where[pl.pat_col[t]] = t - p0; terms[pl.pat_col[t]].clear(); }
			for (int t = p0; t < p1; ++t) {
				const int i = pl.pat_col[t];
				for (long q = pl.Lp[i]; q < pl.pat_slot[t]; ++q) { // entries of column i stored by earlier rows
					const int j = (int)pl.Li[q];
					if (where[j] < 0 || where[j] <= t - p0) { prog.error = "inconsistent plan: update outside the row pattern"; return false; }
					terms[j].push_back({L.Lx + (int)q, t - p0});
				}
			}
			for (int t = p0; t < p1; ++t) emit_chain(E, row_init[pl.pat_col[t]], terms[pl.pat_col[t]], F2_SHARED, ST_SHARED, t - p0);
			E.op(SOP_INITD, diag_init, F2_NONE, 0, ST_NONE, 0); // no diagonal entry: d = the constant +0.0
			for (int t = p0; t < p1; ++t) E.op(SOP_DIVACC, L.D + pl.pat_col[t], F2_SHARED, t - p0, ST_HBM, L.Lx + pl.pat_slot[t]);
			E.op(SOP_ENDD, -1, F2_NONE, 0, ST_HBM, L.D + k);
			for (int s = pl.scat_ptr[k]; s < pl.scat_ptr[k + 1]; ++s)
				if (pl.scat_dst[s] != k) row_init[pl.scat_dst[s]] = -1;
			for (int t = p0; t < p1; ++t) where[pl.pat_col[t]] = -1;
		}
	}
	E.end();

	// ---- LDL_perm, lsolve, dsolve, ltsolve, permt -------------------------------------------------------
	// forward chains leave w in the shared window; the quotients w/D go to HBM (in place over the
	// right-hand side, which the forward sweep has consumed) so that the backward chains can start
	// from a prefetched operand; the solution is copied out of the window at the end.
	E.begin(SSEG_SOLVE, false);
	{
		std::vector<std::vector<Term>> rows(N);
		for (int j = 0; j < N; ++j)
			for (long q = pl.Lp[j]; q < pl.Lp[j + 1]; ++q) rows[pl.Li[q]].push_back({L.Lx + (int)q, j});
		for (int r = 0; r < N; ++r) emit_chain(E, L.rhs + (int)pl.P[r], rows[r], F2_SHARED, ST_SHARED, r);
		for (int j = 0; j < N; ++j) {
			const bool leaf = pl.Lp[j] == pl.Lp[j + 1]; // empty column: the quotient is the solution
			E.op(SOP_DIV, L.D + j, F2_SHARED, j, ST_HBM, L.rhs + (int)pl.P[j]);
			if (leaf) E.op(SOP_NOP, -1, F2_NONE, 0, ST_SHARED, j);
		}
		for (int j = N - 1; j >= 0; --j) {
			if (pl.Lp[j] == pl.Lp[j + 1]) continue;
			std::vector<Term> t;
			for (long q = pl.Lp[j]; q < pl.Lp[j + 1]; ++q) t.push_back({L.Lx + (int)q, (int)pl.Li[q]});
			emit_chain(E, L.rhs + (int)pl.P[j], t, F2_SHARED, ST_SHARED, j);
		}
		for (int j = 0; j < N; ++j)
			if (pl.Lp[j] != pl.Lp[j + 1]) E.op(SOP_LOADS, -1, F2_SHARED, j, ST_HBM, L.rhs + (int)pl.P[j]);
	}
	E.end();

	// ---- computeresiduals (+ the -(P x) of obj_value) ---------------------------------------------------
	// The kernel copies x | y | z into the window first: every x[j], y[i], z[i] is used once per nonzero
	// of its column / row, so only the matrix values travel from HBM.
	const int wx = 0, wy = n, wz = n + p;
	auto rz_chains = [&](bool in_place) {
		for (int i = 0; i < m; ++i) {
			std::vector<Term> t;
			for (int q = grow_ptr[i]; q < grow_ptr[i + 1]; ++q) t.push_back({L.V + grow_val[q], wx + grow_col[q]});
			emit_chain(E, -1, t, F2_SHARED, ST_NONE, 0);
			E.op(SOP_TERM, L.hs + i, F2_ONE, 0, ST_HBM, in_place ? L.hs + i : L.r + n + p + i); // += (h - s)
		}
	};
	E.begin(SSEG_RESID, true);
	for (int i = 0; i < n; ++i) {
		// rx[i] = 0 - (P x)_i [kept as px[i] for obj_value, Auxilary.c:1136-1140] - (G'z)_i - (A'y)_i - c[i]
		std::vector<Term> t;
		for (int q = prow_ptr[i]; q < prow_ptr[i + 1]; ++q) t.push_back({L.V + prow_val[q], wx + prow_col[q]});
		emit_chain(E, -1, t, F2_SHARED, ST_HBM, L.px + i);
		for (int q = gcol_ptr[i]; q < gcol_ptr[i + 1]; ++q) E.op(SOP_TERM, L.V + offG + q, F2_SHARED, wz + gcol_row[q], ST_NONE, 0);
		if (p)
			for (int q = acol_ptr[i]; q < acol_ptr[i + 1]; ++q) E.op(SOP_TERM, L.V + offA + q, F2_SHARED, wy + acol_row[q], ST_NONE, 0);
		E.op(SOP_TERM, L.c + i, F2_ONE, 0, ST_HBM, L.r + i);
	}
	for (int i = 0; i < p; ++i) {
		std::vector<Term> t;
		for (int q = arow_ptr[i]; q < arow_ptr[i + 1]; ++q) t.push_back({L.V + arow_val[q], wx + arow_col[q]});
		emit_chain(E, -1, t, F2_SHARED, ST_NONE, 0);
		E.op(SOP_TERM, L.nb + i, F2_ONE, 0, ST_HBM, L.r + n + i); // += b
	}
	rz_chains(false);
	E.end();

	E.begin(SSEG_RESZ, true);
	rz_chains(true);
	E.end();

	// the kernel fetches words up to four chunks ahead
	for (int i = 0; i < 4 * 32; ++i) prog.words.push_back(((uint64_t)SOP_NOP << 48) | SF_SLOW | (uint64_t)L.zero);
	L.scratch = even(std::max(maxpat, N));
	if (E.overflow) { prog.error = "stream program: operand index out of range"; return false; }
	prog.ok = true;
	return true;
}

uint64_t stream_program_key(const StreamProgram &prog)
{
	uint64_t h = 1469598103934665603ull;
	auto mix = [&](uint64_t v) { h ^= v; h *= 1099511628211ull; };
	const int32_t *lay = reinterpret_cast<const int32_t *>(&prog.lay);
	for (size_t i = 0; i < sizeof(StreamLayout) / sizeof(int32_t); ++i) mix((uint64_t)(uint32_t)lay[i]);
	mix((uint64_t)prog.group);
	for (int s = 0; s < SSEG_COUNT; ++s) { mix((uint64_t)prog.seg[s].first); mix((uint64_t)prog.seg[s].groups); }
	for (uint64_t w : prog.words) mix(w);
	return h;
}

} // namespace qpb

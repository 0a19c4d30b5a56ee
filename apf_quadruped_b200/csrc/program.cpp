// Host compiler for the static dataflow programs described in program.h.
//
// Reference order that every chain reproduces (dogbot_controller/src/qpSWIFT/ldl.c):
//   LDL_numeric 276-322: row k scatters column P[k] of K into Y (Y[i] += value, from Y[i] == 0),
//     then walks Pattern[top..] (topological); for each column i of that list
//     "Y[Li[p]] -= Lx[p]*yi" for the entries already stored in column i, "l = yi/D[i]",
//     "D[k] -= l*yi".  Hence   y(k,j) = (0 + K(k,j)) - sum_{i before j in Pattern, L(j,i) != 0} L(j,i)*y(k,i)
//     with the terms in Pattern order,  L(k,i) = y(k,i)/D[i],  D[k] = (0 + K(k,k)) - sum_i L(k,i)*y(k,i)
//     in Pattern order, followed by the dynamic regularisation of 319-320.
//   LDL_lsolve 495-513: w[r] = b[P[r]] - sum_{j<r, L(r,j) != 0} L(r,j)*w[j], j ascending.
//   LDL_dsolve 520-532: w[j] /= D[j].
//   LDL_ltsolve 539-557: w[j] -= Lx[p]*w[Li[p]] for p ascending inside column j, j descending.
//   LDL_perm / LDL_permt 564-597 are folded into the addresses (b[P[r]] read, x[P[j]] written).

#include "program.h"

#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace qpb {
namespace {

const int INF = 1 << 30;
const int TEMP_BASE = 1 << 20;  // value ids >= TEMP_BASE are scheduler-allocated temporaries
const int ALIAS_BASE = 1 << 19; // ids in [ALIAS_BASE, TEMP_BASE): a second value living at a fixed address (solution over rhs)
const int DIV_COST = 2;         // a quotient wave is about twice as long as a multiply-subtract wave
const int SUM_MIN_TERMS = 3;
// Width of a super-wave in warp-widths: a lane executes `width` independent terms between two
// __syncwarp()s, so the shared-memory and FP64 latencies of one term hide behind the others.  The
// factorisation has the parallelism for 4; the triangular solves are critical-path bound.
static int env_width(const char *name, int dflt)
{
	const char *v = getenv(name);
	const int w = v ? atoi(v) : dflt;
	return (w == 1 || w == 2 || w == 4) ? w : dflt;
}
static const int FACTOR_WIDTH = env_width("QPB_FACTOR_WIDTH", 4);
static const int SOLVE_WIDTH = env_width("QPB_SOLVE_WIDTH", 2);    // backward-sweep columns at least this long use products + ordered sum

int sum_cost(int count) { return 2 + count / 8; }

struct Chain
{
	int init = 0;
	bool negz = false; // true: t = init exactly; false: t = 0.0 + init (the reference's "Y[i] += value")
	std::vector<std::pair<int, int>> terms;
	std::vector<int> gates; // values that must be available before the chain may start
	int dst = 0;
	bool reg = false;
	bool add = false;
	int height = 0; // waves from the end of this chain to the end of the program
	// scheduler state
	int pos = 0;       // next term
	int acc = -1;      // value holding the accumulator while the chain is not in a lane (-1: not started, use init)
	bool done = false;
};

// A quotient S[dst] = S[a] / S[b], or an ordered sum S[dst] = ((S[a] - S[b]) - S[b+1]) - ... (count terms).
struct Unit
{
	bool sum = false;
	int dst = 0, a = 0, b = 0, count = 0;
	int prio = 0;
};

struct OpRef
{
	bool is_unit;
	int idx;
};

struct Issue
{
	uint8_t kind = 0; // 0 idle, 1 term, 2 quotient, 3 ordered sum
	bool first = false, last = false, negz = false, reg = false, add = false;
	int a = 0, b = 0, ci = 0, cs = 0;
};

struct Wave
{
	int type = 0; // 0 multiply-subtract, 1 quotient, 2 ordered sum
	std::vector<Issue> lane; // 32 * width virtual lanes: lane l, slot u -> index u * 32 + l
};

struct Builder
{
	std::vector<Chain> chains;
	std::vector<Unit> units;
	std::vector<OpRef> order; // creation order == a topological order
	int ntemps = 0;
	int fixed_size = 0; // size of the fixed address space
	std::vector<int> alias_phys;
	std::vector<std::pair<int, int>> blocks; // (first temp, length): temporaries that must be contiguous
	std::vector<int> block_of;               // per temp

	int new_block(int len)
	{
		const int first = ntemps;
		blocks.push_back({first, len});
		for (int i = 0; i < len; ++i) block_of.push_back((int)blocks.size() - 1);
		ntemps += len;
		return TEMP_BASE + first;
	}
	int new_temp() { return new_block(1); }
	int new_alias(int phys) { alias_phys.push_back(phys); return ALIAS_BASE + (int)alias_phys.size() - 1; }
	int add_chain(Chain &&c)
	{
		chains.push_back(std::move(c));
		order.push_back({false, (int)chains.size() - 1});
		return (int)chains.size() - 1;
	}
	void add_div(int dst, int a, int b)
	{
		Unit d;
		d.dst = dst; d.a = a; d.b = b;
		units.push_back(d);
		order.push_back({true, (int)units.size() - 1});
	}
	void add_sum(int dst, int init, int block, int count)
	{
		Unit d;
		d.sum = true; d.dst = dst; d.a = init; d.b = block; d.count = count;
		units.push_back(d);
		order.push_back({true, (int)units.size() - 1});
	}
};

struct ValueTable
{
	std::vector<int> fixed, alias, temp;
	int fill;
	ValueTable(const Builder &B, int v) : fixed(B.fixed_size, v), alias(B.alias_phys.size(), v), temp(B.ntemps, v), fill(v) {}
	int &at(int id)
	{
		if (id >= TEMP_BASE) {
			if ((size_t)(id - TEMP_BASE) >= temp.size()) temp.resize(id - TEMP_BASE + 1, fill);
			return temp[id - TEMP_BASE];
		}
		return id >= ALIAS_BASE ? alias[id - ALIAS_BASE] : fixed[id];
	}
};

// List scheduling.  Lanes never idle while holding a chain: a chain whose next term is not ready is
// SUSPENDED (its accumulator is stored to a scratch double in the slot of its last issued term and
// reloaded, exactly, by whichever lane resumes it), so no set of waiting chains can block the warp.
// Returns false on an internal inconsistency (never expected).
bool schedule(Builder &B, int NL, int zero_addr, std::vector<Wave> &waves, SegmentStats &st, std::vector<int> &temp_slot, int &temp_count)
{
	// ---- priorities: longest path to the end, in waves ---------------------------------------------
	ValueTable H(B, 0);
	for (size_t o = B.order.size(); o-- > 0;) {
		const OpRef &op = B.order[o];
		if (op.is_unit) {
			Unit &d = B.units[op.idx];
			d.prio = (d.sum ? sum_cost(d.count) : DIV_COST) + H.at(d.dst);
			H.at(d.a) = std::max(H.at(d.a), d.prio);
			if (d.sum)
				for (int i = 0; i < d.count; ++i) H.at(d.b + i) = std::max(H.at(d.b + i), d.prio);
			else
				H.at(d.b) = std::max(H.at(d.b), d.prio);
		} else {
			Chain &c = B.chains[op.idx];
			c.height = H.at(c.dst);
			const int L = (int)c.terms.size();
			for (int j = 0; j < L; ++j) {
				const int need = (L - j) + c.height;
				H.at(c.terms[j].first) = std::max(H.at(c.terms[j].first), need);
				H.at(c.terms[j].second) = std::max(H.at(c.terms[j].second), need);
			}
			const int whole = std::max(L, 1) + c.height;
			H.at(c.init) = std::max(H.at(c.init), whole);
			for (int g : c.gates) H.at(g) = std::max(H.at(g), whole);
		}
	}

	// ---- availability: wave index from which a value may be read -------------------------------------
	ValueTable A(B, 0);
	for (const Chain &c : B.chains) A.at(c.dst) = INF;
	for (const Unit &d : B.units) A.at(d.dst) = INF;
	ValueTable def(B, -1), last_use(B, -1);
	auto use = [&](int v, int wave) { last_use.at(v) = std::max(last_use.at(v), wave); };

	std::vector<int> pend_chain(B.chains.size()), pend_unit(B.units.size());
	for (size_t i = 0; i < pend_chain.size(); ++i) pend_chain[i] = (int)i;
	for (size_t i = 0; i < pend_unit.size(); ++i) pend_unit[i] = (int)i;
	size_t remaining = B.chains.size() + B.units.size();

	std::vector<int> lane_chain(NL);
	for (int l = 0; l < NL; ++l) lane_chain[l] = -1;

	st = SegmentStats();
	st.chains = (int64_t)B.chains.size();
	for (const Unit &d : B.units) (d.sum ? st.sums : st.divs) += 1;
	for (const Chain &c : B.chains) st.terms += (int64_t)c.terms.size();
	for (const Unit &d : B.units) st.terms += d.sum ? d.count : 0; // the subtractions of an ordered sum

	auto term_ready = [&](const Chain &c, int pos, int wave) {
		return A.at(c.terms[pos].first) <= wave && A.at(c.terms[pos].second) <= wave;
	};
	auto unit_ready = [&](const Unit &d, int wave) {
		if (A.at(d.a) > wave) return false;
		if (!d.sum) return A.at(d.b) <= wave;
		for (int i = 0; i < d.count; ++i)
			if (A.at(d.b + i) > wave) return false;
		return true;
	};
	auto chain_prio = [&](const Chain &c) { return std::max((int)c.terms.size() - c.pos, 1) + c.height; };

	int wave = 0;
	std::vector<int> ready_unit[2], cand;
	while (remaining > 0) {
		ready_unit[0].clear();
		ready_unit[1].clear();
		cand.clear();
		int pu[2] = {-1, -1}, pm = -1, nfree = 0;
		bool any_cont = false;
		for (int d : pend_unit)
			if (unit_ready(B.units[d], wave)) {
				const int k = B.units[d].sum ? 1 : 0;
				ready_unit[k].push_back(d);
				pu[k] = std::max(pu[k], B.units[d].prio);
			}
		for (int l = 0; l < NL; ++l) {
			if (lane_chain[l] < 0) { ++nfree; continue; }
			any_cont = true;
			pm = std::max(pm, chain_prio(B.chains[lane_chain[l]]));
		}
		for (int ci : pend_chain) { // not in a lane: never started, or suspended
			const Chain &c = B.chains[ci];
			if (A.at(c.acc >= 0 ? c.acc : c.init) > wave) continue;
			if (c.pos < (int)c.terms.size() && !term_ready(c, c.pos, wave)) continue;
			bool open = true;
			if (c.acc < 0)
				for (int g : c.gates)
					if (A.at(g) > wave) { open = false; break; }
			if (open) cand.push_back(ci);
		}
		if (nfree > 0)
			for (int ci : cand) pm = std::max(pm, chain_prio(B.chains[ci]));
		const bool can_mul = any_cont || (nfree > 0 && !cand.empty());
		if (!can_mul && ready_unit[0].empty() && ready_unit[1].empty()) return false;

		Wave W;
		W.lane.assign(NL, Issue());
		for (int l = 0; l < NL; ++l) { W.lane[l].a = W.lane[l].b = W.lane[l].ci = W.lane[l].cs = zero_addr; }
		int type = 0;
		{
			int best = can_mul ? pm : -1;
			for (int k = 0; k < 2; ++k)
				if (!ready_unit[k].empty() && (pu[k] > best || ready_unit[k].size() >= (size_t)(k == 1 ? NL / 2 : NL))) { best = std::max(best, pu[k]); type = k + 1; }
		}
		W.type = type;
		if (type != 0) {
			std::vector<int> &ready = ready_unit[type - 1];
			std::stable_sort(ready.begin(), ready.end(), [&](int x, int y) { return B.units[x].prio > B.units[y].prio; });
			const size_t take = std::min<size_t>(NL, ready.size());
			std::vector<char> gone(B.units.size(), 0);
			for (size_t l = 0; l < take; ++l) {
				const Unit &d = B.units[ready[l]];
				Issue &is = W.lane[l];
				if (d.sum) {
					is.kind = 3; is.ci = d.a; is.a = d.b; is.b = d.count; is.cs = d.dst;
					use(d.a, wave);
					for (int i = 0; i < d.count; ++i) use(d.b + i, wave);
				} else {
					is.kind = 2; is.a = d.a; is.b = d.b; is.cs = d.dst;
					use(d.a, wave);
					use(d.b, wave);
				}
				A.at(d.dst) = wave + 1;
				def.at(d.dst) = wave;
				gone[ready[l]] = 1;
				--remaining;
			}
			pend_unit.erase(std::remove_if(pend_unit.begin(), pend_unit.end(), [&](int d) { return gone[d]; }), pend_unit.end());
			++(type == 1 ? st.div_waves : st.sum_waves);
		} else {
			std::vector<char> fresh(NL, 0); // lane freed in this wave: its slot is taken
			auto issue_term = [&](int l, Chain &c) {
				Issue &is = W.lane[l];
				is.kind = 1; is.add = c.add;
				if (c.pos < (int)c.terms.size()) {
					is.a = c.terms[c.pos].first; is.b = c.terms[c.pos].second;
					use(is.a, wave);
					use(is.b, wave);
					++c.pos;
				}
				++st.lane_slots;
				if (c.pos >= (int)c.terms.size()) {
					is.last = true; is.reg = c.reg; is.cs = c.dst;
					A.at(c.dst) = wave + 1;
					def.at(c.dst) = wave;
					c.done = true;
					lane_chain[l] = -1;
					fresh[l] = true;
					--remaining;
				}
			};
			for (int l = 0; l < NL; ++l)
				if (lane_chain[l] >= 0) issue_term(l, B.chains[lane_chain[l]]);
			std::stable_sort(cand.begin(), cand.end(), [&](int x, int y) { return chain_prio(B.chains[x]) > chain_prio(B.chains[y]); });
			size_t next = 0;
			std::vector<char> gone(B.chains.size(), 0);
			for (int l = 0; l < NL && next < cand.size(); ++l) {
				if (lane_chain[l] >= 0 || fresh[l]) continue;
				const int ci = cand[next++];
				Chain &c = B.chains[ci];
				gone[ci] = 1;
				Issue &is = W.lane[l];
				is.first = true;
				if (c.acc >= 0) { is.ci = c.acc; is.negz = true; } // resume: reload the accumulator exactly
				else { is.ci = c.init; is.negz = c.negz; }
				use(is.ci, wave);
				lane_chain[l] = ci;
				issue_term(l, c);
			}
			pend_chain.erase(std::remove_if(pend_chain.begin(), pend_chain.end(), [&](int c) { return gone[c]; }), pend_chain.end());
			// suspend every chain that cannot continue in the next wave
			for (int l = 0; l < NL; ++l) {
				if (lane_chain[l] < 0) continue;
				Chain &c = B.chains[lane_chain[l]];
				if (A.at(c.terms[c.pos].first) <= wave + 1 && A.at(c.terms[c.pos].second) <= wave + 1) continue;
				Issue &is = W.lane[l];
				c.acc = B.new_temp();
				is.last = true; is.reg = false; is.cs = c.acc;
				A.at(c.acc) = wave + 1;
				def.at(c.acc) = wave;
				pend_chain.push_back(lane_chain[l]);
				lane_chain[l] = -1;
				++st.suspensions;
			}
			++st.mul_waves;
		}
		waves.push_back(W);
		++wave;
	}

	// ---- temporaries: first fit of contiguous blocks over live ranges [first def, last use] ------------
	temp_slot.assign(B.ntemps, -1);
	const int nblocks = (int)B.blocks.size();
	std::vector<int> bdef(nblocks, INF), bend(nblocks, -1), by_def(nblocks);
	for (int b = 0; b < nblocks; ++b) {
		by_def[b] = b;
		for (int i = 0; i < B.blocks[b].second; ++i) {
			const int t = TEMP_BASE + B.blocks[b].first + i;
			if (def.at(t) < 0) return false;
			bdef[b] = std::min(bdef[b], def.at(t));
			bend[b] = std::max(bend[b], std::max(last_use.at(t), def.at(t)));
		}
	}
	std::stable_sort(by_def.begin(), by_def.end(), [&](int x, int y) { return bdef[x] < bdef[y]; });
	std::vector<int> free_after; // per scratch double: last wave in which its current value is read
	for (int b : by_def) {
		const int len = B.blocks[b].second;
		int found = -1, run = 0;
		for (size_t s = 0; s < free_after.size(); ++s) {
			run = free_after[s] < bdef[b] ? run + 1 : 0;
			if (run == len) { found = (int)s - len + 1; break; }
		}
		if (found < 0) {
			found = (int)free_after.size() - run; // extend a free run at the end
			free_after.resize(found + len, -1);
		}
		for (int i = 0; i < len; ++i) {
			free_after[found + i] = bend[b];
			temp_slot[B.blocks[b].first + i] = found + i;
		}
	}
	temp_count = (int)free_after.size();
	st.temps = temp_count;
	if (getenv("QPB_SCHED_DEBUG")) {
		std::vector<int> live(waves.size() + 2, 0);
		for (int b = 0; b < nblocks; ++b)
			for (int w = bdef[b]; w <= bend[b]; ++w) live[w] += B.blocks[b].second;
		int mx = 0, at = 0;
		for (size_t w = 0; w < live.size(); ++w) if (live[w] > mx) { mx = live[w]; at = (int)w; }
		fprintf(stderr, "[sched] waves %zu temps allocated %d, max live %d at wave %d\n", waves.size(), temp_count, mx, at);
		std::string types;
		for (const Wave &W : waves) { int act = 0; for (int l = 0; l < NL; ++l) act += W.lane[l].kind != 0; types += (W.type == 0 ? "m" : W.type == 1 ? "D" : "S"); types += std::to_string(act) + " "; }
		fprintf(stderr, "[sched] %s\n", types.c_str());
	}
	st.waves = (int)waves.size();
	return true;
}

uint64_t encode(const Issue &is, int type, int tbase, const std::vector<int> &temp_slot, const std::vector<int> &alias_phys)
{
	auto addr = [&](int id) -> uint64_t { return (uint64_t)(id >= TEMP_BASE ? tbase + temp_slot[id - TEMP_BASE] : id >= ALIAS_BASE ? alias_phys[id - ALIAS_BASE] : id); };
	auto pack = [](uint64_t a, uint64_t ci, uint64_t b, uint64_t cs) { return a | (ci << 16) | (b << 32) | (cs << 48); };
	if (type != 0) {
		uint64_t w = type == 1 ? INS_DIVWAVE : INS_SUMWAVE;
		if (is.kind == 2) w |= INS_ACTIVE | pack(addr(is.a), addr(is.a), addr(is.b), addr(is.cs));
		else if (is.kind == 3) w |= INS_ACTIVE | pack(addr(is.a), addr(is.ci), (uint64_t)is.b, addr(is.cs));
		else w |= pack(addr(is.a), addr(is.a), type == 1 ? addr(is.a) : 0, addr(is.a));
		return w;
	}
	uint64_t w = pack(addr(is.a), addr(is.ci), addr(is.b), addr(is.cs));
	if (is.first) w |= INS_FIRST;
	if (is.last) w |= INS_LAST;
	if (is.negz) w |= INS_NEGZ;
	if (is.add) w |= INS_ADD;
	return w;
}

// Words are stored [super-wave][lane][slot]: a lane reads its `width` instruction words with one load.
void emit(const std::vector<Wave> &waves, int width, int zero_addr, int tbase, const std::vector<int> &temp_slot, const std::vector<int> &alias_phys, std::vector<uint64_t> &words, Segment &seg)
{
	seg.first = (int32_t)(words.size() / 32);
	seg.width = width;
	for (const Wave &W : waves)
		for (int l = 0; l < 32; ++l)
			for (int u = 0; u < width; ++u) words.push_back(encode(W.lane[u * 32 + l], W.type, tbase, temp_slot, alias_phys));
	Issue idle;
	idle.a = idle.b = idle.ci = idle.cs = zero_addr;
	seg.waves = (int32_t)waves.size();
	while (seg.waves % PROGRAM_WAVE_PAD != 0) {
		for (int l = 0; l < 32 * width; ++l) words.push_back(encode(idle, 0, tbase, temp_slot, alias_phys));
		++seg.waves;
	}
}

int even(long v) { return (int)((v + 1) & ~1L); }

} // namespace

bool build_program(const Plan &pl, Program &prog)
{
	prog = Program();
	const int n = (int)pl.n, m = (int)pl.m, p = (int)pl.p, N = (int)pl.N;
	const int nV = (int)(pl.nnzP + pl.nnzA + pl.nnzG);
	const int lnz = (int)pl.Lp[N];
	SliceLayout &L = prog.lay;
	int at = 0;
	auto take = [&](long count) { const int here = at; at += even(count); return here; };
	L.zero = take(2);
	L.V = take(nV); L.c = take(n); L.h = take(m); L.b = take(p);
	L.zd = take(m);
	L.Lx = take(lnz); L.D = take(N);
	L.T = at;

	// ---- factor program -------------------------------------------------------------------------------
	Builder F;
	F.fixed_size = L.T;
	{
		std::vector<int> row_init(N, -1), chain_of(N, -1), ytemp(N, -1);
		for (int k = 0; k < N; ++k) {
			int diag_init = -1;
			for (int s = pl.scat_ptr[k]; s < pl.scat_ptr[k + 1]; ++s) {
				const int dst = pl.scat_dst[s], src = pl.scat_src[s];
				const int addr = src < nV ? L.V + src : L.zd + (src - nV);
				int &where = dst == k ? diag_init : row_init[dst];
				if (where >= 0) { prog.error = "duplicate entries in the KKT pattern"; return false; }
				where = addr;
			}
			const int p0 = pl.pat_ptr[k], p1 = pl.pat_ptr[k + 1];
			for (int t = p0; t < p1; ++t) {
				const int i = pl.pat_col[t];
				Chain c;
				c.init = row_init[i] >= 0 ? row_init[i] : L.zero;
				c.negz = false;
				c.dst = F.new_temp();
				ytemp[i] = c.dst;
				chain_of[i] = F.add_chain(std::move(c));
			}
			for (int t = p0; t < p1; ++t) {
				const int i = pl.pat_col[t];
				for (long q = pl.Lp[i]; q < pl.pat_slot[t]; ++q) { // entries of column i stored by earlier rows
					const int j = (int)pl.Li[q];
					if (chain_of[j] < 0) { prog.error = "inconsistent plan: update outside the row pattern"; return false; }
					F.chains[chain_of[j]].terms.push_back({L.Lx + (int)q, ytemp[i]});
				}
			}
			for (int t = p0; t < p1; ++t) {
				const int i = pl.pat_col[t];
				F.add_div(L.Lx + pl.pat_slot[t], ytemp[i], L.D + i);
			}
			Chain d;
			d.init = diag_init >= 0 ? diag_init : L.zero;
			d.negz = false;
			d.reg = true;
			d.dst = L.D + k;
			for (int t = p0; t < p1; ++t) d.terms.push_back({L.Lx + pl.pat_slot[t], ytemp[pl.pat_col[t]]});
			F.add_chain(std::move(d));
			for (int s = pl.scat_ptr[k]; s < pl.scat_ptr[k + 1]; ++s)
				if (pl.scat_dst[s] != k) row_init[pl.scat_dst[s]] = -1;
			for (int t = p0; t < p1; ++t) { chain_of[pl.pat_col[t]] = -1; ytemp[pl.pat_col[t]] = -1; }
		}
	}
	std::vector<Wave> fwaves;
	std::vector<int> ftemp_slot;
	int ftemps = 0;
	if (!schedule(F, 32 * FACTOR_WIDTH, L.zero, fwaves, prog.factor.stats, ftemp_slot, ftemps)) { prog.error = "factor schedule failed"; return false; }

	// The solve's work vectors wf, wd sit at the start of T and its scheduler temporaries behind them;
	// T is sized for the larger of the two programs.  The schedule does not depend on addresses, so
	// it is built twice: once to learn how much scratch it needs, once with the final layout.
	Builder S;
	std::vector<Wave> swaves;
	std::vector<int> stemp_slot;
	int stemps = 0;
	for (int pass = 0; pass < 2; ++pass) {
	L.tsize = even(std::max(ftemps, 2 * N + stemps));
	at = L.T + L.tsize;
	L.rhs = take(N);
	L.x = take(n); L.y = take(p); L.z = take(m); L.s = take(m);
	L.r = take(N);
	L.lam = take(std::max(m, n)); L.ds = take(m); L.dz = take(m); L.dsl = take(m);
	L.total = at;
	if (L.total > (1 << INS_ADDR_BITS)) { prog.error = "pattern too large: a QP's working set exceeds 16384 doubles"; return false; }

	// ---- solve program ---------------------------------------------------------------------------------
	S = Builder();
	swaves.clear();
	S.fixed_size = L.total;
	{
		const int wf0 = L.T, wd0 = L.T + N;
		// row view of L: for row r the (column j, position q) pairs, j ascending
		std::vector<std::vector<std::pair<int, int>>> rows(N);
		for (int j = 0; j < N; ++j)
			for (long q = pl.Lp[j]; q < pl.Lp[j + 1]; ++q) rows[pl.Li[q]].push_back({j, (int)q});
		// The solution overwrites the right-hand side in place.  It is a different VALUE at the same
		// address: every reader of b[P[r]] precedes, through the dependencies, the chain (or quotient)
		// that stores x[P[r]], so the only thing to keep apart is the readiness bookkeeping.
		std::vector<int> wf(N), wd(N), sol(N);
		for (int j = 0; j < N; ++j) sol[j] = S.new_alias(L.rhs + (int)pl.P[j]);
		for (int r = 0; r < N; ++r) {
			const int src = L.rhs + (int)pl.P[r];
			if (rows[r].empty()) { wf[r] = src; continue; } // nothing to subtract: read b[P[r]] in place
			Chain c;
			c.init = src;
			c.negz = true;
			c.dst = wf0 + r;
			for (auto &e : rows[r]) c.terms.push_back({L.Lx + e.second, wf[e.first]});
			wf[r] = c.dst;
			S.add_chain(std::move(c));
		}
		for (int j = 0; j < N; ++j) {
			const bool leaf = pl.Lp[j] == pl.Lp[j + 1]; // empty column: the quotient is already x[P[j]]
			wd[j] = leaf ? sol[j] : wd0 + j;
			S.add_div(wd[j], wf[j], L.D + j);
		}
		for (int j = N - 1; j >= 0; --j) {
			const int len = (int)(pl.Lp[j + 1] - pl.Lp[j]);
			if (len == 0) continue;
			if (len < SUM_MIN_TERMS) { // short column: an ordinary chain
				Chain c;
				c.init = wd[j];
				c.negz = true;
				c.dst = sol[j];
				for (long q = pl.Lp[j]; q < pl.Lp[j + 1]; ++q) c.terms.push_back({L.Lx + (int)q, sol[pl.Li[q]]});
				S.add_chain(std::move(c));
				continue;
			}
			// ltsolve subtracts L(r,j)*x[r] for r ASCENDING, but x[r] becomes known for r DESCENDING: the
			// first term needs the value that arrives last.  The products are therefore formed in
			// parallel (one lane each, exact: -0.0 + a*b) once every x[r] of the column is known, and one
			// lane then subtracts them in the reference's order at register speed (ordered sum).
			const int block = S.new_block(len);
			for (int i = 0; i < len; ++i) {
				const long q = pl.Lp[j] + i;
				Chain c;
				c.init = L.zero + 1; // -0.0
				c.negz = true;
				c.add = true;
				c.dst = block + i;
				c.terms.push_back({L.Lx + (int)q, sol[pl.Li[q]]});
				for (long g = pl.Lp[j]; g < pl.Lp[j + 1]; ++g) c.gates.push_back(sol[pl.Li[g]]);
				S.add_chain(std::move(c));
			}
			S.add_sum(sol[j], wd[j], block, len);
		}
	}
	const int before = stemps;
	if (!schedule(S, 32 * SOLVE_WIDTH, L.zero, swaves, prog.solve.stats, stemp_slot, stemps)) { prog.error = "solve schedule failed"; return false; }
	if (pass == 1 && stemps != before) { prog.error = "solve schedule is not reproducible"; return false; }
	}

	emit(fwaves, FACTOR_WIDTH, L.zero, L.T, ftemp_slot, F.alias_phys, prog.words, prog.factor);
	emit(swaves, SOLVE_WIDTH, L.zero, L.T + 2 * N, stemp_slot, S.alias_phys, prog.words, prog.solve);
	{ // the kernel prefetches PROGRAM_WAVE_PAD waves ahead without a bounds check
		Issue idle;
		idle.a = idle.b = idle.ci = idle.cs = L.zero;
		for (int i = 0; i < PROGRAM_WAVE_PAD * 32 * std::max(FACTOR_WIDTH, SOLVE_WIDTH); ++i) prog.words.push_back(encode(idle, 0, L.T, ftemp_slot, F.alias_phys));
	}
	prog.ok = true;
	return true;
}

} // namespace qpb

// C ABI over the host-side symbolic plan (no CUDA in this file).  See include/qpswift_b200.h.

#include "capi_internal.h"

#include <cstdlib>
#include <vector>

using namespace qpb;

bool qpb_build(QPBPlan &h, qp_int n, qp_int m, qp_int p, const qp_int *Pjc, const qp_int *Pir, const qp_int *Ajc, const qp_int *Air, const qp_int *Gjc, const qp_int *Gir, const qp_int *Permut)
{
	if (!build_plan(h.plan, n, m, p, Pjc, Pir, Ajc, Air, Gjc, Gir, Permut)) return false;
	const bool ok = build_program(h.plan, h.program);
	const char *g = getenv("QPB_STREAM_GROUP"); // experiments: micro-ops per group (4 or 8)
	return build_stream_program(h.plan, h.stream, g && atoi(g) == 4 ? 4 : 8) && ok;
}

extern "C" {

QPBPlan *QPB_PLAN_CREATE(qp_int n, qp_int m, qp_int p, const qp_int *Pjc, const qp_int *Pir, const qp_int *Ajc, const qp_int *Air, const qp_int *Gjc, const qp_int *Gir, const qp_int *Permut)
{
	QPBPlan *h = new QPBPlan;
	qpb_build(*h, n, m, p, Pjc, Pir, Ajc, Air, Gjc, Gir, Permut);
	return h;
}

void QPB_PLAN_DESTROY(QPBPlan *plan) { delete plan; }

const qp_int *QPB_PLAN_ARRAY(const QPBPlan *h, int which, qp_int *count)
{
	const std::vector<long> *v = nullptr;
	switch (which) {
	case QPB_PLAN_KJC: v = &h->plan.Kjc; break;
	case QPB_PLAN_KIR: v = &h->plan.Kir; break;
	case QPB_PLAN_P: v = &h->plan.P; break;
	case QPB_PLAN_PINV: v = &h->plan.Pinv; break;
	case QPB_PLAN_PARENT: v = &h->plan.Parent; break;
	case QPB_PLAN_LNZ: v = &h->plan.Lnz; break;
	case QPB_PLAN_LP: v = &h->plan.Lp; break;
	case QPB_PLAN_LI: v = &h->plan.Li; break;
	default: break;
	}
	if (!v) { if (count) *count = 0; return nullptr; }
	if (count) *count = (qp_int)v->size();
	return v->data();
}

static qp_int qpb_segment_stat(const Segment &s, int which)
{
	switch (which) {
	case 0: return s.waves;           // padded wave count the kernel executes
	case 1: return s.stats.mul_waves; // multiply-subtract waves
	case 2: return s.stats.div_waves; // quotient waves
	case 3: return s.stats.terms;     // multiply-subtract operations (== the reference's)
	case 4: return s.stats.divs;
	case 5: return s.stats.chains;
	case 6: return s.stats.lane_slots;
	case 7: return s.stats.temps;
	case 8: return s.stats.sum_waves;   // ordered-sum waves
	case 9: return s.stats.sums;
	case 10: return s.stats.suspensions; // accumulators parked in scratch and resumed
	case 11: return s.first;             // offset inside the program, in units of 32 words
	case 12: return s.width;             // instruction words per lane per super-wave
	default: return -1;
	}
}

qp_int QPB_PLAN_INFO(const QPBPlan *h, int what)
{
	const Plan &pl = h->plan;
	switch (what) {
	case QPB_INFO_N: return pl.N;
	case QPB_INFO_NNZK: return (qp_int)pl.Kir.size();
	case QPB_INFO_LNZ: return pl.Lp[pl.N];
	case QPB_INFO_AMD_STATUS: return pl.amd_status;
	case QPB_INFO_FACTOR_FLOPS: return pl.factor_flops();
	case QPB_INFO_NNZU: return (qp_int)pl.scat_dst.size();
	case QPB_INFO_VALID: return pl.valid ? 1 : 0;
	case QPB_INFO_PLAN_WORDS: return (qp_int)pl.blob.size();
	case QPB_INFO_PROGRAM_OK: return h->program.ok ? 1 : 0;
	case QPB_INFO_SLICE_DOUBLES: return h->program.lay.total;
	case QPB_INFO_PROGRAM_WORDS: return (qp_int)h->program.words.size();
	default: break;
	}
	if (what >= QPB_INFO_FACTOR_STATS && what < QPB_INFO_FACTOR_STATS + 16) return qpb_segment_stat(h->program.factor, what - QPB_INFO_FACTOR_STATS);
	if (what >= QPB_INFO_SOLVE_STATS && what < QPB_INFO_SOLVE_STATS + 16) return qpb_segment_stat(h->program.solve, what - QPB_INFO_SOLVE_STATS);
	return -1;
}

const unsigned long long *QPB_PLAN_PROGRAM(const QPBPlan *h, qp_int *words)
{
	if (words) *words = (qp_int)h->program.words.size();
	return reinterpret_cast<const unsigned long long *>(h->program.words.data());
}

qp_int QPB_PLAN_LAYOUT(const QPBPlan *h, int *out, int capacity)
{
	const int count = (int)(sizeof(SliceLayout) / sizeof(int32_t));
	const int32_t *src = reinterpret_cast<const int32_t *>(&h->program.lay);
	for (int i = 0; i < count && i < capacity; ++i) out[i] = src[i];
	return count;
}

const unsigned long long *QPB_PLAN_STREAM(const QPBPlan *h, qp_int *words)
{
	if (words) *words = (qp_int)h->stream.words.size();
	return reinterpret_cast<const unsigned long long *>(h->stream.words.data());
}

unsigned long long QPB_PLAN_STREAM_KEY(const QPBPlan *h) { return stream_program_key(h->stream); }

qp_int QPB_PLAN_STREAM_INFO(const QPBPlan *h, long long *out, int capacity)
{
	std::vector<long long> v;
	const int32_t *lay = reinterpret_cast<const int32_t *>(&h->stream.lay);
	for (size_t i = 0; i < sizeof(StreamLayout) / sizeof(int32_t); ++i) v.push_back(lay[i]);
	v.push_back(h->stream.ok ? 1 : 0);
	v.push_back(h->stream.group);
	for (int s = 0; s < SSEG_COUNT; ++s) {
		const StreamSegment &g = h->stream.seg[s];
		const long long row[10] = {g.first, g.groups, g.ops, g.terms, g.divs, g.late, g.hbm_loads, g.hbm_stores, g.nops, g.slow_groups};
		v.insert(v.end(), row, row + 10);
	}
	for (size_t i = 0; i < v.size() && (int)i < capacity; ++i) out[i] = v[i];
	return (qp_int)v.size();
}

qp_int QPB_AMD_ORDER(qp_int n, const qp_int *Ap, const qp_int *Ai, qp_int *Perm)
{
	return amd_order(n, Ap, Ai, Perm);
}

} // extern "C"

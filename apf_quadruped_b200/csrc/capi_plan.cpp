// C ABI over the host-side symbolic plan (no CUDA in this file).  See include/qpswift_b200.h.

#include "capi_internal.h"

using namespace qpb;

extern "C" {

QPBPlan *QPB_PLAN_CREATE(qp_int n, qp_int m, qp_int p, const qp_int *Pjc, const qp_int *Pir, const qp_int *Ajc, const qp_int *Air, const qp_int *Gjc, const qp_int *Gir, const qp_int *Permut)
{
	QPBPlan *h = new QPBPlan;
	build_plan(h->plan, n, m, p, Pjc, Pir, Ajc, Air, Gjc, Gir, Permut);
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
	default: return -1;
	}
}

qp_int QPB_AMD_ORDER(qp_int n, const qp_int *Ap, const qp_int *Ai, qp_int *Perm)
{
	return amd_order(n, Ap, Ai, Perm);
}

} // extern "C"

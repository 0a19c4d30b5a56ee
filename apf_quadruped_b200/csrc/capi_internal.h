// Private definitions behind the opaque handles of include/qpswift_b200.h.
#ifndef QPB_CAPI_INTERNAL_H
#define QPB_CAPI_INTERNAL_H

#include "qpswift_b200.h"
#include "plan.h"
#include "program.h"
#include "stream_program.h"

struct QPBPlan
{
	qpb::Plan plan;
	qpb::Program program; // factor / solve schedules compiled from the plan (program.h): latency kernel
	qpb::StreamProgram stream; // micro-programs of the throughput kernel (stream_program.h)
};

// build_plan + build_program; returns plan.valid && program.ok
bool qpb_build(QPBPlan &h, qp_int n, qp_int m, qp_int p, const qp_int *Pjc, const qp_int *Pir, const qp_int *Ajc, const qp_int *Air, const qp_int *Gjc, const qp_int *Gir, const qp_int *Permut);

#endif

// Private definitions behind the opaque handles of include/qpswift_b200.h.
#ifndef QPB_CAPI_INTERNAL_H
#define QPB_CAPI_INTERNAL_H

#include "qpswift_b200.h"
#include "plan.h"

struct QPBPlan
{
	qpb::Plan plan;
};

#endif

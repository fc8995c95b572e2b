#include "common.h"
using namespace mmr;
extern "C" int mmr_wgrad_plan_create(const MmrWgradDesc*, void**) { return fail("wgrad: not built yet"); }
extern "C" int mmr_wgrad_plan_run(void*, int, mmr_stream_t) { return fail("wgrad: not built yet"); }
extern "C" int mmr_wgrad_plan_destroy(void*) { return 0; }

// Library identity queries.
#include "common.cuh"

extern "C" int dcgp_version(void) { return 100; }  // 0.1.0

extern "C" int dcgp_arch(void) { return 1000; }  // built for sm_100a only (see build.py)

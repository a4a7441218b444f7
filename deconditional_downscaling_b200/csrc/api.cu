// Library identity queries.
#include "common.cuh"

extern "C" int dcgp_version(void) { return 100; }  // 0.1.0

extern "C" int dcgp_arch(void) { return 1000; }  // built for sm_100a only (see build.py)

#include <atomic>
static std::atomic<unsigned long long> g_launches{0};
void dcgp_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" unsigned long long dcgp_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

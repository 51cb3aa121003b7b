// Runtime glue of the C ABI: error strings and the launch counter.
#include <atomic>
#include <cstdio>
#include <cstring>

#include "common.cuh"

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

extern "C" int vb200_set_cuda_error(cudaError_t e) {
    snprintf(g_err, sizeof(g_err), "CUDA error %d: %s", (int)e, cudaGetErrorString(e));
    return VB200_ECUDA;
}
extern "C" int vb200_set_error(int code, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}
extern "C" void vb200_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" int vb200_abi_version(void) { return VB200_ABI_VERSION; }
extern "C" const char* vb200_last_error(void) { return g_err; }
extern "C" int64_t vb200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" void vb200_reset_launch_count(void) { g_launches.store(0, std::memory_order_relaxed); }

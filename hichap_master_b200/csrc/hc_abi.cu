// ABI glue: version, thread-local last-error string, launch counter.
#include <stdarg.h>
#include <atomic>
#include "hc_common.cuh"

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void hc_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void hc_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" int hc_version(void) { return HC_ABI_VERSION; }
extern "C" const char* hc_last_error(void) { return g_err; }
extern "C" int64_t hc_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

// Keep stream-ordered scratch (cudaMallocAsync) cached in the pool between calls: with the
// default release threshold of 0 every synchronisation hands the memory back to the driver and
// the next call pays for a fresh allocation.
extern "C" int hc_init(void) {
    int dev = 0;
    HC_CUDA(cudaGetDevice(&dev));
    cudaMemPool_t pool;
    HC_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    unsigned long long thr = ~0ull;
    HC_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    return HC_OK;
}

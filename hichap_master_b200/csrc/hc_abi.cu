// ABI glue: version, thread-local last-error string, launch counter.
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include <mutex>
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

// A second stream-ordered pool for the few multi-GB scratch blocks (the column-blocked CSR entries): in the default pool a
// 1 KB allocation of another call can be carved out of the big free block, and the next big request then pays ~200 ms for
// fresh mappings although "enough" memory is free (measured in bench.py: C4 ICE 238 instead of 59 ms).  Nothing small is
// ever allocated here, so a block of the previous size is found whole.
cudaMemPool_t hc_big_pool(void) {
    static std::mutex mu;
    static cudaMemPool_t pools[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!pools[dev]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        if (cudaMemPoolCreate(&pools[dev], &props) != cudaSuccess) { (void)cudaGetLastError(); pools[dev] = nullptr; return nullptr; }
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &thr);
    }
    return pools[dev];
}

// Bytes the default stream-ordered pool holds but does not use right now (what a cudaMallocAsync can get without asking the
// driver): lets a caller that shares the device with another allocator decide whether that one has to give memory back.
extern "C" int64_t hc_mempool_free_bytes(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) != cudaSuccess) return 0;
    unsigned long long reserved = 0, used = 0;
    if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) != cudaSuccess) return 0;
    if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) != cudaSuccess) return 0;
    int64_t freeb = reserved > used ? (int64_t)(reserved - used) : 0;
    if (cudaMemPool_t big = hc_big_pool()) {
        unsigned long long r2 = 0, u2 = 0;
        if (cudaMemPoolGetAttribute(big, cudaMemPoolAttrReservedMemCurrent, &r2) == cudaSuccess &&
            cudaMemPoolGetAttribute(big, cudaMemPoolAttrUsedMemCurrent, &u2) == cudaSuccess && r2 > u2) freeb += (int64_t)(r2 - u2);
    }
    return freeb;
}

// Small device -> host reads (a convergence counter, a few offsets) that must not queue behind a
// large cudaMemcpy on the device-to-host copy engine -- the upper-triangular records (1.4 GB for
// C2) are copied out while the ICE loop runs, and a 4-byte cudaMemcpyAsync poll issued meanwhile
// waited ~25 ms for that copy (tools/e2e_timeline.py).  A one-CTA kernel stores the words into a
// pinned, mapped host mailbox straight over PCIe; the host waits for the stream and reads it.
__global__ void mailbox_store_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int nwords) {
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) dst[i] = src[i];
}

namespace {
struct Mailbox { uint32_t* host = nullptr; uint32_t* dev = nullptr; };
constexpr size_t kMailboxBytes = 64 << 10;
thread_local Mailbox g_mailbox;
}  // namespace

cudaError_t hc_read_small(void* dst, const void* src, size_t bytes, cudaStream_t s) {
    if (bytes == 0) return cudaStreamSynchronize(s);
    if (bytes > kMailboxBytes || (bytes & 3u) || (reinterpret_cast<uintptr_t>(src) & 3u)) {   // not mailbox-shaped
        cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s);
        return e == cudaSuccess ? cudaStreamSynchronize(s) : e;
    }
    Mailbox& m = g_mailbox;
    if (!m.host) {
        cudaError_t e = cudaHostAlloc((void**)&m.host, kMailboxBytes, cudaHostAllocMapped | cudaHostAllocPortable);
        if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&m.dev, m.host, 0);
        if (e != cudaSuccess) { m.host = nullptr; return e; }
    }
    mailbox_store_kernel<<<1, 256, 0, s>>>(reinterpret_cast<const uint32_t*>(src), m.dev, (int)(bytes >> 2));
    hc_count_launch();
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) memcpy(dst, m.host, bytes);
    return e;
}

// dst += src (replicate merge of dense tiles, matrixBuilding.py:1700-1719)
__global__ void __launch_bounds__(256) add_i32_kernel(int32_t* __restrict__ dst, const int32_t* __restrict__ src, long long n) {
    const long long nv = n >> 2, stride = (long long)gridDim.x * blockDim.x;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += stride) {
        int4 a = reinterpret_cast<const int4*>(dst)[v];
        const int4 b = ld_stream_v4(src + 4 * v);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        reinterpret_cast<int4*>(dst)[v] = a;
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n - (nv << 2))) dst[(nv << 2) + threadIdx.x] += src[(nv << 2) + threadIdx.x];
}

extern "C" int hc_add_i32(int32_t* dst, const int32_t* src, int64_t n, void* stream) {
    HC_REQUIRE(n >= 0, "n>=0");
    if (n == 0) return HC_OK;
    HC_REQUIRE(((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15u) == 0, "16-byte alignment");
    long long blocks = (n / 4 + 255) / 256;
    const long long cap = (long long)hc_num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    add_i32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dst, src, n);
    HC_LAUNCH_CHECK();
    return HC_OK;
}

// chromosome ids travel over PCIe as uint8 (255 = filtered) and are widened on the device
__global__ void __launch_bounds__(256) widen_u8_i32_kernel(const uint8_t* __restrict__ src, int32_t* __restrict__ dst, long long n) {
    const long long nv = n >> 2, stride = (long long)gridDim.x * blockDim.x;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += stride) {
        const uint32_t w = reinterpret_cast<const uint32_t*>(src)[v];
        int4 o;
        o.x = (w & 255u) == 255u ? -1 : (int)(w & 255u);
        o.y = ((w >> 8) & 255u) == 255u ? -1 : (int)((w >> 8) & 255u);
        o.z = ((w >> 16) & 255u) == 255u ? -1 : (int)((w >> 16) & 255u);
        o.w = (w >> 24) == 255u ? -1 : (int)(w >> 24);
        reinterpret_cast<int4*>(dst)[v] = o;
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n - (nv << 2))) {
        const uint8_t b = src[(nv << 2) + threadIdx.x];
        dst[(nv << 2) + threadIdx.x] = b == 255 ? -1 : (int)b;
    }
}

extern "C" int hc_widen_u8_i32(const uint8_t* src, int32_t* dst, int64_t n, void* stream) {
    HC_REQUIRE(n >= 0, "n>=0");
    if (n == 0) return HC_OK;
    HC_REQUIRE((reinterpret_cast<uintptr_t>(src) & 3u) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0, "alignment");
    long long blocks = (n / 4 + 255) / 256;
    const long long cap = (long long)hc_num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    widen_u8_i32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, n);
    HC_LAUNCH_CHECK();
    return HC_OK;
}

// runtime.cu -- process-wide helpers behind the C ABI (errors, launch counter, device queries).
#include "runtime.cuh"

#include <atomic>
#include <cstring>

namespace mirb200 {

static thread_local std::string g_error;
static std::atomic<unsigned long long> g_launches{0};

void set_error(const std::string& msg) { g_error = msg; }
void clear_error() { g_error.clear(); }
void count_launch(long long n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

int check_cuda(cudaError_t e, const char* what)
{
    if (e == cudaSuccess) return MIR_B200_OK;
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    cudaGetLastError();   // clear the sticky-less error state
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return MIR_B200_ENODEVICE;
    return MIR_B200_ECUDA;
}

// Stream-ordered allocations come from the device's default pool; by default the pool hands freed memory back to
// the OS at the next synchronisation, so every solve would pay cudaMalloc-class latency for its (multi-GB) J / sample
// buffers again.  Keep freed blocks cached: one-time setting per device.
static void keep_pool_cached()
{
    static std::atomic<unsigned long long> doneMask{0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
    const unsigned long long bit = 1ull << dev;
    if (doneMask.load(std::memory_order_relaxed) & bit) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    doneMask.fetch_or(bit, std::memory_order_relaxed);
}

int require_device(int device)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        cudaGetLastError();
        set_error("mir_optim_b200: no usable CUDA device (this library has no CPU fallback): " +
                  std::string(e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e)));
        return MIR_B200_ENODEVICE;
    }
    if (device >= 0) {
        if (device >= count) { set_error("mir_optim_b200: device index out of range"); return MIR_B200_EINVAL; }
        const int rc = check_cuda(cudaSetDevice(device), "cudaSetDevice");
        if (rc) return rc;
    }
    keep_pool_cached();
    return MIR_B200_OK;
}

// Small pinned host scratch, one grow-only buffer per host thread, never freed before process exit:
// cudaMallocHost / cudaFreeHost synchronise the context and go through the OS, and were measured to add anything from
// 1 ms to several hundred ms to a solve on busy hosts when done per call.
void* pinned_scratch(size_t bytes)
{
    static thread_local void* buf = nullptr;
    static thread_local size_t cap = 0;
    if (bytes <= cap) return buf;
    void* nb = nullptr;
    const size_t want = bytes < 4096 ? 4096 : bytes;
    if (cudaMallocHost(&nb, want) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    // (the previous, smaller buffer is intentionally leaked: an in-flight copy may still target it)
    buf = nb; cap = want;
    return buf;
}

int sm_count()
{
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    return n;
}

}  // namespace mirb200

extern "C" {

const char* mir_b200_last_error(void) { return mirb200::g_error.c_str(); }
uint64_t mir_b200_kernel_launches(void) { return mirb200::g_launches.load(); }
int mir_b200_device_count(void)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
    return count;
}
const char* mir_b200_version(void) { return "mir_optim_b200 0.1 (sm_100a)"; }

}  // extern "C"

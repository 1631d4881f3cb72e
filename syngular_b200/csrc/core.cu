// Library lifecycle + error channel of libsyngular_b200.so.
#include "common.cuh"

#include <cstring>
#include <atomic>

namespace syn {

static thread_local char g_error[512] = "";
static std::atomic<long long> g_launches{0};

void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    return 1;
}

int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) return 0;
    return dev < SYN_MAX_DEVICES ? dev : SYN_MAX_DEVICES - 1;
}

int sm_count() {
    static PerDevice cache;
    const int dev = current_device();
    int n = 0;
    if (cache.get(dev, &n)) return n;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache.set(dev, n);
    return n;
}

}  // namespace syn

extern "C" int syn_version(void) { return 100; }
extern "C" const char* syn_last_error(void) { return syn::g_error; }
extern "C" int syn_device_sm_count(void) { return syn::sm_count(); }
extern "C" long long syn_launch_count(void) { return syn::g_launches.load(); }

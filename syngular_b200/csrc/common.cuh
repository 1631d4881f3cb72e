// Shared helpers for the sm_100a kernels of libsyngular_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/syngular_b200.h"

namespace syn {

// ---- error channel (never throw across the C ABI) -------------------------------------------------------
void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define SYN_CUDA(call)                                                        \
    do {                                                                      \
        cudaError_t _e = (call);                                              \
        if (_e != cudaSuccess) return ::syn::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define SYN_REQUIRE(cond, ...)                    \
    do {                                          \
        if (!(cond)) {                            \
            ::syn::set_error(__VA_ARGS__);        \
            return 2;                             \
        }                                         \
    } while (0)

void note_launch();   // counts kernel launches made by this library (bench.py's gpu_launches)

inline int launch_status(const char* what) {
    note_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, what, __FILE__, __LINE__);
    return 0;
}

int sm_count();          // of the CURRENT device (cached per device)
int current_device();    // cudaGetDevice, clamped to [0, SYN_MAX_DEVICES)

// Per-device cached configuration.  cudaFuncSetAttribute, occupancy numbers and the SM count are per DEVICE, and a process may
// switch devices (torch.cuda.set_device) or call from several host threads: every launcher keeps one of these as a function-local
// static instead of a plain bool.  value = 0 means "not configured on this device yet"; the configuration code is idempotent, so
// two threads racing on the first call at worst both run it (release/acquire orders their writes before the flag).
constexpr int SYN_MAX_DEVICES = 64;
struct PerDevice {
    int v[SYN_MAX_DEVICES];
    bool get(int dev, int* out = nullptr) const {
        int x = __atomic_load_n(&v[dev], __ATOMIC_ACQUIRE);
        if (!x) return false;
        if (out) *out = x - 1;
        return true;
    }
    void set(int dev, int value = 0) { __atomic_store_n(&v[dev], value + 1, __ATOMIC_RELEASE); }
};

// ---- device helpers -------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t idx2(const syn_index_t& ix, int x) {
    // two-level index: (x / div) * outer + (x % div) * inner ; div >= extent collapses to x * inner
    if ((unsigned)x < (unsigned)ix.div) return (int64_t)x * ix.inner;      // single level: no integer division on the hot path
    unsigned q = (unsigned)x / (unsigned)ix.div;
    unsigned r = (unsigned)x - q * (unsigned)ix.div;
    return (int64_t)q * ix.outer + (int64_t)r * ix.inner;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    // D(8x8) += A(8x4,row) * B(4x8,col); SASS: DMMA.8x8x4
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gmem_src, bool pred) {
    // LDGSTS with zero-fill when !pred (src-size 0)
    uint32_t d = smem_u32(smem_dst);
    int sz = pred ? BYTES : 0;
    if constexpr (BYTES == 16) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(sz));
    } else {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gmem_src), "r"(sz));
    }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) + mbarrier ------------------------------------------------------------
// Contiguous runs (a tile line of the GEMM, a row block of the Jacobi exchange) move between global and shared memory with
// no register staging and no per-element address math; completion is tracked by an mbarrier (loads) or the bulk
// async-group (stores).  Bulk copies go through L2 and bypass L1.
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (unsigned spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 26)) __trap();           // a lost transaction would otherwise hang the GPU
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait_all() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- MUFU-seeded FP64 reciprocal / rsqrt (seed 2^-22, each Newton step doubles the bits) -------------------------------
__device__ __forceinline__ double rcp_newton1(double x) {      // ~2^-44
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return fma(r, fma(-x, r, 1.0), r);
}
__device__ __forceinline__ double rsqrt_newton1(double x) {    // ~2^-43
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return fma(y, fma(-0.5 * x * y, y, 0.5), y);
}
__device__ __forceinline__ double rsqrt_newton2(double x) {    // full double precision
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double h = 0.5 * x;
    y = fma(y, fma(-h * y, y, 0.5), y);
    y = fma(y, fma(-h * y, y, 0.5), y);
    return y;
}

__device__ __forceinline__ double rcp_newton2(double x) {      // full double precision
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    return fma(r, fma(-x, r, 1.0), r);
}

}  // namespace syn

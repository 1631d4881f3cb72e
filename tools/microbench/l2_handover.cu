// Hand-over of a BYTES block between two CTAs (on different SMs) through L2, ping-pong: cycles per one-way hand-over.
//   mode 0: TMA store -> wait_group 0 -> fence.proxy.async -> st.release flag ; ld.acquire poll -> TMA load -> mbarrier wait
//   mode 1: all threads st.cg 16 B -> __syncthreads -> thread0 __threadfence + st.release ; poll -> all threads ld.cg -> smem
//   mode 2: like 0 but wait_group.read (smem reusable, NOT a valid completion for publishing: lower bound only)
//   mode 3: like 0 without fence.proxy.async
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_handover l2_handover.cu
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned ld_acq(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_rel(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

template <int MODE>
__global__ void k(int iters, int bytes, unsigned char* buf, unsigned* flag, long long* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long mbar;
    const int me = blockIdx.x;                 // 0 or 1
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    for (int i = threadIdx.x; i < bytes / 8; i += blockDim.x) reinterpret_cast<double*>(smem)[i] = i;
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();
    unsigned phase = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        const bool sender = ((it & 1) == me);
        if (sender) {
            if (MODE == 1) {
                for (int i = threadIdx.x; i < bytes / 16; i += blockDim.x) __stcg(reinterpret_cast<double2*>(buf) + i, reinterpret_cast<const double2*>(smem)[i]);
                __syncthreads();
                if (threadIdx.x == 0) { __threadfence(); st_rel(flag, it + 1); }
            } else {
                asm volatile("fence.proxy.async;" ::: "memory");
                __syncthreads();
                if (threadIdx.x == 0) {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(buf), "r"(smem_u32(smem)), "r"(bytes) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    if (MODE == 2) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    else asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                    if (MODE != 3) asm volatile("fence.proxy.async;" ::: "memory");
                    st_rel(flag, it + 1);
                }
            }
        } else {
            if (MODE == 1) {
                if (threadIdx.x == 0) while (ld_acq(flag) < (unsigned)(it + 1)) {}
                __syncthreads();
                for (int i = threadIdx.x; i < bytes / 16; i += blockDim.x) reinterpret_cast<double2*>(smem)[i] = __ldcg(reinterpret_cast<const double2*>(buf) + i);
                __syncthreads();
            } else {
                if (threadIdx.x == 0) {
                    while (ld_acq(flag) < (unsigned)(it + 1)) {}
                    asm volatile("fence.proxy.async;" ::: "memory");
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(bytes) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem)), "l"(buf),
                                 "r"(bytes), "r"(smem_u32(&mbar)) : "memory");
                }
                unsigned done = 0;
                while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&mbar)), "r"(phase) : "memory");
                phase ^= 1;
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && me == 0) out[0] = (t1 - t0) / iters;
}

template <int MODE>
void run(const char* name, int bytes) {
    unsigned char* buf; unsigned* flag; long long* out;
    cudaMalloc(&buf, bytes); cudaMalloc(&flag, 4); cudaMalloc(&out, 8); cudaMemset(flag, 0, 4);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    k<MODE><<<2, 256, bytes>>>(400, bytes, buf, flag, out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("%-52s bytes=%6d : %6lld cycles per hand-over (%s)\n", name, bytes, h, cudaGetErrorString(e));
    cudaFree(buf); cudaFree(flag); cudaFree(out);
}

int main() {
    for (int bytes : {4096, 16384, 32768, 65536}) {
        run<0>("TMA store, wait, proxy fence, release | poll, TMA load", bytes);
        run<3>("  ... without the proxy fence", bytes);
        run<2>("  ... wait_group.read (lower bound, not valid)", bytes);
        run<1>("st.cg by 256 threads, fence, release | poll, ld.cg", bytes);
    }
    return 0;
}

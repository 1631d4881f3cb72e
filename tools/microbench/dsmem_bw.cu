// DSMEM block hand-over cost on sm_100a: every CTA of a cluster sends BYTES to rank+1 (and optionally rank-1) with
// cp.async.bulk.shared::cluster (completion on the receiver's mbarrier), double-buffered, one cluster barrier per round.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_bw dsmem_bw.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdint>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank)); return r; }

template <int TWO_WAY, int THREAD_COPY>
__global__ void k(int iters, int bytes, long long* out, double* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long mbar[2];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = cluster.block_rank(), CL = cluster.num_blocks();
    // layout: send[2 dirs][bytes] | recv[2 parity][2 dirs][bytes]
    unsigned char* send = smem;
    unsigned char* recv = smem + 2 * bytes;
    if (threadIdx.x < 2) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[threadIdx.x])));
    for (int i = threadIdx.x; i < 2 * bytes / 8; i += blockDim.x) reinterpret_cast<double*>(send)[i] = i + rank;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    cluster.sync();
    const int nd = TWO_WAY ? 2 : 1;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        const int par = it & 1;
        if (!THREAD_COPY) {
            if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar[par])), "r"(nd * bytes) : "memory");
            if (threadIdx.x < nd) {
                const int d = threadIdx.x, peer = d ? (rank + CL - 1) % CL : (rank + 1) % CL;
                uint32_t dst = mapa(smem_u32(recv + (par * 2 + d) * bytes), peer), rmb = mapa(smem_u32(&mbar[par]), peer);
                asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                             "r"(smem_u32(send + d * bytes)), "r"(bytes), "r"(rmb) : "memory");
            }
            unsigned done = 0; const unsigned ph = (it >> 1) & 1;
            while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&mbar[par])), "r"(ph) : "memory");
            cluster.sync();        // receivers are done with parity par^1 of the previous round before it is overwritten next round
        } else {
            for (int d = 0; d < nd; d++) {
                const int peer = d ? (rank + CL - 1) % CL : (rank + 1) % CL;
                double2* dst = reinterpret_cast<double2*>(cluster.map_shared_rank(recv + (par * 2 + d) * bytes, peer));
                const double2* src = reinterpret_cast<const double2*>(send + d * bytes);
                for (int i = threadIdx.x; i < bytes / 16; i += blockDim.x) dst[i] = src[i];
            }
            cluster.sync();
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / iters;
    sink[blockIdx.x] = reinterpret_cast<double*>(recv)[threadIdx.x];
    cluster.sync();
}

template <int TWO_WAY, int THREAD_COPY>
void run(const char* name, int cl, int bytes, int threads) {
    long long* out; double* sink;
    cudaMalloc(&out, 8); cudaMalloc(&sink, 8 * 64);
    auto kern = k<TWO_WAY, THREAD_COPY>;
    size_t smem = (size_t)6 * bytes;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cl); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int iters = 500;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, iters, bytes, out, sink);
    cudaError_t e2 = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("%-34s cluster=%2d bytes/dir=%6d : %7lld cycles/round  -> %5.1f B/clk out per CTA  (%s / %s)\n", name, cl, bytes, h,
           h ? (double)(TWO_WAY ? 2 : 1) * bytes / h : 0.0, cudaGetErrorString(e), cudaGetErrorString(e2));
    cudaFree(out); cudaFree(sink);
}

int main() {
    for (int cl : {8, 16}) for (int bytes : {4096, 16384, 32768}) {
        run<0, 0>("bulk copy one-way", cl, bytes, 256);
        run<1, 0>("bulk copy two-way", cl, bytes, 256);
        run<1, 1>("thread st.shared::cluster two-way", cl, bytes, 256);
    }
    return 0;
}

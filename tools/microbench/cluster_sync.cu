// Fixed costs of the cluster-QR step on sm_100a: cluster barrier flavours, DSMEM all-gather, __syncthreads, per CTA size.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_sync cluster_sync.cu
#include <cooperative_groups.h>
#include <cstdio>
namespace cg = cooperative_groups;

constexpr int CL = 8;
struct __align__(128) Sh { double part[2][CL][256]; double red[4][256]; double stage[2][256]; unsigned long long mbar[2]; int flags[CL]; };

template <int MODE>
__global__ void __cluster_dims__(CL, 1, 1) k(int iters, long long* out, double* sink) {
    __shared__ Sh sh;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = cluster.block_rank();
    const int c = threadIdx.x % 256, rg = threadIdx.x / 256;
    const int nrg = blockDim.x / 256;
    Sh* peer[CL];
    for (int q = 0; q < CL; q++) peer[q] = cluster.map_shared_rank(&sh, q);
    double acc = threadIdx.x;
    if (threadIdx.x < CL) sh.flags[threadIdx.x] = 0;
    if (threadIdx.x < 2) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"((unsigned)__cvta_generic_to_shared(&sh.mbar[threadIdx.x]))); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    cluster.sync();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        const int par = it & 1;
        if (MODE == 0) {                       // __syncthreads only
            sh.red[rg][c] = acc; __syncthreads(); acc += sh.red[(rg + 1) % nrg][c];
        } else if (MODE == 1) {                // bare cluster.sync
            cluster.sync();
        } else if (MODE == 2) {                // all-gather (each row group serves CL/nrg peers) + cluster.sync + read back
            for (int q = rg; q < CL; q += nrg) peer[q]->part[par][rank][c] = acc;
            cluster.sync();
            double d = 0; for (int q = 0; q < CL; q++) d += sh.part[par][q][c]; acc = d * 0.125;
        } else if (MODE == 3) {                // same with explicit arrive.release / wait.acquire split by the local read
            for (int q = rg; q < CL; q += nrg) peer[q]->part[par][rank][c] = acc;
            asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
            double d = 0; for (int q = 0; q < CL; q++) d += sh.part[par][q][c]; acc = d * 0.125;
        } else if (MODE == 4) {                // only warp 0..7 (rg 0) store, everyone barriers
            if (rg == 0) for (int q = 0; q < CL; q++) peer[q]->part[par][rank][c] = acc;
            cluster.sync();
            double d = 0; for (int q = 0; q < CL; q++) d += sh.part[par][q][c]; acc = d * 0.125;
        } else if (MODE == 5) {                // flag-based: each CTA's warp 0 lane 0.. polls per-source counters instead of a barrier
            // stores then a release-store of the step number into every peer's flag; consumers spin on their 8 local flags
            for (int q = rg; q < CL; q += nrg) peer[q]->part[par][rank][c] = acc;
            __syncthreads();
            if (threadIdx.x < CL) {
                __threadfence();                                   // order this CTA's remote stores before the flag
                atomicExch(&peer[threadIdx.x]->flags[rank], it + 1);
                volatile int* f = sh.flags;
                while (f[threadIdx.x] < it + 1) {}
            }
            __syncthreads();
            double d = 0; for (int q = 0; q < CL; q++) d += sh.part[par][q][c]; acc = d * 0.125;
        } else if (MODE == 7) {                // TMA DSMEM all-gather: one 2 KB bulk copy per peer, completion on the receiver's mbarrier
            if (rg == 0) sh.stage[par][c] = acc;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            const unsigned mb = (unsigned)__cvta_generic_to_shared(&sh.mbar[par]);
            if (threadIdx.x == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mb), "r"(CL * 2048) : "memory");
            if (threadIdx.x < CL) {
                unsigned dst = (unsigned)__cvta_generic_to_shared(&sh.part[par][rank][0]), rdst, rmb;
                unsigned src = (unsigned)__cvta_generic_to_shared(&sh.stage[par][0]);
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rdst) : "r"(dst), "r"(threadIdx.x));
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rmb) : "r"(mb), "r"(threadIdx.x));
                asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(rdst), "r"(src), "r"(2048), "r"(rmb) : "memory");
            }
            {
                unsigned done = 0; const unsigned ph = (it >> 1) & 1;
                while (!done)
                    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                 : "=r"(done) : "r"(mb), "r"(ph) : "memory");
            }
            double d = 0; for (int q = 0; q < CL; q++) d += sh.part[par][q][c]; acc = d * 0.125;
        } else if (MODE == 6) {                // FP64 reflector scalar chain executed by every thread (sqrt + 2 div)
            double x = fma(acc, acc, 1.0);
            double n = sqrt(x), beta = -n, tau = (beta - acc) / beta, vs = 1.0 / (acc - beta);
            acc = tau + vs * 1e-3;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / iters;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    cluster.sync();
}

template <int MODE>
void run(const char* name, int threads) {
    long long* out; double* sink;
    cudaMalloc(&out, 8); cudaMalloc(&sink, 8 * CL * 1024);
    k<MODE><<<CL, threads>>>(2000, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("%-58s threads=%4d : %6lld cycles/iter  (%s)\n", name, threads, h, cudaGetErrorString(e));
    cudaFree(out); cudaFree(sink);
}

int main() {
    for (int t : {256, 512, 1024}) {
        run<0>("__syncthreads + smem round trip", t);
        run<1>("cluster.sync", t);
        run<2>("all-gather + cluster.sync + reduce", t);
        run<3>("all-gather + arrive.release/wait.acquire + reduce", t);
        run<4>("all-gather by rg0 only + cluster.sync + reduce", t);
        run<5>("all-gather + flag handshake (atomicExch into peers)", t);
        run<6>("sqrt + 2 div FP64 per thread", t);
        run<7>("TMA DSMEM all-gather (bulk copy + mbarrier) + reduce", t);
    }
    return 0;
}

// Microbenchmark: FP64 pipe throughput on sm_100a (DFMA vs DMMA shapes), cluster-16 launch feasibility.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d: %s\n",#x,__LINE__,cudaGetErrorString(e));}}while(0)

__global__ void k_dfma(double* out, int iters) {
    double a0=1.0+threadIdx.x*1e-9,a1=1.1,a2=1.2,a3=1.3,a4=1.4,a5=1.5,a6=1.6,a7=1.7;
    double b=1.0000001, c=1e-9;
    for (int i=0;i<iters;i++){
        a0=fma(a0,b,c);a1=fma(a1,b,c);a2=fma(a2,b,c);a3=fma(a3,b,c);
        a4=fma(a4,b,c);a5=fma(a5,b,c);a6=fma(a6,b,c);a7=fma(a7,b,c);
    }
    out[blockIdx.x*blockDim.x+threadIdx.x]=a0+a1+a2+a3+a4+a5+a6+a7;
}
__global__ void k_dmma884(double* out, int iters) {
    double a=1.0+threadIdx.x*1e-9,b=1e-3;
    double c[8][2];
    #pragma unroll
    for(int j=0;j<8;j++){c[j][0]=0;c[j][1]=0;}
    for (int i=0;i<iters;i++){
        #pragma unroll
        for(int j=0;j<8;j++)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
            : "+d"(c[j][0]),"+d"(c[j][1]) : "d"(a),"d"(b));
    }
    double s=0; for(int j=0;j<8;j++) s+=c[j][0]+c[j][1];
    out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void k_dmma16816(double* out, int iters) {
    double a[8],b[4];
    for(int j=0;j<8;j++)a[j]=1.0+threadIdx.x*1e-9+j; for(int j=0;j<4;j++)b[j]=1e-3*j;
    double c[4][4];
    #pragma unroll
    for(int j=0;j<4;j++){c[j][0]=0;c[j][1]=0;c[j][2]=0;c[j][3]=0;}
    for (int i=0;i<iters;i++){
        #pragma unroll
        for(int j=0;j<4;j++)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
            : "+d"(c[j][0]),"+d"(c[j][1]),"+d"(c[j][2]),"+d"(c[j][3])
            : "d"(a[0]),"d"(a[1]),"d"(a[2]),"d"(a[3]),"d"(a[4]),"d"(a[5]),"d"(a[6]),"d"(a[7]),
              "d"(b[0]),"d"(b[1]),"d"(b[2]),"d"(b[3]));
    }
    double s=0; for(int j=0;j<4;j++) s+=c[j][0]+c[j][1]+c[j][2]+c[j][3];
    out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void k_dmma1684(double* out, int iters) {
    double a[2],b[1];
    a[0]=1.0+threadIdx.x*1e-9; a[1]=2.0; b[0]=1e-3;
    double c[8][4];
    #pragma unroll
    for(int j=0;j<8;j++){c[j][0]=0;c[j][1]=0;c[j][2]=0;c[j][3]=0;}
    for (int i=0;i<iters;i++){
        #pragma unroll
        for(int j=0;j<8;j++)
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
            : "+d"(c[j][0]),"+d"(c[j][1]),"+d"(c[j][2]),"+d"(c[j][3])
            : "d"(a[0]),"d"(a[1]),"d"(b[0]));
    }
    double s=0; for(int j=0;j<8;j++) s+=c[j][0]+c[j][1]+c[j][2]+c[j][3];
    out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

__global__ void k_cluster_probe(int* out) {
    cg::cluster_group cl = cg::this_cluster();
    extern __shared__ double sm[];
    sm[threadIdx.x] = cl.block_rank();
    cl.sync();
    unsigned peer = (cl.block_rank()+1) % cl.num_blocks();
    double* p = cl.map_shared_rank(sm, peer);
    double v = p[threadIdx.x];
    cl.sync();
    if (threadIdx.x==0) out[blockIdx.x] = (int)v;
}
__global__ void k_coop_probe(int* out, int rounds) {
    cg::grid_group g = cg::this_grid();
    for (int r=0;r<rounds;r++) { g.sync(); }
    if (threadIdx.x==0) out[blockIdx.x]=blockIdx.x;
}

template<typename F> float timeit(F f, int rep=3){
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best=1e30f;
    for(int r=0;r<rep;r++){ cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
    return best;
}

int main(){
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
    printf("device %s sm_%d%d SMs=%d smemOptin=%zu clock=%d kHz\n",p.name,p.major,p.minor,p.multiProcessorCount,p.sharedMemPerBlockOptin,p.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double)*148*8*1024));
    int nsm=p.multiProcessorCount;
    for (int wps : {4,8,16,32}) {
        int threads=wps*32; int blocks=nsm; int iters=20000;
        float ms=timeit([&]{k_dfma<<<blocks,threads>>>(out,iters);});
        double fl=2.0*8*iters*(double)threads*blocks;
        printf("DFMA      warps/SM=%2d : %.2f TFLOP/s\n",wps,fl/ms*1e-9);
        ms=timeit([&]{k_dmma884<<<blocks,threads>>>(out,iters);});
        fl=512.0*8*iters*(double)wps*blocks;
        printf("DMMA884   warps/SM=%2d : %.2f TFLOP/s\n",wps,fl/ms*1e-9);
        ms=timeit([&]{k_dmma1684<<<blocks,threads>>>(out,iters);});
        fl=1024.0*8*iters*(double)wps*blocks;
        printf("DMMA1684  warps/SM=%2d : %.2f TFLOP/s\n",wps,fl/ms*1e-9);
        ms=timeit([&]{k_dmma16816<<<blocks,threads>>>(out,iters/4);});
        fl=4096.0*4*(iters/4)*(double)wps*blocks;
        printf("DMMA16816 warps/SM=%2d : %.2f TFLOP/s\n",wps,fl/ms*1e-9);
    }
    // cluster-16 probe
    int* iout; CK(cudaMalloc(&iout, 4096*sizeof(int)));
    for (int cs : {8,16}) {
        for (size_t smem : {(size_t)64*1024,(size_t)140*1024,(size_t)200*1024}) {
        cudaLaunchConfig_t cfg{}; cfg.gridDim=dim3(cs*4); cfg.blockDim=dim3(256); cfg.dynamicSmemBytes=smem;
        cudaLaunchAttribute at[1]; at[0].id=cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x=cs; at[0].val.clusterDim.y=1; at[0].val.clusterDim.z=1;
        cfg.attrs=at; cfg.numAttrs=1;
        cudaError_t e1=cudaFuncSetAttribute(k_cluster_probe,cudaFuncAttributeNonPortableClusterSizeAllowed,1);
        cudaError_t e2=cudaFuncSetAttribute(k_cluster_probe,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem);
        int ncl=0; cudaError_t e3=cudaOccupancyMaxActiveClusters(&ncl,k_cluster_probe,&cfg);
        cudaError_t e4=cudaLaunchKernelEx(&cfg,k_cluster_probe,iout);
        cudaError_t e5=cudaDeviceSynchronize();
        printf("cluster=%d smem=%zuKB: attr %d/%d maxActiveClusters=%d (%d) launch=%s sync=%s\n",cs,smem/1024,e1,e2,ncl,e3,cudaGetErrorString(e4),cudaGetErrorString(e5));
        cudaGetLastError();
        }
    }
    // cooperative grid sync latency
    for (int blocks : {16, 74, 148}) {
        int rounds=1000; void* args[]={&iout,&rounds};
        int threads=256;
        float ms=timeit([&]{cudaLaunchCooperativeKernel((void*)k_coop_probe,dim3(blocks),dim3(threads),args,0,0);});
        printf("coop grid.sync blocks=%d: %.3f us/sync (%s)\n",blocks,ms*1e3/rounds,cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}

// Fused "W sandwich" of the right-environment update (density-matrix rounding), sm_100a.
//
// The environment of bond k is  E_k = C_k E_{k+1} C_k^T  with C_k the MPO x MPS product core; built through the (X, W) structure
// it is four contractions (syngular/tensor/_sweeps.py: right_environments).  The two in the middle contract the SAME small MPO core
// W (l, i, o, r) = (16, 2, 2, 16) on both sides of a large intermediate:
//         P2[a, (l,o), (r',b')]   = sum_{(i,r)}  W[l,i,o,r]     P1[a, (i,r), (r',b')]          K = 32
//         Z [a, l, (l',i'), b']   = sum_{(o,r')} W[l',i',o,r']  P2[a, (l,o), (r',b')]           K = 32
// As two GEMM launches they are HBM-bound (0.5 flop/byte each: P1, P2 and Z are 268 MB apiece on the C2 plateau) and ran at
// 2 TB/s.  Fused, P2 never leaves the SM: a work item is one `a` and 16 consecutive b' -- a (32 x 256) slab of P1, 64 KB, fetched by
// cp.async into a double-buffered ring while the previous slab is being multiplied -- and both products run on DMMA from shared
// memory with W resident in two layouts.  Traffic drops to P1 in + Z out (0.54 GB per site), arithmetic intensity to 8 flop/byte.
// Stands for the two np.tensordot calls of the restated oracle (oracle/svd_numpy.py: apply_round_density_matrix_structured, P2 and Z).
#include "common.cuh"

namespace syn {

constexpr int EV_THREADS = 256, EV_TB = 16, EV_K = 32, EV_R = 16, EV_L = 16;
// 64-bit fragment loads are served per half-warp (g in 0..3, t in 0..3): every stride that a fragment index multiplies must be
// 4 (mod 16) doubles for the 16 lanes to hit 16 distinct bank pairs (first ncu capture: 43 % of the shared wavefronts were conflicts
// with strides of 8 mod 16)
constexpr int EV_LDS1 = EV_R * EV_TB + 4;          // 260: slab row stride (k-major B operand of stage 1: index t)
constexpr int EV_LDW = EV_K + 4;                   // 36:  W rows (A operand: index g)
constexpr int EV_RS = EV_TB + 4;                   // 20:  r' stride inside a P2 row (B operand of stage 2: index t)
constexpr int EV_LDP = EV_R * EV_RS + 8;           // 328: P2 row stride (8 mod 16: the 128-bit stores of two rows g fill one wavefront)
constexpr int EV_S1 = EV_K * EV_LDS1;              // doubles per slab buffer
constexpr size_t EV_SMEM = sizeof(double) * (2 * EV_S1 + 8 * EV_LDP + 2 * EV_K * EV_LDW);

// P1: (a, 32, 16, b) contiguous [a][(i,r)][r'][b'] ; W: (16, 2, 2, 16) contiguous ; Z: (a, 16, 32, b) contiguous [a][l][(l',i')][b']
__global__ void __launch_bounds__(EV_THREADS, 1)
env_sandwich_kernel(const double* __restrict__ P1, const double* __restrict__ W, double* __restrict__ Z, int na, int b) {
    extern __shared__ __align__(16) double ev_smem[];
    double* S1 = ev_smem;                           // [2][32][264]
    double* S2 = S1 + 2 * EV_S1;                    // [8][16][20]   P2 rows of the current group of 4 l
    double* W1 = S2 + 8 * EV_LDP;                   // [(l,o)][(i,r)]
    double* W2 = W1 + EV_K * EV_LDW;                // [(l',i')][(o,r')]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int tiles_b = b / EV_TB;
    const int64_t items = (int64_t)na * tiles_b;
    const int64_t D = (int64_t)EV_R * b;

    for (int idx = tid; idx < 1024; idx += EV_THREADS) {
        const int l = idx >> 6, i = (idx >> 5) & 1, o = (idx >> 4) & 1, r = idx & 15;
        const double w = W[idx];
        W1[(l * 2 + o) * EV_LDW + i * 16 + r] = w;
        W2[(l * 2 + i) * EV_LDW + o * 16 + r] = w;
    }

    auto fetch = [&](int64_t item, int buf) {
        const int64_t a = item / tiles_b;
        const int bt = (int)(item - a * tiles_b);
        const double* src = P1 + a * (EV_K * D) + (int64_t)bt * EV_TB;
        double* dst = S1 + buf * EV_S1;
#pragma unroll
        for (int p = 0; p < 16; p++) {
            const int c = tid + EV_THREADS * p;     // 4096 16-byte chunks: k (32) | r' (16) | chunk (8)
            const int k = c >> 7, rp = (c >> 3) & 15, ch = c & 7;
            cp_async<16>(dst + k * EV_LDS1 + rp * EV_TB + ch * 2, src + k * D + (int64_t)rp * b + ch * 2, true);
        }
    };

    int64_t item = blockIdx.x;
    int buf = 0;
    if (item < items) fetch(item, 0);
    cp_async_commit();
    for (; item < items; item += gridDim.x, buf ^= 1) {
        const int64_t next = item + gridDim.x;
        if (next < items) fetch(next, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();                             // slab `buf` (and W on the first pass) visible to all warps
        const double* s1 = S1 + buf * EV_S1;
        const int64_t a = item / tiles_b;
        const int bt = (int)(item - a * tiles_b);
        for (int grp = 0; grp < 4; ++grp) {
            // ---- stage 1: P2[(l,o) in group (8 rows)][n = (r',b') in this warp's 32 columns] ---------------------------------------
            double c1[4][2];
#pragma unroll
            for (int j = 0; j < 4; j++) c1[j][0] = c1[j][1] = 0.0;
#pragma unroll
            for (int kk = 0; kk < EV_K; kk += 4) {
                const double af = W1[(grp * 8 + g) * EV_LDW + kk + t];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const double bf = s1[(kk + t) * EV_LDS1 + warp * 32 + j * 8 + g];
                    dmma884(c1[j][0], c1[j][1], af, bf);
                }
            }
            __syncthreads();                         // the previous group's stage 2 has finished reading S2
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int rp = warp * 2 + (j >> 1), b0 = (j & 1) * 8;
                *reinterpret_cast<double2*>(S2 + g * EV_LDP + rp * EV_RS + b0 + 2 * t) = make_double2(c1[j][0], c1[j][1]);
            }
            __syncthreads();
            // ---- stage 2: Z[l][(l',i')][b'] for the 4 l of the group: warp -> (l, half of the 32 output rows) ---------------------------
            const int ll = warp >> 1, half = warp & 1;
            double c2[2][2][2];
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int j = 0; j < 2; j++) c2[i][j][0] = c2[i][j][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < 8; ks++) {
                const int o = ks >> 2, rp0 = (ks & 3) * 4;
                double af[2], bf[2];
#pragma unroll
                for (int i = 0; i < 2; i++) af[i] = W2[((half * 2 + i) * 8 + g) * EV_LDW + ks * 4 + t];
#pragma unroll
                for (int j = 0; j < 2; j++) bf[j] = S2[(ll * 2 + o) * EV_LDP + (rp0 + t) * EV_RS + j * 8 + g];
#pragma unroll
                for (int i = 0; i < 2; i++)
#pragma unroll
                    for (int j = 0; j < 2; j++) dmma884(c2[i][j][0], c2[i][j][1], af[i], bf[j]);
            }
            const int l = grp * 4 + ll;
            double* zb = Z + ((a * EV_L + l) * 32) * (int64_t)b + (int64_t)bt * EV_TB;
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const int m2 = (half * 2 + i) * 8 + g;
                    *reinterpret_cast<double2*>(zb + (int64_t)m2 * b + j * 8 + 2 * t) = make_double2(c2[i][j][0], c2[i][j][1]);
                }
        }
        __syncthreads();                             // every warp is done with slab `buf` before it is refilled
    }
    cp_async_wait<0>();
}

int env_sandwich_f64(const double* P1, const double* W, double* Z, int na, int l, int i, int o, int r, int b, cudaStream_t st) {
    SYN_REQUIRE(P1 && W && Z, "syn_env_sandwich_f64: null argument");
    SYN_REQUIRE(l == EV_L && r == EV_R && i == 2 && o == 2 && b >= EV_TB && b % EV_TB == 0 && na >= 1,
                "syn_env_sandwich_f64: shape (l,i,o,r)=(%d,%d,%d,%d), b=%d is not covered by the fused kernel", l, i, o, r, b);
    SYN_REQUIRE(((((uintptr_t)P1) | ((uintptr_t)Z)) & 15) == 0, "syn_env_sandwich_f64: operands must be 16-byte aligned");
    static PerDevice configured;
    const int dev_ = current_device();
    if (!configured.get(dev_)) {
        SYN_CUDA(cudaFuncSetAttribute(env_sandwich_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EV_SMEM));
        configured.set(dev_);
    }
    const int64_t items = (int64_t)na * (b / EV_TB);
    const int grid = (int)(items < sm_count() ? items : sm_count());
    env_sandwich_kernel<<<grid, EV_THREADS, EV_SMEM, st>>>(P1, W, Z, na, b);
    return launch_status("env_sandwich_kernel");
}

// The environment E[(a,l),(l',a')] is symmetric under (a,l) <-> (a',l').  Its last GEMM only forms the block-lower part (a' in the
// same or an earlier block of `ab` values than a); this kernel fills the strictly upper blocks: for every pair (l, l') the 32 x 32
// tile (a, a') of plane (l, l') is the transpose of tile (a', a) of plane (l', l) -- read and written along contiguous a through
// shared memory.
__global__ void __launch_bounds__(256) env_mirror_kernel(double* __restrict__ E, int na, int L, int ab) {
    __shared__ double tile[32][33];
    const int ja = blockIdx.x, ia = blockIdx.y;
    if ((ja * 32) / ab <= (ia * 32) / ab) return;
    const int l = blockIdx.z / L, lp = blockIdx.z - l * L;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t row = (int64_t)L * na;
    for (int r = ty; r < 32; r += 8) {
        const int as = ja * 32 + r, ac = ia * 32 + tx;             // source element (a'' = as, l') x (l, a = ac)
        tile[r][tx] = E[((int64_t)as * L + lp) * row + (int64_t)l * na + ac];
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int a = ia * 32 + r, ap = ja * 32 + tx;
        E[((int64_t)a * L + l) * row + (int64_t)lp * na + ap] = tile[tx][r];
    }
}

int env_mirror_f64(double* E, int na, int L, int ab, cudaStream_t st) {
    SYN_REQUIRE(E && na >= 32 && na % 32 == 0 && L >= 1 && ab >= 32 && ab % 32 == 0 && na % ab == 0,
                "syn_env_mirror_f64: na=%d must be a multiple of the block ab=%d, both multiples of 32", na, ab);
    SYN_REQUIRE((int64_t)L * L <= 65535, "syn_env_mirror_f64: MPO bond too large");
    dim3 grid(na / 32, na / 32, L * L);
    env_mirror_kernel<<<grid, 256, 0, st>>>(E, na, L, ab);
    return launch_status("env_mirror_kernel");
}

}  // namespace syn

extern "C" int syn_env_mirror_f64(double* E, int na, int L, int ab, void* stream) { return syn::env_mirror_f64(E, na, L, ab, (cudaStream_t)stream); }

extern "C" int syn_env_sandwich_fits(int l, int i, int o, int r, int b) {
    return (l == syn::EV_L && r == syn::EV_R && i == 2 && o == 2 && b >= syn::EV_TB && b % syn::EV_TB == 0) ? 1 : 0;
}

extern "C" int syn_env_sandwich_f64(const double* P1, const double* W, double* Z, int na, int l, int i, int o, int r, int b, void* stream) {
    return syn::env_sandwich_f64(P1, W, Z, na, l, i, o, r, b, (cudaStream_t)stream);
}

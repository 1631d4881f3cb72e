// Batched dominant-eigenspace solver for SMALL bond problems: one CTA per problem, everything in shared memory.
//
// Batches of independent chains (BASELINE configs[3](ii): a shared MPO applied to 8192 chi = 64 states and rounded back) give
// one (n x n) Gram matrix per state and bond, n = kept rank x physical dimension <= 128.  The one-sided Jacobi route spends 3.2 ms per
// launch on 256 such problems (69 % of the batched sweep, profiles/r02_launches_c4_apply_svd_summary.txt).  This kernel is the single-chain
// projection solver of purify.cu (SP2 trace-correcting purification, then Newton-Schulz orthonormalisation of the projector's leading
// columns, convergence decided in the kernel, same 8 info doubles for the host's verdict) re-cut for that size:
//
//   * the iterate X (n x n, row stride n + 4: conflict-free DMMA fragment loads along rows and along columns) never leaves shared memory;
//   * a step's product X X^T is held in registers (64 accumulators per thread at n = 128: 8 warps as 2 x 4, warp tile n/2 x n/4), the CTA
//     meets at a barrier and the update (X^2 or 2X - X^2 by the trace rule) is written over X in place -- two CTA barriers per step and no
//     grid barrier, no global traffic: the 85 grid barriers per call of the single-chain kernel do not exist here;
//   * Newton-Schulz works on U = X[:, :ne] in place: G = U^T U is read along columns of U ("TN"), U G along rows, G (ne x ne) next to X.
//
// A persistent grid walks the batch.  Reference behaviour replaced: the truncation of every bond to the target dimension
// (tensor/matrix_product_state.py:432-468, truncated there by QR; here the optimal SVD truncation of the density-matrix sweep).
#include "common.cuh"

namespace syn {

void note_launch();

constexpr int PB_THREADS = 256;

struct PurifyBatchedArgs {
    const double* A;         // [batch][n][n], symmetric positive semi-definite
    double* U;               // [batch][n][ne]
    double* info;            // [batch][8]
    int batch, n, ne, sp2_max, ns_max;
};

// acc[i][j] = R1 rows x R2 rows over K: element (row r, k) of operand x is Rx[r * RSx + k * KSx]; everything but the pointers is a compile-time
// constant, so the fragment addresses fold into the LDS immediates and the k loop unrolls
template <int MT, int NT, int RS1, int KS1, int RS2, int KS2, int K>
__device__ __forceinline__ void pb_mma(double (&acc)[MT][NT][2], const double* __restrict__ R1, const double* __restrict__ R2) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int i = 0; i < MT; i++)
#pragma unroll
        for (int j = 0; j < NT; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    const double* p1 = R1 + g * RS1 + t * KS1;
    const double* p2 = R2 + g * RS2 + t * KS2;
#pragma unroll 4
    for (int k = 0; k < K; k += 4) {
        double af[MT], bf[NT];
#pragma unroll
        for (int i = 0; i < MT; i++) af[i] = p1[8 * i * RS1 + k * KS1];
#pragma unroll
        for (int j = 0; j < NT; j++) bf[j] = p2[8 * j * RS2 + k * KS2];
#pragma unroll
        for (int i = 0; i < MT; i++)
#pragma unroll
            for (int j = 0; j < NT; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
}

// every thread receives the two block-wide sums; `red` holds 16 doubles; two CTA barriers
__device__ __forceinline__ void pb_sum2(double& a, double& b, double* red) {
    a = warp_sum(a);
    b = warp_sum(b);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { red[warp] = a; red[8 + warp] = b; }
    __syncthreads();
    a = b = 0.0;
#pragma unroll
    for (int w = 0; w < PB_THREADS / 32; w++) { a += red[w]; b += red[8 + w]; }
    __syncthreads();
}

__device__ __forceinline__ double pb_max(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    v = 0.0;
#pragma unroll
    for (int w = 0; w < PB_THREADS / 32; w++) v = fmax(v, red[w]);
    __syncthreads();
    return v;
}

// n = 32 NB, ne = 32 EB
template <int NB, int EB>
__global__ void __launch_bounds__(PB_THREADS, 1) purify_batched_kernel(const PurifyBatchedArgs a) {
    extern __shared__ __align__(16) double pb_smem[];
    constexpr int n = 32 * NB, ne = 32 * EB, LD = n + 4, LDG = ne + 4;
    double* X = pb_smem;
    double* G = X + n * LD;
    double* red = G + ne * LDG;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wr = warp >> 2, wc = warp & 3;
    // warp tiles: X X^T and U G take rows [i0, i0 + n/2); X X^T columns [j0, j0 + n/4); G = U^T U is (ne/2) x (ne/4) per warp
    constexpr int mt = 2 * NB, nt = NB, gmt = 2 * EB, gnt = EB;
    const int i0 = wr * (n / 2), j0 = wc * (n / 4);
    const int gi0 = wr * (ne / 2), gj0 = wc * (ne / 4);
    const double dne = (double)ne;

    for (int p = blockIdx.x; p < a.batch; p += gridDim.x) {
        const double* Ap = a.A + (size_t)p * n * n;
        // ---- load, |A|_F^2 and tr A ------------------------------------------------------------------------------------------------
        double fro2 = 0.0, tra = 0.0;
        for (int i = tid; i < n * n / 2; i += PB_THREADS) {
            const double2 v = reinterpret_cast<const double2*>(Ap)[i];
            const int r = (2 * i) / n, c = (2 * i) - r * n;
            *reinterpret_cast<double2*>(X + r * LD + c) = v;
            fro2 = fma(v.x, v.x, fma(v.y, v.y, fro2));
            if (r == c) tra += v.x;
            if (r == c + 1) tra += v.y;
        }
        pb_sum2(fro2, tra, red);                               // (also orders the stores to X before the scaling pass)
        const double inv = fro2 > 0.0 ? rsqrt_newton2(fro2) : 0.0;
        for (int i = tid; i < n * n; i += PB_THREADS) {
            const int r = i / n, c = i - r * n;
            X[r * LD + c] *= inv;
        }
        __syncthreads();
        // ---- SP2: X <- X^2 or 2X - X^2 by the trace rule (see purify.cu for the stopping rule) -----------------------------------------
        double tr0 = tra * inv, f0 = fro2 * inv * inv;
        int it = 0, extra = 0, lift = 0;
        bool lifting = true;
        for (; it < a.sp2_max; ++it) {
            const bool conv = fabs(tr0 - dne) < 1e-11 * dne && fabs(tr0 - f0) < 1e-11 * dne;
            if (conv || extra) {
                if (extra == 2) break;
                ++extra;
            }
            const bool square = extra == 1 ? true : (extra == 2 ? false : fabs(f0 - dne) < fabs(2.0 * tr0 - f0 - dne));
            if (lifting && !square) ++lift; else lifting = false;
            double acc[mt][nt][2];
            pb_mma<mt, nt, LD, 1, LD, 1, n>(acc, X + i0 * LD, X + j0 * LD);
            __syncthreads();                                   // every warp has read the old iterate
            double tr = 0.0, f2 = 0.0;
#pragma unroll
            for (int i = 0; i < mt; i++)
#pragma unroll
                for (int j = 0; j < nt; j++) {
                        const int r = i0 + 8 * i + g, c = j0 + 8 * j + 2 * t;
                        double2* px = reinterpret_cast<double2*>(X + r * LD + c);
                        double2 x = *px;
                        x.x = square ? acc[i][j][0] : fma(2.0, x.x, -acc[i][j][0]);
                        x.y = square ? acc[i][j][1] : fma(2.0, x.y, -acc[i][j][1]);
                        *px = x;
                        f2 = fma(x.x, x.x, fma(x.y, x.y, f2));
                        if (r == c) tr += x.x;
                        if (r == c + 1) tr += x.y;
                    }
            pb_sum2(tr, f2, red);                              // (its barriers publish the new iterate)
            tr0 = tr;
            f0 = f2;
        }
        const int sp2_used = it;
        // ---- kept weight sum A o P ----------------------------------------------------------------------------------------------------
        double kept = 0.0, unused = 0.0;
        for (int i = tid; i < n * n / 2; i += PB_THREADS) {
            const double2 v = reinterpret_cast<const double2*>(Ap)[i];
            const int r = (2 * i) / n, c = (2 * i) - r * n;
            const double2 x = *reinterpret_cast<const double2*>(X + r * LD + c);
            kept = fma(v.x, x.x, fma(v.y, x.y, kept));
        }
        pb_sum2(kept, unused, red);
        // ---- Newton-Schulz on U = P[:, :ne], in place ------------------------------------------------------------------------------------
        int ns = 0;
        double dev0 = 0.0, lo = 1.0;
        bool last_steep = true;
        for (;; ++ns) {
            double gacc[gmt][gnt][2];
            pb_mma<gmt, gnt, 1, LD, 1, LD, n>(gacc, X + gi0, X + gj0);         // G = U^T U: both operands read down the columns of U
            double trg = 0.0, dmax = 0.0;
#pragma unroll
            for (int i = 0; i < gmt; i++)
#pragma unroll
                for (int j = 0; j < gnt; j++) {
                        const int r = gi0 + 8 * i + g, c = gj0 + 8 * j + 2 * t;
                        *reinterpret_cast<double2*>(G + r * LDG + c) = make_double2(gacc[i][j][0], gacc[i][j][1]);
                        if (r == c) trg += gacc[i][j][0];
                        if (r == c + 1) trg += gacc[i][j][1];
                        dmax = fmax(dmax, fmax(fabs(gacc[i][j][0] - (r == c ? 1.0 : 0.0)), fabs(gacc[i][j][1] - (r == c + 1 ? 1.0 : 0.0))));
                    }
            double unused2 = 0.0;
            pb_sum2(trg, unused2, red);                        // (its barriers publish G)
            if (ns == 0) dev0 = pb_max(dmax, red);             // how far the start is from orthonormal decides whether the steep map is used
            if (ns >= 1 && !last_steep && fabs(dne - trg) < 8e-13) break;
            if (ns >= a.ns_max) break;
            // scaled Newton-Schulz (see purify.cu): the best cubic for singular values in [lo, 1], lo an estimate tracked through the maps
            if (ns == 0) lo = dev0 > 0.5 ? 1.0 / 512.0 : 1.0;
            const double g2 = 3.0 / (1.0 + lo + lo * lo), gs = sqrt(g2);
            const double ca = 1.5 * gs, cb = -0.5 * gs * g2;
            lo = fmin(1.0, 0.5 * gs * lo * (3.0 - g2 * lo * lo));
            const bool steep = false;
            double acc[mt][gnt][2];
            pb_mma<mt, gnt, LD, 1, LDG, 1, ne>(acc, X + i0 * LD, G + gj0 * LDG);   // U G (G symmetric: its rows are its columns)
            __syncthreads();                                   // every warp has read the old U
#pragma unroll
            for (int i = 0; i < mt; i++)
#pragma unroll
                for (int j = 0; j < gnt; j++) {
                        const int r = i0 + 8 * i + g, c = gj0 + 8 * j + 2 * t;
                        double2* pu = reinterpret_cast<double2*>(X + r * LD + c);
                        double2 u = *pu;
                        u.x = fma(ca, u.x, cb * acc[i][j][0]);
                        u.y = fma(ca, u.y, cb * acc[i][j][1]);
                        *pu = u;
                    }
            __syncthreads();
            last_steep = steep;
        }
        // max |G - I| of the final iterate (G is its Gram matrix)
        double dv = 0.0;
        for (int i = tid; i < ne * ne; i += PB_THREADS) {
            const int r = i / ne, c = i - r * ne;
            dv = fmax(dv, fabs(G[r * LDG + c] - (r == c ? 1.0 : 0.0)));
        }
        dv = pb_max(dv, red);
        // ---- results -------------------------------------------------------------------------------------------------------------------
        double* Up = a.U + (size_t)p * n * ne;
        for (int i = tid; i < n * ne; i += PB_THREADS) {
            const int r = i / ne, c = i - r * ne;
            Up[i] = X[r * LD + c];
        }
        if (tid == 0) {
            double* h = a.info + (size_t)p * 8;
            const double f = sqrt(fro2);
            h[0] = tr0;
            h[1] = f0;
            h[2] = kept;
            h[3] = f;
            h[4] = dv;
            h[5] = tra;
            h[6] = tr0 - f0;
            h[7] = (double)(sp2_used + 1000 * ns) + 1e6 * lift;
        }
        __syncthreads();                                       // X and G are rewritten by the next problem
    }
}

static size_t pb_smem_bytes(int n, int ne) { return ((size_t)n * (n + 4) + (size_t)ne * (ne + 4) + 16) * sizeof(double); }

static bool purify_batched_fits(int n, int ne) {
    return n % 32 == 0 && ne % 32 == 0 && n >= 64 && n <= 128 && ne >= 32 && ne < n && pb_smem_bytes(n, ne) <= 227 * 1024;
}

}  // namespace syn

extern "C" int syn_dominant_subspace_batched_fits(int n, int ne) { return syn::purify_batched_fits(n, ne) ? 1 : 0; }

extern "C" int syn_dominant_subspace_batched_f64(const double* A, int batch, int n, int ne, int sp2_max, int ns_max, double* U, double* info,
                                                 void* stream) {
    using namespace syn;
    if (batch <= 0) return 0;
    SYN_REQUIRE(A && U && info, "syn_dominant_subspace_batched_f64: null argument");
    SYN_REQUIRE(purify_batched_fits(n, ne), "syn_dominant_subspace_batched_f64: n, ne multiples of 32 with 32 <= ne < n <= 128 (n=%d ne=%d)", n, ne);
    SYN_REQUIRE(sp2_max >= 1 && sp2_max <= 400 && ns_max >= 0 && ns_max <= 400, "syn_dominant_subspace_batched_f64: bad iteration limits");
    SYN_REQUIRE((((uintptr_t)A) & 15) == 0, "syn_dominant_subspace_batched_f64: A must be 16-byte aligned");
    void (*kern)(const PurifyBatchedArgs) = nullptr;
    switch ((n / 32) * 10 + ne / 32) {
        case 21: kern = purify_batched_kernel<2, 1>; break;
        case 31: kern = purify_batched_kernel<3, 1>; break;
        case 32: kern = purify_batched_kernel<3, 2>; break;
        case 41: kern = purify_batched_kernel<4, 1>; break;
        case 42: kern = purify_batched_kernel<4, 2>; break;
        case 43: kern = purify_batched_kernel<4, 3>; break;
    }
    SYN_REQUIRE(kern != nullptr, "syn_dominant_subspace_batched_f64: no kernel for n=%d ne=%d", n, ne);
    static PerDevice configured[6];
    const int dev = current_device(), slot = (n / 32 - 2) * (n / 32 - 1) / 2 + ne / 32 - 1;
    if (!configured[slot].get(dev)) {
        SYN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured[slot].set(dev);
    }
    PurifyBatchedArgs a;
    a.A = A; a.U = U; a.info = info;
    a.batch = batch; a.n = n; a.ne = ne; a.sp2_max = sp2_max; a.ns_max = ns_max;
    const int sms = sm_count();
    // balanced waves: ceil(batch / waves) CTAs for the smallest number of waves that covers the batch
    const int waves = (batch + sms - 1) / sms;
    const int grid = (batch + waves - 1) / waves;
    kern<<<grid, PB_THREADS, pb_smem_bytes(n, ne), (cudaStream_t)stream>>>(a);
    note_launch();
    return launch_status("purify_batched_kernel");
}

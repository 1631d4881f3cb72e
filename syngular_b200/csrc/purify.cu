// Dominant invariant subspace of a symmetric PSD matrix by trace-steered spectral projection (SP2 purification) followed by a
// Newton-Schulz orthonormalisation -- the GEMM-bound replacement of the Jacobi eigen-solver inside the density-matrix rounding
// sweep (north_star: "truncated SVD of each bond matrix"; the only written-down precedent in the reference is the unfinished
// np.linalg.eigh of matrix_product_operator.py:228).
//
// The sweep needs, per bond, an orthonormal basis U of the span of the `ne` dominant eigenvectors of A = M E M^T (n x n) -- the kept
// left singular vectors are only defined up to a gauge, so individual eigenvectors are not required.  With X0 = A / |A|_F
// (spectrum in [0, 1]) the iteration
//         X <- X^2            if  |tr(X^2) - ne| < |tr(2X - X^2) - ne|
//         X <- 2X - X^2       otherwise
// maps every eigenvalue monotonically and converges quadratically to the spectral projector P with tr P = ne: eigenvalues above
// the ne-th gap go to 1, the others to 0 (Niklasson's second-order spectral projection; no eigenvalue estimate is needed, the
// trace steers it).  tr(X^2) = |X|_F^2 is known BEFORE the product is formed, so the branch is chosen on the device from two
// scalars that the previous step accumulated: the whole iteration is a fixed sequence of launches with no host round trip.
// One step = one DMMA GEMM (split in two along K so that 128 CTAs work on a 512^3 product; the partial sums are added by the
// step kernel) + one elementwise kernel.  Then U0 = P[:, :ne] (its columns are the projections of the leading coordinate vectors;
// rows are ordered by the previous bond's singular values, so they are well conditioned) is orthonormalised by
// U <- U (1.5 I - 0.5 U^T U): two GEMMs per step.  All matrices stay in L2 (n = 512: 2 MB each).
//
// The caller checks idempotency, trace and orthonormality (info[]) and falls back to the Jacobi path when the spectrum has no
// gap at ne (rank-deficient bonds near the chain ends).
#include <climits>

#include "common.cuh"

namespace syn {

int gemm_f64(const syn_gemm_desc_t& d, const double* A, const double* B, double* C, cudaStream_t st);

static inline syn_index_t PIX(int64_t stride) {
    syn_index_t i;
    i.outer = 0; i.inner = stride; i.div = INT_MAX; i._pad = 0;
    return i;
}

static syn_gemm_desc_t pdesc(int M, int N, int K, int batch, int64_t a_m, int64_t a_k, int64_t a_b, int64_t b_k, int64_t b_n, int64_t b_b,
                             int64_t c_m, int64_t c_n, int64_t c_b, double alpha = 1.0, double beta = 0.0) {
    syn_gemm_desc_t d = {};
    d.M = M; d.N = N; d.K = K; d.batch = batch;
    d.a_m = PIX(a_m); d.a_k = PIX(a_k); d.a_b = PIX(a_b);
    d.b_k = PIX(b_k); d.b_n = PIX(b_n); d.b_b = PIX(b_b);
    d.c_m = PIX(c_m); d.c_n = PIX(c_n); d.c_b = PIX(c_b);
    d.alpha = alpha; d.beta = beta;
    return d;
}

__device__ __forceinline__ void block_accumulate(double a, double b, double* out0, double* out1) {
    __shared__ double r0[32], r1[32];
    __syncthreads();          // warp 0 may still be reading r0 / r1 of a previous call (two calls with no CTA barrier in between: a CTA
                              // without Gram items goes from the kept-weight sum straight to the trace sum; found by racecheck)
    a = warp_sum(a);
    b = warp_sum(b);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { r0[warp] = a; r1[warp] = b; }
    __syncthreads();
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        a = lane < nw ? r0[lane] : 0.0;
        b = lane < nw ? r1[lane] : 0.0;
        a = warp_sum(a);
        b = warp_sum(b);
        if (lane == 0) {
            if (out0) atomicAdd(out0, a);
            if (out1) atomicAdd(out1, b);
        }
    }
}

// out[0] += sum A_i B_i
__global__ void __launch_bounds__(256) purify_dot_kernel(const double* __restrict__ A, const double* __restrict__ B, int64_t count,
                                                         double* __restrict__ out) {
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) s = fma(A[i], B[i], s);
    block_accumulate(s, 0.0, out, nullptr);
}

// X = A / sqrt(fro2[0]);  ctrl[0] += tr X, ctrl[1] += |X|_F^2
__global__ void __launch_bounds__(256) sp2_init_kernel(const double* __restrict__ A, double* __restrict__ X, int n, const double* __restrict__ fro2,
                                                       double* __restrict__ ctrl) {
    const double f = fro2[0];
    const double inv = f > 0.0 ? rsqrt(f) : 0.0;
    const int64_t count = (int64_t)n * n;
    double tr = 0.0, f2 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = A[i] * inv;
        X[i] = x;
        f2 = fma(x, x, f2);
        if (i / n == i % n) tr += x;
    }
    block_accumulate(tr, f2, ctrl, ctrl + 1);
}

// One SP2 step.  Y = sum of `parts` partial products of X X (row-major n x n, stride part_stride); symmetrised on the fly.
// cin = (tr X, |X|_F^2) of the current iterate, cout accumulates the same for the new one.
__global__ void __launch_bounds__(256) sp2_step_kernel(const double* __restrict__ Y, int64_t part_stride, int parts, double* __restrict__ X, int n,
                                                       double ne, const double* __restrict__ cin, double* __restrict__ cout) {
    const double tr0 = cin[0], f0 = cin[1];
    const bool square = fabs(f0 - ne) < fabs(2.0 * tr0 - f0 - ne);
    // 32 x 32 tile per block; the transposed partner tile is read through shared memory so both reads are coalesced
    __shared__ double tt[32][33];
    const int bi = blockIdx.y * 32, bj = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 256 threads: 8 rows per pass
    for (int r = ty; r < 32; r += 8) {
        const int i = bj + r, j = bi + tx;                            // element (i, j) of the partner tile
        double y = 0.0;
        if (i < n && j < n)
            for (int p = 0; p < parts; ++p) y += Y[p * part_stride + (int64_t)i * n + j];
        tt[r][tx] = y;
    }
    __syncthreads();
    double tr = 0.0, f2 = 0.0;
    for (int r = ty; r < 32; r += 8) {
        const int i = bi + r, j = bj + tx;
        if (i < n && j < n) {
            double y = 0.0;
            for (int p = 0; p < parts; ++p) y += Y[p * part_stride + (int64_t)i * n + j];
            y = 0.5 * (y + tt[tx][r]);
            const int64_t o = (int64_t)i * n + j;
            const double xn = square ? y : fma(2.0, X[o], -y);
            X[o] = xn;
            f2 = fma(xn, xn, f2);
            if (i == j) tr += xn;
        }
    }
    block_accumulate(tr, f2, cout, cout + 1);
}

// H = a I + b * sum_p G_p   (k x k)
__global__ void __launch_bounds__(256) ns_poly_kernel(const double* __restrict__ G, int64_t part_stride, int parts, double* __restrict__ H, int k,
                                                      double a, double b) {
    const int64_t count = (int64_t)k * k;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        double g = 0.0;
        for (int p = 0; p < parts; ++p) g += G[p * part_stride + i];
        H[i] = fma(b, g, (i / k == i % k) ? a : 0.0);
    }
}

__global__ void __launch_bounds__(256) sum_parts_kernel(const double* __restrict__ parts, int64_t part_stride, int nparts, double* __restrict__ out,
                                                        int64_t count) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int p = 0; p < nparts; ++p) s += parts[p * part_stride + i];
        out[i] = s;
    }
}

// out[0] = max |H - I|   (single block)
__global__ void __launch_bounds__(1024) purify_dev_kernel(const double* __restrict__ H, int k, double* __restrict__ out) {
    __shared__ double red[32];
    double m = 0.0;
    for (int64_t i = threadIdx.x; i < (int64_t)k * k; i += 1024) {
        const double v = fabs(H[i] - ((i / k == i % k) ? 1.0 : 0.0));
        m = v > m ? v : m;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = red[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (threadIdx.x == 0) out[0] = v;
    }
}

__global__ void purify_finish_kernel(const double* __restrict__ ctrl, int iters, double ne, const double* __restrict__ scal, double* __restrict__ info) {
    // scal: [0] = |A|_F^2, [1] = sum A o P ; ctrl[2 it .. 2 it + 1] = (tr X, |X|_F^2) of iterate `it` ; info[4] is written by purify_dev_kernel
    info[0] = ctrl[2 * iters];
    info[1] = ctrl[2 * iters + 1];
    info[2] = scal[1];
    info[3] = sqrt(scal[0]);
    info[5] = ctrl[0] * sqrt(scal[0]);          // tr A
    info[6] = ctrl[2 * (iters - 1)] - ctrl[2 * (iters - 1) + 1];   // idempotency defect tr(X - X^2) one step before the end
    int lift = 0;                                                   // leading 2X - X^2 steps ~ log2(|A|_F / lambda_cut)
    for (int it = 0; it < iters; ++it) {
        const double tr0 = ctrl[2 * it], f0 = ctrl[2 * it + 1];
        if (fabs(f0 - ne) < fabs(2.0 * tr0 - f0 - ne)) break;
        ++lift;
    }
    info[7] = (double)iters + 1e6 * lift;
}

static inline int split_k(int rows_out, int cols_out, int K) {
    // enough 64 x 64 output tiles to fill the machine?  otherwise split the contraction (partials are summed by the consumer kernel)
    const int tiles = ((rows_out + 63) / 64) * ((cols_out + 63) / 64);
    int parts = 1;
    while (parts < 8 && tiles * parts * 2 <= 160 && K % (parts * 2 * 32) == 0 && K / (parts * 2) >= 64) parts *= 2;
    return parts;
}

static size_t purify_ws_doubles(int n, int ne, int sp2_iters) {
    const int pk = split_k(n, n, n), pg = split_k(ne, ne, n);
    size_t d = 0;
    d += (size_t)n * n;                       // X
    d += (size_t)pk * n * n;                  // Y partials
    d += (size_t)2 * (sp2_iters + 1) + 8;     // ctrl + scalars
    d += (size_t)pg * ne * ne;                // Gram partials
    d += (size_t)ne * ne;                     // H
    d += (size_t)2 * n * ne;                  // U ping-pong
    return d + 64;
}

int dominant_subspace_f64(const double* A, int n, int ne, int sp2_iters, int ns_iters, double* U, double* ws, size_t ws_bytes, double* info,
                          cudaStream_t st) {
    SYN_REQUIRE(A && U && ws && info, "syn_dominant_subspace_f64: null argument");
    SYN_REQUIRE(n >= 2 && ne >= 1 && ne < n && sp2_iters >= 1 && sp2_iters <= 200 && ns_iters >= 0 && ns_iters <= 200,
                "syn_dominant_subspace_f64: bad sizes n=%d ne=%d iters=%d/%d", n, ne, sp2_iters, ns_iters);
    SYN_REQUIRE(ws_bytes >= purify_ws_doubles(n, ne, sp2_iters) * sizeof(double), "syn_dominant_subspace_f64: workspace too small");
    SYN_REQUIRE((((uintptr_t)ws) & 15) == 0, "syn_dominant_subspace_f64: workspace must be 16-byte aligned");
    const int pk = split_k(n, n, n), pg = split_k(ne, ne, n);
    const int64_t nn = (int64_t)n * n, kk = (int64_t)ne * ne;
    double* X = ws;
    double* Y = X + nn;
    double* ctrl = Y + (int64_t)pk * nn;
    double* scal = ctrl + 2 * (sp2_iters + 1);
    double* G = scal + 8;                                          // 2 (iters + 1) + 8 doubles: the matrices stay 16-byte aligned
    double* H = G + (int64_t)pg * kk;
    double* Ua = H + kk;
    double* Ub = Ua + (int64_t)n * ne;
    SYN_CUDA(cudaMemsetAsync(ctrl, 0, sizeof(double) * (2 * (sp2_iters + 1) + 8), st));
    const int eb = (int)((nn + 255) / 256 < 592 ? (nn + 255) / 256 : 592);
    purify_dot_kernel<<<eb, 256, 0, st>>>(A, A, nn, scal);
    if (int rc = launch_status("purify_dot_kernel")) return rc;
    sp2_init_kernel<<<eb, 256, 0, st>>>(A, X, n, scal, ctrl);
    if (int rc = launch_status("sp2_init_kernel")) return rc;
    const int kc = n / pk;
    const syn_gemm_desc_t dxx = pdesc(n, n, kc, pk, n, 1, kc, n, 1, (int64_t)kc * n, n, 1, nn);
    const dim3 tgrid((n + 31) / 32, (n + 31) / 32);
    for (int it = 0; it < sp2_iters; ++it) {
        if (int rc = gemm_f64(dxx, X, X, Y, st)) return rc;
        sp2_step_kernel<<<tgrid, 256, 0, st>>>(Y, nn, pk, X, n, (double)ne, ctrl + 2 * it, ctrl + 2 * (it + 1));
        if (int rc = launch_status("sp2_step_kernel")) return rc;
    }
    purify_dot_kernel<<<eb, 256, 0, st>>>(A, X, nn, scal + 1);
    if (int rc = launch_status("purify_dot_kernel")) return rc;
    purify_finish_kernel<<<1, 1, 0, st>>>(ctrl, sp2_iters, (double)ne, scal, info);
    if (int rc = launch_status("purify_finish_kernel")) return rc;
    // U0 = P[:, :ne]
    SYN_CUDA(cudaMemcpy2DAsync(Ua, sizeof(double) * ne, X, sizeof(double) * n, sizeof(double) * ne, n, cudaMemcpyDeviceToDevice, st));
    const int gc = n / pg;
    const syn_gemm_desc_t dg = pdesc(ne, ne, gc, pg, 1, ne, (int64_t)gc * ne, ne, 1, (int64_t)gc * ne, ne, 1, kk);
    const syn_gemm_desc_t du = pdesc(n, ne, ne, 1, ne, 1, 0, ne, 1, 0, ne, 1, 0);
    const int hb = (int)((kk + 255) / 256 < 592 ? (kk + 255) / 256 : 592);
    for (int it = 0; it < ns_iters; ++it) {
        if (int rc = gemm_f64(dg, Ua, Ua, G, st)) return rc;
        ns_poly_kernel<<<hb, 256, 0, st>>>(G, kk, pg, H, ne, 1.5, -0.5);
        if (int rc = launch_status("ns_poly_kernel")) return rc;
        if (int rc = gemm_f64(du, Ua, H, Ub, st)) return rc;
        double* t = Ua; Ua = Ub; Ub = t;
    }
    if (int rc = gemm_f64(dg, Ua, Ua, G, st)) return rc;
    ns_poly_kernel<<<hb, 256, 0, st>>>(G, kk, pg, H, ne, 0.0, 1.0);
    if (int rc = launch_status("ns_poly_kernel")) return rc;
    purify_dev_kernel<<<1, 1024, 0, st>>>(H, ne, info + 4);
    if (int rc = launch_status("purify_dev_kernel")) return rc;
    SYN_CUDA(cudaMemcpyAsync(U, Ua, sizeof(double) * n * ne, cudaMemcpyDeviceToDevice, st));
    return 0;
}


// =====================================================================================================================================
// Fused variant: the WHOLE solver in one persistent cooperative kernel.
//
// The multi-launch sequence above is launch- and latency-bound (every product is a 512^3 GEMM of ~10 us of math behind ~25 us of
// launch gaps, pipeline fill and partial-sum traffic).  Here one CTA per SM stays resident, the iterates live in L2 (2 MB each), and
// a step is: every CTA forms its 32 x 32 output tiles with DMMA from a 3-stage cp.async ring, applies the step's epilogue, and meets
// the others at a grid barrier (one atomic + acquire poll).  All three products are arranged as  C = R1 R2^T  with both operands read
// along contiguous rows ("NT"), so one tile routine serves them:
//     SP2:     Y = X X^T (X symmetric)            -- only the 136 lower tiles are computed, the mirror image is written with them;
//              epilogue X' = Y or 2X - Y by the trace rule, and the new trace / Frobenius norm are accumulated for the next rule
//     Gram:    G = V V^T  with V = U^T kept alongside U          -- lower tiles + mirror, epilogue max |G - I|
//     update:  U' = a U + b U G^T (G symmetric)    -- epilogue writes U' and its transpose V'
// Convergence is decided INSIDE the kernel from the accumulated scalars (every CTA reads the same values after the barrier), so the
// iteration counts adapt to the spectrum with no host round trip: SP2 stops two steps after tr(X - X^2) < 1e-11 ne, Newton-Schulz when
// max |U^T U - I| < 1e-13.  While the smallest singular value of U is still far from 1 the steeper map 2x - x^3 replaces 1.5x - 0.5x^3.
constexpr int PF_THREADS = 256, PF_BK = 128, PF_STAGES = 3, PF_LD = PF_BK + 4, PF_T = 32, PF_STEEP = 8;
constexpr bool PF_SCALED_NS = true;
#ifndef SYN_PF_L0_ORTHO_INV
#define SYN_PF_L0_ORTHO_INV 32.0
#endif
#ifndef SYN_PF_L0_PROJ_INV
#define SYN_PF_L0_PROJ_INV 512.0
#endif
constexpr double PF_L0_ORTHO = 1.0 / SYN_PF_L0_ORTHO_INV, PF_L0_PROJ = 1.0 / SYN_PF_L0_PROJ_INV;
constexpr size_t PF_SMEM = (size_t)PF_STAGES * 2 * PF_T * PF_LD * sizeof(double);

struct PurifyArgs {
    const double* A;
    double *X0, *X1, *Ua, *Ub, *Va, *Vb, *G, *G2, *Uout, *ctrl, *info;
    unsigned* bar;
    int n, ne, sp2_max, ns_max;
    int ns_only;             // 1: A is an (n x ne) matrix with row stride ld0 whose columns are to be orthonormalised (no projection phase)
    int64_t ld0;
    int cx;                  // 1: A is the interleaved real embedding of a complex Hermitian matrix (every entry a + ib as [[a, -b], [b, a]]):
                             //    every iterate is an embedding too, so only the EVEN rows of each product are computed (a thread's accumulator
                             //    pair is one complex entry (a, -b)) and the odd rows are written from them -- half the DMMA work
};

__device__ __forceinline__ unsigned pf_ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void pf_grid_barrier(unsigned* ctr, unsigned& target) {
    // arrive with a release reduction, wait with acquire loads: the CTA barrier orders the other threads' stores before the release and
    // their later loads after the acquire (all cross-CTA data is read through L2: ld.cg / cp.async.cg), so no separate fences are needed
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        for (unsigned spin = 0; pf_ld_acquire(ctr) < target; ++spin)
            if (spin > (1u << 23)) __trap();        // a lost CTA would otherwise hang the GPU
    }
    __syncthreads();
}

// acc[i][j][0..1] of this warp's 16 x 16 sub-tile of  C(32 x 32) = R1[i0.., :K] R2[j0.., :K]^T ; warps 4..7 take the upper half of
// every PF_BK-wide k-stage and are folded into warps 0..3 through shared memory at the end.  Returns true for the warps that own
// the result.  K even (the tail of the last stage is zero-filled), rows 16-byte aligned.  128-wide stages: four barriers per 512-long
// contraction instead of eight, and three quarters of it in flight in the ring.
__device__ __forceinline__ bool pf_tile_nt(const double* __restrict__ R1, int64_t ld1, const double* __restrict__ R2, int64_t ld2, int i0, int j0, int K,
                                           double* smem, double (&acc)[2][2][2]) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wg = warp >> 2, q = warp & 3, wm0 = (q >> 1) * 16, wn0 = (q & 1) * 16;
    constexpr int TILE = PF_T * PF_LD;                       // doubles per operand tile
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    auto fill = [&](int stage, int k0) {
        double* sa = smem + (size_t)stage * 2 * TILE;
        constexpr int CPR = PF_BK / 2;                       // 16-byte chunks per row of a stage
#pragma unroll
        for (int p = 0; p < (2 * PF_T * CPR) / PF_THREADS; p++) {
            const int c = tid + PF_THREADS * p;              // 16-byte chunks: operand | row | chunk
            const int op = c / (PF_T * CPR), row = (c / CPR) % PF_T, ch = c % CPR;
            const bool ok = k0 + ch * 2 < K;                 // K may end inside a stage (K % 32 == 0): the tail is zero-filled
            const double* base = op ? (R2 + (int64_t)(j0 + row) * ld2) : (R1 + (int64_t)(i0 + row) * ld1);
            cp_async<16>(sa + op * TILE + row * PF_LD + ch * 2, base + (ok ? k0 + ch * 2 : 0), ok);
        }
    };
    const int KT = (K + PF_BK - 1) / PF_BK;
    __syncthreads();                                          // the previous user of the ring is done
#pragma unroll
    for (int s = 0; s < PF_STAGES - 1; s++) {
        if (s < KT) fill(s, s * PF_BK);
        cp_async_commit();
    }
    for (int kt = 0; kt < KT; kt++) {
        cp_async_wait<PF_STAGES - 2>();
        __syncthreads();
        const int nk = kt + PF_STAGES - 1;
        if (nk < KT) fill(nk % PF_STAGES, nk * PF_BK);
        cp_async_commit();
        const double* a_s = smem + (size_t)(kt % PF_STAGES) * 2 * TILE;
        const double* b_s = a_s + TILE;
        constexpr int KH = PF_BK / 2;                        // k-extent of a warp group within a stage
        const int kb = wg * KH;
        double af[2][2], bf[2][2];
        auto frags = [&](int buf, int kk) {
#pragma unroll
            for (int i = 0; i < 2; i++) af[buf][i] = a_s[(wm0 + i * 8 + g) * PF_LD + kb + kk + t];
#pragma unroll
            for (int j = 0; j < 2; j++) bf[buf][j] = b_s[(wn0 + j * 8 + g) * PF_LD + kb + kk + t];
        };
        frags(0, 0);
#pragma unroll
        for (int kk = 0; kk < KH; kk += 4) {
            const int cur = (kk >> 2) & 1;
            if (kk + 4 < KH) frags(cur ^ 1, kk + 4);
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int j = 0; j < 2; j++) dmma884(acc[i][j][0], acc[i][j][1], af[cur][i], bf[cur][j]);
        }
    }
    cp_async_wait<0>();
    __syncthreads();                                          // ring free: reuse its head for the k-half reduction
    double* red = smem;                                       // [4 warps][8 values][32 lanes]
    if (wg == 1) {
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 2; j++) {
                red[(q * 8 + (i * 2 + j) * 2 + 0) * 32 + lane] = acc[i][j][0];
                red[(q * 8 + (i * 2 + j) * 2 + 1) * 32 + lane] = acc[i][j][1];
            }
    }
    __syncthreads();
    if (wg == 0) {
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 2; j++) {
                acc[i][j][0] += red[(q * 8 + (i * 2 + j) * 2 + 0) * 32 + lane];
                acc[i][j][1] += red[(q * 8 + (i * 2 + j) * 2 + 1) * 32 + lane];
            }
    }
    return wg == 0;
}

__device__ __forceinline__ void pf_lower_tile(int idx, int& ti, int& tj) {
    int r = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
    while ((r + 1) * (r + 2) / 2 <= idx) ++r;
    while (r * (r + 1) / 2 > idx) --r;
    ti = r;
    tj = idx - r * (r + 1) / 2;
}

__device__ __forceinline__ void pf_atomic_max(double* addr, double v) {       // non-negative doubles order like their bit patterns
    atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

// ctrl layout (doubles, zeroed by the host): [0] |A|_F^2  [1] sum A o P  [2..3] unused  [4 + 2 it] tr X_it  [5 + 2 it] |X_it|_F^2
//                                            [4 + 2 (sp2_max + 2) + it] max |G_it - I| (final iterate only)
//                                            [4 + 2 (sp2_max + 2) + (ns_max + 2) + it] tr G_it
__global__ void __launch_bounds__(PF_THREADS, 1) purify_fused_kernel(const PurifyArgs a) {
    extern __shared__ __align__(16) double pf_smem[];
    const int tid = threadIdx.x, lane = tid & 31, g = lane >> 2, t = lane & 3, q = (tid >> 5) & 3;
    const int wm0 = (q >> 1) * 16, wn0 = (q & 1) * 16;
    const int n = a.n, ne = a.ne;
    const int64_t nn = (int64_t)n * n;
    unsigned target = 0;
    double* ctrl = a.ctrl;
    double* trf = ctrl + 4;
    double* devs = ctrl + 4 + 2 * (a.sp2_max + 2);
    double* trs = devs + (a.ns_max + 2);
    const int64_t gstride = (int64_t)gridDim.x * PF_THREADS, gid = (int64_t)blockIdx.x * PF_THREADS + tid;

    // both Gram accumulators of the Newton-Schulz phase start at zero (they are accumulated into with atomics)
    for (int64_t i = gid; i < (int64_t)ne * ne; i += gstride) { a.G[i] = 0.0; a.G2[i] = 0.0; }
    // ---- |A|_F^2 -------------------------------------------------------------------------------------------------------------------
    {
        double s = 0.0;
        if (!a.ns_only) {
            for (int64_t i = gid; i < nn; i += gstride) { const double x = a.A[i]; s = fma(x, x, s); }
        } else {
            for (int64_t i = gid; i < (int64_t)n * ne; i += gstride) { const double x = a.A[(i / ne) * a.ld0 + (i % ne)]; s = fma(x, x, s); }
        }
        block_accumulate(s, 0.0, ctrl, nullptr);
    }
    pf_grid_barrier(a.bar, target);
    if (!a.ns_only) {
        // ---- X0 = A / |A|_F --------------------------------------------------------------------------------------------------------
        const double f = __ldcg(ctrl);
        const double inv = f > 0.0 ? rsqrt(f) : 0.0;
        double tr = 0.0, f2 = 0.0;
        for (int64_t i = gid; i < nn; i += gstride) {
            const double x = a.A[i] * inv;
            a.X0[i] = x;
            f2 = fma(x, x, f2);
            if (i / n == i % n) tr += x;
        }
        block_accumulate(tr, f2, trf, trf + 1);
    } else {
        // ---- U = A / |A|_F (all singular values <= 1) and its transpose -------------------------------------------------------------
        const double f = __ldcg(ctrl);
        const double inv = f > 0.0 ? rsqrt(f) : 0.0;
        for (int64_t i = gid; i < (int64_t)n * ne; i += gstride) {
            const int64_t r = i / ne, c = i - r * ne;
            const double x = a.A[r * a.ld0 + c] * inv;
            a.Ub[i] = x;
            a.Vb[c * n + r] = x;
        }
    }
    pf_grid_barrier(a.bar, target);

    // ---- SP2 -----------------------------------------------------------------------------------------------------------------------
    double* Xc = a.X0;
    double* Xn = a.X1;
    const int T = n / PF_T, lower = T * (T + 1) / 2;
    int it = 0;
    int extra = 0;
    int lift = 0;            // leading 2X - X^2 steps: each doubles the small eigenvalues, so lift ~ log2(|A|_F / lambda_cut)
    bool lifting = true;
    double tr0 = 0.0, f0 = 0.0;
    const int sp2_limit = a.ns_only ? 0 : a.sp2_max;
    for (; it < sp2_limit; ++it) {
        tr0 = __ldcg(trf + 2 * it);
        f0 = __ldcg(trf + 2 * it + 1);
        // a step squares the distance to {0, 1} of the eigenvalues on one side and doubles it on the other, so the defect tr(X - X^2) is
        // dominated by the side that was just doubled: two more steps, one of each kind, finish both sides
        const bool conv = fabs(tr0 - ne) < 1e-11 * ne && fabs(tr0 - f0) < 1e-11 * ne;
        if (conv || extra) {
            if (extra == 2) break;
            ++extra;
        }
        // once converged the trace rule only sees rounding noise: the two finishing steps are forced to be one of each kind
        // (x^2 squares the distance of the eigenvalues near 0, 2x - x^2 of those near 1; the pair leaves 2 d^2 and 4 d^2)
        const bool square = extra == 1 ? true : (extra == 2 ? false : fabs(f0 - ne) < fabs(2.0 * tr0 - f0 - ne));
        if (lifting && !square) ++lift; else lifting = false;
        double tr = 0.0, f2 = 0.0;
        if (a.cx) {
            // tall tiles: 32 EVEN rows of the 64-row block I against the 32 columns of block J, J <= 2 I + 1 (the lower triangle in 64-blocks)
            const int TI = n / 64, tall = TI * (TI + 1);
            for (int tile = blockIdx.x; tile < tall; tile += gridDim.x) {
                int I = (int)((sqrt(4.0 * tile + 1.0) - 1.0) * 0.5);
                while ((I + 1) * (I + 2) <= tile) ++I;
                while (I * (I + 1) > tile) --I;
                const int J = tile - I * (I + 1);
                double acc[2][2][2];
                const bool owner = pf_tile_nt(Xc, 2 * (int64_t)n, Xc, n, I * PF_T, J * PF_T, n, pf_smem, acc);
                if (owner) {
                    const bool mirror = J < 2 * I;
                    const double w = mirror ? 4.0 : 2.0;
#pragma unroll
                    for (int i = 0; i < 2; i++)
#pragma unroll
                        for (int j = 0; j < 2; j++) {
                            const int r = 2 * (I * PF_T + wm0 + i * 8 + g), c = J * PF_T + wn0 + j * 8 + 2 * t;
                            double x0 = acc[i][j][0], x1 = acc[i][j][1];
                            if (!square) {
                                x0 = fma(2.0, __ldcg(Xc + (int64_t)r * n + c), -x0);
                                x1 = fma(2.0, __ldcg(Xc + (int64_t)r * n + c + 1), -x1);
                            }
                            *reinterpret_cast<double2*>(Xn + (int64_t)r * n + c) = make_double2(x0, x1);
                            *reinterpret_cast<double2*>(Xn + (int64_t)(r + 1) * n + c) = make_double2(-x1, x0);
                            if (mirror) {
                                *reinterpret_cast<double2*>(Xn + (int64_t)c * n + r) = make_double2(x0, -x1);
                                *reinterpret_cast<double2*>(Xn + (int64_t)(c + 1) * n + r) = make_double2(x1, x0);
                            }
                            f2 = fma(w * x0, x0, fma(w * x1, x1, f2));
                            if (r == c) tr += 2.0 * x0;
                        }
                }
            }
        } else
        for (int tile = blockIdx.x; tile < lower; tile += gridDim.x) {
            int ti, tj;
            pf_lower_tile(tile, ti, tj);
            double acc[2][2][2];
            const bool owner = pf_tile_nt(Xc, n, Xc, n, ti * PF_T, tj * PF_T, n, pf_smem, acc);
            if (owner) {
                const double w = (ti == tj) ? 1.0 : 2.0;
#pragma unroll
                for (int i = 0; i < 2; i++)
#pragma unroll
                    for (int j = 0; j < 2; j++)
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int r = ti * PF_T + wm0 + i * 8 + g, c = tj * PF_T + wn0 + j * 8 + 2 * t + e;
                            const double y = acc[i][j][e];
                            const double xn = square ? y : fma(2.0, __ldcg(Xc + (int64_t)r * n + c), -y);
                            Xn[(int64_t)r * n + c] = xn;
                            if (ti != tj) Xn[(int64_t)c * n + r] = xn;
                            f2 = fma(w * xn, xn, f2);
                            if (r == c) tr += xn;
                        }
            }
        }
        block_accumulate(tr, f2, trf + 2 * (it + 1), trf + 2 * (it + 1) + 1);
        pf_grid_barrier(a.bar, target);
        double* s = Xc; Xc = Xn; Xn = s;
    }
    const int sp2_used = it;
    tr0 = __ldcg(trf + 2 * it);
    f0 = __ldcg(trf + 2 * it + 1);
    // Xc = P.  kept weight sum A o P (no barrier needed before the Gram phase: different outputs)
    if (!a.ns_only) {
        double s = 0.0;
        for (int64_t i = gid; i < nn; i += gstride) s = fma(a.A[i], __ldcg(Xc + i), s);
        block_accumulate(s, 0.0, ctrl + 1, nullptr);
    }

    // ---- Newton-Schulz on U = P[:, :ne]  (V = U^T = P[:ne, :]) -----------------------------------------------------------------------
    const double* U = a.ns_only ? a.Ub : Xc;
    int64_t ldu = a.ns_only ? ne : n;
    const double* V = a.ns_only ? a.Vb : Xc;                  // (ne x n), ld n
    double* Un = a.Ua;  double* Vn = a.Va;
    const int TG = ne / PF_T, lower_g = TG * (TG + 1) / 2, tiles_u = (n / PF_T) * TG;
    // The Gram matrix G = U^T U = V V^T of every step is formed by ALL CTAs: its lower tiles are split KS ways along the contraction
    // (36 tiles x 4 splits at 512 -> 256 instead of 36 busy CTAs of 136) and the partial tiles are accumulated with FP64 atomics into one of
    // two buffers (the other one is cleared during the update phase).  Convergence is read from tr G = sum sigma_i^2, which the same
    // atomics accumulate: after a standard step x -> 1.5 x - 0.5 x^3 every singular value is <= 1, so ne - tr G = sum (1 - sigma_i^2)
    // bounds every 1 - sigma_i^2; max |G - I| of the final iterate (what the host checks) is taken once at the end.
    int KS = 1;
    const int gram_items = a.cx ? (ne / 64) * (ne / 64 + 1) : lower_g;
    while (gram_items * KS * 2 <= (int)gridDim.x + (int)gridDim.x / 8 && (n / (KS * 2)) % 2 == 0 && n / (KS * 2) >= 64) KS *= 2;
    const int Kc = n / KS;
    int ns = 0;
    double dev = 0.0, dev0 = 0.0;
    int steep_steps = PF_STEEP;
    double lo = 1.0;                     // running lower bound (estimate) of the singular values of the iterate, see the scaled step below
    bool last_steep = true;              // the map that produced the current iterate (the start counts as "not yet standard")
    double* Gc = a.G;
    for (;; ++ns) {
        Gc = (ns & 1) ? a.G2 : a.G;
        double* Gnext = (ns & 1) ? a.G : a.G2;
        double trp = 0.0;
        if (a.cx) {
            const int TI = ne / 64, tall = TI * (TI + 1);
            for (int item = blockIdx.x; item < tall * KS; item += gridDim.x) {
                const int tile = item % tall, ks = item / tall;
                int I = (int)((sqrt(4.0 * tile + 1.0) - 1.0) * 0.5);
                while ((I + 1) * (I + 2) <= tile) ++I;
                while (I * (I + 1) > tile) --I;
                const int J = tile - I * (I + 1);
                double acc[2][2][2];
                const bool owner = pf_tile_nt(V + (int64_t)ks * Kc, 2 * (int64_t)n, V + (int64_t)ks * Kc, n, I * PF_T, J * PF_T, Kc, pf_smem, acc);
                if (owner) {
                    const bool mirror = J < 2 * I;
#pragma unroll
                    for (int i = 0; i < 2; i++)
#pragma unroll
                        for (int j = 0; j < 2; j++) {
                            const int r = 2 * (I * PF_T + wm0 + i * 8 + g), c = J * PF_T + wn0 + j * 8 + 2 * t;
                            const double g0 = acc[i][j][0], g1 = acc[i][j][1];
                            atomicAdd(Gc + (int64_t)r * ne + c, g0);
                            atomicAdd(Gc + (int64_t)r * ne + c + 1, g1);
                            atomicAdd(Gc + (int64_t)(r + 1) * ne + c, -g1);
                            atomicAdd(Gc + (int64_t)(r + 1) * ne + c + 1, g0);
                            if (mirror) {
                                atomicAdd(Gc + (int64_t)c * ne + r, g0);
                                atomicAdd(Gc + (int64_t)c * ne + r + 1, -g1);
                                atomicAdd(Gc + (int64_t)(c + 1) * ne + r, g1);
                                atomicAdd(Gc + (int64_t)(c + 1) * ne + r + 1, g0);
                            }
                            if (r == c) trp += 2.0 * g0;
                        }
                }
            }
        } else
        for (int item = blockIdx.x; item < lower_g * KS; item += gridDim.x) {
            const int tile = item % lower_g, ks = item / lower_g;
            int ti, tj;
            pf_lower_tile(tile, ti, tj);
            double acc[2][2][2];
            const bool owner = pf_tile_nt(V + (int64_t)ks * Kc, n, V + (int64_t)ks * Kc, n, ti * PF_T, tj * PF_T, Kc, pf_smem, acc);
            if (owner) {
#pragma unroll
                for (int i = 0; i < 2; i++)
#pragma unroll
                    for (int j = 0; j < 2; j++)
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int r = ti * PF_T + wm0 + i * 8 + g, c = tj * PF_T + wn0 + j * 8 + 2 * t + e;
                            const double gv = acc[i][j][e];
                            atomicAdd(Gc + (int64_t)r * ne + c, gv);
                            if (ti != tj) atomicAdd(Gc + (int64_t)c * ne + r, gv);
                            if (r == c) trp += gv;
                        }
            }
        }
        block_accumulate(trp, 0.0, trs + ns, nullptr);
        pf_grid_barrier(a.bar, target);
        const double tr = __ldcg(trs + ns);
        double gamma = 1.0;
        if (ns == 0) {
            // One scan of the first Gram matrix per CTA (every CTA computes the same numbers from the same data, no extra barrier):
            //   dev0 = max |G - I| decides whether the steep map is used at all (a start that is already near-orthonormal skips it);
            //   rho = max column sum of |G| >= sigma_max^2 (Gershgorin).  The Frobenius scaling of the NS-only start leaves
            //   sigma_max ~ 1 / sqrt(ne) for a generic matrix: gamma = rho^(-1/2) is folded into the coefficients of the first step and
            //   the steep phase is shortened by the log2(gamma) doublings it replaces.
            __shared__ double s_col[PF_THREADS], s_dev[PF_THREADS];
            double best = 0.0, dbest = 0.0;
            for (int c0 = tid; c0 < ne; c0 += PF_THREADS) {
                double sum = 0.0;
                for (int r = 0; r < ne; ++r) {
                    const double v = __ldcg(Gc + (int64_t)r * ne + c0);
                    sum += fabs(v);
                    dbest = fmax(dbest, fabs(v - (r == c0 ? 1.0 : 0.0)));
                }
                best = fmax(best, sum);
            }
            s_col[tid] = best;
            s_dev[tid] = dbest;
            __syncthreads();
            for (int o = PF_THREADS / 2; o > 0; o >>= 1) {
                if (tid < o) { s_col[tid] = fmax(s_col[tid], s_col[tid + o]); s_dev[tid] = fmax(s_dev[tid], s_dev[tid + o]); }
                __syncthreads();
            }
            const double rho = s_col[0];
            dev0 = s_dev[0];
            __syncthreads();
            if (a.ns_only && rho > 0.0 && rho < 1.0) {
                gamma = rsqrt(rho);
                int cut = 0;
                for (double x = gamma; x >= 2.0; x *= 0.5) ++cut;
                steep_steps = PF_STEEP - cut < 2 ? 2 : PF_STEEP - cut;
            }
        }
        if (ns == 0) lo = dev0 > 0.5 ? (a.ns_only ? PF_L0_ORTHO : PF_L0_PROJ) : 1.0;
        if (ns >= 1 && !last_steep && fabs((double)ne - tr) < 8e-13) break;
        if (ns >= a.ns_max) break;
        // U' = ca U + cb U G   (G symmetric: rows of G are its columns), V' = U'^T
        bool steep;
        double ca, cb;
        if (PF_SCALED_NS) {
            // Scaled Newton-Schulz: with the singular values in [lo, 1] the cubic x -> g x (3 - g^2 x^2) / 2, g^2 = 3 / (1 + lo + lo^2), maps
            // both ends to the same value and its maximum to exactly 1 -- the best cubic for that interval (the small end grows ~2.5x per
            // step instead of 2x for the fixed steep map 2x - x^3, and nothing ever exceeds 1, so tr G bounds every 1 - sigma^2 after any
            // step).  lo is only an estimate (1/32 for the blocks of the QR sweep, whose condition numbers are 12-23; 1/512 for the
            // leading columns of a projector): tracked through the maps it reaches 1, where the map is the standard 1.5 x - 0.5 x^3; a
            // smaller true minimum just finishes with standard steps.
            const double g2 = 3.0 / (1.0 + lo + lo * lo), g = sqrt(g2);
            ca = 1.5 * g * gamma;
            cb = -0.5 * g * g2 * gamma * gamma * gamma;
            lo = fmin(1.0, 0.5 * g * lo * (3.0 - g2 * lo * lo));
            steep = false;
        } else {
            // the first steps use x -> 2x - x^3 (slope 2 at 0 instead of 1.5; values near 1 stay within [0.88, 1.09]): the smallest
            // singular values of P[:, :ne] (1e-3 .. 1e-1 on the C2 chain) reach the quadratic region in half the steps
            steep = ns < steep_steps && dev0 > 0.5;
            ca = (steep ? 2.0 : 1.5) * gamma;
            cb = (steep ? -1.0 : -0.5) * gamma * gamma * gamma;
        }
        if (a.cx) {
            for (int tile = blockIdx.x; tile < (n / 64) * TG; tile += gridDim.x) {
                const int I = tile / TG, J = tile - I * TG;
                double acc[2][2][2];
                const bool owner = pf_tile_nt(U, 2 * ldu, Gc, ne, I * PF_T, J * PF_T, ne, pf_smem, acc);
                if (owner) {
#pragma unroll
                    for (int i = 0; i < 2; i++)
#pragma unroll
                        for (int j = 0; j < 2; j++) {
                            const int r = 2 * (I * PF_T + wm0 + i * 8 + g), c = J * PF_T + wn0 + j * 8 + 2 * t;
                            const double u0 = fma(ca, __ldcg(U + (int64_t)r * ldu + c), cb * acc[i][j][0]);
                            const double u1 = fma(ca, __ldcg(U + (int64_t)r * ldu + c + 1), cb * acc[i][j][1]);
                            *reinterpret_cast<double2*>(Un + (int64_t)r * ne + c) = make_double2(u0, u1);
                            *reinterpret_cast<double2*>(Un + (int64_t)(r + 1) * ne + c) = make_double2(-u1, u0);
                            *reinterpret_cast<double2*>(Vn + (int64_t)c * n + r) = make_double2(u0, -u1);
                            *reinterpret_cast<double2*>(Vn + (int64_t)(c + 1) * n + r) = make_double2(u1, u0);
                        }
                }
            }
        } else
        for (int tile = blockIdx.x; tile < tiles_u; tile += gridDim.x) {
            const int ti = tile / TG, tj = tile - ti * TG;
            double acc[2][2][2];
            const bool owner = pf_tile_nt(U, ldu, Gc, ne, ti * PF_T, tj * PF_T, ne, pf_smem, acc);
            if (owner) {
#pragma unroll
                for (int i = 0; i < 2; i++)
#pragma unroll
                    for (int j = 0; j < 2; j++)
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int r = ti * PF_T + wm0 + i * 8 + g, c = tj * PF_T + wn0 + j * 8 + 2 * t + e;
                            const double un = fma(ca, __ldcg(U + (int64_t)r * ldu + c), cb * acc[i][j][e]);
                            Un[(int64_t)r * ne + c] = un;
                            Vn[(int64_t)c * n + r] = un;
                        }
            }
        }
        for (int64_t i = gid; i < (int64_t)ne * ne; i += gstride) Gnext[i] = 0.0;       // the accumulator of the next step
        pf_grid_barrier(a.bar, target);
        last_steep = steep;
        U = Un; ldu = ne; V = Vn;
        Un = (Un == a.Ua) ? a.Ub : a.Ua;
        Vn = (Vn == a.Va) ? a.Vb : a.Va;
    }
    // max |G - I| of the final iterate (Gc is complete: it was read after the barrier that closed its accumulation)
    {
        double dmax = 0.0;
        for (int64_t i = gid; i < (int64_t)ne * ne; i += gstride) {
            const int64_t r = i / ne, c = i - r * ne;
            dmax = fmax(dmax, fabs(__ldcg(Gc + i) - (r == c ? 1.0 : 0.0)));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
        if (lane == 0 && dmax > 0.0) pf_atomic_max(devs + ns, dmax);
        pf_grid_barrier(a.bar, target);
        dev = __ldcg(devs + ns);
    }
    // ---- results -------------------------------------------------------------------------------------------------------------------
    for (int64_t i = gid; i < (int64_t)n * ne; i += gstride) {
        const int64_t r = i / ne, c = i - r * ne;
        a.Uout[i] = __ldcg(U + r * ldu + c);
    }
    if (blockIdx.x == 0 && tid == 0) {
        const double f = __ldcg(ctrl);
        a.info[0] = tr0;
        a.info[1] = f0;
        a.info[2] = __ldcg(ctrl + 1);
        a.info[3] = sqrt(f);
        a.info[4] = dev;
        a.info[5] = __ldcg(trf) * sqrt(f);
        a.info[6] = tr0 - f0;
        a.info[7] = (double)(sp2_used + 1000 * ns) + 1e6 * lift;
    }
}

static size_t purify_fused_ws_doubles(int n, int ne, int sp2_max, int ns_max) {
    return (size_t)2 * n * n + (size_t)4 * n * ne + (size_t)2 * ne * ne + 4 + 2 * (sp2_max + 2) + 2 * (ns_max + 2) + 64;
}

static bool ortho_fused_fits(int m, int q) { return m % PF_T == 0 && q % PF_T == 0 && m >= 128 && q >= 32 && q <= m; }

static bool purify_fused_fits(int n, int ne) { return n % PF_T == 0 && ne % PF_T == 0 && n >= 128 && ne >= 32 && ne < n; }

int dominant_subspace_fused_f64(const double* A, int n, int ne, int sp2_max, int ns_max, double* U, double* ws, size_t ws_bytes, double* info,
                                cudaStream_t st, int ns_only = 0, int64_t ld0 = 0, int cx = 0) {
    SYN_REQUIRE(A && U && ws && info, "syn_dominant_subspace_f64: null argument");
    SYN_REQUIRE(!cx || (!ns_only && n % 64 == 0 && ne % 64 == 0), "syn_dominant_subspace_c128: embedded sizes must be multiples of 64 (n=%d ne=%d)", n, ne);
    SYN_REQUIRE(ns_only ? ortho_fused_fits(n, ne) : purify_fused_fits(n, ne),
                "syn_dominant_subspace_f64 (fused): n and ne must be multiples of 32, n >= 128 (n=%d ne=%d)", n, ne);
    SYN_REQUIRE(sp2_max >= 1 && sp2_max <= 400 && ns_max >= 0 && ns_max <= 400, "syn_dominant_subspace_f64: bad iteration limits");
    SYN_REQUIRE(ws_bytes >= purify_fused_ws_doubles(n, ne, sp2_max, ns_max) * sizeof(double), "syn_dominant_subspace_f64: workspace too small");
    SYN_REQUIRE(((((uintptr_t)ws) | ((uintptr_t)A)) & 15) == 0, "syn_dominant_subspace_f64: A and the workspace must be 16-byte aligned");
    auto kern = purify_fused_kernel;
    static PerDevice configured;          // value = co-resident CTAs of this kernel on the device
    const int dev_ = current_device();
    int max_ctas = 0;
    if (!configured.get(dev_, &max_ctas)) {
        SYN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PF_SMEM));
        int per_sm = 0;
        SYN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PF_THREADS, PF_SMEM));
        max_ctas = per_sm * sm_count();
        configured.set(dev_, max_ctas);
    }
    SYN_REQUIRE(max_ctas >= 1, "syn_dominant_subspace_f64: the fused kernel does not fit this device");
    const int64_t nn = (int64_t)n * n, nk = (int64_t)n * ne;
    PurifyArgs a;
    a.A = A;
    a.X0 = ws; a.X1 = a.X0 + nn;
    a.Ua = a.X1 + nn; a.Ub = a.Ua + nk; a.Va = a.Ub + nk; a.Vb = a.Va + nk;
    a.G = a.Vb + nk;
    a.G2 = a.G + (int64_t)ne * ne;
    a.ctrl = a.G2 + (int64_t)ne * ne;
    const size_t ctrl_doubles = 4 + 2 * (sp2_max + 2) + 2 * (ns_max + 2);
    a.bar = reinterpret_cast<unsigned*>(a.ctrl + ctrl_doubles);
    a.Uout = U; a.info = info;
    a.n = n; a.ne = ne; a.sp2_max = sp2_max; a.ns_max = ns_max;
    a.ns_only = ns_only; a.ld0 = ld0; a.cx = cx;
    SYN_CUDA(cudaMemsetAsync(a.ctrl, 0, sizeof(double) * (ctrl_doubles + 2), st));
    const int T = n / PF_T, lower = cx ? (n / 64) * (n / 64 + 1) : T * (T + 1) / 2;
    int grid = lower < max_ctas ? lower : max_ctas;
    void* args[] = {&a};
    SYN_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(PF_THREADS), args, PF_SMEM, st));
    note_launch();
    return 0;
}

// ---- complex Hermitian problems: planar in / planar out around the embedded solver ----------------------------------------------------
// E (2m x 2m) = interleaved real embedding of H = Hre + i Him: entry (i, j) -> [[re, -im], [im, re]]
__global__ void __launch_bounds__(256) embed_hermitian_kernel(const double* __restrict__ Hre, const double* __restrict__ Him, int m,
                                                              double* __restrict__ E) {
    const int64_t total = (int64_t)m * m;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / m, j = e - i * m;
        const double re = Hre[e], im = Him[e];
        *reinterpret_cast<double2*>(E + (2 * i) * 2 * m + 2 * j) = make_double2(re, -im);
        *reinterpret_cast<double2*>(E + (2 * i + 1) * 2 * m + 2 * j) = make_double2(im, re);
    }
}

// even columns of the embedded basis Ue (2m x 2k): complex column j = Ue[0::2, 2j] + i Ue[1::2, 2j]
__global__ void __launch_bounds__(256) unembed_columns_kernel(const double* __restrict__ Ue, int m, int k, double* __restrict__ Ure,
                                                              double* __restrict__ Uim) {
    const int64_t total = (int64_t)m * k;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / k, j = e - i * k;
        Ure[e] = Ue[(2 * i) * 2 * k + 2 * j];
        Uim[e] = Ue[(2 * i + 1) * 2 * k + 2 * j];
    }
}

static bool purify_c128_fits(int m, int k) { return m % 32 == 0 && k % 32 == 0 && m >= 64 && k >= 32 && k < m; }

static size_t purify_c128_ws_doubles(int m, int k, int sp2_max, int ns_max) {
    return purify_fused_ws_doubles(2 * m, 2 * k, sp2_max, ns_max) + (size_t)4 * m * m + (size_t)4 * m * k;
}

}  // namespace syn

extern "C" int syn_dominant_subspace_c128_fits(int m, int k) { return syn::purify_c128_fits(m, k) ? 1 : 0; }

extern "C" size_t syn_dominant_subspace_c128_workspace(int m, int k, int sp2_max, int ns_max) {
    return syn::purify_c128_ws_doubles(m, k, sp2_max, ns_max) * sizeof(double);
}

extern "C" int syn_dominant_subspace_c128(const double* Hre, const double* Him, int m, int k, int sp2_max, int ns_max, double* Ure, double* Uim,
                                          void* ws, size_t ws_bytes, double* info, void* stream) {
    using namespace syn;
    SYN_REQUIRE(Hre && Him && Ure && Uim && ws && info, "syn_dominant_subspace_c128: null argument");
    SYN_REQUIRE(purify_c128_fits(m, k), "syn_dominant_subspace_c128: m and k must be multiples of 32 with 32 <= k < m, m >= 64 (m=%d k=%d)", m, k);
    SYN_REQUIRE(ws_bytes >= purify_c128_ws_doubles(m, k, sp2_max, ns_max) * sizeof(double), "syn_dominant_subspace_c128: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    double* E = (double*)ws;                                   // (2m x 2m) embedding
    double* Ue = E + (size_t)4 * m * m;                        // (2m x 2k) embedded basis
    double* inner = Ue + (size_t)4 * m * k;
    const size_t inner_bytes = ws_bytes - ((size_t)4 * m * m + (size_t)4 * m * k) * sizeof(double);
    const int blocks = (int)(((int64_t)m * m + 255) / 256 < 4096 ? ((int64_t)m * m + 255) / 256 : 4096);
    embed_hermitian_kernel<<<blocks, 256, 0, st>>>(Hre, Him, m, E);
    if (int rc = launch_status("embed_hermitian_kernel")) return rc;
    if (int rc = dominant_subspace_fused_f64(E, 2 * m, 2 * k, sp2_max, ns_max, Ue, inner, inner_bytes, info, st, 0, 0, 1)) return rc;
    const int blocks2 = (int)(((int64_t)m * k + 255) / 256 < 4096 ? ((int64_t)m * k + 255) / 256 : 4096);
    unembed_columns_kernel<<<blocks2, 256, 0, st>>>(Ue, m, k, Ure, Uim);
    return launch_status("unembed_columns_kernel");
}

extern "C" size_t syn_dominant_subspace_workspace_f64(int n, int ne, int sp2_iters) {
    if (n < 2 || ne < 1 || sp2_iters < 1) return 0;
    return syn::purify_ws_doubles(n, ne, sp2_iters) * sizeof(double);
}

extern "C" size_t syn_dominant_subspace_fused_workspace_f64(int n, int ne, int sp2_max, int ns_max) {
    if (n < 2 || ne < 1 || sp2_max < 1 || ns_max < 0) return 0;
    return syn::purify_fused_ws_doubles(n, ne, sp2_max, ns_max) * sizeof(double);
}

// out[i] = sum_p parts[p * part_stride + i]: the reduction that follows a split-K GEMM (the rows x rows Gram product of the sweep)
extern "C" int syn_sum_parts_f64(const double* parts, int64_t part_stride, int nparts, double* out, int64_t count, void* stream) {
    using namespace syn;
    SYN_REQUIRE(parts && out && nparts >= 1 && count >= 1 && count <= (int64_t)46340 * 46340, "syn_sum_parts_f64: bad arguments");
    const int blocks = (int)((count + 255) / 256 < 1184 ? (count + 255) / 256 : 1184);
    sum_parts_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(parts, part_stride, nparts, out, count);
    return launch_status("sum_parts_kernel");
}

extern "C" int syn_orthonormalize_columns_fits(int m, int q) { return syn::ortho_fused_fits(m, q) ? 1 : 0; }

extern "C" size_t syn_orthonormalize_columns_workspace_f64(int m, int q, int ns_max) {
    if (m < 2 || q < 1 || ns_max < 0) return 0;
    return syn::purify_fused_ws_doubles(m, q, 1, ns_max) * sizeof(double);
}

extern "C" int syn_orthonormalize_columns_f64(const double* A, int64_t lda, int m, int q, int ns_max, double* Q, void* ws, size_t ws_bytes,
                                              double* info, void* stream) {
    return syn::dominant_subspace_fused_f64(A, m, q, 1, ns_max, Q, (double*)ws, ws_bytes, info, (cudaStream_t)stream, 1, lda);
}

extern "C" int syn_dominant_subspace_fused_fits(int n, int ne) { return syn::purify_fused_fits(n, ne) ? 1 : 0; }

extern "C" int syn_dominant_subspace_fused_f64(const double* A, int n, int ne, int sp2_max, int ns_max, double* U, void* ws, size_t ws_bytes,
                                               double* info, void* stream) {
    return syn::dominant_subspace_fused_f64(A, n, ne, sp2_max, ns_max, U, (double*)ws, ws_bytes, info, (cudaStream_t)stream);
}

extern "C" int syn_dominant_subspace_f64(const double* A, int n, int ne, int sp2_iters, int ns_iters, double* U, void* ws, size_t ws_bytes,
                                         double* info, void* stream) {
    return syn::dominant_subspace_f64(A, n, ne, sp2_iters, ns_iters, U, (double*)ws, ws_bytes, info, (cudaStream_t)stream);
}

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
    syn_gemm_desc_t d;
    d.M = M; d.N = N; d.K = K; d.batch = batch;
    d.a_m = PIX(a_m); d.a_k = PIX(a_k); d.a_b = PIX(a_b);
    d.b_k = PIX(b_k); d.b_n = PIX(b_n); d.b_b = PIX(b_b);
    d.c_m = PIX(c_m); d.c_n = PIX(c_n); d.c_b = PIX(c_b);
    d.alpha = alpha; d.beta = beta;
    return d;
}

__device__ __forceinline__ void block_accumulate(double a, double b, double* out0, double* out1) {
    __shared__ double r0[32], r1[32];
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

__global__ void purify_finish_kernel(const double* __restrict__ ctrl, int iters, const double* __restrict__ scal, double* __restrict__ info) {
    // scal: [0] = |A|_F^2, [1] = sum A o P ; ctrl[2 it .. 2 it + 1] = (tr X, |X|_F^2) of iterate `it` ; info[4] is written by purify_dev_kernel
    info[0] = ctrl[2 * iters];
    info[1] = ctrl[2 * iters + 1];
    info[2] = scal[1];
    info[3] = sqrt(scal[0]);
    info[5] = ctrl[0] * sqrt(scal[0]);          // tr A
    info[6] = ctrl[2 * (iters - 1)] - ctrl[2 * (iters - 1) + 1];   // idempotency defect tr(X - X^2) one step before the end
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
    purify_finish_kernel<<<1, 1, 0, st>>>(ctrl, sp2_iters, scal, info);
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

}  // namespace syn

extern "C" size_t syn_dominant_subspace_workspace_f64(int n, int ne, int sp2_iters) {
    if (n < 2 || ne < 1 || sp2_iters < 1) return 0;
    return syn::purify_ws_doubles(n, ne, sp2_iters) * sizeof(double);
}

extern "C" int syn_dominant_subspace_f64(const double* A, int n, int ne, int sp2_iters, int ns_iters, double* U, void* ws, size_t ws_bytes,
                                         double* info, void* stream) {
    return syn::dominant_subspace_f64(A, n, ne, sp2_iters, ns_iters, U, (double*)ws, ws_bytes, info, (cudaStream_t)stream);
}

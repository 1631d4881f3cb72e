// Shifted Cholesky factorisation of the density-matrix Gram matrices, sm_100a.
//
// The rounding sweep diagonalises G = M E M^T (symmetric PSD, n = chi d <= 1024).  One-sided Jacobi applied to the rows of G itself
// works on G^2: the spectrum is squared once more and convergence takes 13 sweeps on the C2 plateau, 27 on rank-deficient sites.
// Applied to the rows of B = L^T, where G + delta I = L L^T, the same kernel works on a matrix similar to G itself (one step of
// the LR-Cholesky iteration closer to diagonal): 10 and 15 sweeps (measured, tools/chol_experiment.py), and the eigenvalues come
// out on the sigma scale: row_i -> sqrt(lambda_i + delta) q_i^T with q_i the eigenvectors of G  (Veselic & Hari).
// delta = 2 n eps max_i G_ii keeps the factorisation defined when G is singular; it shifts every eigenvalue by exactly delta
// (subtracted again by the finalize kernel) and leaves the eigenvectors alone.
//
// Blocked right-looking, 64 columns per panel: one kernel factors the diagonal block (every CTA redundantly, in registers)
// and solves its slab of the panel below it with the row held in registers; the trailing update is one strided DMMA GEMM.
// No counterpart in the reference (its only eigh is the unfinished MPO:228); this is part of the north_star SVD rounding.
#include "common.cuh"

namespace syn {

int gemm_f64(const syn_gemm_desc_t& d, const double* A, const double* B, double* C, cudaStream_t st);

constexpr int CH_NB = 64;          // panel width
constexpr int CH_SLAB = 128;       // rows of the panel solved by one CTA (one thread per row)
constexpr int CH_THREADS = 256;
constexpr int CH_LD = CH_NB + 1;   // shared-memory row stride: conflict-free for "thread t reads row t"

__global__ void __launch_bounds__(1024, 1) chol_shift_kernel(double* __restrict__ G, int64_t ld, int n, double* __restrict__ shift) {
    __shared__ double red[32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double mx = 0.0;
    for (int i = tid; i < n; i += 1024) mx = fmax(mx, G[(int64_t)i * ld + i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
    for (int k = 1; k < 32; k++) mx = fmax(mx, red[k]);
    const double delta = 2.0 * (double)n * 2.220446049250313e-16 * mx;
    for (int i = tid; i < n; i += 1024) G[(int64_t)i * ld + i] += delta;
    if (tid == 0) shift[0] = delta;
}

static syn_index_t CIX(int64_t stride) {
    syn_index_t i;
    i.outer = 0; i.inner = stride; i.div = INT_MAX; i._pad = 0;
    return i;
}

// Cholesky factor of a (16 BS) x (16 BS) symmetric positive definite block held in REGISTERS: 16 x 16 threads own BS x BS blocks
// (the lower triangle of blocks works).  Step j: the owners of column j publish it unscaled (its entry j is the pivot) in a
// double-buffered shared vector -- one barrier per step --, everybody derives 1/pivot itself and applies the rank-1 update to its
// block.  On return a holds L (garbage above the diagonal of the diagonal blocks), invd[j] = 1 / L_jj.
template <int BS>
__device__ __forceinline__ void potrf_regs(double (&a)[BS][BS], double* __restrict__ col /* [2][16 BS] */, double* __restrict__ invd,
                                           double floor_piv, int ty, int tx, int block_cols) {
    constexpr int N = 16 * BS;
    for (int jb = 0; jb < block_cols; ++jb) {
#pragma unroll
        for (int jj = 0; jj < BS; ++jj) {
            const int j = BS * jb + jj;
            double* cb = col + (j & 1) * N;
            const bool col_owner = (tx == jb && ty >= jb);
            if (col_owner) {
#pragma unroll
                for (int r = 0; r < BS; ++r) cb[BS * ty + r] = a[r][jj];
            }
            __syncthreads();
            double piv = cb[j];
            piv = piv > floor_piv ? piv : floor_piv;                    // only reachable through rounding: G + delta I is positive definite
            if (ty >= tx && tx >= jb) {
                const double rinv = rcp_newton2(piv);
                double cr[BS], cc[BS];
#pragma unroll
                for (int r = 0; r < BS; ++r) cr[r] = cb[BS * ty + r];
#pragma unroll
                for (int c = 0; c < BS; ++c) cc[c] = cb[BS * tx + c] * rinv;
#pragma unroll
                for (int r = 0; r < BS; ++r)
#pragma unroll
                    for (int c = 0; c < BS; ++c)
                        if (BS * tx + c > j && BS * ty + r > j) a[r][c] = fma(-cr[r], cc[c], a[r][c]);
            }
            if (col_owner) {                                            // column j of L
                const double rs = rsqrt_newton2(piv);
                double l = piv * rs;
                l = fma(fma(-l, l, piv), 0.5 * rs, l);                  // sqrt(piv) to the last bit or two
#pragma unroll
                for (int r = 0; r < BS; ++r) {
                    if (BS * ty + r > j) a[r][jj] *= rs;
                    else if (BS * ty + r == j) { a[r][jj] = l; invd[j] = rs; }
                }
            }
        }
    }
}

// Small problems (n <= 128), one CTA per problem, batched: shift, factor and transposed write in one launch.
template <int BS>
__global__ void __launch_bounds__(CH_THREADS, 1)
chol_small_kernel(const double* __restrict__ G, int64_t ld, int64_t bs, int n, double* __restrict__ B, int64_t ldb, int64_t bbs,
                  double* __restrict__ shift) {
    constexpr int N = 16 * BS;
    __shared__ double col[2 * N];
    __shared__ double invd[N];
    __shared__ double red[CH_THREADS / 32];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    G += (int64_t)blockIdx.x * bs;
    B += (int64_t)blockIdx.x * bbs;
    double a[BS][BS];
    double mx = 0.0;
#pragma unroll
    for (int r = 0; r < BS; ++r)
#pragma unroll
        for (int c = 0; c < BS; ++c) {
            const int i = BS * ty + r, j = BS * tx + c;
            double v = (i == j) ? 1.0 : 0.0;
            if (i < n && j < n) {
                v = (j <= i) ? G[(int64_t)i * ld + j] : 0.0;
                if (i == j) mx = fmax(mx, v);
            }
            a[r][c] = v;
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int k = 1; k < CH_THREADS / 32; k++) mx = fmax(mx, red[k]);
    const double delta = 2.0 * (double)n * 2.220446049250313e-16 * mx;
    if (tid == 0) shift[blockIdx.x] = delta;
    if (ty == tx) {
#pragma unroll
        for (int r = 0; r < BS; ++r)
            if (BS * ty + r < n) a[r][r] += delta;
    }
    potrf_regs<BS>(a, col, invd, 0.25 * delta, ty, tx, (n + BS - 1) / BS);
#pragma unroll
    for (int c = 0; c < BS; ++c)
#pragma unroll
        for (int r = 0; r < BS; ++r) {
            const int i = BS * ty + r, j = BS * tx + c;       // L[i][j] -> B[j][i]; zeros below the diagonal of B
            if (i < n && j < n) B[(int64_t)j * ldb + i] = (j <= i) ? a[r][c] : 0.0;
        }
}

// Panel k0: D = G[k0:k0+nb, k0:k0+nb] -> L11 (lower); rows below: L21 = A21 L11^-T.  L is written back into G (lower part, the
// trailing GEMM reads L21 from there) and, transposed, into B.
__global__ void __launch_bounds__(CH_THREADS, 1)
chol_panel_kernel(double* __restrict__ G, int64_t ld, int n, int k0, int nb, double* __restrict__ B, int64_t ldb, const double* __restrict__ shift) {
    __shared__ __align__(16) double D[CH_NB * CH_NB];                  // L11, row-major, read as broadcast double2 by the panel solve
    __shared__ double invd[CH_NB];
    extern __shared__ __align__(16) unsigned char chol_raw[];
    double* T = reinterpret_cast<double*>(chol_raw);                    // [CH_SLAB][CH_LD]
    const int tid = threadIdx.x;
    const double floor_piv = 0.25 * shift[0];
    // Diagonal block, padded to 64 x 64 with the identity, factored in REGISTERS: 16 x 16 threads own 4 x 4 blocks (the lower
    // triangle of blocks works).  Step j: the owners of column j publish it unscaled (its entry j is the pivot) in a double-buffered
    // shared vector -- one barrier per step --, everybody derives 1/pivot itself and applies the rank-1 update to its block.
    __shared__ double col[2 * CH_NB];
    const int ty = tid >> 4, tx = tid & 15;
    double a[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int i = 4 * ty + r, j = 4 * tx + c;
            double v = (i == j) ? 1.0 : 0.0;
            if (i < nb && j < nb) v = (j <= i) ? G[(int64_t)(k0 + i) * ld + k0 + j] : 0.0;
            a[r][c] = v;
        }
    potrf_regs<4>(a, col, invd, floor_piv, ty, tx, CH_NB / 4);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) D[(4 * ty + r) * CH_NB + 4 * tx + c] = (4 * tx + c <= 4 * ty + r) ? a[r][c] : 0.0;
    __syncthreads();
    if (blockIdx.x == 0) {
        for (int idx = tid; idx < nb * nb; idx += CH_THREADS) {
            const int i = idx / nb, j = idx - i * nb;
            if (j <= i) {
                const double v = D[i * CH_NB + j];
                G[(int64_t)(k0 + i) * ld + k0 + j] = v;
                B[(int64_t)(k0 + j) * ldb + k0 + i] = v;
            }
        }
    }
    const int below = n - k0 - nb;
    const int r0 = k0 + nb + blockIdx.x * CH_SLAB;
    const int rows = min(CH_SLAB, n - r0);
    if (below <= 0 || rows <= 0) return;
    for (int idx = tid; idx < CH_SLAB * CH_NB; idx += CH_THREADS) {
        const int t = idx / CH_NB, j = idx - t * CH_NB;
        T[t * CH_LD + j] = (t < rows && j < nb) ? G[(int64_t)(r0 + t) * ld + k0 + j] : 0.0;
    }
    __syncthreads();
    if (tid < CH_SLAB) {
        // x L11^T = a for my row, 8 columns at a time (left-looking over 8 x 8 blocks: a loop, not 2016 unrolled FMAs -- straight-line
        // code that runs once is bound by instruction fetch)
        double* row = T + tid * CH_LD;
        for (int J = 0; J < CH_NB / 8; ++J) {
            double x[8];
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) x[kk] = row[8 * J + kk];
            for (int I = 0; I < J; ++I) {
                double xi[8];
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) xi[jj] = row[8 * I + jj];
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const double2* d = reinterpret_cast<const double2*>(D + (8 * J + kk) * CH_NB + 8 * I);
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        const double2 v = d[h];
                        x[kk] = fma(-xi[2 * h], v.x, x[kk]);
                        x[kk] = fma(-xi[2 * h + 1], v.y, x[kk]);
                    }
                }
            }
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                x[jj] *= invd[8 * J + jj];
#pragma unroll
                for (int kk = jj + 1; kk < 8; ++kk) x[kk] = fma(-x[jj], D[(8 * J + kk) * CH_NB + 8 * J + jj], x[kk]);
            }
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) row[8 * J + kk] = x[kk];
        }
    }
    __syncthreads();
    for (int idx = tid; idx < CH_SLAB * CH_NB; idx += CH_THREADS) {       // L21 in place (row-major: j fastest)
        const int t = idx / CH_NB, j = idx - t * CH_NB;
        if (t < rows && j < nb) G[(int64_t)(r0 + t) * ld + k0 + j] = T[t * CH_LD + j];
    }
    for (int idx = tid; idx < CH_SLAB * CH_NB; idx += CH_THREADS) {       // and transposed into B (t fastest)
        const int j = idx / CH_SLAB, t = idx - j * CH_SLAB;
        if (t < rows && j < nb) B[(int64_t)(k0 + j) * ldb + r0 + t] = T[t * CH_LD + j];
    }
}

static int chol_upper_one(double* G, int64_t ld, int n, double* B, int64_t ldb, double* shift, cudaStream_t st) {
    static PerDevice configured;
    const int dev_ = current_device();
    const size_t smem = (size_t)CH_SLAB * CH_LD * sizeof(double);
    if (!configured.get(dev_)) {
        SYN_CUDA(cudaFuncSetAttribute(chol_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.set(dev_);
    }
    SYN_CUDA(cudaMemset2DAsync(B, (size_t)ldb * sizeof(double), 0, (size_t)n * sizeof(double), (size_t)n, st));
    chol_shift_kernel<<<1, 1024, 0, st>>>(G, ld, n, shift);
    if (int rc = launch_status("chol_shift_kernel")) return rc;
    for (int k0 = 0; k0 < n; k0 += CH_NB) {
        const int nb = (n - k0) < CH_NB ? (n - k0) : CH_NB;
        const int below = n - k0 - nb;
        const int ctas = below > 0 ? (below + CH_SLAB - 1) / CH_SLAB : 1;
        chol_panel_kernel<<<ctas, CH_THREADS, smem, st>>>(G, ld, n, k0, nb, B, ldb, shift);
        if (int rc = launch_status("chol_panel_kernel")) return rc;
        if (below > 0) {                 // A22 -= L21 L21^T  (full square: the upper half is not read again, but stays symmetric)
            syn_gemm_desc_t d = {};
            d.M = below; d.N = below; d.K = nb; d.batch = 1;
            d.a_m = CIX(ld); d.a_k = CIX(1); d.a_b = CIX(0);
            d.b_k = CIX(1); d.b_n = CIX(ld); d.b_b = CIX(0);
            d.c_m = CIX(ld); d.c_n = CIX(1); d.c_b = CIX(0);
            d.alpha = -1.0; d.beta = 1.0;
            const double* L21 = G + (int64_t)(k0 + nb) * ld + k0;
            if (int rc = gemm_f64(d, L21, L21, G + (int64_t)(k0 + nb) * ld + (k0 + nb), st)) return rc;
        }
    }
    return 0;
}

int chol_upper_f64(double* G, int64_t ld, int64_t bs, int n, int batch, double* B, int64_t ldb, int64_t bbs, double* shift, cudaStream_t st) {
    SYN_REQUIRE(n >= 1 && n <= 4096 && ld >= n && ldb >= n && batch >= 1, "syn_chol_upper_f64: n=%d ld=%lld ldb=%lld batch=%d out of range", n,
                (long long)ld, (long long)ldb, batch);
    SYN_REQUIRE(G && B && shift, "syn_chol_upper_f64: null pointer");
    if (n <= 128) {                      // one CTA per problem, everything in registers
        for (int b0 = 0; b0 < batch; b0 += 65535) {
            const int nb = (batch - b0) < 65535 ? (batch - b0) : 65535;
            if (n <= 64) chol_small_kernel<4><<<nb, CH_THREADS, 0, st>>>(G + (int64_t)b0 * bs, ld, bs, n, B + (int64_t)b0 * bbs, ldb, bbs, shift + b0);
            else chol_small_kernel<8><<<nb, CH_THREADS, 0, st>>>(G + (int64_t)b0 * bs, ld, bs, n, B + (int64_t)b0 * bbs, ldb, bbs, shift + b0);
            if (int rc = launch_status("chol_small_kernel")) return rc;
        }
        return 0;
    }
    for (int b = 0; b < batch; ++b)
        if (int rc = chol_upper_one(G + (int64_t)b * bs, ld, n, B + (int64_t)b * bbs, ldb, shift + b, st)) return rc;
    return 0;
}

}  // namespace syn

extern "C" int syn_chol_upper_f64(double* G, int64_t ld, int64_t bs, int n, int batch, double* B, int64_t ldb, int64_t bbs, double* shift,
                                  void* stream) {
    return syn::chol_upper_f64(G, ld, bs, n, batch, B, ldb, bbs, shift, (cudaStream_t)stream);
}

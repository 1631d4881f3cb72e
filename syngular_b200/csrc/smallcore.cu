// Applying a SMALL matrix to the middle index of a large 3-index tensor:   Y[q][ro][x] = sum_ri W[ro][ri] X[q][ri][x]
//
// The contractions of a batch of chains with ONE shared MPO core (BASELINE configs[3](ii): chi_W = 4, d = 2 -> an 8 x 8 matrix against
// 4.2 million columns; reference: the per-site einsum of MatrixProductOperator.__matmul__, matrix_product_operator.py:181-192, in a Python
// loop over the states) have K = N = 8: as GEMMs they ran at 0.7 TFLOP/s = 0.7 TB/s on the DMMA tile kernel (27 % of the batched SVD
// sweep).  They are pure streaming: 16 doubles moved per 64 FMAs.  Here one thread owns one (q, x) column: RIN coalesced loads (x is the
// unit-stride index), ROUT x RIN FMAs against W in shared memory, ROUT stores; the output's ro index is a two-level strided index so that the
// caller's layouts ((o, r) split around x) need no transposition pass.  Bound: HBM.
#include "common.cuh"

namespace syn {

struct SmallCoreArgs {
    const double* X;
    const double* W;         // element (ro, ri) at W[ro * w_ro + ri * w_ri]
    int64_t w_ro, w_ri;
    double* Y;
    int64_t Q;
    int L;
    int64_t x_q, x_r, x_l;   // strides of X over q, ri, x
    int64_t y_q, y_l;        // strides of Y over q, x
    int64_t y_ro_outer, y_ro_inner;
    int y_ro_div;            // ro -> (ro / div) * outer + (ro % div) * inner
};

template <int RIN, int ROUT>
__global__ void __launch_bounds__(256) small_core_kernel(const SmallCoreArgs a) {
    __shared__ double w[ROUT * RIN];
    for (int e = threadIdx.x; e < ROUT * RIN; e += blockDim.x) w[e] = a.W[(e / RIN) * a.w_ro + (e % RIN) * a.w_ri];
    __syncthreads();
    const int64_t total = a.Q * a.L;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = e / a.L;
        const int x = (int)(e - q * a.L);
        const double* xp = a.X + q * a.x_q + x * a.x_l;
        double in[RIN];
#pragma unroll
        for (int ri = 0; ri < RIN; ++ri) in[ri] = __ldcs(xp + ri * a.x_r);            // streamed once
        double* yp = a.Y + q * a.y_q + x * a.y_l;
#pragma unroll
        for (int ro = 0; ro < ROUT; ++ro) {
            double s = 0.0;
#pragma unroll
            for (int ri = 0; ri < RIN; ++ri) s = fma(w[ro * RIN + ri], in[ri], s);
            yp[(int64_t)(ro / a.y_ro_div) * a.y_ro_outer + (int64_t)(ro % a.y_ro_div) * a.y_ro_inner] = s;
        }
    }
}

template <int RIN>
static int small_core_launch_rout(const SmallCoreArgs& a, int rout, int grid, cudaStream_t st) {
    switch (rout) {
        case 2: small_core_kernel<RIN, 2><<<grid, 256, 0, st>>>(a); break;
        case 4: small_core_kernel<RIN, 4><<<grid, 256, 0, st>>>(a); break;
        case 8: small_core_kernel<RIN, 8><<<grid, 256, 0, st>>>(a); break;
        case 16: small_core_kernel<RIN, 16><<<grid, 256, 0, st>>>(a); break;
        default: return 2;
    }
    return 0;
}

// 32 x 32 was measured too: 40 us against 22 us for the DMMA GEMM on 65536 columns (1024 dependent FMAs and shared-memory reads per thread) -- not offered
static bool small_core_size(int r) { return r == 2 || r == 4 || r == 8 || r == 16; }

}  // namespace syn

extern "C" int syn_apply_small_core_fits(int rin, int rout) { return (syn::small_core_size(rin) && syn::small_core_size(rout)) ? 1 : 0; }

namespace syn {
int apply_small_core(const double* X, const double* W, int64_t w_ro, int64_t w_ri, double* Y, int64_t Q, int rin, int rout, int L, int64_t x_q,
                     int64_t x_r, int64_t x_l, int64_t y_q, int64_t y_ro_outer, int64_t y_ro_inner, int y_ro_div, int64_t y_l, cudaStream_t stream) {
    if (Q <= 0 || L <= 0) return 0;
    SYN_REQUIRE(X && W && Y, "syn_apply_small_core_f64: null argument");
    SYN_REQUIRE(small_core_size(rin) && small_core_size(rout), "syn_apply_small_core_f64: rin and rout must be 2, 4, 8 or 16 (got %d, %d)", rin, rout);
    SYN_REQUIRE(y_ro_div >= 1, "syn_apply_small_core_f64: y_ro_div must be positive");
    SmallCoreArgs a;
    a.X = X; a.W = W; a.w_ro = w_ro; a.w_ri = w_ri; a.Y = Y; a.Q = Q; a.L = L;
    a.x_q = x_q; a.x_r = x_r; a.x_l = x_l; a.y_q = y_q; a.y_l = y_l;
    a.y_ro_outer = y_ro_outer; a.y_ro_inner = y_ro_inner; a.y_ro_div = y_ro_div;
    const int64_t blocks = (Q * L + 255) / 256;
    const int cap = sm_count() * 32;
    const int grid = (int)(blocks < cap ? blocks : cap);
    cudaStream_t st = stream;
    int rc = 2;
    switch (rin) {
        case 2: rc = small_core_launch_rout<2>(a, rout, grid, st); break;
        case 4: rc = small_core_launch_rout<4>(a, rout, grid, st); break;
        case 8: rc = small_core_launch_rout<8>(a, rout, grid, st); break;
        case 16: rc = small_core_launch_rout<16>(a, rout, grid, st); break;
    }
    SYN_REQUIRE(rc == 0, "syn_apply_small_core_f64: no kernel for rin=%d rout=%d", rin, rout);
    note_launch();
    return launch_status("small_core_kernel");
}
bool small_core_fits(int rin, int rout) { return small_core_size(rin) && small_core_size(rout); }
}  // namespace syn

extern "C" int syn_apply_small_core_f64(const double* X, const double* W, double* Y, int64_t Q, int rin, int rout, int L, int64_t x_q, int64_t x_r,
                                        int64_t x_l, int64_t y_q, int64_t y_ro_outer, int64_t y_ro_inner, int y_ro_div, int64_t y_l, void* stream) {
    return syn::apply_small_core(X, W, rin, 1, Y, Q, rin, rout, L, x_q, x_r, x_l, y_q, y_ro_outer, y_ro_inner, y_ro_div, y_l, (cudaStream_t)stream);
}

/* the same with a strided small matrix: element (ro, ri) at W[ro * w_ro + ri * w_ri] (an MPO core read in place, transposed or not) */
extern "C" int syn_apply_small_core_strided_f64(const double* X, const double* W, int64_t w_ro, int64_t w_ri, double* Y, int64_t Q, int rin, int rout,
                                                int L, int64_t x_q, int64_t x_r, int64_t x_l, int64_t y_q, int64_t y_ro_outer, int64_t y_ro_inner,
                                                int y_ro_div, int64_t y_l, void* stream) {
    return syn::apply_small_core(X, W, w_ro, w_ri, Y, Q, rin, rout, L, x_q, x_r, x_l, y_q, y_ro_outer, y_ro_inner, y_ro_div, y_l, (cudaStream_t)stream);
}

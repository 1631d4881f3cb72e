// Small data-movement / elementwise kernels of the hot path (FP64): block assembly for `+`, Kronecker bonds for `*`,
// Frobenius norm and in-place scaling for `normalize`.  All are HBM-bound: one thread per output element, the
// fastest-varying output index on threadIdx.x so stores (and the dominant loads) are coalesced.
#include "common.cuh"

namespace syn {

// `A + B` site (matrix_product_state.py:82-96, matrix_product_operator.py:90-106): direct sum of the bond spaces,
// self's block first.  Cores are viewed as (l, phys, r) with all physical legs flattened.
__global__ void add_site_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ out, int la, int ra, int lb,
                                int rb, int phys, int first, int last) {
    const int lo = first ? la : la + lb, ro = last ? ra : ra + rb;
    const int64_t total = (int64_t)lo * phys * ro;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int r = (int)(idx % ro);
        int64_t t = idx / ro;
        int p = (int)(t % phys);
        int l = (int)(t / phys);
        double v = 0.0;
        if (first) {
            v = r < ra ? A[((int64_t)l * phys + p) * ra + r] : B[((int64_t)l * phys + p) * rb + (r - ra)];
        } else if (last) {
            v = l < la ? A[((int64_t)l * phys + p) * ra + r] : B[((int64_t)(l - la) * phys + p) * rb + r];
        } else if (l < la && r < ra) {
            v = A[((int64_t)l * phys + p) * ra + r];
        } else if (l >= la && r >= ra) {
            v = B[((int64_t)(l - la) * phys + p) * rb + (r - ra)];
        }
        out[idx] = v;
    }
}

// `A * B` site (matrix_product_operator.py:140-152): out[(la,lb), p, (ra,rb)] = A[la,p,ra] * B[lb,p,rb]
__global__ void kron_site_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ out, int la, int ra, int lb,
                                 int rb, int phys) {
    const int ro = ra * rb, lo = la * lb;
    const int64_t total = (int64_t)lo * phys * ro;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int r = (int)(idx % ro);
        int64_t t = idx / ro;
        int p = (int)(t % phys);
        int l = (int)(t / phys);
        int ia = l / lb, ib = l - ia * lb, ja = r / rb, jb = r - ja * rb;
        out[idx] = A[((int64_t)ia * phys + p) * ra + ja] * B[((int64_t)ib * phys + p) * rb + jb];
    }
}

// out[0] = sum x_i^2 (single CTA, deterministic order)
__global__ void __launch_bounds__(1024, 1) sumsq_kernel(const double* __restrict__ x, int64_t n, double* __restrict__ out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 1024) s = fma(x[i], x[i], s);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = warp_sum(red[threadIdx.x]);
        if (threadIdx.x == 0) out[0] = v;
    }
}

// x *= 1/sqrt(*sumsq)  (normalize: matrix_product_state.py:252-256 divides the last core by its Frobenius norm)
__global__ void scale_rsqrt_kernel(double* __restrict__ x, int64_t n, const double* __restrict__ sumsq) {
    const double inv = 1.0 / sqrt(sumsq[0]);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] *= inv;
}

static inline int grid_for(int64_t total, int threads) {
    int64_t b = (total + threads - 1) / threads;
    int64_t cap = (int64_t)sm_count() * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace syn

extern "C" int syn_add_site_f64(const double* A, const double* B, double* out, int la, int ra, int lb, int rb, int phys, int first, int last,
                                void* stream) {
    using namespace syn;
    SYN_REQUIRE(la >= 1 && ra >= 1 && lb >= 1 && rb >= 1 && phys >= 1, "syn_add_site_f64: bad extents");
    SYN_REQUIRE(!(first && last), "syn_add_site_f64: single-site chains cannot be added (the reference crashes there too)");
    if (first) SYN_REQUIRE(la == lb, "syn_add_site_f64: first cores must share the left bond");
    if (last) SYN_REQUIRE(ra == rb, "syn_add_site_f64: last cores must share the right bond");
    int64_t total = (int64_t)(first ? la : la + lb) * phys * (last ? ra : ra + rb);
    add_site_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(A, B, out, la, ra, lb, rb, phys, first, last);
    return launch_status("add_site_kernel");
}

extern "C" int syn_kron_site_f64(const double* A, const double* B, double* out, int la, int ra, int lb, int rb, int phys, void* stream) {
    using namespace syn;
    SYN_REQUIRE(la >= 1 && ra >= 1 && lb >= 1 && rb >= 1 && phys >= 1, "syn_kron_site_f64: bad extents");
    int64_t total = (int64_t)la * lb * phys * ra * rb;
    kron_site_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(A, B, out, la, ra, lb, rb, phys);
    return launch_status("kron_site_kernel");
}

extern "C" int syn_sumsq_f64(const double* x, int64_t n, double* out, void* stream) {
    using namespace syn;
    SYN_REQUIRE(n >= 0, "syn_sumsq_f64: negative length");
    sumsq_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(x, n, out);
    return launch_status("sumsq_kernel");
}

extern "C" int syn_scale_rsqrt_f64(double* x, int64_t n, const double* sumsq, void* stream) {
    using namespace syn;
    scale_rsqrt_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, sumsq);
    return launch_status("scale_rsqrt_kernel");
}

// y[r, c] = act(y[r, c] + bias[c])   act: 0 = identity, 1 = relu.  TensorDense epilogue (layers/TensorDense.py:139-142).
__global__ void bias_act_kernel(double* __restrict__ y, const double* __restrict__ bias, int64_t rows, int cols, int act) {
    const int64_t total = rows * cols;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        double v = y[idx] + (bias ? bias[idx % cols] : 0.0);
        if (act == 1) v = v > 0.0 ? v : 0.0;
        y[idx] = v;
    }
}

extern "C" int syn_bias_act_f64(double* y, const double* bias, int64_t rows, int cols, int act, void* stream) {
    using namespace syn;
    SYN_REQUIRE(rows >= 0 && cols >= 1 && (act == 0 || act == 1), "syn_bias_act_f64: bad arguments");
    if (rows == 0) return 0;
    bias_act_kernel<<<grid_for(rows * cols, 256), 256, 0, (cudaStream_t)stream>>>(y, bias, rows, cols, act);
    return launch_status("bias_act_kernel");
}

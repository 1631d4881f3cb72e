// Fused transfer-matrix inner product `A | B` for batches of matrix-product states, sm_100a.
//
//   E_0 = 1;   E_{k+1}[b, b'] = sum_{a, a', i} A_k[a, i, b] E_k[a, a'] B_k[a', i, b'];   <A|B> = E_N        (bilinear, no conjugation)
//
// Replaces the single 2N-operand opt_einsum.contract of MatrixProductState.__or__ (matrix_product_state.py:116-129; the
// left-to-right order on its index spec IS this transfer-matrix order) and `dot` (:243), for every state of a batch.
//
// Design (B200): one CTA walks one pair of chains from the first site to the last; the transfer matrix E (<= 64 x 64) and the
// half-contracted F = E B_k never leave shared memory, so HBM sees each core exactly once (the GEMM-per-site path re-reads and
// re-writes E and F through L2/HBM every site: 320 KB instead of 128 KB per plateau site of BASELINE configs[3]).  The cores
// stream through a 3-stage cp.async ring of K-chunks (16-byte LDGSTS along the contiguous dimension, rows padded to a stride
// = 4 (mod 16) doubles so that the DMMA fragment loads are bank-conflict free) that runs ahead across the two steps of a site
// and across sites; both steps run on the FP64 tensor pipe (DMMA.8x8x4), 16 warps (four per SM sub-partition: with eight the pipe was 49 % busy,
// 38 % of the stall samples on fixed-latency waits) as 2 x 8 (step 0, warp tile 32 x 16) and 4 x 4 (step 1, warp tile 16 x 16):
//   step 0:  F[a, (i, b')]  = sum_a'    E[a, a']  Bk[a', (i, b')]      M = la, N = d rb, K = lb      (A operand: E, shared memory)
//   step 1:  E'[b, b']      = sum_(a,i) Ak[(a,i), b] F[(a,i), b']      M = ra, N = rb,   K = la d    (B operand: F, shared memory)
// The grid is persistent: min(batch, #SM) CTAs stride over the states.  Per plateau site of configs[3] (chi = 64, d = 2):
// 2.1e6 flop against 128 KB of HBM reads -> FP64-tensor bound (8.4 us per site and SM at the 37.1 TFLOP/s DMMA peak).
#include <cstdlib>

#include "common.cuh"

namespace syn {

constexpr int OV_MAXB = SYN_OVERLAP_MAX_BOND;       // largest bond of either chain
constexpr int OV_MAXK = 2 * OV_MAXB;                // largest round8(la) * d and d * rb
constexpr int OV_ES = OV_MAXB + 4;                  // row stride of E and F in shared memory (= 4 mod 16)
constexpr int OV_STAGE_ELEMS = 64 * OV_ES;          // one K-chunk: 64 rows of a 64-wide core or 32 rows of a 128-wide one (half the CTA barriers of 32-row chunks)
constexpr int OV_STAGES = 3;
constexpr int OV_THREADS = 512;
constexpr int OV_SMEM_DOUBLES = OV_MAXB * OV_ES + OV_MAXK * OV_ES + OV_STAGES * OV_STAGE_ELEMS + 8;   // + slack: tiles read up to round8(len)

struct OverlapParams {
    syn_overlap_site_t site[SYN_OVERLAP_MAX_SITES];
    const double* E_in;
    double* E_out;
    int n_sites, batch;
};

__host__ __device__ __forceinline__ int ov_pad_stride(int len) {      // smallest s >= len with s = 4 (mod 16)
    return len <= 4 ? 4 : ((len - 4 + 15) / 16) * 16 + 4;
}
__host__ __device__ __forceinline__ int ov_chunk_rows(int stride) {   // rows of a K-chunk: a multiple of 4 that fits one stage
    return (OV_STAGE_ELEMS / stride) & ~3;
}

struct OvCursor {        // position in the flat (site, step, K-chunk) sequence of one state
    int site, step, row0;
};

// Full-tile inner loop: MT x NT DMMA tiles per warp over ROWS k-rows, everything a compile-time constant (fully unrolled,
// fragments double-buffered in registers like the GEMM kernel).  A fragment (i, kk): a0[i * A_MS + kk * A_KS]; B fragment
// (j, kk): b0[kk * B_KS + j * 8].
template <int MT, int NT, int ROWS, int A_MS, int A_KS, int B_KS>
__device__ __forceinline__ void ov_mma_full(double (&acc)[4][2][2], const double* __restrict__ a0, const double* __restrict__ b0) {
    double af[2][MT], bf[2][NT];
    auto load_frags = [&](int buf, int kk) {
#pragma unroll
        for (int i = 0; i < MT; ++i) af[buf][i] = a0[i * A_MS + kk * A_KS];
#pragma unroll
        for (int j = 0; j < NT; ++j) bf[buf][j] = b0[kk * B_KS + j * 8];
    };
    load_frags(0, 0);
#pragma unroll
    for (int kk = 0; kk < ROWS; kk += 4) {
        const int cur = (kk >> 2) & 1;
        if (kk + 4 < ROWS) load_frags(cur ^ 1, kk + 4);
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[cur][i], bf[cur][j]);
    }
}

// Edge-tile inner loop: runtime tile counts, strides and rows (first / last sites of a chain, odd bonds).
template <int NTMAX>
__device__ __forceinline__ void ov_mma_edge(double (&acc)[4][2][2], const double* a0, const double* b0, int mt, int nt, int rows, int a_ms,
                                            int a_ks, int b_ks) {
    double af[2][4], bf[2][NTMAX];
    auto load_frags = [&](int buf, int kk) {
#pragma unroll
        for (int i = 0; i < 4; ++i) af[buf][i] = (i < mt) ? a0[i * a_ms + kk * a_ks] : 0.0;
#pragma unroll
        for (int j = 0; j < NTMAX; ++j) bf[buf][j] = (j < nt) ? b0[kk * b_ks + j * 8] : 0.0;
    };
    auto mma = [&](int cur) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < NTMAX; ++j)
                if (i < mt && j < nt) dmma884(acc[i][j][0], acc[i][j][1], af[cur][i], bf[cur][j]);
    };
    load_frags(0, 0);
    for (int kk = 0; kk < rows; kk += 8) {      // two k-steps per trip: the fragment buffers stay compile-time registers
        const bool more = kk + 4 < rows;
        if (more) load_frags(1, kk + 4);
        mma(0);
        if (more) {
            if (kk + 8 < rows) load_frags(0, kk + 8);
            mma(1);
        }
    }
}

__global__ void __launch_bounds__(OV_THREADS, 1) overlap_chain_kernel(const __grid_constant__ OverlapParams p) {
    extern __shared__ __align__(16) double ov_smem[];
    double* sE = ov_smem;                               // [OV_MAXB][OV_ES]
    double* sF = sE + OV_MAXB * OV_ES;                  // [OV_MAXK][OV_ES]   F viewed as ((a, i), b')
    double* sS = sF + OV_MAXK * OV_ES;                  // stages
    __shared__ int4 s_geo[SYN_OVERLAP_MAX_SITES][2];    // per site and step: K, row length, padded row stride, rows per K-chunk

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wmi0 = warp >> 3, wni0 = warp & 7;        // step 0: 2 x 8 warps, tile 32 x 16
    const int wmi1 = warp >> 2, wni1 = warp & 3;        // step 1: 4 x 4 warps, tile 16 x 16

    for (int idx = tid; idx < 2 * p.n_sites; idx += OV_THREADS) {
        const syn_overlap_site_t& s = p.site[idx >> 1];
        int4 q;
        if ((idx & 1) == 0) { q.x = s.lb; q.y = s.d * s.rb; }
        else { q.x = s.la * s.d; q.y = s.ra; }
        q.z = ov_pad_stride(q.y);
        q.w = ov_chunk_rows(q.z);
        s_geo[idx >> 1][idx & 1] = q;
    }
    __syncthreads();

    auto advance = [&](OvCursor& c) {
        const int4 q = s_geo[c.site][c.step];
        c.row0 += q.w;
        if (c.row0 >= q.x) {
            c.row0 = 0;
            if (++c.step == 2) { c.step = 0; ++c.site; }
        }
    };
    // producer side: rows [row0, row0 + R) of the (K x len) contiguous core into a stage, rows padded to `stride`, rows past K zero
    auto issue = [&](const OvCursor& c, int state, double* stage) {
        const syn_overlap_site_t& s = p.site[c.site];
        const int4 q = s_geo[c.site][c.step];
        const int len = q.y, stride = q.z;
        const int64_t bstride = c.step == 0 ? s.b_stride : s.a_stride;
        const double* src = (c.step == 0 ? s.b : s.a) + (int64_t)state * bstride + (int64_t)c.row0 * len;
        const int valid = min(q.w, q.x - c.row0);
        const int rows = (valid + 3) & ~3;
        const bool vec2 = ((len & 1) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
        const int U = vec2 ? (len >> 1) : len;          // copies per row
        if (vec2 && valid == rows && ((U == 64 && rows == 32) || (U == 32 && rows == 64))) {
            // plateau chunks (full 64-row x 64-wide or 32-row x 128-wide K-chunks of aligned cores): exactly four 16-byte copies per thread
            // at constant pointer increments (the general loops below cost 21 instructions per copy).  Tried instead: a dedicated producer
            // warp issuing every copy, so that the DMMA warps run no copy code at all -- no faster (503 k vs 511 k states/s): the issue
            // phase was not what idles the tensor pipe.
            const int sh = U == 64 ? 6 : 5;
            const int col = (tid & (U - 1)) * 2, r0 = tid >> sh, rstep = OV_THREADS >> sh;
            const double* from = src + (int64_t)r0 * len + col;
            double* to = stage + r0 * stride + col;
            const int64_t fstep = (int64_t)rstep * len;
            const int tstep = rstep * stride;
#pragma unroll
            for (int k = 0; k < 4; ++k) cp_async<16>(to + k * tstep, from + k * fstep, true);
        } else if (U <= OV_THREADS && (U & (U - 1)) == 0) {    // power-of-two rows: a thread keeps its column and walks down the rows
            const int sh = 31 - __clz(U);
            const int col = (tid & (U - 1)) * (vec2 ? 2 : 1), step = OV_THREADS >> sh;
            for (int r = tid >> sh; r < rows; r += step) {
                const bool ok = r < valid;
                const double* from = ok ? src + (int64_t)r * len + col : src;
                if (vec2) cp_async<16>(stage + r * stride + col, from, ok);
                else cp_async<8>(stage + r * stride + col, from, ok);
            }
        } else {
            for (int e = tid; e < rows * U; e += OV_THREADS) {
                const int r = e / U, col = (e - r * U) * (vec2 ? 2 : 1);
                const bool ok = r < valid;
                const double* from = ok ? src + (int64_t)r * len + col : src;
                if (vec2) cp_async<16>(stage + r * stride + col, from, ok);
                else cp_async<8>(stage + r * stride + col, from, ok);
            }
        }
    };

    double acc[4][2][2];
    const int first = p.site[0].la * p.site[0].lb;
    const syn_overlap_site_t& last = p.site[p.n_sites - 1];

    for (int state = blockIdx.x; state < p.batch; state += gridDim.x) {
        // E_0: zero everywhere (pad rows / columns must be exact zeros), then the caller's matrix or the scalar 1
        for (int idx = tid; idx < OV_MAXB * OV_ES; idx += OV_THREADS) sE[idx] = 0.0;
        __syncthreads();
        if (p.E_in) {
            const int lb0 = p.site[0].lb;
            for (int idx = tid; idx < first; idx += OV_THREADS) sE[(idx / lb0) * OV_ES + idx % lb0] = p.E_in[(int64_t)state * first + idx];
        } else if (tid == 0) {
            sE[0] = 1.0;
        }
        // (visibility: the barrier at the top of the first chunk iteration)

        OvCursor prod = {0, 0, 0}, cons = {0, 0, 0};
#pragma unroll
        for (int s = 0; s < OV_STAGES - 1; ++s) {
            if (prod.site < p.n_sites) { issue(prod, state, sS + s * OV_STAGE_ELEMS); advance(prod); }
            cp_async_commit();
        }
        for (int chunk = 0; cons.site < p.n_sites; ++chunk) {
            cp_async_wait<OV_STAGES - 2>();
            __syncthreads();            // chunk `chunk` has landed for everybody; everybody is done with the stage refilled below
            if (prod.site < p.n_sites) { issue(prod, state, sS + ((chunk + OV_STAGES - 1) % OV_STAGES) * OV_STAGE_ELEMS); advance(prod); }
            cp_async_commit();

            const double* st = sS + (chunk % OV_STAGES) * OV_STAGE_ELEMS;
            const syn_overlap_site_t& s = p.site[cons.site];
            const int4 q = s_geo[cons.site][cons.step];
            const int K = q.x, len = q.y, stride = q.z, R = q.w;
            const int rows = (min(R, K - cons.row0) + 3) & ~3;
            const bool last_chunk = cons.row0 + R >= K;
            if (cons.row0 == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
            }
            if (cons.step == 0) {
                // F = E Bk : warp tile 32 x 16 of (round8(la) x round8(d rb))
                const int wm0 = wmi0 * 32, wn0 = wni0 * 16;
                const int Mp = (s.la + 7) & ~7, Np = (len + 7) & ~7;
                const int mt = min(4, max(0, (Mp - wm0) >> 3)), nt = min(2, max(0, (Np - wn0) >> 3));
                if (mt > 0 && nt > 0) {
                    const double* eA = sE + (wm0 + g) * OV_ES + cons.row0 + t;
                    const double* bB = st + t * stride + wn0 + g;
                    if (mt == 4 && nt == 2 && stride == 2 * OV_MAXB + 4 && rows == 32)
                        ov_mma_full<4, 2, 32, 8 * OV_ES, 1, 2 * OV_MAXB + 4>(acc, eA, bB);
                    else
                        ov_mma_edge<2>(acc, eA, bB, mt, nt, rows, 8 * OV_ES, 1, stride);
                    if (last_chunk) {       // F[a, n = (i, b')] -> sF[(a d + i)][b'];  columns n >= d rb do not exist
                        int off[2][2];      // the split of n does not depend on the row tile: 4 divisions per lane and site
#pragma unroll
                        for (int j = 0; j < 2; ++j)
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const int n = wn0 + j * 8 + 2 * t + h;
                                const int ii = n / s.rb;
                                off[j][h] = (j < nt && n < len) ? ii * OV_ES + (n - ii * s.rb) : -1;
                            }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            if (i < mt) {
                                double* row = sF + (wm0 + i * 8 + g) * s.d * OV_ES;
#pragma unroll
                                for (int j = 0; j < 2; ++j)
#pragma unroll
                                    for (int h = 0; h < 2; ++h)
                                        if (off[j][h] >= 0) row[off[j][h]] = acc[i][j][h];
                            }
                        }
                    }
                }
            } else {
                // E' = Ak^T F : warp tile 16 x 16 of (round8(ra) x round8(rb))
                const int wm0 = wmi1 * 16, wn0 = wni1 * 16;
                const int Mp = (s.ra + 7) & ~7, Np = (s.rb + 7) & ~7;
                const int mt = min(2, max(0, (Mp - wm0) >> 3)), nt = min(2, max(0, (Np - wn0) >> 3));
                if (mt > 0 && nt > 0) {
                    const double* aA = st + t * stride + wm0 + g;
                    const double* fB = sF + (cons.row0 + t) * OV_ES + wn0 + g;
                    if (mt == 2 && nt == 2 && stride == OV_ES && rows == 64)
                        ov_mma_full<2, 2, 64, 8, OV_ES, OV_ES>(acc, aA, fB);
                    else
                        ov_mma_edge<2>(acc, aA, fB, mt, nt, rows, 8, stride, OV_ES);
                    if (last_chunk) {       // nobody reads E during step 1; pad rows / columns are written as exact zeros
#pragma unroll
                        for (int i = 0; i < 2; ++i)
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                if (i < mt && j < nt) {
                                    const int b = wm0 + i * 8 + g, c0 = wn0 + j * 8 + 2 * t;
                                    const bool rok = b < s.ra;
                                    double2 v;
                                    v.x = (rok && c0 < s.rb) ? acc[i][j][0] : 0.0;
                                    v.y = (rok && c0 + 1 < s.rb) ? acc[i][j][1] : 0.0;
                                    *reinterpret_cast<double2*>(sE + b * OV_ES + c0) = v;
                                }
                            }
                    }
                }
            }
            advance(cons);
        }
        cp_async_wait<0>();
        __syncthreads();
        const int nout = last.ra * last.rb;
        for (int idx = tid; idx < nout; idx += OV_THREADS) p.E_out[(int64_t)state * nout + idx] = sE[(idx / last.rb) * OV_ES + idx % last.rb];
        __syncthreads();                // E is re-initialised for the next state
    }
}

static bool overlap_site_fits(const syn_overlap_site_t& s) {
    if (s.la < 1 || s.ra < 1 || s.lb < 1 || s.rb < 1 || s.d < 1) return false;
    if (s.la > OV_MAXB || s.ra > OV_MAXB || s.lb > OV_MAXB || s.rb > OV_MAXB) return false;
    if (((s.la + 7) & ~7) * s.d > OV_MAXK || s.d * s.rb > OV_MAXK) return false;
    return true;
}

}  // namespace syn

extern "C" int syn_overlap_batched_fits(const syn_overlap_site_t* sites, int n_sites) {
    if (!sites || n_sites < 1 || n_sites > SYN_OVERLAP_MAX_SITES) return 0;
    for (int k = 0; k < n_sites; ++k) {
        if (!syn::overlap_site_fits(sites[k])) return 0;
        if (k > 0 && (sites[k].la != sites[k - 1].ra || sites[k].lb != sites[k - 1].rb)) return 0;
    }
    return 1;
}

extern "C" int syn_overlap_batched_f64(const syn_overlap_site_t* sites, int n_sites, int batch, const double* E_in, double* E_out,
                                       void* stream) {
    using namespace syn;
    SYN_REQUIRE(sites && E_out, "syn_overlap_batched_f64: null argument");
    SYN_REQUIRE(n_sites >= 1 && n_sites <= SYN_OVERLAP_MAX_SITES, "syn_overlap_batched_f64: n_sites %d not in 1..%d (split the chain and pass E_in)",
                n_sites, SYN_OVERLAP_MAX_SITES);
    SYN_REQUIRE(batch >= 1, "syn_overlap_batched_f64: batch %d", batch);
    SYN_REQUIRE(syn_overlap_batched_fits(sites, n_sites),
                "syn_overlap_batched_f64: chain does not fit the fused kernel (bonds <= %d, round8(l) * d <= %d, d * r <= %d, matching bonds)",
                OV_MAXB, OV_MAXK, OV_MAXK);
    SYN_REQUIRE(E_in || (sites[0].la == 1 && sites[0].lb == 1), "syn_overlap_batched_f64: E_in is required when the first bonds are not 1");
    for (int k = 0; k < n_sites; ++k) SYN_REQUIRE(sites[k].a && sites[k].b, "syn_overlap_batched_f64: null core pointer at site %d", k);
    OverlapParams q;
    for (int k = 0; k < n_sites; ++k) q.site[k] = sites[k];
    for (int k = n_sites; k < SYN_OVERLAP_MAX_SITES; ++k) q.site[k] = sites[n_sites - 1];
    q.E_in = E_in;
    q.E_out = E_out;
    q.n_sites = n_sites;
    q.batch = batch;
    const size_t smem = (size_t)OV_SMEM_DOUBLES * sizeof(double);
    static PerDevice configured;
    const int dev_ = current_device();
    if (!configured.get(dev_)) {
        SYN_CUDA(cudaFuncSetAttribute(overlap_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.set(dev_);
    }
    const int grid = batch < sm_count() ? batch : sm_count();
    overlap_chain_kernel<<<grid, OV_THREADS, smem, (cudaStream_t)stream>>>(q);
    return launch_status("overlap_chain_kernel");
}

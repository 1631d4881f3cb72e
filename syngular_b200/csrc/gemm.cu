// Strided tensor-contraction GEMM on the FP64 tensor pipe (DMMA.8x8x4), sm_100a.
//
//   C[m,n] = alpha * sum_k A[m,k] * B[k,n] + beta * C[m,n]      (batched)
//
// Every logical index (m, n, k, batch) of every operand is a two-level strided index (syn_index_t), so the
// per-site contractions of the matrix-product hot path -- which in the reference are opt_einsum/np.tensordot
// calls on reshaped, permuted cores (matrix_product_operator.py:184,280; tensor/utils.py:35,49,62) and the
// `S @ R`, `B @ R`, `L @ U.T` products of the sweeps (matrix_product_state.py:448,561,575;
// matrix_product_operator.py:560,682,705) -- run without any transposition copy.
//
// Design (B200): Blackwell's tcgen05 has no FP64 kind; FP64 matrix math runs on DMMA (measured 37.1 TFLOP/s on
// this part, tools/microbench).  One DMMA.8x8x4 occupies an SM sub-partition for 16 cycles, so the pipe is fed
// comfortably from shared memory: operands are staged by a 4-stage cp.async (LDGSTS) ring, 16-byte chunks along
// whichever logical dimension is contiguous in HBM; tiles are stored in shared memory in the orientation they
// arrive in, padded so that the 64-bit fragment loads of a half-warp hit 16 distinct 8-byte bank pairs.
#include <cstdlib>

#include "common.cuh"

namespace syn {

template <int BM_, int BN_, int BK_, int WM_, int WN_, int STAGES_>
struct GemmCfg {
    static constexpr int BM = BM_, BN = BN_, BK = BK_, WM = WM_, WN = WN_, STAGES = STAGES_;
    static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
    static constexpr int THREADS = WARPS_M * WARPS_N * 32;
    static constexpr int MT = WM / 8, NT = WN / 8;
    static constexpr int PAD = 4;   // strides = 4 (mod 16) doubles -> conflict-free DMMA fragment loads
};

// Loader for one operand tile of ROWS x BK logical elements (ROWS = BM for A, BN for B).
//   ROW_MAJOR_IN_HBM == false : contiguous along k  -> smem [row][k]   (stride BK+PAD)
//   ROW_MAJOR_IN_HBM == true  : contiguous along row-> smem [k][row]   (stride ROWS+PAD)
template <class C, int ROWS, bool ALONG_ROWS, int VEC>
struct OperandLoader {
    static constexpr int T = C::THREADS;
    static constexpr int STRIDE = ALONG_ROWS ? (ROWS + C::PAD) : (C::BK + C::PAD);
    static constexpr int TILE_ELEMS = ALONG_ROWS ? C::BK * STRIDE : ROWS * STRIDE;
    // chunks along the contiguous dimension
    static constexpr int CH = (ALONG_ROWS ? ROWS : C::BK) / VEC;       // chunks per line
    static constexpr int LINES = ALONG_ROWS ? C::BK : ROWS;            // number of lines
    static constexpr int LINES_PER_PASS = T / CH;
    static constexpr int PASSES = LINES_PER_PASS > 0 ? LINES / (LINES_PER_PASS > 0 ? LINES_PER_PASS : 1) : 1;
    static constexpr bool USABLE = (T % CH == 0) && LINES_PER_PASS > 0 && (LINES % (LINES_PER_PASS > 0 ? LINES_PER_PASS : 1) == 0);

    int pos;            // position along the contiguous dimension (element index within tile)
    int line0;          // first line handled by this thread
    // ALONG_ROWS: one fixed row chunk, PASSES k-lines per tile.  else: one fixed k chunk, PASSES rows.
    int64_t fixed_off;  // ALONG_ROWS: offset of the row chunk
    bool fixed_ok;
    int64_t row_off[ALONG_ROWS ? 1 : PASSES];
    unsigned row_ok;

    __device__ __forceinline__ void init(const syn_index_t& rows_ix, int row0, int nrows, int tid) {
        static_assert(USABLE, "cp.async loader: thread count must be a multiple of the chunks per line, lines must divide evenly");
        pos = (tid % CH) * VEC;
        line0 = tid / CH;
        if constexpr (ALONG_ROWS) {
            int r = row0 + pos;
            fixed_ok = r < nrows;
            fixed_off = fixed_ok ? idx2(rows_ix, r) : 0;
            row_ok = 0;
        } else {
            row_ok = 0;
#pragma unroll
            for (int i = 0; i < PASSES; i++) {
                int r = row0 + line0 + i * LINES_PER_PASS;
                bool ok = r < nrows;
                row_off[i] = ok ? idx2(rows_ix, r) : 0;
                row_ok |= (ok ? 1u : 0u) << i;
            }
            fixed_ok = true;
            fixed_off = 0;
        }
    }

    __device__ __forceinline__ void load(double* smem, const double* base, const syn_index_t& k_ix, int k0, int K) const {
        if constexpr (ALONG_ROWS) {
#pragma unroll
            for (int i = 0; i < PASSES; i++) {
                int kl = line0 + i * LINES_PER_PASS;
                int k = k0 + kl;
                bool ok = fixed_ok && (k < K);
                const double* src = base + (ok ? (fixed_off + idx2(k_ix, k)) : 0);
                cp_async<VEC * 8>(smem + kl * STRIDE + pos, src, ok);
            }
        } else {
            int k = k0 + pos;
            bool kok = k < K;
            int64_t koff = kok ? idx2(k_ix, k) : 0;
#pragma unroll
            for (int i = 0; i < PASSES; i++) {
                int rl = line0 + i * LINES_PER_PASS;
                bool ok = kok && ((row_ok >> i) & 1u);
                const double* src = base + (ok ? (row_off[i] + koff) : 0);
                cp_async<VEC * 8>(smem + rl * STRIDE + pos, src, ok);
            }
        }
    }
};

// TMA staging: one cp.async.bulk (UBLKCP) per contiguous tile line, completion on the stage's mbarrier.  Eligible when the
// tile is interior (M, N, K multiples of the tile) and every line is one contiguous, 16-byte aligned run in HBM.
template <class C, int ROWS, bool ALONG_ROWS>
struct BulkLoader {
    static constexpr int STRIDE = ALONG_ROWS ? (ROWS + C::PAD) : (C::BK + C::PAD);
    static constexpr int LINES = ALONG_ROWS ? C::BK : ROWS;
    static constexpr uint32_t LINE_BYTES = (uint32_t)((ALONG_ROWS ? ROWS : C::BK) * sizeof(double));
    // source of line `l` of the tile whose first row is row0 and first k is k0
    __device__ __forceinline__ static const double* src(const double* base, const syn_index_t& rows_ix, const syn_index_t& k_ix, int row0,
                                                        int k0, int l) {
        return ALONG_ROWS ? base + idx2(rows_ix, row0) + idx2(k_ix, k0 + l) : base + idx2(rows_ix, row0 + l) + idx2(k_ix, k0);
    }
};

// One k-stage (BK deep) of a warp's WM x WN sub-tile from shared memory: fragments are double-buffered in registers -- the loads of
// k-step kk+4 are issued before the DMMAs of step kk, into registers the in-flight DMMAs do not read (no WAR/RAW serialisation between
// the LDS and the tensor pipe).
template <class C, bool A_ALONG_M, bool B_ALONG_N, int SA, int SB>
__device__ __forceinline__ void gemm_stage(const double* __restrict__ a_s, const double* __restrict__ b_s, int wm0, int wn0, int g, int t,
                                           double (&acc)[C::MT][C::NT][2]) {
    double af[2][C::MT], bf[2][C::NT];
    auto load_frags = [&](int buf, int kk) {
#pragma unroll
        for (int i = 0; i < C::MT; i++) {
            int r = wm0 + i * 8 + g;
            af[buf][i] = A_ALONG_M ? a_s[(kk + t) * SA + r] : a_s[r * SA + kk + t];
        }
#pragma unroll
        for (int j = 0; j < C::NT; j++) {
            int c = wn0 + j * 8 + g;
            bf[buf][j] = B_ALONG_N ? b_s[(kk + t) * SB + c] : b_s[c * SB + kk + t];
        }
    };
    load_frags(0, 0);
#pragma unroll
    for (int kk = 0; kk < C::BK; kk += 4) {
        const int cur = (kk >> 2) & 1;
        if (kk + 4 < C::BK) load_frags(cur ^ 1, kk + 4);
#pragma unroll
        for (int i = 0; i < C::MT; i++)
#pragma unroll
            for (int j = 0; j < C::NT; j++) dmma884(acc[i][j][0], acc[i][j][1], af[cur][i], bf[cur][j]);
    }
}

// Epilogue of one warp sub-tile whose first row / column is (mw, nw): the thread owns C[row = g][col = 2t, 2t+1] of every 8x8 block.
template <class C>
__device__ __forceinline__ void gemm_store_tile(const syn_gemm_desc_t& d, double* __restrict__ Cmat, const double (&acc)[C::MT][C::NT][2], int mw,
                                                int nw, int g, int t, int c_vec) {
    const double alpha = d.alpha, beta = d.beta;
#pragma unroll
    for (int i = 0; i < C::MT; i++) {
        int m = mw + i * 8 + g;
        if (m >= d.M) continue;
        int64_t mo = idx2(d.c_m, m);
#pragma unroll
        for (int j = 0; j < C::NT; j++) {
            int n = nw + j * 8 + 2 * t;
            if (n >= d.N) continue;
            double v0 = alpha * acc[i][j][0], v1 = alpha * acc[i][j][1];
            int64_t o0 = mo + idx2(d.c_n, n);
            if (c_vec && (n + 1 < d.N)) {
                double2* p = reinterpret_cast<double2*>(Cmat + o0);
                if (beta != 0.0) {
                    double2 old = *p;
                    v0 += beta * old.x;
                    v1 += beta * old.y;
                }
                *p = make_double2(v0, v1);
            } else {
                if (beta != 0.0) v0 += beta * Cmat[o0];
                Cmat[o0] = v0;
                if (n + 1 < d.N) {
                    int64_t o1 = mo + idx2(d.c_n, n + 1);
                    if (beta != 0.0) v1 += beta * Cmat[o1];
                    Cmat[o1] = v1;
                }
            }
        }
    }
}

template <class C, bool A_ALONG_M, bool B_ALONG_N, int VEC, bool TMA>
__global__ void __launch_bounds__(C::THREADS)
gemm_f64_kernel(const syn_gemm_desc_t d, const double* __restrict__ A, const double* __restrict__ B,
                double* __restrict__ Cmat, int tiles_m, int tiles_n, int c_vec) {
    using LA = OperandLoader<C, C::BM, A_ALONG_M, VEC>;
    using LB = OperandLoader<C, C::BN, B_ALONG_N, VEC>;
    extern __shared__ __align__(16) double smem[];
    double* sA = smem;
    double* sB = smem + C::STAGES * LA::TILE_ELEMS;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp / C::WARPS_N) * C::WM;
    const int wn0 = (warp % C::WARPS_N) * C::WN;

    // tile order: the dimension with FEWER tiles varies fastest, so the CTAs that share a tile of the long operand are
    // co-scheduled and it is fetched from HBM once (ncu: DRAM traffic == algorithmic bytes)
    const int tile = blockIdx.x;
    int tm, tn;
    if (tiles_m <= tiles_n) { tm = tile % tiles_m; tn = tile / tiles_m; }
    else { tn = tile % tiles_n; tm = tile / tiles_n; }
    const int m0 = tm * C::BM;
    const int n0 = tn * C::BN;
    const int batch = blockIdx.y + gridDim.y * blockIdx.z;
    if (batch >= d.batch) return;
    // block-lower mask: a tile that lies entirely above the kept staircase has nothing to do (the whole CTA leaves before any barrier)
    if (d.mask_rows > 0 && n0 >= (m0 / d.mask_rows + 1) * d.mask_cols) return;

    A += idx2(d.a_b, batch);
    B += idx2(d.b_b, batch);
    Cmat += idx2(d.c_b, batch);

    using TA = BulkLoader<C, C::BM, A_ALONG_M>;
    using TB = BulkLoader<C, C::BN, B_ALONG_N>;
    __shared__ __align__(8) uint64_t full_bar[C::STAGES];
    LA la;
    LB lb;
    if constexpr (!TMA) {
        la.init(d.a_m, m0, d.M, tid);
        lb.init(d.b_n, n0, d.N, tid);
    } else {
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < C::STAGES; s++) mbar_init(&full_bar[s], TA::LINES + TB::LINES);
        }
        __syncthreads();
    }
    // TMA producer side: every tile line is owned by one thread, which arms the stage barrier with its bytes and issues the copy.
    // A B tile whose lines run along n may hang over the right edge of the matrix (the 128 x 112 configuration: N need not be a
    // multiple of 112): its lines are copied short; the stale tail of the stage only feeds output columns the epilogue masks.
    uint32_t b_line_bytes = TB::LINE_BYTES;
    if constexpr (TMA && B_ALONG_N) {
        const int left = d.N - n0;
        if (left < C::BN) b_line_bytes = (uint32_t)left * (uint32_t)sizeof(double);
    }
    auto tma_fill = [&](int stage, int k0) {
        for (int line = tid; line < TA::LINES + TB::LINES; line += C::THREADS) {
            if (line < TA::LINES) {
                mbar_expect_tx(&full_bar[stage], TA::LINE_BYTES);
                bulk_g2s(sA + stage * LA::TILE_ELEMS + line * TA::STRIDE, TA::src(A, d.a_m, d.a_k, m0, k0, line), TA::LINE_BYTES, &full_bar[stage]);
            } else {
                const int l = line - TA::LINES;
                mbar_expect_tx(&full_bar[stage], b_line_bytes);
                bulk_g2s(sB + stage * LB::TILE_ELEMS + l * TB::STRIDE, TB::src(B, d.b_n, d.b_k, n0, k0, l), b_line_bytes, &full_bar[stage]);
            }
        }
    };

    const int KT = (d.K + C::BK - 1) / C::BK;

    double acc[C::MT][C::NT][2];
#pragma unroll
    for (int i = 0; i < C::MT; i++)
#pragma unroll
        for (int j = 0; j < C::NT; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < C::STAGES - 1; s++) {
        if (s < KT) {
            if constexpr (TMA) {
                tma_fill(s, s * C::BK);
            } else {
                la.load(sA + s * LA::TILE_ELEMS, A, d.a_k, s * C::BK, d.K);
                lb.load(sB + s * LB::TILE_ELEMS, B, d.b_k, s * C::BK, d.K);
            }
        }
        if constexpr (!TMA) cp_async_commit();
    }

    for (int kt = 0; kt < KT; kt++) {
        if constexpr (TMA) {
            mbar_wait(&full_bar[kt % C::STAGES], (uint32_t)((kt / C::STAGES) & 1));
        } else {
            cp_async_wait<C::STAGES - 2>();
        }
        __syncthreads();
        {
            int nk = kt + C::STAGES - 1;
            if (nk < KT) {
                int s = nk % C::STAGES;
                if constexpr (TMA) {
                    tma_fill(s, nk * C::BK);
                } else {
                    la.load(sA + s * LA::TILE_ELEMS, A, d.a_k, nk * C::BK, d.K);
                    lb.load(sB + s * LB::TILE_ELEMS, B, d.b_k, nk * C::BK, d.K);
                }
            }
            if constexpr (!TMA) cp_async_commit();
        }
        gemm_stage<C, A_ALONG_M, B_ALONG_N, LA::STRIDE, LB::STRIDE>(sA + (kt % C::STAGES) * LA::TILE_ELEMS, sB + (kt % C::STAGES) * LB::TILE_ELEMS,
                                                                       wm0, wn0, g, t, acc);
    }
    if constexpr (!TMA) cp_async_wait<0>();

    gemm_store_tile<C>(d, Cmat, acc, m0 + wm0, n0 + wn0, g, t, c_vec);
}

// Persistent TMA variant for problems of many tiles with a short contraction (the X.E product of the environment update: 2048 tiles of
// 8 k-stages): one CTA per SM walks its tiles and the operand ring runs ON ACROSS TILE BOUNDARIES -- the first stages of the next tile
// are in flight while the last stages of this one are multiplied and its accumulators are stored, so the per-tile pipeline fill and
// the epilogue's exposed latency (15 % of a K = 256 tile) disappear.  Tiles are numbered batch-major, then with the shorter tile
// dimension fastest, and CTA c takes tiles c, c + grid, ...: CTAs that run together share operand tiles through L2 as before.
template <class C, bool A_ALONG_M, bool B_ALONG_N>
__global__ void __launch_bounds__(C::THREADS)
gemm_f64_persistent_kernel(const syn_gemm_desc_t d, const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ Cmat,
                           int tiles_m, int tiles_n, int c_vec) {
    using LA = OperandLoader<C, C::BM, A_ALONG_M, 2>;
    using LB = OperandLoader<C, C::BN, B_ALONG_N, 2>;
    using TA = BulkLoader<C, C::BM, A_ALONG_M>;
    using TB = BulkLoader<C, C::BN, B_ALONG_N>;
    extern __shared__ __align__(16) double smem[];
    double* sA = smem;
    double* sB = smem + C::STAGES * LA::TILE_ELEMS;
    __shared__ __align__(8) uint64_t full_bar[C::STAGES];

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp / C::WARPS_N) * C::WM;
    const int wn0 = (warp % C::WARPS_N) * C::WN;
    const int per_batch = tiles_m * tiles_n;
    const long long total = (long long)per_batch * d.batch;
    const int KT = d.K / C::BK;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < C::STAGES; s++) mbar_init(&full_bar[s], TA::LINES + TB::LINES);
    }
    __syncthreads();

    auto coords = [&](long long tile, int& batch, int& m0, int& n0) {
        batch = (int)(tile / per_batch);
        const int local = (int)(tile - (long long)batch * per_batch);
        int tm, tn;
        if (tiles_m <= tiles_n) { tm = local % tiles_m; tn = local / tiles_m; }
        else { tn = local % tiles_n; tm = local / tiles_n; }
        m0 = tm * C::BM;
        n0 = tn * C::BN;
    };
    // producer cursor: (tile, k-stage) of the next fill and its running count (stage = count % STAGES)
    long long ptile = blockIdx.x;
    int pkt = 0;
    unsigned pcount = 0;
    auto fill_next = [&]() {
        if (ptile >= total) return;
        int batch, m0, n0;
        coords(ptile, batch, m0, n0);
        const double* Ab = A + idx2(d.a_b, batch);
        const double* Bb = B + idx2(d.b_b, batch);
        const int stage = (int)(pcount % C::STAGES), k0 = pkt * C::BK;
        uint32_t b_line_bytes = TB::LINE_BYTES;
        if constexpr (B_ALONG_N) {
            const int left = d.N - n0;
            if (left < C::BN) b_line_bytes = (uint32_t)left * (uint32_t)sizeof(double);
        }
        for (int line = tid; line < TA::LINES + TB::LINES; line += C::THREADS) {
            if (line < TA::LINES) {
                mbar_expect_tx(&full_bar[stage], TA::LINE_BYTES);
                bulk_g2s(sA + stage * LA::TILE_ELEMS + line * TA::STRIDE, TA::src(Ab, d.a_m, d.a_k, m0, k0, line), TA::LINE_BYTES, &full_bar[stage]);
            } else {
                const int l = line - TA::LINES;
                mbar_expect_tx(&full_bar[stage], b_line_bytes);
                bulk_g2s(sB + stage * LB::TILE_ELEMS + l * TB::STRIDE, TB::src(Bb, d.b_n, d.b_k, n0, k0, l), b_line_bytes, &full_bar[stage]);
            }
        }
        ++pcount;
        if (++pkt == KT) { pkt = 0; ptile += gridDim.x; }
    };
#pragma unroll
    for (int s = 0; s < C::STAGES - 1; s++) fill_next();

    unsigned ccount = 0;
    for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
        double acc[C::MT][C::NT][2];
#pragma unroll
        for (int i = 0; i < C::MT; i++)
#pragma unroll
            for (int j = 0; j < C::NT; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
        for (int kt = 0; kt < KT; kt++, ccount++) {
            const int stage = (int)(ccount % C::STAGES);
            mbar_wait(&full_bar[stage], (ccount / C::STAGES) & 1u);
            __syncthreads();          // every warp is done with the stage consumed one iteration ago: it is refilled now
            fill_next();
            gemm_stage<C, A_ALONG_M, B_ALONG_N, LA::STRIDE, LB::STRIDE>(sA + stage * LA::TILE_ELEMS, sB + stage * LB::TILE_ELEMS, wm0, wn0, g, t, acc);
        }
        int batch, m0, n0;
        coords(tile, batch, m0, n0);
        gemm_store_tile<C>(d, Cmat + idx2(d.c_b, batch), acc, m0 + wm0, n0 + wn0, g, t, c_vec);
    }
}

// ---------------------------------------------------------------------------------------------------------
using CfgL = GemmCfg<128, 128, 32, 64, 32, 3>;   // 256 threads, 1 CTA/SM, 212 KB smem
using CfgS = GemmCfg<64, 64, 32, 32, 32, 3>;     // 128 threads, 107 KB smem, 2 CTAs/SM
using CfgW = GemmCfg<128, 128, 32, 32, 32, 3>;   // 512 threads (4 warps per SM sub-partition), 32x32 warp tiles

// 128 x 112 tiles, 16 warps of 16 x 56: an output of 512 x 4096 (the M.E product of the density-matrix sweep) is 4 x 37 = 148 tiles -- one
// full wave of the 148 SMs -- where 128 x 128 tiles give 128 (13 % of the machine idle).  TMA-staged only (no cp.async loader exists
// for 112-wide lines), A k-contiguous, B n-contiguous.
using CfgT = GemmCfg<128, 112, 32, 16, 56, 3>;

// 128 x 64 tiles, 16 warps of 32 x 16: outputs with too few 128 x 128 tiles for the 148 SMs (the carry product U^T M: 256 x 4096 = 64 tiles)
// and the block-lower environment product, whose 64-column staircase wastes a quarter of a 128 x 128 tiling.
using CfgH = GemmCfg<128, 64, 32, 32, 16, 3>;

using CfgN = GemmCfg<128, 32, 32, 32, 32, 2>;    // skinny outputs (N <= 32): 128 threads, 92 KB smem
using CfgM = GemmCfg<32, 128, 32, 32, 32, 2>;    // flat outputs (M <= 32)

static int gemm_env_cfg() {   // experiment knob: SYN_GEMM_CFG=L|W|S forces a tile configuration
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SYN_GEMM_CFG");
        v = !e ? 0 : (e[0] == 'L' ? 1 : (e[0] == 'W' ? 2 : (e[0] == 'S' ? 3 : 0)));
    }
    return v;
}

template <class C, bool AM, bool BN, int VEC, bool TMA = false>
static int launch_gemm(const syn_gemm_desc_t& d, const double* A, const double* B, double* Cm, int c_vec, cudaStream_t st) {
    using LA = OperandLoader<C, C::BM, AM, VEC>;
    using LB = OperandLoader<C, C::BN, BN, VEC>;
    constexpr size_t smem = (size_t)C::STAGES * (LA::TILE_ELEMS + LB::TILE_ELEMS) * sizeof(double);
    auto kern = gemm_f64_kernel<C, AM, BN, VEC, TMA>;
    static PerDevice configured;
    const int dev_ = current_device();
    if (!configured.get(dev_)) {
        SYN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.set(dev_);
    }
    int tiles_m = (d.M + C::BM - 1) / C::BM, tiles_n = (d.N + C::BN - 1) / C::BN;
    long long tiles = (long long)tiles_m * tiles_n;
    SYN_REQUIRE(tiles < (1ll << 31), "syn_gemm_f64: too many tiles");
    int by = d.batch < 65535 ? d.batch : 65535;
    int bz = (d.batch + by - 1) / by;
    dim3 grid((unsigned)tiles, by, bz);
    kern<<<grid, C::THREADS, smem, st>>>(d, A, B, Cm, tiles_m, tiles_n, c_vec);
    return launch_status("gemm_f64_kernel");
}

template <class C, bool AM, bool BN>
static int launch_gemm_persistent(const syn_gemm_desc_t& d, const double* A, const double* B, double* Cm, int c_vec, cudaStream_t st) {
    using LA = OperandLoader<C, C::BM, AM, 2>;
    using LB = OperandLoader<C, C::BN, BN, 2>;
    constexpr size_t smem = (size_t)C::STAGES * (LA::TILE_ELEMS + LB::TILE_ELEMS) * sizeof(double);
    auto kern = gemm_f64_persistent_kernel<C, AM, BN>;
    static PerDevice configured;
    const int dev_ = current_device();
    if (!configured.get(dev_)) {
        SYN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.set(dev_);
    }
    const int tiles_m = (d.M + C::BM - 1) / C::BM, tiles_n = (d.N + C::BN - 1) / C::BN;
    const long long total = (long long)tiles_m * tiles_n * d.batch;
    const int grid = (int)(total < sm_count() ? total : sm_count());
    kern<<<grid, C::THREADS, smem, st>>>(d, A, B, Cm, tiles_m, tiles_n, c_vec);
    return launch_status("gemm_f64_persistent_kernel");
}

static bool gemm_env_persistent() {   // SYN_GEMM_PERSISTENT=0 disables the persistent variant (A/B comparisons)
    static int v = -1;
    if (v < 0) { const char* e = getenv("SYN_GEMM_PERSISTENT"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

static bool gemm_env_tma() {   // SYN_GEMM_TMA=0 disables the TMA-staged variant (A/B comparisons)
    static int v = -1;
    if (v < 0) { const char* e = getenv("SYN_GEMM_TMA"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

// can every line of a (rows x BK) tile be fetched by one bulk copy?  (`along`: is the line along this index?)
static bool tma_line_ok(const syn_index_t& along, int ext, int tile) {
    return along.inner == 1 && ext % tile == 0 && (along.div >= ext || along.div % tile == 0);
}

template <class C>
static bool tma_eligible(const syn_gemm_desc_t& d, bool am, bool bn, int vec) {
    if (!gemm_env_tma() || vec != 2) return false;                       // vec == 2 already implies 16-byte aligned, even offsets
    if (C::BM * C::BN < 128 * 128) return false;                         // measured: small tiles (256-512 B lines) are faster with LDGSTS
    if (d.M % C::BM || d.N % C::BN || d.K % C::BK) return false;
    const bool a_ok = am ? tma_line_ok(d.a_m, d.M, C::BM) : tma_line_ok(d.a_k, d.K, C::BK);
    const bool b_ok = bn ? tma_line_ok(d.b_n, d.N, C::BN) : tma_line_ok(d.b_k, d.K, C::BK);
    return a_ok && b_ok;
}

// time of the 128 x BN tiling in units of "columns per SM": waves * BN
static long long wave_cost(const syn_gemm_desc_t& d, int bn) {
    const long long tiles = (long long)((d.M + 127) / 128) * ((d.N + bn - 1) / bn) * d.batch;
    const long long sms = sm_count();
    return ((tiles + sms - 1) / sms) * bn;
}

// the 128 x 112 TMA configuration: eligible layouts only, and only where it saves a wave
static bool prefer_cfg_t(const syn_gemm_desc_t& d, bool am, bool bn, int vec) {
    if (!gemm_env_tma() || vec != 2 || am || !bn) return false;
    if (d.M % 128 || d.K % 32 || (d.N & 1) || d.N < 112) return false;
    if (!tma_line_ok(d.a_k, d.K, 32)) return false;
    // B lines run along n: every tile start n0 = j * 112 must stay inside one contiguous, 16-byte aligned run
    if (!(d.b_n.inner == 1 && d.b_n.div >= d.N)) return false;
    return wave_cost(d, 112) * 100 < wave_cost(d, 128) * 95;
}

template <class C>
static int dispatch_layout(const syn_gemm_desc_t& d, const double* A, const double* B, double* Cm, bool am, bool bn, int vec,
                           int c_vec, cudaStream_t st) {
    const bool tma = tma_eligible<C>(d, am, bn, vec);
    // several tiles per SM and a contraction short enough for the per-tile fill / drain to show: the persistent ring
    const long long tiles = (long long)((d.M + C::BM - 1) / C::BM) * ((d.N + C::BN - 1) / C::BN) * d.batch;
    const bool persistent = tma && d.mask_rows == 0 && gemm_env_persistent() && C::BM * C::BN >= 128 * 128 && tiles >= 2ll * sm_count() && d.K <= 1024;
#define SYN_GEMM_CASE(AM, BN)                                                                  \
    if (am == AM && bn == BN) {                                                                \
        if (persistent) return launch_gemm_persistent<C, AM, BN>(d, A, B, Cm, c_vec, st);      \
        if (tma) return launch_gemm<C, AM, BN, 2, true>(d, A, B, Cm, c_vec, st);               \
        return vec == 2 ? launch_gemm<C, AM, BN, 2>(d, A, B, Cm, c_vec, st)                    \
                        : launch_gemm<C, AM, BN, 1>(d, A, B, Cm, c_vec, st);                   \
    }
    SYN_GEMM_CASE(false, false)
    SYN_GEMM_CASE(false, true)
    SYN_GEMM_CASE(true, false)
    SYN_GEMM_CASE(true, true)
#undef SYN_GEMM_CASE
    return 2;
}

static inline bool even64(int64_t v) { return (v & 1) == 0; }

// can a 16-byte chunk run along logical index `ix` (extent `ext`)?
static bool vec_ok_along(const syn_index_t& ix, int ext) {
    if (ix.inner != 1) return false;
    if (ix.div >= ext) return even64(ext);               // single level: tiles start at even offsets, extent even
    return even64(ix.div) && even64(ix.outer) && even64(ext);
}
static bool other_even(const syn_index_t& ix, int ext) {
    // every offset produced by this index must be even (16-byte alignment of chunk starts)
    if (ext <= 1) return true;
    if (ix.div >= ext) return even64(ix.inner);
    return even64(ix.inner) && even64(ix.outer);
}

int gemm_f64(const syn_gemm_desc_t& d, const double* A, const double* B, double* C, cudaStream_t st) {
    SYN_REQUIRE(d.M >= 0 && d.N >= 0 && d.K >= 0 && d.batch >= 0, "syn_gemm_f64: negative extent");
    if (d.M == 0 || d.N == 0 || d.batch == 0) return 0;
    const syn_index_t* all[] = {&d.a_m, &d.a_k, &d.a_b, &d.b_k, &d.b_n, &d.b_b, &d.c_m, &d.c_n, &d.c_b};
    for (auto* ix : all) SYN_REQUIRE(ix->div >= 1, "syn_gemm_f64: index div must be >= 1");
    SYN_REQUIRE(A && B && C, "syn_gemm_f64: null operand");
    SYN_REQUIRE((d.mask_rows == 0 && d.mask_cols == 0) || (d.mask_rows > 0 && d.mask_cols > 0 && d.mask_rows % 64 == 0 && d.mask_cols % 64 == 0),
                "syn_gemm_f64: mask_rows / mask_cols must both be 0 or both positive multiples of 64");

    // orientation: which logical dimension is contiguous in HBM (neither -> scalar loads, k-major tile)
    bool a_along_m = (d.a_m.inner == 1) && (d.a_k.inner != 1 || d.K == 1);
    bool b_along_n = (d.b_n.inner == 1) && (d.b_k.inner != 1 || d.K == 1);

    bool a_vec = a_along_m ? (vec_ok_along(d.a_m, d.M) && other_even(d.a_k, d.K))
                           : (vec_ok_along(d.a_k, d.K) && other_even(d.a_m, d.M));
    bool b_vec = b_along_n ? (vec_ok_along(d.b_n, d.N) && other_even(d.b_k, d.K))
                           : (vec_ok_along(d.b_k, d.K) && other_even(d.b_n, d.N));
    a_vec = a_vec && other_even(d.a_b, d.batch) && (((uintptr_t)A & 15) == 0);
    b_vec = b_vec && other_even(d.b_b, d.batch) && (((uintptr_t)B & 15) == 0);
    int vec = (a_vec && b_vec) ? 2 : 1;
    int c_vec = (vec_ok_along(d.c_n, d.N) && other_even(d.c_m, d.M) && other_even(d.c_b, d.batch) &&
                 (((uintptr_t)C & 15) == 0)) ? 1 : 0;

    // tile choice: the large tile when it still fills the machine, else the small one
    long long big_tiles = (long long)((d.M + 127) / 128) * ((d.N + 127) / 128) * d.batch;
    bool use_large = (d.M > 64 && d.N > 64) && big_tiles * 5 >= (long long)sm_count() * 4;   // >= 80% of the SMs get a 128x128 tile
    const int forced = gemm_env_cfg();
    if (forced == 1) return dispatch_layout<CfgL>(d, A, B, C, a_along_m, b_along_n, vec, c_vec, st);
    if (forced == 2) return dispatch_layout<CfgW>(d, A, B, C, a_along_m, b_along_n, vec, c_vec, st);
    if (forced == 3) return dispatch_layout<CfgS>(d, A, B, C, a_along_m, b_along_n, vec, c_vec, st);
    if (d.N <= 32 && d.M >= 128) return dispatch_layout<CfgN>(d, A, B, C, a_along_m, b_along_n, vec, c_vec, st);
    if (d.M <= 32 && d.N >= 128) return dispatch_layout<CfgM>(d, A, B, C, a_along_m, b_along_n, vec, c_vec, st);
    if (d.mask_rows > 0) {
        // masked launch: tiles must not straddle the staircase -- 128 x 64 tiles when the steps are 64 columns wide, 64 x 64 for odd row blocks
        if (d.mask_rows % 128 == 0 && d.M >= 128) return dispatch_layout<CfgH>(d, A, B, C, a_along_m, b_along_n, vec, c_vec, st);
        return dispatch_layout<CfgS>(d, A, B, C, a_along_m, b_along_n, vec, c_vec, st);
    }
    if (!forced && !use_large && d.M >= 128 && d.N >= 128) {
        // too few 128 x 128 tiles: halve the tile width when that fills the machine
        const long long half_tiles = (long long)((d.M + 127) / 128) * ((d.N + 63) / 64) * d.batch;
        if (half_tiles * 5 >= (long long)sm_count() * 4) return dispatch_layout<CfgH>(d, A, B, C, a_along_m, b_along_n, vec, c_vec, st);
    }
    if (use_large && !forced && prefer_cfg_t(d, a_along_m, b_along_n, vec)) return launch_gemm<CfgT, false, true, 2, true>(d, A, B, C, c_vec, st);
    if (use_large) return dispatch_layout<CfgW>(d, A, B, C, a_along_m, b_along_n, vec, c_vec, st);
    return dispatch_layout<CfgS>(d, A, B, C, a_along_m, b_along_n, vec, c_vec, st);
}

}  // namespace syn

extern "C" int syn_gemm_f64(const syn_gemm_desc_t* desc, const double* A, const double* B, double* C, void* stream) {
    if (!desc) {
        syn::set_error("syn_gemm_f64: null descriptor");
        return 2;
    }
    return syn::gemm_f64(*desc, A, B, C, (cudaStream_t)stream);
}

// One-sided (Hestenes) Jacobi on the ROWS of a square FP64 matrix, batched, sm_100a.
//
//   G (n x n, row-major)  ->  J G  with mutually orthogonal rows:   row_i = sigma_i * u_i^T
//
// This is the SVD kernel of the bond-rounding sweep (north_star: "truncated SVD of each bond matrix ... one-sided
// Jacobi kernels, staged in shared memory with warp-shuffle reductions ... fuse the singular-value cutoff").
// The reference itself has no live SVD (its `>>` is a QR truncation); the only written-down precedents are dead code
// (trash/mpo.py:59-190, experimental/layers.py:241-325) and the unfinished density-matrix branch
// (matrix_product_operator.py:228 np.linalg.eigh).  Two uses:
//   * textbook rounding: G = R factor of the (transposed) unfolding  -> sigma_i are the singular values and u_i the
//     left singular vectors of the unfolding;
//   * density-matrix rounding: G = M E M^T (symmetric PSD)            -> rows converge to lambda_i u_i^T, sigma_i = sqrt(lambda_i).
//
// Parallel scheme: the n rows are cut into 2P blocks of w rows.  A problem is owned by P co-resident CTAs (cooperative
// launch); every CTA keeps two blocks (2w rows, <= 160 KB) in shared memory, orthogonalises pairs of rows with one warp
// per pair (dot products by warp shuffles, the pair held in registers between the reduction and the rotation), and the
// blocks are re-paired by a round-robin tournament through L2 with one grid-level barrier per outer round
// (2P-1 rounds per sweep).  Convergence (no rotation above tol in a whole sweep) is detected on the device.
// The finalize kernel sorts sigma descending, normalises the rows, applies chi_max / cutoff and reports the kept rank and
// the discarded weight -- the singular-value cutoff is fused here, not done on the host.
#include <cstdlib>

#include <algorithm>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace syn {

constexpr size_t JAC_SMEM_CAP = 160 * 1024;

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// barrier among the P CTAs of one problem (all co-resident: cooperative launch)
__device__ __forceinline__ void problem_barrier(unsigned* ctr, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        while (ld_acquire_u32(ctr) < target) { __nanosleep(32); }
        __threadfence();
    }
    __syncthreads();
}

// ---- fast scalar math for the rotation ----------------------------------------------------------------------------
// FP64 division and sqrt are long dependent software sequences; the rotation sits on the critical path of every Jacobi
// round, so it is built from the MUFU seeds (rcp / rsqrt.approx.f64, 2^-22) plus Newton steps.  Only c needs full
// precision (s = c*t makes c^2 + s^2 = 1 to rounding whatever t is); a t accurate to ~1e-13 only perturbs the
// convergence rate, never the orthogonality of the accumulated transform.
template <typename R> struct Vec16;
template <> struct Vec16<double> { using type = double2; static constexpr int N = 2; __device__ static double2 zero() { return make_double2(0.0, 0.0); } };
template <> struct Vec16<float> { using type = float4; static constexpr int N = 4; __device__ static float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); } };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// FP32 rotation for the preconditioning sweeps: hardware rsqrt / reciprocal are accurate enough (the FP64 sweeps that follow
// restore full accuracy), and an order of magnitude shorter on the critical path.
__device__ __forceinline__ int rotation(float aa, float bb, float ab, float tol2, float null2, float& c, float& s, float& t_out) {
    const float prod = aa * bb;
    const float ab2 = ab * ab;
    if (!(ab2 > tol2 * prod) || !(prod > 1e-35f)) return 0;
    if (aa < null2 && bb < null2) return 0;
    const float d = bb - aa, g = 2.0f * ab;
    const float h = sqrtf(fmaf(d, d, g * g));
    float t = __fdividef(fabsf(g), fabsf(d) + h);
    if ((d < 0.0f) != (g < 0.0f)) t = -t;
    c = rsqrtf(fmaf(t, t, 1.0f));
    s = c * t;
    t_out = t;
    return 3;
}

// Jacobi rotation (Hestenes / Rutishauser) from aa = a.a, bb = b.b, ab = a.b: a' = c a - s b, b' = s a + c b are orthogonal.
// Returns 0 when the pair is already orthogonal to tolerance (ab^2 <= tol^2 aa bb) or numerically null, 1 for a "small"
// rotation (|cos| <= 1e-9: by quadratic convergence the sweep after a sweep of only small rotations cannot rotate above
// tol any more, so no verification sweep is needed), 3 for a big one.
__device__ __forceinline__ int rotation(double aa, double bb, double ab, double tol2, double null2, double& c, double& s, double& t_out) {
    const double prod = aa * bb;
    const double ab2 = ab * ab;
    if (!(ab2 > tol2 * prod) || !(prod > 1e-280)) return 0;
    if (aa < null2 && bb < null2) return 0;     // both rows are below the rank threshold: they will be discarded, leave them
    const int kind = (ab2 > 1e-18 * prod) ? 3 : 1;
    const double d = bb - aa, g = 2.0 * ab;
    const double h2 = fma(d, d, g * g);
    const double h = h2 * rsqrt_newton1(h2);
    double t = fabs(g) * rcp_newton1(fabs(d) + h);         // |t| = |g| / (|d| + sqrt(d^2 + g^2)) <= 1
    if ((d < 0.0) != (g < 0.0)) t = -t;                   // sign(zeta), zeta = d / g
    c = rsqrt_newton2(fma(t, t, 1.0));
    s = c * t;
    t_out = t;
    return kind;
}

// rotate rows a and b (both in shared memory, length n): used by the all-pairs round (t == 0) and by single-CTA problems
template <typename R, int NREG>
__device__ __forceinline__ int rotate_pair(R* __restrict__ a, R* __restrict__ b, int n, int lane, R tol2, R null2) {
    R ra[NREG], rb[NREG];
    R aa = R(0), bb = R(0), ab = R(0);
#pragma unroll
    for (int k = 0; k < NREG; k++) {
        int c = lane + 32 * k;
        ra[k] = c < n ? a[c] : R(0);
        rb[k] = c < n ? b[c] : R(0);
        aa = fma(ra[k], ra[k], aa);
        bb = fma(rb[k], rb[k], bb);
        ab = fma(ra[k], rb[k], ab);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        aa += __shfl_xor_sync(0xffffffffu, aa, o);
        bb += __shfl_xor_sync(0xffffffffu, bb, o);
        ab += __shfl_xor_sync(0xffffffffu, ab, o);
    }
    R c, s, t;
    const int kind = rotation(aa, bb, ab, tol2, null2, c, s, t);
    if (!kind) return 0;
#pragma unroll
    for (int k = 0; k < NREG; k++) {
        int col = lane + 32 * k;
        if (col < n) {
            a[col] = fma(c, ra[k], -s * rb[k]);
            b[col] = fma(s, ra[k], c * rb[k]);
        }
    }
    return kind;
}

// Cross-round workhorse: row a lives in the registers of its warp for the whole outer round, the squared norms of both rows
// are cached (na in a register, *nb in shared memory) and updated analytically (|a'|^2 = aa - t ab, |b'|^2 = bb + t ab), so a
// round costs ONE dot product, one warp reduction, and one trip of the partner row through shared memory.
// FULL: n == 32 * NREG, no bounds checks.
template <typename R, int NREG, bool FULL>
__device__ __forceinline__ int rotate_cached(R (&ra)[NREG], R& na, R* __restrict__ b, R* __restrict__ nb, int n, int lane,
                                              R tol2, R null2) {
    R rb[NREG];
#pragma unroll
    for (int k = 0; k < NREG; k++) {
        int c = lane + 32 * k;
        rb[k] = (FULL || c < n) ? b[c] : R(0);
    }
    R ab0 = R(0), ab1 = R(0);
#pragma unroll
    for (int k = 0; k < NREG; k += 2) {
        ab0 = fma(ra[k], rb[k], ab0);
        ab1 = fma(ra[k + 1], rb[k + 1], ab1);
    }
    const R ab = warp_sum(ab0 + ab1);
    const R bb = *nb;
    R c, s, t;
    const int kind = rotation(na, bb, ab, tol2, null2, c, s, t);
    if (!kind) return 0;
#pragma unroll
    for (int k = 0; k < NREG; k++) {
        int col = lane + 32 * k;
        const R x = ra[k], y = rb[k];
        ra[k] = fma(c, x, -s * y);
        if (FULL || col < n) b[col] = fma(s, x, c * y);
    }
    na = fma(-t, ab, na);
    __syncwarp();          // every lane has read the cached norm above before lane 0 replaces it (racecheck: intra-warp WAR)
    if (lane == 0) *nb = fma(t, ab, bb);
    __syncwarp();          // ... and it is read by every lane of this warp in its next rotation (racecheck: intra-warp RAW)
    return kind;
}

template <typename R, int NREG, bool FULL>
__device__ __forceinline__ R row_sumsq(const R* __restrict__ a, int n, int lane) {
    R s0 = R(0), s1 = R(0);
#pragma unroll
    for (int k = 0; k < NREG; k += 2) {
        int c0 = lane + 32 * k, c1 = c0 + 32;
        R x = (FULL || c0 < n) ? a[c0] : R(0), y = (FULL || c1 < n) ? a[c1] : R(0);
        s0 = fma(x, x, s0);
        s1 = fma(y, y, s1);
    }
    return warp_sum(s0 + s1);
}

constexpr unsigned JF_LOAD_A = 1u, JF_LOAD_B = 2u, JF_STORE_A = 4u, JF_STORE_B = 8u, JF_ALLPAIRS = 16u;   // per-round flags of the hand-over schedule
constexpr int JAC_XFER = 4;        // pieces a round's hand-over is split into (2 blocks x 2 halves), each with its own thread and version word

// grid = (P, problems in this launch).  ctrl layout per problem: [0] barrier counter, [1 + sweep] rotation flags,
// [1 + max_sweeps] sweeps used.
template <typename R, int NREG, int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
jacobi_rows_kernel(R* __restrict__ G, int64_t ld, int64_t bs, int n, int w, int P, unsigned* __restrict__ ctrl, int ctrl_stride,
                   int max_sweeps, double tol2_, double null_rel2_, const uint2* __restrict__ sched, const unsigned* __restrict__ sched_cnt) {
    const R tol2 = (R)tol2_, null_rel2 = (R)null_rel2_;
    constexpr int WARPS = THREADS / 32;
    extern __shared__ __align__(16) unsigned char jac_smem_raw[];
    R* rows = reinterpret_cast<R*>(jac_smem_raw);   // [2w][LDS]
    __shared__ int s_rot;
    __shared__ R s_nrm[32];
    __shared__ __align__(8) uint64_t s_mbar;
    using V = typename Vec16<R>::type;               // 16-byte vector: double2 / float4
    constexpr int VE = Vec16<R>::N;
    const int LDS = (n + VE - 1) / VE * VE;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int p = blockIdx.x;
    G += (int64_t)blockIdx.y * bs;
    unsigned* bar = ctrl + (int64_t)blockIdx.y * ctrl_stride;
    unsigned* flags = bar + 1;
    const int ctrl_base = (max_sweeps + 7) & ~1;      // words before the per-block version counters
    unsigned* ver = bar + ctrl_base;                  // ver[blk] = number of outer rounds block blk has been through
    unsigned bar_target = 0;
    const int NB = 2 * P;          // number of row blocks
    const int Mr = NB - 1;         // outer rounds per sweep
    const bool vec2 = (n % VE == 0) && (ld % VE == 0) && ((((uintptr_t)G) & 15) == 0);
    const bool full = (n == 32 * NREG);
    // TMA path: multi-CTA problem, every block full, rows 16-byte aligned and dense in shared memory
    const bool tma = (P > 1) && vec2 && (n == 2 * P * w) && (LDS == n);
    // Ring schedule (host-built table, see jac_ring_schedule): slot A (rows [0, w)) keeps its block over many rounds -- it stays in the
    // registers of the warps and never travels --, only slot B (rows [w, 2w)) is handed over: half the bytes of the circle method,
    // and the hand-over is bound by the ~46 B/clk one SM can move to and from L2 (tools/microbench/l2_handover.cu).
    const bool ring = tma && (sched != nullptr) && (w <= WARPS);
    uint32_t mbar_phase = 0;
    if (tma) {
        if (tid == 0) mbar_init(&s_mbar, JAC_XFER);   // one producer per piece
        __syncthreads();
    }

    // blocks travel through L2 (ld.cg / st.cg: L1 is not coherent across the CTAs of a problem), 16 bytes per thread
    auto load_block = [&](int blk, int half) {
        R* dst = rows + half * w * LDS;
        if (vec2) {
            const int n2 = n / VE, L2 = LDS / VE;
#pragma unroll 8
            for (int idx = tid; idx < w * L2; idx += THREADS) {      // all loads of a block in flight at once
                int r = idx / L2, c2 = idx - r * L2;
                int gr = blk * w + r;
                V v = Vec16<R>::zero();
                if (gr < n && c2 < n2) v = __ldcg(reinterpret_cast<const V*>(G + (int64_t)gr * ld) + c2);
                reinterpret_cast<V*>(dst + r * LDS)[c2] = v;
            }
        } else {
            for (int idx = tid; idx < w * LDS; idx += THREADS) {
                int r = idx / LDS, c = idx - r * LDS;
                int gr = blk * w + r;
                dst[r * LDS + c] = (gr < n && c < n) ? __ldcg(G + (int64_t)gr * ld + c) : R(0);
            }
        }
    };
    auto store_block = [&](int blk, int half) {
        const R* src = rows + half * w * LDS;
        if (vec2) {
            const int n2 = n / VE, L2 = LDS / VE;
#pragma unroll 4
            for (int idx = tid; idx < w * L2; idx += THREADS) {
                int r = idx / L2, c2 = idx - r * L2;
                int gr = blk * w + r;
                if (gr < n && c2 < n2)
                    __stcg(reinterpret_cast<V*>(G + (int64_t)gr * ld) + c2, reinterpret_cast<const V*>(src + r * LDS)[c2]);
            }
        } else {
            for (int idx = tid; idx < w * LDS; idx += THREADS) {
                int r = idx / LDS, c = idx - r * LDS;
                int gr = blk * w + r;
                if (gr < n && c < n) __stcg(G + (int64_t)gr * ld + c, src[r * LDS + c]);
            }
        }
    };

    // largest squared row norm of the problem -> absolute floor below which a pair of rows is left alone (both would be
    // discarded by the rank threshold of the finalize kernel; without this, null-space rows of a rank-deficient matrix rotate
    // among themselves for ever: their mutual cosines are O(1) noise)
    R null2 = R(0);
    if (null_rel2 > R(0)) {
        unsigned long long* gmax = reinterpret_cast<unsigned long long*>(bar + ctrl_base - 2);
        R local = R(0);
        for (int r = p * 2 * w + warp; r < n && r < (p + 1) * 2 * w; r += WARPS) {
            R sacc = R(0);
            for (int c = lane; c < n; c += 32) { R x = __ldcg(G + (int64_t)r * ld + c); sacc = fma(x, x, sacc); }
            sacc = warp_sum(sacc);
            local = sacc > local ? sacc : local;
        }
        if (lane == 0 && local > R(0)) atomicMax(gmax, (unsigned long long)__double_as_longlong((double)local));   // positive doubles order like integers
        if (P > 1) {
            bar_target += P;
            problem_barrier(bar, bar_target);
        } else {
            __syncthreads();
        }
        const R mx = (R)__longlong_as_double((long long)__ldcg(gmax));
        null2 = null_rel2 * mx;
    }

#ifdef SYN_JACOBI_TIMING
    long long tk[6] = {0, 0, 0, 0, 0, 0};      // poll, load, prologue, steps, epilogue, store+publish   (thread 0 of CTA 1)
    long long tmark = clock64();
#define JTICK(i) do { if (tid == 0) { long long now_ = clock64(); tk[i] += now_ - tmark; tmark = now_; } } while (0)
#else
#define JTICK(i) do { } while (0)
#endif
    R ra[NREG];                    // stationary row of this warp (ring schedule: lives across rounds)
    R na = R(0);
    bool a_regs = false;           // slot A currently lives in registers, its shared-memory copy is stale
#pragma unroll
    for (int k = 0; k < NREG; k++) ra[k] = R(0);
    int sweep = 0;
    bool converged = false;
    for (; sweep < max_sweeps; ++sweep) {
        int rotated = 0;
        for (int t = 0; t < (P == 1 ? 1 : Mr); ++t) {
            JTICK(5);
            // round-robin tournament over the NB blocks (circle method): CTA p plays (b0, b1) in round t
            int b0, b1;
            unsigned fl = JF_LOAD_A | JF_LOAD_B | JF_STORE_A | JF_STORE_B | (t == 0 ? JF_ALLPAIRS : 0u);
            unsigned need0 = (unsigned)(sweep * Mr + t), need1 = need0, done0 = need0 + 1u, done1 = done0;
            if (P == 1) { b0 = 0; b1 = 1; }
            else if (ring) {
                const uint2 e = sched[t * P + p];
                b0 = (int)(e.x & 0xffu); b1 = (int)((e.x >> 8) & 0xffu); fl = (e.x >> 16) & 0xffu;
                const unsigned c0 = sched_cnt[b0] * (unsigned)sweep, c1 = sched_cnt[b1] * (unsigned)sweep;
                need0 = c0 + (e.y & 0xffffu); need1 = c1 + (e.y >> 16);
                done0 = need0 + 1u; done1 = need1 + 1u;
            }
            else if (p == 0) { b0 = NB - 1; b1 = t % Mr; }
            else { b0 = (t + p) % Mr; b1 = (t - p + Mr) % Mr; }
            if (tma) {
                if (lane == 0 && warp < JAC_XFER) {   // one thread (of different warps) per piece: the hand-overs run in parallel
                    const int h = warp / (JAC_XFER / 2), sub = warp % (JAC_XFER / 2), blk = h ? b1 : b0;
                    const int wr = w / (JAC_XFER / 2);                                  // rows per piece
                    const uint32_t row_bytes = (uint32_t)n * (uint32_t)sizeof(R), blk_bytes = (uint32_t)wr * row_bytes;
                    // point-to-point hand-over instead of a grid barrier: a block is ready when the CTA that held it in the
                    // previous round has published its version (every block takes part in every round)
                    if (fl & (h ? JF_LOAD_B : JF_LOAD_A)) {
                        const unsigned need = h ? need1 : need0;
                        for (unsigned spin = 0; ld_acquire_u32(ver + blk * (JAC_XFER / 2) + sub) < need; ++spin) {
                            if (spin > (1u << 26)) __trap();
                            __nanosleep(20);
                        }
                        JTICK(0);
                        fence_proxy_async();
                        mbar_expect_tx(&s_mbar, blk_bytes);
                        const R* src = G + (int64_t)(blk * w + sub * wr) * ld;
                        R* dst = rows + (h * w + sub * wr) * LDS;
                        if (ld == n) bulk_g2s(dst, src, blk_bytes, &s_mbar);
                        else for (int r = 0; r < wr; ++r) bulk_g2s(dst + r * LDS, src + (int64_t)r * ld, row_bytes, &s_mbar);
                    } else {
                        mbar_arrive(&s_mbar);          // this piece stays where it is
                    }
                }
                mbar_wait(&s_mbar, mbar_phase);
                mbar_phase ^= 1u;
                JTICK(1);
            } else if (P > 1 || sweep == 0) {
                load_block(b0, 0);
                load_block(b1, 1);
                __syncthreads();
            }
            if (fl & JF_ALLPAIRS) {
                if (a_regs) {              // slot A comes back from the registers
                    if (warp < w) {
#pragma unroll
                        for (int k = 0; k < NREG; k++) { int c = lane + 32 * k; if (c < n) rows[warp * LDS + c] = ra[k]; }
                    }
                    a_regs = false;
                    __syncthreads();
                }
                // all pairs among the 2w local rows (within-block pairs are visited once per sweep, here)
                const int items = 2 * w, rounds = items - 1;
                for (int s = 0; s < rounds; ++s) {
                    for (int k = warp; k < w; k += WARPS) {
                        int i, j;
                        if (k == 0) { i = items - 1; j = s; }
                        else { i = (s + k) % rounds; j = (s - k + rounds) % rounds; }
                        rotated |= rotate_pair<R, NREG>(rows + i * LDS, rows + j * LDS, n, lane, tol2, null2);
                    }
                    __syncthreads();
                }
            } else if (w <= WARPS) {
                // cross pairs only: row k of block b0 (held in registers by warp k) with row (k+s) mod w of block b1
                if (warp < w) {
                    if (!a_regs) {
#pragma unroll
                        for (int k = 0; k < NREG; k++) { int c = lane + 32 * k; ra[k] = c < n ? rows[warp * LDS + c] : R(0); }
                    }
                    na = R(0);             // recomputed every round: the analytic updates of the cached norm do not accumulate
#pragma unroll
                    for (int k = 0; k < NREG; k++) na = fma(ra[k], ra[k], na);
                    na = warp_sum(na);
                    const R nbv = full ? row_sumsq<R, NREG, true>(rows + (w + warp) * LDS, n, lane)
                                            : row_sumsq<R, NREG, false>(rows + (w + warp) * LDS, n, lane);
                    if (lane == 0) s_nrm[warp] = nbv;
                }
                __syncthreads();
                JTICK(2);
                for (int s = 0; s < w; ++s) {
                    if (warp < w) {
                        const int j = (warp + s) & (w - 1);           // w is a power of two
                        rotated |= full ? rotate_cached<R, NREG, true>(ra, na, rows + (w + j) * LDS, s_nrm + j, n, lane, tol2, null2)
                                        : rotate_cached<R, NREG, false>(ra, na, rows + (w + j) * LDS, s_nrm + j, n, lane, tol2, null2);
                    }
                    __syncthreads();
                }
                a_regs = true;
                if (!ring || (fl & JF_STORE_A)) {          // slot A leaves (always, under the circle method)
                    if (warp < w) {
#pragma unroll
                        for (int k = 0; k < NREG; k++) { int c = lane + 32 * k; if (c < n) rows[warp * LDS + c] = ra[k]; }
                    }
                    a_regs = false;
                }
                JTICK(3);
                __syncthreads();
            } else {
                for (int s = 0; s < w; ++s) {
                    for (int k = warp; k < w; k += WARPS) {
                        int j = w + ((k + s) % w);
                        rotated |= rotate_pair<R, NREG>(rows + k * LDS, rows + j * LDS, n, lane, tol2, null2);
                    }
                    __syncthreads();
                }
            }
            if (tma) {
                fence_proxy_async();               // every thread: its st.shared results become visible to the async proxy
                __syncthreads();
                JTICK(4);
                if (lane == 0 && warp < JAC_XFER) {
                    const int h = warp / (JAC_XFER / 2), sub = warp % (JAC_XFER / 2), blk = h ? b1 : b0;
                    const int wr = w / (JAC_XFER / 2);
                    const uint32_t row_bytes = (uint32_t)n * (uint32_t)sizeof(R), blk_bytes = (uint32_t)wr * row_bytes;
                    if (fl & (h ? JF_STORE_B : JF_STORE_A)) {
                        R* dst = G + (int64_t)(blk * w + sub * wr) * ld;
                        const R* src = rows + (h * w + sub * wr) * LDS;
                        if (ld == n) bulk_s2g(dst, src, blk_bytes);
                        else for (int r = 0; r < wr; ++r) bulk_s2g(dst + (int64_t)r * ld, src + r * LDS, row_bytes);
                        bulk_commit_wait_all();        // writes complete (and shared memory free) before they are published
                        fence_proxy_async();           // async-proxy writes ordered before the generic-proxy release below
                        st_release_u32(ver + blk * (JAC_XFER / 2) + sub, h ? done1 : done0);
                    }
                }
            } else if (P > 1) {
                store_block(b0, 0);
                store_block(b1, 1);
                bar_target += P;
                problem_barrier(bar, bar_target);
            }
        }
        // sweep-level convergence vote
        if (tid == 0) s_rot = 0;
        __syncthreads();
        if (rotated && lane == 0) atomicOr(&s_rot, rotated);
        __syncthreads();
        if (P > 1) {
            if (tid == 0 && s_rot) atomicOr(flags + sweep, (unsigned)s_rot);
            bar_target += P;
            problem_barrier(bar, bar_target);
            unsigned f = ld_acquire_u32(flags + sweep);
            if ((f & 2u) == 0u) { ++sweep; converged = true; break; }      // no rotation, or only small ones: converged
        } else {
            if ((s_rot & 2) == 0) { ++sweep; converged = true; break; }
            __syncthreads();
        }
    }
    if (P == 1) {
        store_block(0, 0);
        store_block(1, 1);
    }
    if (p == 0 && tid == 0) {
        flags[max_sweeps] = (unsigned)sweep;                  // sweeps used (diagnostic)
        flags[max_sweeps + 1] = converged ? 1u : 0u;          // read by jacobi_finalize_kernel: an unconverged problem is reported in info[1]
    }
#ifdef SYN_JACOBI_TIMING
    if (p == 1 && tid == 0 && blockIdx.y == 0)
        printf("jacobi n=%d w=%d P=%d sweeps=%d cycles: poll %lld  load %lld  prologue %lld  steps %lld  epilogue %lld  store+publish %lld\n", n, w, P,
               sweep, tk[0], tk[1], tk[2], tk[3], tk[4], tk[5]);
#endif
}

// ---- finalize: sort, normalise, cut ------------------------------------------------------------------------------------
// G rows (after jacobi_rows_kernel) -> Ut[k][:] = row_{perm[k]} / |row_{perm[k]}|, sigma[k] (descending).
//   sqrt_mode = 1: sigma = sqrt(|row|)  (density-matrix use: rows are lambda_i u_i^T)
//   sqrt_mode = 2: sigma = sqrt(|row|^2 - shift[0])  (rows of the rotated Cholesky factor of G + shift I: sqrt(lambda_i + shift) q_i^T)
//   info[0] = kept rank = #{k < chi_max : sigma_k > max(cutoff, rank_tol) * sigma_0}, at least 1
//   winfo[0] = discarded weight sum_{k >= kept} sigma_k^2 ; winfo[1] = sigma_0
__global__ void __launch_bounds__(1024, 1)
jacobi_finalize_kernel(const double* __restrict__ G, int64_t ld, int64_t bs, int n, double* __restrict__ Ut, int64_t ldu, int64_t ubs,
                       double* __restrict__ sigma, int64_t sbs, int* __restrict__ info, double* __restrict__ winfo, int chi_max, double cutoff,
                       double rank_tol, int sqrt_mode, const double* __restrict__ shift, const unsigned* __restrict__ ctrl, int ctrl_stride,
                       int max_sweeps) {
    __shared__ double key[1024];
    __shared__ int perm[1024];
    __shared__ double red[32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    G += (int64_t)blockIdx.x * bs;
    Ut += (int64_t)blockIdx.x * ubs;
    sigma += (int64_t)blockIdx.x * sbs;
    info += blockIdx.x * 2;
    winfo += blockIdx.x * 2;
    int npow = 1;
    while (npow < n) npow <<= 1;
    for (int r = warp; r < npow; r += 32) {
        double s = 0.0;
        if (r < n) {
            const double* g = G + (int64_t)r * ld;
            for (int c = lane; c < n; c += 32) { double x = g[c]; s = fma(x, x, s); }
            s = warp_sum(s);
        }
        if (lane == 0) { key[r] = r < n ? sqrt(s) : -1.0; perm[r] = r; }
    }
    __syncthreads();
    // bitonic sort, descending by key
    for (int k = 2; k <= npow; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < npow; i += 1024) {
                int l = i ^ j;
                if (l > i) {
                    bool desc = ((i & k) == 0);
                    double ki = key[i], kl = key[l];
                    if ((ki < kl) == desc) {
                        key[i] = kl; key[l] = ki;
                        int t = perm[i]; perm[i] = perm[l]; perm[l] = t;
                    }
                }
            }
            __syncthreads();
        }
    }
    // normalised rows in sorted order
    for (int k = warp; k < n; k += 32) {
        const double nrm = key[k];
        const double inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
        const double* g = G + (int64_t)perm[k] * ld;
        double* u = Ut + (int64_t)k * ldu;
        for (int c = lane; c < n; c += 32) u[c] = g[c] * inv;
    }
    // singular values, kept rank, discarded weight
    const double delta = (sqrt_mode == 2 && shift) ? shift[blockIdx.x] : 0.0;
    auto sv_of = [&](double k) { return sqrt_mode == 1 ? sqrt(k) : (sqrt_mode == 2 ? sqrt(fmax(fma(k, k, -delta), 0.0)) : k); };
    const double s0 = sv_of(key[0]);
    const double thr = (cutoff > rank_tol ? cutoff : rank_tol) * s0;
    int cnt = 0;
    double wsum = 0.0;
    for (int k = tid; k < n; k += 1024) {
        double sv = sv_of(key[k]);
        sigma[k] = sv;
        if (k < chi_max && sv > thr) cnt++;
    }
    __shared__ int s_cnt;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    if (cnt) atomicAdd(&s_cnt, cnt);
    __syncthreads();
    int keep = s_cnt < 1 ? 1 : s_cnt;
    // singular values are sorted, so the kept set {k < chi_max, sv > thr} is a prefix of length keep
    for (int k = tid; k < n; k += 1024) {
        if (k >= keep) { double sv = sv_of(key[k]); wsum = fma(sv, sv, wsum); }
    }
    wsum = warp_sum(wsum);
    if (lane == 0) red[warp] = wsum;
    __syncthreads();
    if (warp == 0) {
        double v = red[lane];
        v = warp_sum(v);
        // info[1] = n, negated when the rows kernel ran out of sweeps before its convergence vote passed (ctrl word max_sweeps + 2)
        const bool conv = ctrl == nullptr || ctrl[(int64_t)blockIdx.x * ctrl_stride + max_sweeps + 2] != 0u;
        if (lane == 0) { info[0] = keep; info[1] = conv ? n : -n; winfo[0] = v; winfo[1] = s0; }
    }
}

// ---- host drivers --------------------------------------------------------------------------------------------------------
// ---- ring schedule ----------------------------------------------------------------------------------------------------------
// Block-level ordering of one sweep for NB = 2P blocks (P a power of two), as a table [round][CTA] -> (block in slot A, block in
// slot B, flags, versions to wait for).  Recursive: the first half of the blocks stays put (slot A of P CTAs), the second half
// rotates past them (slot B) for NB/2 rounds; then both halves are treated the same way on half the CTAs each, down to pairs of
// blocks, whose round also visits the pairs inside the two blocks.  NB - 1 rounds like the circle method, every pair of rows once,
// same or fewer sweeps to converge (simulated on Gram matrices of the density-matrix algorithm) -- but slot A only moves when the
// recursion descends, log2(P) times per sweep instead of every round.
struct RingRound { std::vector<int> a, b; bool allpairs; };

static void ring_rounds(const std::vector<int>& blocks, const std::vector<int>& ctas, size_t first, std::vector<RingRound>& out, int P) {
    const size_t m = blocks.size();
    auto at = [&](size_t r) -> RingRound& {
        while (out.size() <= r) { RingRound rr; rr.a.assign(P, -1); rr.b.assign(P, -1); rr.allpairs = false; out.push_back(rr); }
        return out[r];
    };
    if (m == 2) {
        RingRound& rr = at(first);
        rr.a[ctas[0]] = blocks[0]; rr.b[ctas[0]] = blocks[1]; rr.allpairs = true;
        return;
    }
    const size_t h = m / 2;
    for (size_t t = 0; t < h; ++t) {
        RingRound& rr = at(first + t);
        for (size_t q = 0; q < h; ++q) { rr.a[ctas[q]] = blocks[q]; rr.b[ctas[q]] = blocks[h + (q + t) % h]; }
    }
    std::vector<int> s(blocks.begin(), blocks.begin() + h), mv(blocks.begin() + h, blocks.end());
    std::vector<int> c0(ctas.begin(), ctas.begin() + h / 2), c1(ctas.begin() + h / 2, ctas.end());
    ring_rounds(s, c0, first + h, out, P);
    ring_rounds(mv, c1, first + h, out, P);
}

// entries: x = blkA | blkB << 8 | flags << 16 ; y = needA | needB << 16 (stores of that block earlier in the sweep); cnt[blk] = stores
// of the block per sweep (the version counters keep counting across sweeps)
static void jac_ring_schedule(int P, std::vector<uint2>& table, std::vector<unsigned>& cnt) {
    const int NB = 2 * P;
    std::vector<int> blocks(NB), ctas(P);
    for (int i = 0; i < NB; ++i) blocks[i] = i;
    for (int i = 0; i < P; ++i) ctas[i] = i;
    std::vector<RingRound> rounds;
    ring_rounds(blocks, ctas, 0, rounds, P);
    const int R = (int)rounds.size();           // NB - 1
    table.assign((size_t)R * P, make_uint2(0u, 0u));
    cnt.assign(NB, 0u);
    for (int r = 0; r < R; ++r) {
        for (int q = 0; q < P; ++q) {
            const int a = rounds[r].a[q], b = rounds[r].b[q];
            unsigned fl = rounds[r].allpairs ? JF_ALLPAIRS : 0u;
            if (r == 0 || rounds[r - 1].a[q] != a) fl |= JF_LOAD_A;
            if (r == 0 || rounds[r - 1].b[q] != b) fl |= JF_LOAD_B;
            if (r == R - 1 || rounds[r + 1].a[q] != a) fl |= JF_STORE_A;
            if (r == R - 1 || rounds[r + 1].b[q] != b) fl |= JF_STORE_B;
            table[(size_t)r * P + q] = make_uint2((unsigned)a | ((unsigned)b << 8) | (fl << 16), cnt[a] | (cnt[b] << 16));
        }
        for (int q = 0; q < P; ++q) {            // versions advance after the round
            const unsigned fl = (table[(size_t)r * P + q].x >> 16) & 0xffu;
            if (fl & JF_STORE_A) cnt[rounds[r].a[q]]++;
            if (fl & JF_STORE_B) cnt[rounds[r].b[q]]++;
        }
    }
}

constexpr int JAC_SCHED_MAX_LOG2P = 6;                     // P <= 64 (n = 1024, w = 8)
constexpr int JAC_SCHED_ENTRIES = 10800;                   // sum over P = 2 .. 64 of (2P - 1) P  = 10794
__device__ uint2 g_jac_sched[JAC_SCHED_ENTRIES];
__device__ unsigned g_jac_sched_cnt[JAC_SCHED_MAX_LOG2P * 128];

static bool jac_env_ring() {   // experiment knob: SYN_JACOBI_RING=0 falls back to the circle method
    static int v = -1;
    if (v < 0) { const char* e = getenv("SYN_JACOBI_RING"); v = e ? (atoi(e) != 0) : 1; }
    return v != 0;
}

// device pointers to the schedule of a P-CTA problem (uploaded once per device), or nulls when P is not a power of two
static int jac_schedule_ptrs(int P, const uint2** sched, const unsigned** cnt) {
    *sched = nullptr; *cnt = nullptr;
    int lg = 0;
    while ((1 << lg) < P) ++lg;
    if (P < 2 || (1 << lg) != P || lg > JAC_SCHED_MAX_LOG2P || !jac_env_ring()) return 0;
    static std::mutex mu;
    static bool uploaded[64] = {false};
    int dev = 0;
    SYN_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (dev >= 0 && dev < 64 && !uploaded[dev]) {
        std::vector<uint2> all(JAC_SCHED_ENTRIES, make_uint2(0u, 0u));
        std::vector<unsigned> allc(JAC_SCHED_MAX_LOG2P * 128, 0u);
        size_t off = 0;
        for (int k = 1; k <= JAC_SCHED_MAX_LOG2P; ++k) {
            std::vector<uint2> t; std::vector<unsigned> c;
            jac_ring_schedule(1 << k, t, c);
            std::copy(t.begin(), t.end(), all.begin() + off);
            std::copy(c.begin(), c.end(), allc.begin() + (size_t)(k - 1) * 128);
            off += t.size();
        }
        SYN_CUDA(cudaMemcpyToSymbol(g_jac_sched, all.data(), all.size() * sizeof(uint2)));
        SYN_CUDA(cudaMemcpyToSymbol(g_jac_sched_cnt, allc.data(), allc.size() * sizeof(unsigned)));
        uploaded[dev] = true;
    }
    size_t off = 0;
    for (int k = 1; k < lg; ++k) off += (size_t)((2 << k) - 1) * (1 << k);
    uint2* base = nullptr; unsigned* cbase = nullptr;
    SYN_CUDA(cudaGetSymbolAddress((void**)&base, g_jac_sched));
    SYN_CUDA(cudaGetSymbolAddress((void**)&cbase, g_jac_sched_cnt));
    *sched = base + off;
    *cnt = cbase + (size_t)(lg - 1) * 128;
    return 0;
}
struct JacPlan { int w, P, nreg; size_t smem; };

static int jac_env_w() {   // experiment knob: SYN_JACOBI_W=8 forces the block height of multi-CTA problems
    static int v = -1;
    if (v < 0) { const char* e = getenv("SYN_JACOBI_W"); v = e ? atoi(e) : 0; }
    return v;
}

static int jac_plan(int n, JacPlan& pl, size_t esize = sizeof(double)) {
    SYN_REQUIRE(n >= 1 && n <= 1024, "syn_jacobi_rows_f64: n=%d out of range (1..1024)", n);
    const int ve = (int)(16 / esize);
    const int LDS = (n + ve - 1) / ve * ve;
    pl.nreg = n <= 128 ? 4 : (n <= 256 ? 8 : (n <= 512 ? 16 : 32));
    int w;
    if (n <= 128) {                 // whole problem in one CTA (P = 1): no grid barriers, good for large batches
        w = 1;
        while (2 * w < n) w <<= 1;
    } else {                        // w rows per block, one warp per stationary row in the cross rounds
        w = jac_env_w() > 0 ? jac_env_w() : ((n > 256 && esize == sizeof(double)) ? 8 : 16);   // FP64, n > 256: 8 warps of 255 registers (a row is 16 registers per lane)
        while (w > 1 && (size_t)2 * w * LDS * esize > JAC_SMEM_CAP) w >>= 1;
    }
    pl.w = w;
    pl.P = (n + 2 * w - 1) / (2 * w);
    pl.smem = (size_t)2 * w * LDS * esize;
    return 0;
}

template <typename R, int NREG, int THREADS>
static int launch_jacobi(R* G, int64_t ld, int64_t bs, int n, int batch, const JacPlan& pl, unsigned* ctrl, int ctrl_stride,
                         int max_sweeps, double tol, double null_rel, cudaStream_t st) {
    auto kern = jacobi_rows_kernel<R, NREG, THREADS>;
    static PerDevice configured;          // value = co-resident CTAs of this kernel on the device
    const int dev_ = current_device();
    int max_ctas = 0;
    if (!configured.get(dev_, &max_ctas)) {
        SYN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)JAC_SMEM_CAP));
        int per_sm = 0;
        SYN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, JAC_SMEM_CAP));
        max_ctas = per_sm * sm_count();
        configured.set(dev_, max_ctas);
    }
    SYN_REQUIRE(pl.P <= max_ctas, "syn_jacobi_rows_f64: problem needs %d co-resident CTAs, device fits %d", pl.P, max_ctas);
    int chunk = max_ctas / pl.P;
    if (chunk > 65535) chunk = 65535;
    const uint2* sched = nullptr;
    const unsigned* sched_cnt = nullptr;
    if (pl.P > 1) { if (int rc = jac_schedule_ptrs(pl.P, &sched, &sched_cnt)) return rc; }
    double tol2 = tol * tol, null_rel2 = null_rel * null_rel;
    for (int b0 = 0; b0 < batch; b0 += chunk) {
        int nb = (batch - b0) < chunk ? (batch - b0) : chunk;
        R* Gb = G + (int64_t)b0 * bs;
        unsigned* cb = ctrl + (int64_t)b0 * ctrl_stride;
        int w = pl.w, P = pl.P;
        void* args[] = {&Gb, &ld, &bs, &n, &w, &P, &cb, &ctrl_stride, &max_sweeps, &tol2, &null_rel2, &sched, &sched_cnt};
        dim3 grid(pl.P, nb), block(THREADS);
        if (pl.P > 1) {
            SYN_CUDA(cudaLaunchCooperativeKernel((void*)kern, grid, block, args, pl.smem, st));
            note_launch();
        } else {
            kern<<<grid, block, pl.smem, st>>>(Gb, ld, bs, n, w, P, cb, ctrl_stride, max_sweeps, tol2, null_rel2, sched, sched_cnt);
            if (int rc = launch_status("jacobi_rows_kernel")) return rc;
        }
    }
    return 0;
}

int jacobi_ctrl_stride(int max_sweeps) { return ((max_sweeps + 7) & ~1) + 128 * (JAC_XFER / 2); }   // words: barrier, flags[max_sweeps], sweeps used, pad, 64-bit max norm, versions of up to 128 blocks x pieces
size_t jacobi_ctrl_bytes(int batch, int max_sweeps) { return (size_t)batch * jacobi_ctrl_stride(max_sweeps) * sizeof(unsigned); }

int jacobi_rows_f64(double* G, int64_t ld, int64_t bs, int n, int batch, void* ctrl, size_t ctrl_bytes, int max_sweeps, double tol,
                    double null_rel, cudaStream_t st) {
    SYN_REQUIRE(batch >= 1 && max_sweeps >= 1, "syn_jacobi_rows_f64: bad batch / max_sweeps");
    SYN_REQUIRE(ctrl_bytes >= jacobi_ctrl_bytes(batch, max_sweeps), "syn_jacobi_rows_f64: control buffer too small");
    JacPlan pl;
    if (int rc = jac_plan(n, pl)) return rc;
    SYN_CUDA(cudaMemsetAsync(ctrl, 0, jacobi_ctrl_bytes(batch, max_sweeps), st));
    const int stride = jacobi_ctrl_stride(max_sweeps);
    switch (pl.nreg) {
        case 4: return launch_jacobi<double, 4, 512>(G, ld, bs, n, batch, pl, (unsigned*)ctrl, stride, max_sweeps, tol, null_rel, st);
        case 8: return launch_jacobi<double, 8, 512>(G, ld, bs, n, batch, pl, (unsigned*)ctrl, stride, max_sweeps, tol, null_rel, st);
        case 16:
            if (pl.w > 8) return launch_jacobi<double, 16, 512>(G, ld, bs, n, batch, pl, (unsigned*)ctrl, stride, max_sweeps, tol, null_rel, st);
            return launch_jacobi<double, 16, 256>(G, ld, bs, n, batch, pl, (unsigned*)ctrl, stride, max_sweeps, tol, null_rel, st);   // w = 8: one warp per stationary row, 255 registers each
        default: return launch_jacobi<double, 32, 256>(G, ld, bs, n, batch, pl, (unsigned*)ctrl, stride, max_sweeps, tol, null_rel, st);
    }
}

int jacobi_finalize_f64(const double* G, int64_t ld, int64_t bs, int n, int batch, double* Ut, int64_t ldu, int64_t ubs, double* sigma,
                        int64_t sbs, int* info, double* winfo, int chi_max, double cutoff, double rank_tol, int sqrt_mode, const double* shift,
                        const void* ctrl, int max_sweeps, cudaStream_t st) {
    SYN_REQUIRE(n >= 1 && n <= 1024 && batch >= 1, "syn_jacobi_finalize_f64: n=%d batch=%d out of range", n, batch);
    SYN_REQUIRE(ctrl == nullptr || max_sweeps >= 1, "syn_jacobi_finalize_f64: max_sweeps must be the value given to syn_jacobi_rows_f64");
    jacobi_finalize_kernel<<<batch, 1024, 0, st>>>(G, ld, bs, n, Ut, ldu, ubs, sigma, sbs, info, winfo, chi_max, cutoff, rank_tol, sqrt_mode, shift,
                                                   (const unsigned*)ctrl, ctrl ? jacobi_ctrl_stride(max_sweeps) : 0, max_sweeps);
    return launch_status("jacobi_finalize_kernel");
}

}  // namespace syn

namespace syn {
// out[0] = max_ij |X[i][j] - delta_ij|   (single CTA)
__global__ void __launch_bounds__(1024, 1) identity_dev_kernel(const double* __restrict__ X, int n, double* __restrict__ out) {
    __shared__ double red[32];
    double m = 0.0;
    for (int64_t i = threadIdx.x; i < (int64_t)n * n; i += 1024) {
        int r = (int)(i / n), c = (int)(i - (int64_t)r * n);
        double v = fabs(X[i] - (r == c ? 1.0 : 0.0));
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
}  // namespace syn

extern "C" int syn_identity_deviation_f64(const double* X, int n, double* out, void* stream) {
    using namespace syn;
    SYN_REQUIRE(n >= 1, "syn_identity_deviation_f64: bad n");
    identity_dev_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(X, n, out);
    return launch_status("identity_dev_kernel");
}

extern "C" int syn_jacobi_ring_schedule(int P, uint32_t* table, uint32_t* cnt) {
    int lg = 0;
    while ((1 << lg) < P) ++lg;
    SYN_REQUIRE(P >= 2 && P <= 64 && (1 << lg) == P && table && cnt, "syn_jacobi_ring_schedule: P=%d must be a power of two in 2..64", P);
    std::vector<uint2> t;
    std::vector<unsigned> c;
    syn::jac_ring_schedule(P, t, c);
    for (size_t i = 0; i < t.size(); ++i) { table[2 * i] = t[i].x; table[2 * i + 1] = t[i].y; }
    for (size_t i = 0; i < c.size(); ++i) cnt[i] = c[i];
    return 0;
}
extern "C" size_t syn_jacobi_ctrl_bytes(int batch, int max_sweeps) { return syn::jacobi_ctrl_bytes(batch, max_sweeps); }
extern "C" int syn_jacobi_ctrl_stride(int max_sweeps) { return syn::jacobi_ctrl_stride(max_sweeps); }

extern "C" int syn_jacobi_rows_f64(double* G, int64_t ld, int64_t bs, int n, int batch, void* ctrl, size_t ctrl_bytes, int max_sweeps,
                                   double tol, double null_rel, void* stream) {
    return syn::jacobi_rows_f64(G, ld, bs, n, batch, ctrl, ctrl_bytes, max_sweeps, tol, null_rel, (cudaStream_t)stream);
}

extern "C" int syn_jacobi_finalize_f64(const double* G, int64_t ld, int64_t bs, int n, int batch, double* Ut, int64_t ldu, int64_t ubs,
                                       double* sigma, int64_t sbs, int* info, double* winfo, int chi_max, double cutoff, double rank_tol,
                                       int sqrt_mode, const double* shift, const void* ctrl, int max_sweeps, void* stream) {
    return syn::jacobi_finalize_f64(G, ld, bs, n, batch, Ut, ldu, ubs, sigma, sbs, info, winfo, chi_max, cutoff, rank_tol, sqrt_mode, shift,
                                    ctrl, max_sweeps, (cudaStream_t)stream);
}

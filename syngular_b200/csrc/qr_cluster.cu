// Cluster-resident Householder QR for the plateau shapes of the rounding sweep (m <= ~800 rows, <= 256 kept columns), sm_100a.
//
// The blocked path of qr.cu spends its time in ~40 small dependent launches per factorisation (panel kernel + two GEMMs per
// panel, twice: factor and form Q).  Here ONE launch does both: a thread-block cluster of 8 CTAs keeps the whole m x q working
// matrix in distributed shared memory (each CTA owns m/8 rows, all columns; 128 KB for 512 x 256), one thread per column.
// Every Householder step is: per-CTA partial dot products of the pivot column with all later columns (no intra-CTA
// reduction: a thread walks its own column), a reduce-scatter of the partial sums to the CTA owning each group of 32 columns and
// an all-gather of the resulting coefficients through DSMEM (st.shared::cluster; ~6 KB per CTA and step, two cluster barriers),
// and the rank-1 update of the local rows.  Q is then formed in place
// (LAPACK dorg2r order, last reflector first) with the same step structure; extra columns (q > n: the reference's bond
// "inflation") start as identity columns and come out as the orthonormal completion.
//
// Same arithmetic as dgeqr2 + dorg2r (unblocked Householder), i.e. what np.linalg.qr(L, mode="complete")[:, :q] spans
// (matrix_product_state.py:443-446, matrix_product_operator.py:555-558).
#include <cooperative_groups.h>

#include <cstdlib>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace syn {

constexpr int QRC_CLUSTER = 8;
constexpr int QRC_COLS = 256;         // one thread column per matrix column: q <= 256
constexpr int QRC_RG = 4;             // row groups: 4 threads share a column, each walks every 4th local row
constexpr int QRC_THREADS = QRC_COLS * QRC_RG;
constexpr size_t QRC_SMEM_CAP = 200 * 1024;
static_assert(QRC_CLUSTER % QRC_RG == 0, "each row group sends to CLUSTER / RG peers");

struct QrcShared {                    // fixed-size part of the shared memory image (the row slice follows)
    double part[2][QRC_CLUSTER][QRC_COLS];   // all-gathered partial dot products, double-buffered by step parity
    double prow[2][QRC_COLS];                // pivot row, broadcast by the CTA that holds global row j
    double red[QRC_RG][QRC_COLS];            // intra-CTA combination of the row groups' partial sums
    double tau[QRC_COLS];                    // reflector scalars kept for the form-Q phase:  H_j = I - tau_j v v^T,
    double vscale[QRC_COLS];                 //   v = (1, vscale_j * stored column below the diagonal)
};

// tau and 1/(alpha - beta) of dlarfg from alpha = a_jj and s = sum of squares below the diagonal.  FP64 sqrt and division are
// ~100-instruction dependent software sequences; this sits on the critical path of every step, so the normal range is served by
// MUFU seeds + Newton steps (full double precision) and only extreme magnitudes take the library path.
__device__ __forceinline__ void reflector_scalars(double alpha, double s, double& tau, double& vscale) {
    if (s == 0.0) { tau = 0.0; vscale = 0.0; return; }
    const double x = fma(alpha, alpha, s);
    if (x > 1e-280 && x < 1e280) {
        const double rs = rsqrt_newton2(x);
        double nrm = x * rs;
        nrm = fma(fma(-nrm, nrm, x), 0.5 * rs, nrm);
        const double beta = alpha >= 0.0 ? -nrm : nrm;
        const double rbeta = alpha >= 0.0 ? -rs : rs;
        tau = (beta - alpha) * rbeta;
        vscale = rcp_newton2(alpha - beta);
    } else {
        const double nrm = sqrt(x);
        const double beta = alpha >= 0.0 ? -nrm : nrm;
        tau = (beta - alpha) / beta;
        vscale = 1.0 / (alpha - beta);
    }
}

// Rows are dealt to the CTAs cyclically (global row g lives in CTA g % 8 at local index g / 8) so that the work of every
// CTA shrinks evenly as the factorisation advances; inside a CTA, 4 threads share a column (local rows lr % 4 == rg).
// One step = partial dots -> all-gather through DSMEM -> ONE cluster barrier -> every CTA reduces redundantly and updates.
__global__ void __cluster_dims__(QRC_CLUSTER, 1, 1) __launch_bounds__(QRC_THREADS, 1)
qr_cluster_kernel(const double* __restrict__ A, int64_t a_rs, int64_t a_cs, int64_t a_bs, int m, int kf, int qk, double* __restrict__ Q,
                  int64_t q_rs, int64_t q_cs, int64_t q_bs, int rows_per) {
    extern __shared__ __align__(16) unsigned char qrc_raw[];
    QrcShared* sh = reinterpret_cast<QrcShared*>(qrc_raw);
    double* P = reinterpret_cast<double*>(qrc_raw + sizeof(QrcShared));     // [rows_per][qk]
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int c = threadIdx.x % QRC_COLS;            // my column
    const int rg = threadIdx.x / QRC_COLS;           // my row group
    const int prob = blockIdx.x / QRC_CLUSTER;
    A += (int64_t)prob * a_bs;
    Q += (int64_t)prob * q_bs;
    const int nloc = (m > rank) ? (m - rank + QRC_CLUSTER - 1) / QRC_CLUSTER : 0;   // local row lr <-> global row lr * 8 + rank
    constexpr int PEERS_PER_RG = QRC_CLUSTER / QRC_RG;
    QrcShared* peer[PEERS_PER_RG];
#pragma unroll
    for (int k = 0; k < PEERS_PER_RG; k++) peer[k] = cluster.map_shared_rank(sh, rg * PEERS_PER_RG + k);

    // ---- load: columns < kf from A, completion columns as identity columns ---------------------------------------------
    for (int idx = threadIdx.x; idx < nloc * qk; idx += QRC_THREADS) {
        int lr = idx / qk, cc = idx - lr * qk;
        int gr = lr * QRC_CLUSTER + rank;
        P[lr * qk + cc] = cc < kf ? A[(int64_t)gr * a_rs + (int64_t)cc * a_cs] : (gr == cc ? 1.0 : 0.0);
    }
    cluster.sync();          // also: every CTA of the cluster is running before the first remote store

    // One Householder step on pivot column j acting on columns [c_lo, c_hi).  FACTOR: the scalars come from the data and are kept;
    // otherwise (form Q) they are read back.  The stored column j stays unscaled (vscale is carried instead), so nobody writes
    // column j while the others read it.
    auto step = [&](int j, int c_lo, int c_hi, bool factor, int par) {
        const bool active = (c >= c_lo && c < c_hi);
        const bool owner = (j % QRC_CLUSTER) == rank;
        const int jl = j / QRC_CLUSTER;
        const int lo = j >= rank ? (j - rank) / QRC_CLUSTER + 1 : 0;       // first local row strictly below global row j
        const int start = lo + ((rg - lo) & (QRC_RG - 1));                // ... that belongs to my row group
        {
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            if (active) {
                int lr = start;
                for (; lr + 3 * QRC_RG < nloc; lr += 4 * QRC_RG) {
                    s0 = fma(P[lr * qk + j], P[lr * qk + c], s0);
                    s1 = fma(P[(lr + QRC_RG) * qk + j], P[(lr + QRC_RG) * qk + c], s1);
                    s2 = fma(P[(lr + 2 * QRC_RG) * qk + j], P[(lr + 2 * QRC_RG) * qk + c], s2);
                    s3 = fma(P[(lr + 3 * QRC_RG) * qk + j], P[(lr + 3 * QRC_RG) * qk + c], s3);
                }
                for (; lr < nloc; lr += QRC_RG) s0 = fma(P[lr * qk + j], P[lr * qk + c], s0);
            }
            sh->red[rg][c] = (s0 + s1) + (s2 + s3);
        }
        __syncthreads();
        if (active) {                                                     // all-gather: every row group serves 2 of the 8 peers
            const double s = (sh->red[0][c] + sh->red[1][c]) + (sh->red[2][c] + sh->red[3][c]);
            const double pr = owner ? P[jl * qk + c] : 0.0;
#pragma unroll
            for (int k = 0; k < PEERS_PER_RG; k++) {
                peer[k]->part[par][rank][c] = s;
                if (owner) peer[k]->prow[par][c] = pr;
            }
        }
        cluster.sync();
        double tau, vs;
        if (factor) {
            double s_jj = 0.0;
#pragma unroll
            for (int k = 0; k < QRC_CLUSTER; k++) s_jj += sh->part[par][k][j];
            reflector_scalars(sh->prow[par][j], s_jj, tau, vs);
            if (threadIdx.x == j) { sh->tau[j] = tau; sh->vscale[j] = vs; }
        } else {
            tau = sh->tau[j];
            vs = sh->vscale[j];
        }
        if (active && c != j) {
            double d = 0.0;
#pragma unroll
            for (int k = 0; k < QRC_CLUSTER; k++) d += sh->part[par][k][c];
            const double w = tau * fma(vs, d, sh->prow[par][c]);           // tau * v^T a_c
            const double f = -w * vs;
            int lr = start;
            for (; lr + 3 * QRC_RG < nloc; lr += 4 * QRC_RG) {
                double* p0 = P + lr * qk;
                double* p1 = p0 + QRC_RG * qk;
                double* p2 = p1 + QRC_RG * qk;
                double* p3 = p2 + QRC_RG * qk;
                const double v0 = p0[j], v1 = p1[j], v2 = p2[j], v3 = p3[j];
                const double a0 = p0[c], a1 = p1[c], a2 = p2[c], a3 = p3[c];
                p0[c] = fma(f, v0, a0);
                p1[c] = fma(f, v1, a1);
                p2[c] = fma(f, v2, a2);
                p3[c] = fma(f, v3, a3);
            }
            for (; lr < nloc; lr += QRC_RG) P[lr * qk + c] = fma(f, P[lr * qk + j], P[lr * qk + c]);
            if (!factor && owner && rg == 0) P[jl * qk + c] -= w;           // (the factor phase never reads row j again)
        }
        __syncthreads();                                                  // column j+1 complete before it becomes the pivot
        if (!factor && c == j) {                                          // column j of Q = H_j e_j; only its own threads read it later
            const double g = -tau * vs;
            for (int lr = start; lr < nloc; lr += QRC_RG) P[lr * qk + j] *= g;
            if (owner && rg == (jl & (QRC_RG - 1))) P[jl * qk + j] = 1.0 - tau;   // by the thread that reads it next step
            const int hi = j > rank ? (j - rank + QRC_CLUSTER - 1) / QRC_CLUSTER : 0;       // local rows above the diagonal: zero
            for (int lr = rg; lr < hi && lr < nloc; lr += QRC_RG) P[lr * qk + j] = 0.0;
        }
    };

    int par = 0;
    // ---- factor: dgeqr2 on columns 0 .. kf-1 (completion columns are not touched) ------------------------------------------
    for (int j = 0; j < kf; ++j, par ^= 1) step(j, j, kf, true, par);
    // ---- form Q in place: dorg2r, reflectors last to first, on columns j+1 .. qk-1 ------------------------------------------
    for (int j = kf - 1; j >= 0; --j, par ^= 1) step(j, j + 1, qk, false, par);
    __syncthreads();

    // ---- write Q -----------------------------------------------------------------------------------------------------------------
    for (int idx = threadIdx.x; idx < nloc * qk; idx += QRC_THREADS) {
        int lr = idx / qk, cc = idx - lr * qk;
        Q[(int64_t)(lr * QRC_CLUSTER + rank) * q_rs + (int64_t)cc * q_cs] = P[lr * qk + cc];
    }
    cluster.sync();          // no CTA may exit while a peer can still write into its shared memory
}

// ---- panel-blocked variant ---------------------------------------------------------------------------------------------------------
// The kernel above streams the whole slice through shared memory three times per reflector and all-gathers 16 KB per CTA and step.
// Here reflectors are built 8 at a time on the 8 panel columns only (level-2 work on 1/32 of the matrix, 64-byte all-gathers, one
// cluster barrier per column), and the rest of the matrix sees the panel ONCE, as a block reflector I - V T V^T (compact WY,
// dlarft): partial V^T A per CTA -> reduce-scatter to the CTA owning each group of 32 columns -> T^T (or T) applied there ->
// all-gather of the 8 x 256 coefficients -> rank-8 update of the local rows (8 FMAs per element and pass instead of 1): two cluster
// barriers and 32 KB of DSMEM traffic per CTA per PANEL.  Q is formed the same way (dorgqr order), one block reflector per panel.
constexpr int QRB_PB = 8;                               // panel width
constexpr int QRB_MAXP = QRC_COLS / QRB_PB;             // panels
constexpr int QRB_OWN = QRC_COLS / QRC_CLUSTER;         // columns whose V^T A is reduced by one CTA
constexpr int QRB_MAXROWS = 64;                         // local rows: m <= 512 (the gathered panel must fit the exchange area)

constexpr int QRB_PANROWS = 512;                        // rows of a gathered panel (m <= 512)
struct __align__(16) QrbShared {
    // One 32 KB exchange area with three lives: (1) the whole current panel, all-gathered, column-major pan[k][r] (r = global row - j0);
    // (2) rs_part[src][k][cc]: reduce-scatter of the partial V^T A of my 32 columns from the 8 CTAs; (3) ag_tw[k][c]: all-gather of
    // op(T) V^T A for every column.  (2) and (3) are disjoint halves; cluster barriers separate (1) from them.
    double xch[QRB_PB * QRB_PANROWS];
    double Vp[QRB_MAXROWS][QRB_PB];                     // explicit V of the current panel (unit diagonal, zeros above, scaled below), my rows
    double T[QRB_MAXP][QRB_PB][QRB_PB];                 // triangular factors of all panels (kept for the form-Q phase)
    double Wsm[QRB_PB][QRB_OWN];
    double tau[QRC_COLS], vscale[QRC_COLS];
    double dsm[QRB_PB];
};
static_assert(QRB_PB * QRB_PANROWS == QRC_CLUSTER * QRB_PB * QRB_OWN + QRB_PB * QRC_COLS, "exchange area layout");
__device__ __forceinline__ double& qrb_rs_part(QrbShared* sh, int src, int k, int cc) { return sh->xch[(src * QRB_PB + k) * QRB_OWN + cc]; }
__device__ __forceinline__ double& qrb_ag_tw(QrbShared* sh, int k, int c) { return sh->xch[QRC_CLUSTER * QRB_PB * QRB_OWN + k * QRC_COLS + c]; }
__device__ __forceinline__ double& qrb_pan(QrbShared* sh, int k, int r) { return sh->xch[k * QRB_PANROWS + r]; }

__global__ void __cluster_dims__(QRC_CLUSTER, 1, 1) __launch_bounds__(QRC_THREADS, 1)
qr_cluster_blocked_kernel(const double* __restrict__ A, int64_t a_rs, int64_t a_cs, int64_t a_bs, int m, int kf, int qk,
                          double* __restrict__ Q, int64_t q_rs, int64_t q_cs, int64_t q_bs, int rows_per) {
    extern __shared__ __align__(16) unsigned char qrc_raw[];
    QrbShared* sh = reinterpret_cast<QrbShared*>(qrc_raw);
    double* P = reinterpret_cast<double*>(qrc_raw + sizeof(QrbShared));     // [rows_per][qk]
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c = tid % QRC_COLS;            // my column in the block passes
    const int rg = tid / QRC_COLS;           // my row group (update pass) / pair of panel columns (V^T A pass)
    const int prob = blockIdx.x / QRC_CLUSTER;
    A += (int64_t)prob * a_bs;
    Q += (int64_t)prob * q_bs;
    const int nloc = (m > rank) ? (m - rank + QRC_CLUSTER - 1) / QRC_CLUSTER : 0;   // local row lr <-> global row lr * 8 + rank

    for (int idx = tid; idx < nloc * qk; idx += QRC_THREADS) {
        int lr = idx / qk, cc = idx - lr * qk;
        int gr = lr * QRC_CLUSTER + rank;
        P[lr * qk + cc] = cc < kf ? A[(int64_t)gr * a_rs + (int64_t)cc * a_cs] : (gr == cc ? 1.0 : 0.0);
    }
    for (int idx = tid; idx < QRB_MAXP * QRB_PB * QRB_PB; idx += QRC_THREADS) (&sh->T[0][0][0])[idx] = 0.0;
    for (int idx = tid; idx < QRC_COLS; idx += QRC_THREADS) { sh->tau[idx] = 0.0; sh->vscale[idx] = 0.0; }
    cluster.sync();          // also: every CTA of the cluster is running before the first remote store

#ifdef SYN_QRC_TIMING
    long long tk[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // panel factorisation, write-back / build_vp, pass 1, barrier 1, owner, barrier 2, pass 2, panel gather
    long long tmark = clock64();
#define QTICK(i) do { if (tid == 0) { long long now_ = clock64(); tk[i] += now_ - tmark; tmark = now_; } } while (0)
#else
#define QTICK(i) do { } while (0)
#endif
    auto first_ge = [&](int g) { return g > rank ? (g - rank + QRC_CLUSTER - 1) / QRC_CLUSTER : 0; };        // ... >= g

    // Explicit V of panel [j0, j0 + pb); with `identity` the panel columns of P are replaced by unit columns (form Q).
    auto build_vp = [&](int j0, int pb, bool identity) {
        for (int idx = tid; idx < nloc * QRB_PB; idx += QRC_THREADS) {
            const int lr = idx / QRB_PB, k = idx - lr * QRB_PB;
            const int g = lr * QRC_CLUSTER + rank, jc = j0 + k;
            double v = 0.0;
            if (k < pb) {
                v = g < jc ? 0.0 : (g == jc ? 1.0 : sh->vscale[jc] * P[lr * qk + jc]);
                if (identity) P[lr * qk + jc] = (g == jc) ? 1.0 : 0.0;
            }
            sh->Vp[lr][k] = v;
        }
        __syncthreads();
        QTICK(1);
    };

    // columns [c0, c1)  <-  (I - V op(T) V^T) columns,  op(T) = T^T (factor: H^T A) or T (form Q: H Q)
    auto block_apply = [&](int pnl, int j0, int c0, int c1, bool transpose_t) {
        const int lo0 = first_ge(j0);
        const bool mine = (c >= c0 && c < c1);
        if (mine) {                                                     // pass 1: my column, panel columns 2 rg and 2 rg + 1, all local rows
            double s0 = 0.0, s1 = 0.0, t0 = 0.0, t1 = 0.0;
            int lr = lo0;
            for (; lr + 1 < nloc; lr += 2) {
                const double2 v0 = *reinterpret_cast<const double2*>(&sh->Vp[lr][2 * rg]);
                const double2 v1 = *reinterpret_cast<const double2*>(&sh->Vp[lr + 1][2 * rg]);
                const double p0 = P[lr * qk + c], p1 = P[(lr + 1) * qk + c];
                s0 = fma(v0.x, p0, s0); s1 = fma(v0.y, p0, s1);
                t0 = fma(v1.x, p1, t0); t1 = fma(v1.y, p1, t1);
            }
            if (lr < nloc) {
                const double2 v0 = *reinterpret_cast<const double2*>(&sh->Vp[lr][2 * rg]);
                const double p0 = P[lr * qk + c];
                s0 = fma(v0.x, p0, s0); s1 = fma(v0.y, p0, s1);
            }
            QrbShared* own = cluster.map_shared_rank(sh, c / QRB_OWN);
            qrb_rs_part(own, rank, 2 * rg, c % QRB_OWN) = s0 + t0;
            qrb_rs_part(own, rank, 2 * rg + 1, c % QRB_OWN) = s1 + t1;
        }
        QTICK(2);
        cluster.sync();
        QTICK(3);
        const int ok = tid / QRB_OWN, occ = tid % QRB_OWN, oc = rank * QRB_OWN + occ;     // owner role: (panel column, one of my 32 columns)
        const bool own_active = (tid < QRB_PB * QRB_OWN) && (oc >= c0 && oc < c1);
        if (tid < QRB_PB * QRB_OWN) {
            double w = 0.0;
            if (own_active) {
#pragma unroll
                for (int src = 0; src < QRC_CLUSTER; ++src) w += qrb_rs_part(sh, src, ok, occ);
            }
            sh->Wsm[ok][occ] = w;
        }
        __syncthreads();
        if (own_active) {
            double tw = 0.0;
#pragma unroll
            for (int l = 0; l < QRB_PB; ++l) tw = fma(transpose_t ? sh->T[pnl][l][ok] : sh->T[pnl][ok][l], sh->Wsm[l][occ], tw);
#pragma unroll
            for (int peer = 0; peer < QRC_CLUSTER; ++peer) qrb_ag_tw(cluster.map_shared_rank(sh, peer), ok, oc) = tw;
        }
        QTICK(4);
        cluster.sync();
        QTICK(5);
        if (mine) {                                                     // pass 2: my column, local rows lr = rg (mod 4)
            double tw[QRB_PB];
#pragma unroll
            for (int k = 0; k < QRB_PB; ++k) tw[k] = qrb_ag_tw(sh, k, c);
            for (int lr = lo0 + ((rg - lo0) & (QRC_RG - 1)); lr < nloc; lr += QRC_RG) {
                const double2* v = reinterpret_cast<const double2*>(&sh->Vp[lr][0]);
                const double2 v0 = v[0], v1 = v[1], v2 = v[2], v3 = v[3];
                double x = P[lr * qk + c], y = 0.0;
                x = fma(-v0.x, tw[0], x); y = fma(-v0.y, tw[1], y);
                x = fma(-v1.x, tw[2], x); y = fma(-v1.y, tw[3], y);
                x = fma(-v2.x, tw[4], x); y = fma(-v2.y, tw[5], y);
                x = fma(-v3.x, tw[6], x); y = fma(-v3.y, tw[7], y);
                P[lr * qk + c] = x + y;
            }
        }
        __syncthreads();
        QTICK(6);
    };

    // Panel [j0, j0 + pb): every CTA receives the whole panel (rows >= j0: at most 512 x 8 doubles) and factors it redundantly --
    // 8 warps, one per panel column, two named barriers per reflector and NO cluster traffic inside the panel -- then keeps its own
    // rows: unscaled reflector columns back into P, explicit V into Vp, T / tau / vscale in shared memory.
    // Gathered layout: pan[k][slot], slot = src * 64 + local row of src (contiguous per sender: coalesced DSMEM stores); the global
    // row of a slot is (slot & 63) * 8 + (slot >> 6).
    auto factor_panel = [&](int pnl, int j0, int pb) {
        const int lo0 = first_ge(j0), nrows = nloc - lo0;
        cluster.sync();                                                  // everybody is done with the exchange area (previous block update)
        for (int idx = tid; idx < nrows * QRB_PB * QRC_CLUSTER; idx += QRC_THREADS) {
            const int peer = idx / (nrows * QRB_PB), e = idx - peer * (nrows * QRB_PB);
            const int k = e / nrows, lr = lo0 + (e - k * nrows);
            qrb_pan(cluster.map_shared_rank(sh, peer), k, rank * QRB_MAXROWS + lr) = k < pb ? P[lr * qk + j0 + k] : 0.0;
        }
        cluster.sync();                                                  // the panel is complete everywhere
        QTICK(7);
        if (warp < QRB_PB) {
            const int k = warp;
            double* colk = &qrb_pan(sh, k, 0);
            for (int jj = 0; jj < pb; ++jj) {
                const int j = j0 + jj;
                const int pslot = (j & (QRC_CLUSTER - 1)) * QRB_MAXROWS + (j >> 3);          // slot of the pivot row
                const double* colj = &qrb_pan(sh, jj, 0);
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                auto below = [&](int slot) {
                    const int g = ((slot & (QRB_MAXROWS - 1)) << 3) + (slot >> 6);
                    return (g > j) && (g < m);
                };
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const int q0 = lane + 32 * i, q1 = q0 + 32, q2 = q0 + 64, q3 = q0 + 96;
                    const bool b0 = below(q0), b1 = below(q1), b2 = below(q2), b3 = below(q3);
                    const double a0 = b0 ? colj[q0] : 0.0, a1 = b1 ? colj[q1] : 0.0, a2 = b2 ? colj[q2] : 0.0, a3 = b3 ? colj[q3] : 0.0;
                    const double c0 = b0 ? colk[q0] : 0.0, c1 = b1 ? colk[q1] : 0.0, c2 = b2 ? colk[q2] : 0.0, c3 = b3 ? colk[q3] : 0.0;
                    s0 = fma(a0, c0, s0); s1 = fma(a1, c1, s1); s2 = fma(a2, c2, s2); s3 = fma(a3, c3, s3);
                }
                const double s = warp_sum((s0 + s1) + (s2 + s3));
                if (lane == 0) sh->dsm[k] = s;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                double tau, vs;
                reflector_scalars(colj[pslot], sh->dsm[jj], tau, vs);
                if (k > jj && k < pb) {                                  // a_k <- H_j a_k below the pivot row (R is not kept)
                    const double f = -tau * fma(vs, s, colk[pslot]) * vs;
#pragma unroll 4
                    for (int i = 0; i < 16; ++i) {
                        const int slot = lane + 32 * i;
                        if (below(slot)) colk[slot] = fma(f, colj[slot], colk[slot]);
                    }
                }
                if (k == 0) {                                            // column jj of T (dlarft, forward columnwise) and the scalars
                    const int l8 = lane & (QRB_PB - 1);
                    const double vsl = (l8 < jj) ? sh->vscale[j0 + l8] : 0.0;
                    const double zl = vsl * fma(vs, sh->dsm[l8], qrb_pan(sh, l8, pslot));  // V_l^T v_j
                    double tcol = 0.0;
                    for (int l = 0; l < jj; ++l) {
                        const double zz = __shfl_sync(0xffffffffu, zl, l);
                        if (l8 <= l) tcol = fma(sh->T[pnl][l8][l], zz, tcol);
                    }
                    if (lane < QRB_PB) {
                        if (l8 < jj) sh->T[pnl][l8][jj] = -tau * tcol;
                        else if (l8 == jj) { sh->T[pnl][jj][jj] = tau; sh->tau[j] = tau; sh->vscale[j] = vs; }
                    }
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
        }
        __syncthreads();
        QTICK(0);
        for (int idx = tid; idx < nrows * QRB_PB; idx += QRC_THREADS) {
            const int lr = lo0 + idx / QRB_PB, k = idx % QRB_PB;
            const int g = lr * QRC_CLUSTER + rank, jc = j0 + k;
            double v = 0.0;
            if (k < pb) {
                const double x = qrb_pan(sh, k, rank * QRB_MAXROWS + lr);
                P[lr * qk + jc] = x;
                v = g < jc ? 0.0 : (g == jc ? 1.0 : sh->vscale[jc] * x);
            }
            sh->Vp[lr][k] = v;
        }
        cluster.sync();                                                  // nobody reads the gathered panel any more: the area may be overwritten
        QTICK(1);
    };

    // ---- factor (columns 0 .. kf-1; completion columns are not touched) ---------------------------------------------------------
    for (int j0 = 0; j0 < kf; j0 += QRB_PB) {
        const int pb = (kf - j0) < QRB_PB ? (kf - j0) : QRB_PB, pnl = j0 / QRB_PB;
        factor_panel(pnl, j0, pb);
        if (j0 + pb < kf) block_apply(pnl, j0, j0 + pb, kf, true);
    }
    // ---- form Q in place: block reflectors last to first on columns j0 .. qk-1 ---------------------------------------------------
    for (int pnl = (kf - 1) / QRB_PB; pnl >= 0; --pnl) {
        const int j0 = pnl * QRB_PB, pb = (kf - j0) < QRB_PB ? (kf - j0) : QRB_PB;
        build_vp(j0, pb, true);
        block_apply(pnl, j0, j0, qk, false);
    }

    for (int idx = tid; idx < nloc * qk; idx += QRC_THREADS) {
        int lr = idx / qk, cc = idx - lr * qk;
        Q[(int64_t)(lr * QRC_CLUSTER + rank) * q_rs + (int64_t)cc * q_cs] = P[lr * qk + cc];
    }
#ifdef SYN_QRC_TIMING
    if (tid == 0 && blockIdx.x == 3)
        printf("qr_cluster_blocked m=%d kf=%d qk=%d cycles: panel gather %lld  panel factor %lld  build_vp %lld  pass1 %lld  barrier1 %lld  owner %lld  barrier2 %lld  pass2 %lld\n",
               m, kf, qk, tk[7], tk[0], tk[1], tk[2], tk[3], tk[4], tk[5], tk[6]);
#endif
    cluster.sync();          // no CTA may exit while a peer can still write into its shared memory
}

static int qr_env_cluster_mode() {   // SYN_QR_CLUSTER: 0 = blocked multi-launch path only, 1 = unblocked cluster kernel, 2 (default) = panel-blocked cluster kernel
    static int v = -1;
    if (v < 0) { const char* e = getenv("SYN_QR_CLUSTER"); v = e ? atoi(e) : 2; }
    return v;
}

// Can the cluster kernel take this problem?  (m x qk slice of m/8 rows must fit next to the exchange buffers.)
bool qr_cluster_fits(int m, int qk) {
    if (qk > QRC_COLS || m < QRC_CLUSTER) return false;
    int rows_per = (m + QRC_CLUSTER - 1) / QRC_CLUSTER;
    if (rows_per > QRB_MAXROWS) return false;
    const size_t fixed = qr_env_cluster_mode() == 1 ? sizeof(QrcShared) : sizeof(QrbShared);
    return fixed + (size_t)rows_per * qk * sizeof(double) <= QRC_SMEM_CAP;
}

int qr_cluster_form_q(const double* A, int64_t a_rs, int64_t a_cs, int64_t a_bs, int m, int kf, int qk, int batch, double* Q, int64_t q_rs,
                      int64_t q_cs, int64_t q_bs, cudaStream_t st) {
    const int rows_per = (m + QRC_CLUSTER - 1) / QRC_CLUSTER;
    const bool blocked = qr_env_cluster_mode() != 1;
    const size_t smem = (blocked ? sizeof(QrbShared) : sizeof(QrcShared)) + (size_t)rows_per * qk * sizeof(double);
    static PerDevice configured;
    const int dev_ = current_device();
    if (!configured.get(dev_)) {
        SYN_CUDA(cudaFuncSetAttribute(qr_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QRC_SMEM_CAP));
        SYN_CUDA(cudaFuncSetAttribute(qr_cluster_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QRC_SMEM_CAP));
        configured.set(dev_);
    }
    // gridDim.x must be a multiple of the (compile-time) cluster size; chunk large batches
    const int max_probs = 8192;
    for (int b0 = 0; b0 < batch; b0 += max_probs) {
        const int nb = (batch - b0) < max_probs ? (batch - b0) : max_probs;
        if (blocked)
            qr_cluster_blocked_kernel<<<nb * QRC_CLUSTER, QRC_THREADS, smem, st>>>(A + (int64_t)b0 * a_bs, a_rs, a_cs, a_bs, m, kf, qk,
                                                                                  Q + (int64_t)b0 * q_bs, q_rs, q_cs, q_bs, rows_per);
        else
            qr_cluster_kernel<<<nb * QRC_CLUSTER, QRC_THREADS, smem, st>>>(A + (int64_t)b0 * a_bs, a_rs, a_cs, a_bs, m, kf, qk,
                                                                          Q + (int64_t)b0 * q_bs, q_rs, q_cs, q_bs, rows_per);
        if (int rc = launch_status("qr_cluster_kernel")) return rc;
    }
    return 0;
}

}  // namespace syn

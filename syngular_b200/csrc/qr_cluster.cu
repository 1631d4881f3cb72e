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

#include "common.cuh"

namespace cg = cooperative_groups;

namespace syn {

constexpr int QRC_CLUSTER = 8;
constexpr int QRC_COLS = 256;         // one thread column per matrix column: q <= 256
constexpr int QRC_RG = 4;             // row groups: 4 threads share a column, each walks every 4th local row
constexpr int QRC_THREADS = QRC_COLS * QRC_RG;
constexpr size_t QRC_SMEM_CAP = 200 * 1024;
constexpr int QRC_GROUP = QRC_COLS / QRC_CLUSTER;      // columns whose dot products are summed by one CTA

struct QrcShared {                    // fixed-size part of the shared memory image (the row slice follows)
    double part[QRC_CLUSTER][QRC_COLS];      // partial dot products sent to the CTA that owns the column group
    double prow[QRC_COLS];                   // pivot-row elements of my column group (from the CTA that holds row j)
    double njj[QRC_CLUSTER];                 // partial squared norms of the pivot column (everybody needs the sum)
    double alpha;                            // pivot element P[j][j]
    double coef[QRC_COLS];                   // tau * v^T a_c for every column, gathered from the group owners
    double tau[QRC_COLS];
    double red[QRC_RG][QRC_COLS];            // intra-CTA combination of the row groups' partial sums
};

// Rows are dealt to the CTAs cyclically (global row g lives in CTA g % 8 at local index g / 8) so that the work of every
// CTA shrinks evenly as the factorisation advances; inside a CTA, 4 threads share a column (local rows lr % 4 == rg).
__global__ void __cluster_dims__(QRC_CLUSTER, 1, 1) __launch_bounds__(QRC_THREADS, 1)
qr_cluster_kernel(const double* __restrict__ A, int64_t a_rs, int64_t a_cs, int64_t a_bs, int m, int kf, int qk, double* __restrict__ Q,
                  int64_t q_rs, int64_t q_cs, int64_t q_bs, int rows_per) {
    extern __shared__ __align__(16) unsigned char qrc_raw[];
    QrcShared* sh = reinterpret_cast<QrcShared*>(qrc_raw);
    double* __restrict__ P = reinterpret_cast<double*>(qrc_raw + sizeof(QrcShared));     // [rows_per][qk]
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int c = threadIdx.x % QRC_COLS;            // my column
    const int rg = threadIdx.x / QRC_COLS;           // my row group
    const int grp = c / QRC_GROUP;                   // CTA that reduces my column's dot products
    const int prob = blockIdx.x / QRC_CLUSTER;
    A += (int64_t)prob * a_bs;
    Q += (int64_t)prob * q_bs;
    const int nloc = (m > rank) ? (m - rank + QRC_CLUSTER - 1) / QRC_CLUSTER : 0;   // local row lr <-> global row lr * 8 + rank

    // ---- load: columns < kf from A, completion columns as identity columns ---------------------------------------------
    for (int idx = threadIdx.x; idx < nloc * qk; idx += QRC_THREADS) {
        int lr = idx / qk, cc = idx - lr * qk;
        int gr = lr * QRC_CLUSTER + rank;
        P[lr * qk + cc] = cc < kf ? A[(int64_t)gr * a_rs + (int64_t)cc * a_cs] : (gr == cc ? 1.0 : 0.0);
    }
    __syncthreads();

    // first local row strictly below global row j
    auto first_below = [&](int j) { return j >= rank ? (j - rank) / QRC_CLUSTER + 1 : 0; };
    // first local row >= lo that belongs to my row group
    auto my_start = [&](int lo) { return lo + ((rg - lo) & (QRC_RG - 1)); };

    // partial dot of column j with my column over my rows below the diagonal, combined over the 4 row groups (result in rg 0)
    auto column_dot = [&](int j, bool active) -> double {
        double s0 = 0.0, s1 = 0.0;
        if (active) {
            int lr = my_start(first_below(j));
            for (; lr + QRC_RG < nloc; lr += 2 * QRC_RG) {
                s0 = fma(P[lr * qk + j], P[lr * qk + c], s0);
                s1 = fma(P[(lr + QRC_RG) * qk + j], P[(lr + QRC_RG) * qk + c], s1);
            }
            if (lr < nloc) s0 = fma(P[lr * qk + j], P[lr * qk + c], s0);
        }
        sh->red[rg][c] = s0 + s1;
        __syncthreads();
        return (sh->red[0][c] + sh->red[1][c]) + (sh->red[2][c] + sh->red[3][c]);
    };
    // Reduce-scatter: the partial goes to the CTA that owns the column group; the pivot-column norm and the pivot element go to
    // everybody; the CTA holding row j sends the pivot-row elements to the group owners.  (~4 KB of DSMEM stores per CTA.)
    auto scatter = [&](int j, bool active, double s, bool with_norm) {
        if (rg != 0) return;
        const bool owner = (j % QRC_CLUSTER) == rank;
        const int jl = j / QRC_CLUSTER;
        QrcShared* dst = cluster.map_shared_rank(sh, grp);
        if (active) dst->part[rank][c] = s;
        if (owner && active) dst->prow[c] = P[jl * qk + c];
        if (c == j) {
            for (int k = 0; k < QRC_CLUSTER; k++) {
                QrcShared* pk = cluster.map_shared_rank(sh, k);
                if (with_norm) pk->njj[rank] = s;
                if (owner) pk->alpha = P[jl * qk + j];
            }
        }
    };
    auto gather = [&](int c_, double tw) {
        for (int k = 0; k < QRC_CLUSTER; k++) cluster.map_shared_rank(sh, k)->coef[c_] = tw;
    };

    // ---- factor: dgeqr2 on columns 0 .. kf-1 (completion columns are not touched) ------------------------------------------
    for (int j = 0; j < kf; ++j) {
        const bool active = (c >= j && c < kf);
        const double s = column_dot(j, active);
        scatter(j, active, s, true);
        cluster.sync();
        double s_jj = 0.0;
#pragma unroll
        for (int k = 0; k < QRC_CLUSTER; k++) s_jj += sh->njj[k];
        const double alpha = sh->alpha;
        double beta, tau, scale;
        if (s_jj == 0.0) {
            tau = 0.0; beta = alpha; scale = 0.0;
        } else {
            const double nrm = sqrt(fma(alpha, alpha, s_jj));
            beta = alpha >= 0.0 ? -nrm : nrm;
            tau = (beta - alpha) / beta;
            scale = 1.0 / (alpha - beta);
        }
        if (rg == 0 && grp == rank && c > j && c < kf) {                 // all-gather of tau * v^T a_c   (v_j = 1)
            double d = 0.0;
#pragma unroll
            for (int k = 0; k < QRC_CLUSTER; k++) d += sh->part[k][c];
            gather(c, tau * fma(scale, d, sh->prow[c]));
        }
        cluster.sync();
        const bool owner = (j % QRC_CLUSTER) == rank;
        const int jl = j / QRC_CLUSTER;
        const int lo = first_below(j);
        if (c > j && c < kf) {
            const double tw = sh->coef[c];
            const double f = tw * scale;
            for (int lr = my_start(lo); lr < nloc; lr += QRC_RG) P[lr * qk + c] = fma(-f, P[lr * qk + j], P[lr * qk + c]);
            if (owner && rg == 0) P[jl * qk + c] -= tw;
        }
        if (threadIdx.x == j) sh->tau[j] = tau;
        __syncthreads();                                                 // column j is read by every thread above
        if (c == j) {
            for (int lr = my_start(lo); lr < nloc; lr += QRC_RG) P[lr * qk + j] *= scale;      // store v_j below the diagonal
            if (owner && rg == 0) P[jl * qk + j] = beta;
        }
        __syncthreads();
    }

    // ---- form Q in place: dorg2r, reflectors last to first, on columns j .. qk-1 ----------------------------------------------
    for (int j = kf - 1; j >= 0; --j) {
        const bool active = (c > j && c < qk);
        const double s = column_dot(j, active);
        scatter(j, active, s, false);
        cluster.sync();
        const double tau = sh->tau[j];
        if (rg == 0 && grp == rank && active) {
            double d = 0.0;
#pragma unroll
            for (int k = 0; k < QRC_CLUSTER; k++) d += sh->part[k][c];
            gather(c, tau * (sh->prow[c] + d));                          // tau * v^T q_c   (v stored, v_j = 1)
        }
        cluster.sync();
        const bool owner = (j % QRC_CLUSTER) == rank;
        const int jl = j / QRC_CLUSTER;
        const int lo = first_below(j);
        if (active) {
            const double tw = sh->coef[c];
            for (int lr = my_start(lo); lr < nloc; lr += QRC_RG) P[lr * qk + c] = fma(-tw, P[lr * qk + j], P[lr * qk + c]);
            if (owner && rg == 0) P[jl * qk + c] -= tw;
        }
        __syncthreads();
        if (c == j) {                                                    // column j of Q: H_j e_j
            for (int lr = my_start(lo); lr < nloc; lr += QRC_RG) P[lr * qk + j] *= -tau;
            if (owner && rg == 0) P[jl * qk + j] = 1.0 - tau;
            const int hi = j > rank ? (j - rank + QRC_CLUSTER - 1) / QRC_CLUSTER : 0;       // local rows above the diagonal: zero
            for (int lr = rg; lr < hi && lr < nloc; lr += QRC_RG) P[lr * qk + j] = 0.0;
        }
        __syncthreads();
    }

    // ---- write Q -----------------------------------------------------------------------------------------------------------------
    for (int idx = threadIdx.x; idx < nloc * qk; idx += QRC_THREADS) {
        int lr = idx / qk, cc = idx - lr * qk;
        Q[(int64_t)(lr * QRC_CLUSTER + rank) * q_rs + (int64_t)cc * q_cs] = P[lr * qk + cc];
    }
    cluster.sync();          // no CTA may exit while a peer can still write into its shared memory
}

// Can the cluster kernel take this problem?  (m x qk slice of m/8 rows must fit next to the exchange buffers.)
bool qr_cluster_fits(int m, int qk) {
    if (qk > QRC_COLS || m < QRC_CLUSTER) return false;
    int rows_per = (m + QRC_CLUSTER - 1) / QRC_CLUSTER;
    return sizeof(QrcShared) + (size_t)rows_per * qk * sizeof(double) <= QRC_SMEM_CAP;
}

int qr_cluster_form_q(const double* A, int64_t a_rs, int64_t a_cs, int64_t a_bs, int m, int kf, int qk, int batch, double* Q, int64_t q_rs,
                      int64_t q_cs, int64_t q_bs, cudaStream_t st) {
    const int rows_per = (m + QRC_CLUSTER - 1) / QRC_CLUSTER;
    const size_t smem = sizeof(QrcShared) + (size_t)rows_per * qk * sizeof(double);
    static bool configured = false;
    if (!configured) {
        SYN_CUDA(cudaFuncSetAttribute(qr_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QRC_SMEM_CAP));
        configured = true;
    }
    // gridDim.x must be a multiple of the (compile-time) cluster size; chunk large batches
    const int max_probs = 8192;
    for (int b0 = 0; b0 < batch; b0 += max_probs) {
        const int nb = (batch - b0) < max_probs ? (batch - b0) : max_probs;
        qr_cluster_kernel<<<nb * QRC_CLUSTER, QRC_THREADS, smem, st>>>(A + (int64_t)b0 * a_bs, a_rs, a_cs, a_bs, m, kf, qk,
                                                                      Q + (int64_t)b0 * q_bs, q_rs, q_cs, q_bs, rows_per);
        if (int rc = launch_status("qr_cluster_kernel")) return rc;
    }
    return 0;
}

}  // namespace syn

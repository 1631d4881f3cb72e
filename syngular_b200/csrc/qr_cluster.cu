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

// Householder QR for the bond-rounding / canonicalisation sweeps (FP64, sm_100a), batched.
//
// One primitive covers every QR call site of the reference:
//
//   qrt(A (m x n), q)  ->  Q (m x qk), S = Q^T A (qk x n),  qk = min(q, m)
//
//   * `>>` / decompose truncation step: np.linalg.qr(L, mode="complete"); Q[:, :q]; R[:q, :]
//       (matrix_product_state.py:307-311,443-446; matrix_product_operator.py:436-440,555-558).  Only the first
//       min(q, n) columns need factoring: the kept Q spans them, and R[:q, :] == Q^T L.  When q > n (the
//       reference's bond "inflation") the extra columns are the orthonormal completion the same reflectors give.
//       The reference builds the full m x m Q; we never do.
//   * reduced QR of the canonicalisation sweeps: np.linalg.qr(L) / np.linalg.qr(R.T)
//       (matrix_product_state.py:560,574; matrix_product_operator.py:680,703) == qrt(A, min(m, n)).
//
// Algorithm: blocked Householder, compact-WY.  A panel (<= 32 columns) is factored by ONE CTA with the panel
// resident in shared memory, one warp per column: each step needs a single __syncthreads because every warp
// derives the reflector (norm, beta, tau) redundantly from the pivot column while it accumulates its own dot
// product with it (warp-shuffle reductions).  The panel kernel also emits V (explicit) and V*T, so trailing
// updates and the formation of Q are two strided DMMA GEMMs per panel (gemm.cu).  Inputs/outputs take arbitrary
// row/column strides, so transposed unfoldings (R^T) and core layouts are consumed and produced in place.
#include <climits>
#include <cstdlib>

#include "common.cuh"

namespace syn {

int gemm_f64(const syn_gemm_desc_t& d, const double* A, const double* B, double* C, cudaStream_t st);
bool qr_cluster_fits(int m, int qk);
int qr_cluster_form_q(const double* A, int64_t a_rs, int64_t a_cs, int64_t a_bs, int m, int kf, int qk, int batch, double* Q, int64_t q_rs,
                      int64_t q_cs, int64_t q_bs, cudaStream_t st);

static bool qr_env_cluster() {   // SYN_QR_CLUSTER=0 forces the blocked multi-launch path (A/B comparisons)
    static int v = -1;
    if (v < 0) { const char* e = getenv("SYN_QR_CLUSTER"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

static inline syn_index_t IX(int64_t stride) {
    syn_index_t i;
    i.outer = 0;
    i.inner = stride;
    i.div = INT_MAX;
    i._pad = 0;
    return i;
}

static syn_gemm_desc_t mk_desc(int M, int N, int K, int batch, int64_t a_m, int64_t a_k, int64_t a_b, int64_t b_k, int64_t b_n,
                               int64_t b_b, int64_t c_m, int64_t c_n, int64_t c_b, double alpha, double beta) {
    syn_gemm_desc_t d = {};
    d.M = M; d.N = N; d.K = K; d.batch = batch;
    d.a_m = IX(a_m); d.a_k = IX(a_k); d.a_b = IX(a_b);
    d.b_k = IX(b_k); d.b_n = IX(b_n); d.b_b = IX(b_b);
    d.c_m = IX(c_m); d.c_n = IX(c_n); d.c_b = IX(c_b);
    d.alpha = alpha; d.beta = beta;
    return d;
}

// dst[b][i][j] (row-major, ld = n) = src[b*s_bs + i*s_rs + j*s_cs]; 32x32 tiles through shared memory so that both
// sides are coalesced whichever of s_rs / s_cs is the unit stride.
__global__ void copy_strided_kernel(const double* __restrict__ src, int64_t s_rs, int64_t s_cs, int64_t s_bs, double* __restrict__ dst,
                                    int64_t d_bs, int m, int n) {
    __shared__ double tile[32][33];
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    src += (int64_t)blockIdx.z * s_bs;
    dst += (int64_t)blockIdx.z * d_bs;
    const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
    if (s_cs <= s_rs) {   // columns contiguous (or equal): read with tx along j
        for (int r = ty; r < 32; r += 8) {
            int i = i0 + r, j = j0 + tx;
            if (i < m && j < n) tile[r][tx] = src[i * s_rs + j * s_cs];
        }
    } else {              // rows contiguous: read with tx along i
        for (int c = ty; c < 32; c += 8) {
            int i = i0 + tx, j = j0 + c;
            if (i < m && j < n) tile[tx][c] = src[i * s_rs + j * s_cs];
        }
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        int i = i0 + r, j = j0 + tx;
        if (i < m && j < n) dst[(int64_t)i * n + j] = tile[r][tx];
    }
}

__global__ void set_identity_kernel(double* __restrict__ Q, int64_t q_rs, int64_t q_cs, int64_t q_bs, int m, int q) {
    Q += (int64_t)blockIdx.z * q_bs;
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i < m && j < q) Q[i * q_rs + j * q_cs] = (i == j) ? 1.0 : 0.0;
}

constexpr int PANEL_THREADS = 1024;
constexpr int PANEL_MAX_NB = 32;

// Factor the panel P[j0:m, j0:j0+nbp] (row-major, leading dim ldp) of every batch member.
//   P  : overwritten LAPACK-style (R on/above the diagonal, Householder vectors below)
//   V  : explicit reflector block (unit diagonal, zeros above) at the same position of a (m x ldp) array
//   VT : V * T (compact-WY triangular factor folded in)
__global__ void __launch_bounds__(PANEL_THREADS, 1)
house_panel_kernel(double* __restrict__ P, double* __restrict__ V, double* __restrict__ VT, int64_t bs, int ldp, int m, int j0, int nbp,
                   int LD, double* __restrict__ gpan, int64_t gpan_bs) {
    extern __shared__ __align__(16) double dyn[];
    __shared__ double s_tau[PANEL_MAX_NB], s_beta[PANEL_MAX_NB], s_scale[PANEL_MAX_NB];
    __shared__ double s_T[PANEL_MAX_NB][PANEL_MAX_NB + 1];
    __shared__ double s_G[PANEL_MAX_NB][PANEL_MAX_NB + 1];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int mp = m - j0;
    P += (int64_t)blockIdx.x * bs;
    V += (int64_t)blockIdx.x * bs;
    VT += (int64_t)blockIdx.x * bs;
    double* pan = gpan ? (gpan + (int64_t)blockIdx.x * gpan_bs) : dyn;   // [nbp][LD], column-major panel

    for (int idx = tid; idx < mp * nbp; idx += PANEL_THREADS) {
        int r = idx / nbp, c = idx - r * nbp;
        pan[c * LD + r] = P[(int64_t)(j0 + r) * ldp + j0 + c];
    }
    for (int idx = tid; idx < PANEL_MAX_NB * (PANEL_MAX_NB + 1); idx += PANEL_THREADS) {
        (&s_T[0][0])[idx] = 0.0;
        (&s_G[0][0])[idx] = 0.0;
    }
    __syncthreads();

    for (int j = 0; j < nbp; ++j) {
        if (warp >= j && warp < nbp) {
            const double* pj = pan + j * LD;
            double* pc = pan + warp * LD;
            double s_jj = 0.0, s_jc = 0.0;
            for (int i = j + 1 + lane; i < mp; i += 32) {
                double x = pj[i];
                s_jj = fma(x, x, s_jj);
                s_jc = fma(x, pc[i], s_jc);
            }
            s_jj = warp_sum(s_jj);
            s_jc = warp_sum(s_jc);
            const double alpha = pj[j];
            double beta, tau, scale;
            if (s_jj == 0.0) {          // dlarfg: H = I
                tau = 0.0; beta = alpha; scale = 0.0;
            } else {
                double nrm = sqrt(fma(alpha, alpha, s_jj));
                beta = alpha >= 0.0 ? -nrm : nrm;
                tau = (beta - alpha) / beta;
                scale = 1.0 / (alpha - beta);
            }
            if (warp == j) {
                if (lane == 0) { s_tau[j] = tau; s_beta[j] = beta; s_scale[j] = scale; }
            } else {
                const double cj = pc[j];
                __syncwarp();
                const double tw = tau * fma(scale, s_jc, cj);       // tau * (v^T a_c), v_j = 1
                if (lane == 0) pc[j] = cj - tw;
                const double f = tw * scale;
                for (int i = j + 1 + lane; i < mp; i += 32) pc[i] = fma(-f, pj[i], pc[i]);
            }
        }
        __syncthreads();
        if (warp == j) {                // now nobody reads column j any more in this loop: store v and beta
            const double sc = s_scale[j];
            double* pj = pan + j * LD;
            for (int i = j + 1 + lane; i < mp; i += 32) pj[i] *= sc;
            if (lane == 0) pj[j] = s_beta[j];
        }
    }
    __syncthreads();

    // Gram of the reflectors: G[i][j] = v_i . v_j  (i < j)
    for (int j = 1; j < nbp; ++j) {
        for (int i = 0; i < j; ++i) {
            int idx = j * (j - 1) / 2 + i;
            if ((idx & 31) != warp) continue;
            const double* pi = pan + i * LD;
            const double* pj = pan + j * LD;
            double s = 0.0;
            for (int r = j + 1 + lane; r < mp; r += 32) s = fma(pi[r], pj[r], s);
            s = warp_sum(s);
            if (lane == 0) s_G[i][j] = s + (j < mp ? pi[j] : 0.0);
        }
    }
    __syncthreads();
    // T (upper triangular): T[j][j] = tau_j, T[0:j, j] = -tau_j * T[0:j, 0:j] * G[0:j, j]; lane i owns row i
    if (warp == 0) {
        for (int j = 0; j < nbp; ++j) {
            const double tj = s_tau[j];
            if (lane < j) {
                double acc = 0.0;
                for (int k = lane; k < j; ++k) acc = fma(s_T[lane][k], s_G[k][j], acc);
                s_T[lane][j] = -tj * acc;
            } else if (lane == j) {
                s_T[j][j] = tj;
            }
            __syncwarp();
        }
    }
    __syncthreads();

    // write back: P (LAPACK layout), V (explicit), VT = V * T
    for (int idx = tid; idx < mp * nbp; idx += PANEL_THREADS) {
        int r = idx / nbp, c = idx - r * nbp;
        int64_t o = (int64_t)(j0 + r) * ldp + j0 + c;
        double x = pan[c * LD + r];
        P[o] = x;
        V[o] = r > c ? x : (r == c ? 1.0 : 0.0);
        double acc = 0.0;
        int kmax = c < r ? c : r;      // V[r][k] != 0 only for k <= r ; T[k][c] != 0 only for k <= c
        for (int k = 0; k <= kmax; ++k) {
            double vrk = (k == r) ? 1.0 : pan[k * LD + r];
            acc = fma(vrk, s_T[k][c], acc);
        }
        VT[o] = acc;
    }
}

static inline int64_t even_up(int64_t x) { return (x + 1) & ~(int64_t)1; }

size_t qrt_workspace_doubles(int m, int n, int q, int batch) {
    int qk = q < m ? q : m;
    int kf = qk < n ? qk : n;
    int64_t LD = ((m + 15) / 16) * 16 + 1;
    int64_t wcols = kf > qk ? kf : qk;
    return (size_t)(3 * even_up((int64_t)m * kf * batch) + even_up((int64_t)PANEL_MAX_NB * wcols * batch) +
                    (int64_t)PANEL_MAX_NB * LD * batch);
}

struct QrPlan {
    int m, n, qk, kf, batch, nb, wcols;
    int64_t LD, pbs, wbs;
    bool use_global;
    size_t smem;
    double *P, *V, *VT, *W, *G;
};

static int qr_plan(QrPlan& p, int m, int n, int q, int batch, double* ws, size_t ws_doubles, const char* who) {
    SYN_REQUIRE(m >= 1 && n >= 1 && q >= 1 && batch >= 1, "%s: bad extents m=%d n=%d q=%d batch=%d", who, m, n, q, batch);
    p.m = m; p.n = n; p.batch = batch;
    p.qk = q < m ? q : m;
    p.kf = p.qk < n ? p.qk : n;
    SYN_REQUIRE(ws_doubles >= qrt_workspace_doubles(m, n, q, batch), "%s: workspace too small (%zu < %zu doubles)", who, ws_doubles,
                qrt_workspace_doubles(m, n, q, batch));
    SYN_REQUIRE((((uintptr_t)ws) & 15) == 0, "%s: workspace must be 16-byte aligned", who);
    p.LD = ((m + 15) / 16) * 16 + 1;
    p.wcols = p.kf > p.qk ? p.kf : p.qk;
    p.pbs = (int64_t)m * p.kf;
    p.wbs = (int64_t)PANEL_MAX_NB * p.wcols;
    p.P = ws;
    p.V = p.P + even_up(p.pbs * batch);
    p.VT = p.V + even_up(p.pbs * batch);
    p.W = p.VT + even_up(p.pbs * batch);
    p.G = p.W + even_up(p.wbs * batch);
    const size_t smem_cap = 200 * 1024;
    int nb = PANEL_MAX_NB;
    while (nb > 4 && (size_t)nb * p.LD * sizeof(double) > smem_cap) nb >>= 1;
    p.use_global = (size_t)nb * p.LD * sizeof(double) > smem_cap;
    if (p.use_global) nb = 16;
    p.nb = nb;
    p.smem = p.use_global ? 0 : (size_t)nb * p.LD * sizeof(double);
    static PerDevice configured;
    const int dev_ = current_device();
    if (!configured.get(dev_)) {
        SYN_CUDA(cudaFuncSetAttribute(house_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
        configured.set(dev_);
    }
    return 0;
}

// steps 1-2: P = A[:, :kf], blocked Householder factorisation of P (V, VT emitted per panel)
static int house_factor(const QrPlan& p, const double* A, int64_t a_rs, int64_t a_cs, int64_t a_bs, cudaStream_t st) {
    const int m = p.m, kf = p.kf, batch = p.batch, nb = p.nb;
    {
        dim3 grid((kf + 31) / 32, (m + 31) / 32, batch), block(32, 8);
        copy_strided_kernel<<<grid, block, 0, st>>>(A, a_rs, a_cs, a_bs, p.P, p.pbs, m, kf);
        if (int rc = launch_status("copy_strided_kernel")) return rc;
    }
    for (int j0 = 0; j0 < kf; j0 += nb) {
        const int nbp = (kf - j0) < nb ? (kf - j0) : nb;
        house_panel_kernel<<<batch, PANEL_THREADS, p.smem, st>>>(p.P, p.V, p.VT, p.pbs, kf, m, j0, nbp, (int)p.LD,
                                                                 p.use_global ? p.G : nullptr, (int64_t)PANEL_MAX_NB * p.LD);
        if (int rc = launch_status("house_panel_kernel")) return rc;
        const int j1 = j0 + nbp, nc = kf - j1, mp = m - j0;
        if (nc > 0) {
            // W (nbp x nc) = (V T)^T P[j0:, j1:]
            syn_gemm_desc_t d1 = mk_desc(nbp, nc, mp, batch, 1, kf, p.pbs, kf, 1, p.pbs, p.wcols, 1, p.wbs, 1.0, 0.0);
            if (int rc = gemm_f64(d1, p.VT + (int64_t)j0 * kf + j0, p.P + (int64_t)j0 * kf + j1, p.W, st)) return rc;
            // P[j0:, j1:] -= V W
            syn_gemm_desc_t d2 = mk_desc(mp, nc, nbp, batch, kf, 1, p.pbs, p.wcols, 1, p.wbs, kf, 1, p.pbs, -1.0, 1.0);
            if (int rc = gemm_f64(d2, p.V + (int64_t)j0 * kf + j0, p.W, p.P + (int64_t)j0 * kf + j1, st)) return rc;
        }
    }
    return 0;
}

// A: (a_rs, a_cs, a_bs) strides; Q: m x qk; S: qk x n (may be null).  Returns qk via *qk_out.
int qrt_f64(const double* A, int64_t a_rs, int64_t a_cs, int64_t a_bs, int m, int n, int q, int batch, double* Q, int64_t q_rs,
            int64_t q_cs, int64_t q_bs, double* S, int64_t s_rs, int64_t s_cs, int64_t s_bs, double* ws, size_t ws_doubles,
            int* qk_out, cudaStream_t st) {
    QrPlan p;
    if (int rc = qr_plan(p, m, n, q, batch, ws, ws_doubles, "syn_qrt_f64")) return rc;
    if (qk_out) *qk_out = p.qk;
    SYN_REQUIRE(A && Q, "syn_qrt_f64: null operand");
    if (qr_env_cluster() && qr_cluster_fits(m, p.qk)) {
        // plateau shapes: factor + form Q in ONE cluster-resident launch (qr_cluster.cu), then S = Q^T A
        if (int rc = qr_cluster_form_q(A, a_rs, a_cs, a_bs, m, p.kf, p.qk, batch, Q, q_rs, q_cs, q_bs, st)) return rc;
        if (S) {
            syn_gemm_desc_t d = mk_desc(p.qk, n, m, batch, q_cs, q_rs, q_bs, a_rs, a_cs, a_bs, s_rs, s_cs, s_bs, 1.0, 0.0);
            if (int rc = gemm_f64(d, Q, A, S, st)) return rc;
        }
        return 0;
    }
    if (int rc = house_factor(p, A, a_rs, a_cs, a_bs, st)) return rc;
    const int qk = p.qk, kf = p.kf, nb = p.nb;
    {   // 3. Q = first qk columns of the identity
        dim3 block(32, 8), grid((qk + 31) / 32, (m + 7) / 8, batch);
        set_identity_kernel<<<grid, block, 0, st>>>(Q, q_rs, q_cs, q_bs, m, qk);
        if (int rc = launch_status("set_identity_kernel")) return rc;
    }
    // 4. Q = H_1 ... H_k E : apply the block reflectors last to first; block p only touches Q[j0:, j0:]
    const int last0 = ((kf - 1) / nb) * nb;
    for (int j0 = last0; j0 >= 0; j0 -= nb) {
        const int nbp = (kf - j0) < nb ? (kf - j0) : nb;
        const int mp = m - j0, nc = qk - j0;
        double* Qs = Q + (int64_t)j0 * q_rs + (int64_t)j0 * q_cs;
        // W (nbp x nc) = V^T Q[j0:, j0:]
        syn_gemm_desc_t d1 = mk_desc(nbp, nc, mp, batch, 1, kf, p.pbs, q_rs, q_cs, q_bs, p.wcols, 1, p.wbs, 1.0, 0.0);
        if (int rc = gemm_f64(d1, p.V + (int64_t)j0 * kf + j0, Qs, p.W, st)) return rc;
        // Q[j0:, j0:] -= (V T) W
        syn_gemm_desc_t d2 = mk_desc(mp, nc, nbp, batch, kf, 1, p.pbs, p.wcols, 1, p.wbs, q_rs, q_cs, q_bs, -1.0, 1.0);
        if (int rc = gemm_f64(d2, p.VT + (int64_t)j0 * kf + j0, p.W, Qs, st)) return rc;
    }
    // 5. S = Q^T A
    if (S) {
        syn_gemm_desc_t d = mk_desc(qk, n, m, batch, q_cs, q_rs, q_bs, a_rs, a_cs, a_bs, s_rs, s_cs, s_bs, 1.0, 0.0);
        if (int rc = gemm_f64(d, Q, A, S, st)) return rc;
    }
    return 0;
}

__global__ void triu_extract_kernel(const double* __restrict__ P, int64_t pbs, int ldp, double* __restrict__ R, int64_t r_rs, int64_t r_cs,
                                    int64_t r_bs, int k) {
    P += (int64_t)blockIdx.z * pbs;
    R += (int64_t)blockIdx.z * r_bs;
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i < k && j < k) R[i * r_rs + j * r_cs] = (j >= i) ? P[(int64_t)i * ldp + j] : 0.0;
}

// R factor only (k x k upper triangular, k = min(m, n)) of the first k columns of A; used to reduce a tall
// unfolding before the Jacobi SVD.  R may be written with any strides (e.g. transposed).
int qr_r_f64(const double* A, int64_t a_rs, int64_t a_cs, int64_t a_bs, int m, int n, int batch, double* R, int64_t r_rs, int64_t r_cs,
             int64_t r_bs, double* ws, size_t ws_doubles, cudaStream_t st) {
    QrPlan p;
    int k = m < n ? m : n;
    if (int rc = qr_plan(p, m, n, k, batch, ws, ws_doubles, "syn_qr_r_f64")) return rc;
    SYN_REQUIRE(A && R, "syn_qr_r_f64: null operand");
    SYN_REQUIRE(n <= m, "syn_qr_r_f64: needs m >= n (got %d x %d)", m, n);
    if (int rc = house_factor(p, A, a_rs, a_cs, a_bs, st)) return rc;
    dim3 block(32, 8), grid((k + 31) / 32, (k + 7) / 8, batch);
    triu_extract_kernel<<<grid, block, 0, st>>>(p.P, p.pbs, p.kf, R, r_rs, r_cs, r_bs, k);
    return launch_status("triu_extract_kernel");
}

int copy_strided(const double* src, int64_t s_rs, int64_t s_cs, int64_t s_bs, double* dst, int64_t d_bs, int m, int n, int batch,
                 cudaStream_t st) {
    if (m <= 0 || n <= 0 || batch <= 0) return 0;
    dim3 grid((n + 31) / 32, (m + 31) / 32, batch), block(32, 8);
    copy_strided_kernel<<<grid, block, 0, st>>>(src, s_rs, s_cs, s_bs, dst, d_bs, m, n);
    return launch_status("copy_strided_kernel");
}

}  // namespace syn

extern "C" size_t syn_qrt_workspace_f64(int m, int n, int q, int batch) {
    return syn::qrt_workspace_doubles(m, n, q, batch) * sizeof(double);
}

extern "C" int syn_qrt_f64(const double* A, int64_t a_rs, int64_t a_cs, int64_t a_bs, int m, int n, int q, int batch, double* Q,
                           int64_t q_rs, int64_t q_cs, int64_t q_bs, double* S, int64_t s_rs, int64_t s_cs, int64_t s_bs, void* ws,
                           size_t ws_bytes, int* qk_out, void* stream) {
    return syn::qrt_f64(A, a_rs, a_cs, a_bs, m, n, q, batch, Q, q_rs, q_cs, q_bs, S, s_rs, s_cs, s_bs, (double*)ws,
                        ws_bytes / sizeof(double), qk_out, (cudaStream_t)stream);
}

extern "C" int syn_qr_r_f64(const double* A, int64_t a_rs, int64_t a_cs, int64_t a_bs, int m, int n, int batch, double* R, int64_t r_rs,
                            int64_t r_cs, int64_t r_bs, void* ws, size_t ws_bytes, void* stream) {
    return syn::qr_r_f64(A, a_rs, a_cs, a_bs, m, n, batch, R, r_rs, r_cs, r_bs, (double*)ws, ws_bytes / sizeof(double),
                         (cudaStream_t)stream);
}

extern "C" int syn_copy_strided_f64(const double* src, int64_t s_rs, int64_t s_cs, int64_t s_bs, double* dst, int64_t d_bs, int m, int n,
                                    int batch, void* stream) {
    return syn::copy_strided(src, s_rs, s_cs, s_bs, dst, d_bs, m, n, batch, (cudaStream_t)stream);
}

/* syngular_b200 -- C ABI of the sm_100a matrix-product library (libsyngular_b200.so).
 *
 * The reference (antoine311200/Syngular) is pure Python over numpy/LAPACK and has no FFI of its own; the
 * functions below are what a binding for its hot path binds to.  Each entry cites the reference call site it
 * replaces (paths relative to the reference root; "MPS" = tensor/matrix_product_state.py, "MPO" =
 * tensor/matrix_product_operator.py).  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; syn_last_error() gives the message.  Nothing
 *     throws across the ABI.
 *   - all pointers are DEVICE pointers to float64 unless stated; buffers are owned by the caller (PyTorch);
 *     the library never allocates or frees device memory and keeps no state between calls.
 *   - all work is ordered on the `stream` argument (a cudaStream_t passed as void*).
 *   - matrices are C-ordered (row-major); MPS core (l,d,r), MPO core (l,i,o,r) exactly as in the reference.
 */
#ifndef SYNGULAR_B200_H
#define SYNGULAR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- lifecycle --------------------------------------------------------------------------------------- */
int         syn_version(void);
const char* syn_last_error(void);
int         syn_device_sm_count(void);
long long   syn_launch_count(void);      /* kernels launched by the library so far (this process) */

/* ---- strided tensor-contraction GEMM (FP64 DMMA) ------------------------------------------------------ */
/* C[m,n] = alpha * sum_k A[m,k] B[k,n] + beta * C[m,n], batched.  Every logical index may be split in two
 * levels so that a contraction over reshaped/permuted cores needs no transposition copy:
 *    offset(x) = (x / div) * outer + (x % div) * inner            (div >= extent  ->  plain stride `inner`)
 * Replaces: np.tensordot/dgemm behind opt_einsum.contract (MPO:184,280; tensor/utils.py:35,49,62) and the
 * `S @ R`, `B @ R`, `L @ U.T` products of the sweeps (MPS:448,561,575; MPO:560,682,705). */
typedef struct {
    int64_t outer, inner;
    int32_t div;
    int32_t _pad;
} syn_index_t;

typedef struct {
    int32_t M, N, K, batch;
    syn_index_t a_m, a_k, a_b;
    syn_index_t b_k, b_n, b_b;
    syn_index_t c_m, c_n, c_b;
    double alpha, beta;
    /* optional block-lower output mask (0, 0 = none): with mask_rows = R and mask_cols = C only the elements (m, n) with
     * n < (m / R + 1) * C are computed, the rest of C is left untouched -- the symmetric right environment E = C E' C^T of the
     * density-matrix sweep is built from its block-lower part (R = a-block * l rows, C = a-block columns) and mirrored.  Both must be
     * multiples of 64.  Zero-initialise the struct (= {}) when filling it field by field. */
    int32_t mask_rows, mask_cols;
} syn_gemm_desc_t;

int syn_gemm_f64(const syn_gemm_desc_t* desc, const double* A, const double* B, double* C, void* stream);

/* ---- Householder QR truncation step (batched, arbitrary row/column/batch strides) ------------------------- */
/* qrt(A (m x n), q):  Q (m x qk) = orthonormal basis of span(A[:, :q]) (completed with further orthonormal columns when
 * q > n), S = Q^T A (qk x n), qk = min(q, m) written to *qk_out.  S may be NULL.
 * Replaces  Q, R = np.linalg.qr(L, mode="complete"); Q[:, :q]; R[:q, :]   -- the `>>` / decompose truncation step
 *   (MPS:307-311, 332-336, 443-446; MPO:436-440, 555-558; also the non-strict sweeps MPS:378-381,410-413, MPO:488-491,520-523)
 * and the reduced QRs np.linalg.qr(L), np.linalg.qr(R.T) of the canonicalisation sweeps with q = min(m, n)
 *   (MPS:560, 574; MPO:680, 703).  Pass a transposed view (row stride 1) to factor R^T or to write Q^T in place.
 * Workspace: syn_qrt_workspace_f64(m, n, q, batch) bytes, 16-byte aligned, owned by the caller. */
size_t syn_qrt_workspace_f64(int m, int n, int q, int batch);
int syn_qrt_f64(const double* A, int64_t a_rs, int64_t a_cs, int64_t a_bs, int m, int n, int q, int batch,
                double* Q, int64_t q_rs, int64_t q_cs, int64_t q_bs,
                double* S, int64_t s_rs, int64_t s_cs, int64_t s_bs,
                void* ws, size_t ws_bytes, int* qk_out, void* stream);
/* R factor only (n x n upper triangular) of a tall A (m >= n); reduces an unfolding before the Jacobi SVD. */
int syn_qr_r_f64(const double* A, int64_t a_rs, int64_t a_cs, int64_t a_bs, int m, int n, int batch,
                 double* R, int64_t r_rs, int64_t r_cs, int64_t r_bs, void* ws, size_t ws_bytes, void* stream);
/* dst (contiguous m x n per batch) = strided src; a tiled transpose when src is a transposed view
 * (np.reshape / .T of cores: MPS:596-622, MPO:735-769). */
int syn_copy_strided_f64(const double* src, int64_t s_rs, int64_t s_cs, int64_t s_bs, double* dst, int64_t d_bs,
                         int m, int n, int batch, void* stream);

/* ---- one-sided Jacobi SVD (north_star "truncated SVD of each bond matrix"; no live counterpart in the reference:
 *      its only SVD/eigh are dead code -- trash/mpo.py:59-190, experimental/layers.py:241-325, MPO:228) ------------- */
/* In place: rows of G (n x n, row-major ld, n <= 1024) are rotated until mutually orthogonal: row_i -> sigma_i u_i^T. */
size_t syn_jacobi_ctrl_bytes(int batch, int max_sweeps);
int    syn_jacobi_ctrl_stride(int max_sweeps);   /* 32-bit words per problem in ctrl; word [max_sweeps + 1] = sweeps used */
/* null_rel: pairs of rows whose norms are both below null_rel * (largest row norm) are left alone (they fall under the rank
 * threshold of the finalize step); 0 disables the test. */
int syn_jacobi_rows_f64(double* G, int64_t ld, int64_t bs, int n, int batch, void* ctrl, size_t ctrl_bytes,
                        int max_sweeps, double tol, double null_rel, void* stream);
/* Host only (no GPU needed): the block-level ordering of one sweep that the multi-CTA kernel follows for P CTAs (P = 2..64, a
 * power of two; 2P row blocks, 2P - 1 rounds).  table[2 * (round * P + cta)] = blockA | blockB << 8 | flags << 16 with flags
 * 1/2 = load A/B, 4/8 = store A/B, 16 = also rotate the pairs inside both blocks; table[.. + 1] = versionA | versionB << 16 the
 * blocks must have reached (stores earlier in the sweep); cnt[block] = stores of the block per sweep.  Slot A only changes when
 * the recursion descends (log2 P times per sweep), so only slot B is handed from CTA to CTA each round. */
int syn_jacobi_ring_schedule(int P, uint32_t* table, uint32_t* cnt);
int syn_identity_deviation_f64(const double* X, int n, double* out, void* stream);  /* out[0] = max |X - I| */
/* Sort sigma descending, normalise rows into Ut, apply chi_max / relative cutoff on the device.
 * info[2*b] = kept rank, winfo[2*b] = discarded weight sum_{k>=kept} sigma_k^2, winfo[2*b+1] = sigma_0.
 * sqrt_mode = 1 when G was a Gram matrix M E M^T (rows are lambda_i u_i^T, sigma_i = sqrt(lambda_i));
 * sqrt_mode = 2 when G was the factor of syn_chol_upper_f64 (rows are sqrt(lambda_i + shift) q_i^T, sigma_i = sqrt(lambda_i));
 * shift: device array (one per problem) written by syn_chol_upper_f64 (read only when sqrt_mode = 2; may be NULL otherwise).
 * ctrl / max_sweeps: the control buffer and sweep limit of the syn_jacobi_rows_f64 call that produced G (or NULL / 0).  With it,
 * info[2*b+1] = n when that problem's convergence vote passed and -n when the rows kernel ran out of sweeps first (the rows are
 * then not orthogonal to working precision: call syn_jacobi_rows_f64 on the same G again -- it resumes -- and finalize again). */
int syn_jacobi_finalize_f64(const double* G, int64_t ld, int64_t bs, int n, int batch, double* Ut, int64_t ldu, int64_t ubs,
                            double* sigma, int64_t sbs, int* info, double* winfo, int chi_max, double cutoff,
                            double rank_tol, int sqrt_mode, const double* shift, const void* ctrl, int max_sweeps, void* stream);
/* Shifted Cholesky factor of a symmetric PSD Gram matrix (the M E M^T of the density-matrix rounding; finished form of the
 * reference's MATMUL_MODE "opti" branch, matrix_product_operator.py:193-260, whose np.linalg.eigh at :228 this pipeline replaces):
 *   G (batch x n x n, row-major ld, batch stride bs, overwritten when n > 128)  ->  B (row-major ldb, batch stride bbs) upper
 *   triangular with  G_b + shift[b] I = B_b^T B_b;  shift[b] = 2 n eps max_i G_ii (device array of `batch` doubles) keeps the
 *   factorisation defined for singular G.  n <= 128: one CTA per problem, one launch; larger n: 64-column panels.
 * One-sided Jacobi on the rows of B (syn_jacobi_rows_f64) then converges in ~10 sweeps instead of ~13-27 on G itself, and
 * syn_jacobi_finalize_f64(sqrt_mode = 2, shift) returns the eigenvectors of G and sigma_i = sqrt(lambda_i). */
int syn_chol_upper_f64(double* G, int64_t ld, int64_t bs, int n, int batch, double* B, int64_t ldb, int64_t bbs, double* shift, void* stream);

/* ---- fused transfer-matrix inner product (FP64 DMMA, persistent over the batch) ------------------------------------ */
/* <A_s | B_s> for s = 0..batch-1 (bilinear, no conjugation):  E <- sum_{a,a',i} A_k[a,i,b] E[a,a'] B_k[a',i,b'],  k = 0..n_sites-1.
 * Replaces the 2N-operand opt_einsum.contract of MatrixProductState.__or__ (MPS:116-129), `dot` (MPS:243) and syn.mul on two
 * states (tensor/utils.py:22-28) -- in the reference a batch is a Python loop over that call.  One CTA walks one pair of chains;
 * the transfer matrix stays in shared memory and every core is read from HBM exactly once.
 *   sites   HOST array; site k: `a` / `b` = core k of state 0 of either chain, C-ordered (la, d, ra) / (lb, d, rb); state s lives
 *           a_stride / b_stride elements further (0 = one core shared by the whole batch).
 *   E_in    (batch, la_0, lb_0) or NULL (then la_0 = lb_0 = 1 and E starts as the scalar 1);  E_out (batch, ra_last, rb_last).
 * Chains longer than SYN_OVERLAP_MAX_SITES are split by the caller (E_out of one call is E_in of the next).
 * syn_overlap_batched_fits() != 0 when every bond is <= SYN_OVERLAP_MAX_BOND, round8(l) * d and d * r are <= 2 * SYN_OVERLAP_MAX_BOND
 * and consecutive bonds match; other chains take the GEMM-per-site route (syn_gemm_f64). */
#define SYN_OVERLAP_MAX_SITES 64
#define SYN_OVERLAP_MAX_BOND 64
typedef struct {
    const double* a;
    const double* b;
    int64_t a_stride, b_stride;
    int32_t la, ra, lb, rb, d;
    int32_t _pad;
} syn_overlap_site_t;
int syn_overlap_batched_fits(const syn_overlap_site_t* sites, int n_sites);
int syn_overlap_batched_f64(const syn_overlap_site_t* sites, int n_sites, int batch, const double* E_in, double* E_out, void* stream);

/* ---- dominant invariant subspace by spectral projection (GEMM-bound eigen-solver of the density-matrix rounding) ------------
 * U (n x ne, row-major) = orthonormal basis of the span of the ne dominant eigenvectors of the symmetric PSD matrix A (n x n,
 * contiguous, not modified): `sp2_iters` trace-steered SP2 steps X <- X^2 | 2X - X^2 from X = A / |A|_F (one DMMA GEMM + one
 * elementwise kernel each, branch chosen on the device), then `ns_iters` Newton-Schulz steps on P[:, :ne].  No host round trip.
 * Stands where the reference's unfinished density-matrix branch calls np.linalg.eigh (MPO:228) and keeps eigvecs[:, :min_bond].
 * info (device, 8 doubles): [0] tr P, [1] |P|_F^2 (both = ne when converged), [2] sum A o P = kept weight, [3] |A|_F,
 * [4] max |U^T U - I|, [5] tr A, [6] tr(X - X^2) one step before the end, [7] steps + 1e6 * (number of leading 2X - X^2 steps ~
 * log2(|A|_F / lambda_cut): the subspace is accurate to ~eps |A| / gap, so callers reject cuts deep in a decaying spectrum).  The caller decides (and falls back to
 * syn_jacobi_rows_f64 when the spectrum has no gap at ne). */
size_t syn_dominant_subspace_workspace_f64(int n, int ne, int sp2_iters);
int syn_dominant_subspace_f64(const double* A, int n, int ne, int sp2_iters, int ns_iters, double* U, void* ws, size_t ws_bytes,
                              double* info, void* stream);
/* The same solver as ONE persistent cooperative kernel (one CTA per SM, iterates resident in L2, 32 x 32 DMMA tiles from a cp.async
 * ring, symmetric products computed on their lower tiles only, grid barrier per step).  The iteration counts adapt on the device:
 * SP2 stops two steps after tr(X - X^2) < 1e-11 ne (at most sp2_max steps), Newton-Schulz when max |U^T U - I| < 1e-13 (at most
 * ns_max).  Needs n and ne multiples of 32, n >= 128 (syn_dominant_subspace_fused_fits).  info as above, [7] = sp2 steps + 1000 * NS steps + 1e6 * leading lift steps. */
size_t syn_dominant_subspace_fused_workspace_f64(int n, int ne, int sp2_max, int ns_max);
int syn_dominant_subspace_fused_fits(int n, int ne);
int syn_dominant_subspace_fused_f64(const double* A, int n, int ne, int sp2_max, int ns_max, double* U, void* ws, size_t ws_bytes,
                                    double* info, void* stream);

/* The same solver for a complex HERMITIAN positive semi-definite matrix H = Hre + i Him (planar, m x m each, contiguous): orthonormal basis
 * U = Ure + i Uim (m x k planar) of the span of its k dominant eigenvectors.  The kernel runs on the interleaved real embedding (every entry
 * a + ib as [[a, -b], [b, a]]: 2m x 2m, built and taken apart on the device by this call) and uses its structure: every iterate is an
 * embedding again, so only the even rows of each product are formed (one accumulator pair = one complex entry) and the odd rows are written
 * from them -- the DMMA work of a native complex product, half of the plain embedded one.  info as above in the embedded convention
 * (traces count every eigenvalue twice: tr P = 2k).  m, k multiples of 32, k < m.  Caller: the bond SVD of complex chains
 * (reference: tensor/matrix_product_state.py:487-534 apply with the complex gates of quantum/gate.py:9-37). */
int syn_dominant_subspace_c128_fits(int m, int k);
size_t syn_dominant_subspace_c128_workspace(int m, int k, int sp2_max, int ns_max);
int syn_dominant_subspace_c128(const double* Hre, const double* Him, int m, int k, int sp2_max, int ns_max, double* Ure, double* Uim,
                               void* ws, size_t ws_bytes, double* info, void* stream);

/* The projection solver for a BATCH of small bond problems (csrc/purify_batched.cu): one CTA per problem, the iterate and the basis stay in
 * shared memory, products in registers, two CTA barriers per step; a persistent grid walks the batch.  A: batch x n x n contiguous symmetric
 * PSD matrices (16-byte aligned), U: batch x n x ne, info: batch x 8 doubles with the meaning above.  n, ne multiples of 32,
 * 32 <= ne < n <= 128 (syn_dominant_subspace_batched_fits).  Replaces the per-state truncation of a batch of chains
 * (reference: a Python loop over tensor/matrix_product_state.py:432-468). */
int syn_dominant_subspace_batched_fits(int n, int ne);
int syn_dominant_subspace_batched_f64(const double* A, int batch, int n, int ne, int sp2_max, int ns_max, double* U, double* info, void* stream);

/* Y[q][ro][x] = sum_ri W[ro][ri] X[q][ri][x]: a small matrix (rin, rout in {2, 4, 8, 16}; W row-major contiguous, or strided in the
 * _strided form: element (ro, ri) at W[ro * w_ro + ri * w_ri], e.g. an MPO core read in place) applied to the middle index
 * of a large tensor -- the contractions of a BATCH of chains with one shared MPO core (the per-site einsum of MPO:181-192 looped over states),
 * which are pure streaming (csrc/smallcore.cu).  X strides (x_q, x_r, x_l), Y strides (y_q, y_l) and a two-level ro index
 * ((ro / y_ro_div) * y_ro_outer + (ro % y_ro_div) * y_ro_inner), all in elements; x should be the unit-stride index of X. */
int syn_apply_small_core_fits(int rin, int rout);
int syn_apply_small_core_f64(const double* X, const double* W, double* Y, int64_t Q, int rin, int rout, int L, int64_t x_q, int64_t x_r,
                             int64_t x_l, int64_t y_q, int64_t y_ro_outer, int64_t y_ro_inner, int y_ro_div, int64_t y_l, void* stream);
int syn_apply_small_core_strided_f64(const double* X, const double* W, int64_t w_ro, int64_t w_ri, double* Y, int64_t Q, int rin, int rout, int L,
                                     int64_t x_q, int64_t x_r, int64_t x_l, int64_t y_q, int64_t y_ro_outer, int64_t y_ro_inner, int y_ro_div,
                                     int64_t y_l, void* stream);

/* out[i] = sum_p parts[p * part_stride + i], i < count: the reduction after a split-K syn_gemm_f64 (partials as the batch index). */
int syn_sum_parts_f64(const double* parts, int64_t part_stride, int nparts, double* out, int64_t count, void* stream);

/* Q (m x q, contiguous) = orthonormal basis of the column space of A (m x q, row stride lda, full column rank) by the Newton-Schulz
 * phase of the same persistent kernel: Q = A (A^T A)^(-1/2).  For the truncation step `Q, R = np.linalg.qr(L, "complete"); Q[:, :q]`
 * (MPS:443-446, MPO:555-558) any orthonormal basis of span(L[:, :q]) is the same projection, and this one is GEMM-bound (0.2 ms against
 * 1.2 ms for the cluster Householder kernel at 512 x 256).  info[4] = max |Q^T Q - I| at exit, info[7] = 1000 * steps; the caller falls
 * back to syn_qrt_f64 when it did not converge (rank-deficient columns).  m, q multiples of 32, m >= 128, q <= m. */
size_t syn_orthonormalize_columns_workspace_f64(int m, int q, int ns_max);
int syn_orthonormalize_columns_fits(int m, int q);
int syn_orthonormalize_columns_f64(const double* A, int64_t lda, int m, int q, int ns_max, double* Q, void* ws, size_t ws_bytes,
                                   double* info, void* stream);

/* ---- fused W-sandwich of the right-environment update (density-matrix rounding) -------------------------------------------------
 * Z[a, l, (l',i'), b'] = sum_{(o,r')} W[l',i',o,r'] ( sum_{(i,r)} W[l,i,o,r] P1[a, (i,r), (r',b')] ):  the two small-K contractions
 * with the MPO core W (l,i,o,r) on either side of the large intermediate in ONE kernel, the intermediate staying in shared memory
 * (two np.tensordot calls in the restated oracle, oracle/svd_numpy.py:107-110; the reference's unfinished counterpart is MPO:193-260).
 * P1 (a, i*r, r, b), Z (a, l, l*i, b) contiguous.  Covered shapes: (l,i,o,r) = (16,2,2,16), b % 16 == 0 (syn_env_sandwich_fits);
 * anything else takes two syn_gemm_f64 calls. */
/* E[(a,l),(l',a')] (na*L x L*na, contiguous) is symmetric under (a,l) <-> (a',l'): fill the blocks with a' in a LATER block of `ab`
 * values than a from their mirror images (the last GEMM of the environment update then only forms the block-lower part). */
int syn_env_mirror_f64(double* E, int na, int L, int ab, void* stream);
int syn_env_sandwich_fits(int l, int i, int o, int r, int b);
int syn_env_sandwich_f64(const double* P1, const double* W, double* Z, int na, int l, int i, int o, int r, int b, void* stream);

/* ---- block assembly and elementwise kernels ---------------------------------------------------------------- */
/* `A + B` site: direct sum of the bond spaces, self's block first; cores as (l, phys, r) with the physical legs flattened
 * (np.block / scipy.linalg.block_diag loops of MPS:82-96 and MPO:90-106). */
int syn_add_site_f64(const double* A, const double* B, double* out, int la, int ra, int lb, int rb, int phys,
                     int first, int last, void* stream);
/* `A * B` site: out[(la,lb), p, (ra,rb)] = A[la,p,ra] * B[lb,p,rb]  (np.kron loop of MPO:140-152). */
int syn_kron_site_f64(const double* A, const double* B, double* out, int la, int ra, int lb, int rb, int phys, void* stream);
/* out[0] = sum x_i^2 ; x *= 1/sqrt(sumsq[0])   (normalize: np.linalg.norm + in-place divide, MPS:252-256) */
int syn_sumsq_f64(const double* x, int64_t n, double* out, void* stream);
int syn_scale_rsqrt_f64(double* x, int64_t n, const double* sumsq, void* stream);
/* y[r,c] = act(y[r,c] + bias[c]), act 0 = identity, 1 = relu: the `+ bias` / activation epilogue of the TT layer
 * (layers/TensorDense.py:139-142).  The TT contraction itself is a chain of syn_gemm_f64 calls. */
int syn_bias_act_f64(double* y, const double* bias, int64_t rows, int cols, int act, void* stream);

/* ---- whole-chain sweeps: ONE call per `@` + `>>` or per `>>` over arrays of core pointers --------------------------------------- */
/* syn_apply_round_chain_f64: MatrixProductOperator.__matmul__ (matrix_product_operator.py:181-192: per-site contract of X_k (a,i,b) with
 * W_k (l,i,o,r), product bonds flattened MPS-bond major) followed by the strict `>> dim` QR-truncation sweep of
 * MatrixProductState.compress (matrix_product_state.py:432-468), fused: the D = chi * chi_W product cores are never formed, the result is
 * identical (same projections, same column order).  Natural clamp: kept width = min(dim, rows).
 *   X, W        HOST arrays of n_sites DEVICE pointers to contiguous float64 cores; xshape = (a,i,b) per site, wshape = (l,i,o,r) per site
 *   out         HOST array of n_sites DEVICE pointers, core k sized by syn_apply_round_chain_shapes (s_k, o_k, s_{k+1}); the last core
 *               carries the norm and has right bond b * r
 *   ws          device workspace of syn_apply_round_chain_workspace_f64 bytes, 256-byte aligned, owned by the caller
 * syn_round_chain_f64: the same `>>` sweep on the cores of one chain (MPS; an MPO passes its two physical legs flattened into one:
 * matrix_product_operator.py:544-580), shape = (l, d, r) per site.
 * The calls synchronise the stream once per bond (8 doubles: the verdict of the Newton-Schulz orthonormalisation that the truncation
 * step tries before the Householder kernels). */
int syn_apply_round_chain_shapes(int n_sites, const int* xshape, const int* wshape, int dim, int* out_shape);
size_t syn_apply_round_chain_workspace_f64(int n_sites, const int* xshape, const int* wshape, int dim);
int syn_apply_round_chain_f64(int n_sites, const double* const* X, const int* xshape, const double* const* W, const int* wshape, int dim,
                              double* const* out, void* ws, size_t ws_bytes, void* stream);
int syn_round_chain_shapes(int n_sites, const int* shape, int dim, int* out_shape);
size_t syn_round_chain_workspace_f64(int n_sites, const int* shape, int dim);
int syn_round_chain_f64(int n_sites, const double* const* cores, const int* shape, int dim, double* const* out, void* ws, size_t ws_bytes,
                        void* stream);

/* ---- TensorDense forward on the 5th-generation tensor cores (tcgen05.mma.kind::tf32, TMEM, TMA) ------------------------------- */
/* Fused TT-matvec of the MPO-compressed dense layer, reference layers/TensorDense.py:74-142 (`call`: per-sample opt_einsum contract of
 * the reshaped input with the three cores at :103-118, `+ bias` at :139, activation at :142), for three cores with every mode and bond
 * equal to 16 (BASELINE configs[4]: a 4096 -> 4096 layer):
 *   y[s,o1,o2,o3] = act( sum x[s,i1,i2,i3] G1[i1,o1,b1] G2[i2,o2,b1,b2] G3[i3,o3,b2] + bias[o1,o2,o3] ),  float32 in / out, TF32 products
 *   with FP32 accumulation; both 65,536-float intermediates of a sample stay in TMEM / shared memory.
 * syn_tt_dense3_pack_tf32 turns the cores (the reference's layouts, device float32) into the pre-swizzled operand images the kernel
 * stages with bulk copies (`packed`: syn_tt_dense3_packed_floats() device floats, 16-byte aligned; redo it when the weights change).
 * syn_tt_dense3_tf32: x (batch x 4096) and y (batch x 4096) device float32, 16-byte aligned, contiguous; bias 4096 floats or NULL;
 * relu = 0 / 1.  Other layer shapes run on the FP64 strided GEMM (syn_gemm_f64 + syn_bias_act_f64). */
size_t syn_tt_dense3_packed_floats(void);
int syn_tt_dense3_pack_tf32(const float* G1, const float* G2, const float* G3, float* packed, void* stream);
int syn_tt_dense3_tf32(const float* x, const float* packed, const float* bias, float* y, int batch, int relu, void* stream);
/* Variant on CTA pairs (tcgen05.mma.cta_group::2, M = 256: the pair shares one sample, each CTA holds one half of G2 and of the output).
 * Same arguments and tolerance.  Measured slower than the single-CTA kernel at batch 65536 (6.6 vs 6.2 ms), so nothing selects it by
 * default; it is kept as a checked alternative. */
int syn_tt_dense3_tf32_pair(const float* x, const float* packed, const float* bias, float* y, int batch, int relu, void* stream);

#ifdef __cplusplus
}
#endif
#endif

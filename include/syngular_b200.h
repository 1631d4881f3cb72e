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
} syn_gemm_desc_t;

int syn_gemm_f64(const syn_gemm_desc_t* desc, const double* A, const double* B, double* C, void* stream);

#ifdef __cplusplus
}
#endif
#endif

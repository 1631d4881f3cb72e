// Stand-alone consumer of the C ABI (no Python, no PyTorch): what a non-Python binding of the reference would do.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I include tests/abi/abi_smoke.cu -L syngular_b200 -lsyngular_b200 -o abi_smoke
// Checks syn_gemm_f64 (plain + two-level site contraction), syn_qrt_f64 and the Jacobi SVD against CPU loops, and runs a
// C2-shaped apply + round (N = 64, chi = 256, chi_W = 16) with ONE call of syn_apply_round_chain_f64.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "syngular_b200.h"

#define CK(x) do { if ((x) != cudaSuccess) { printf("cuda error line %d\n", __LINE__); return 1; } } while (0)
#define SYN(x) do { if ((x) != 0) { printf("syn error line %d: %s\n", __LINE__, syn_last_error()); return 1; } } while (0)

static syn_index_t IX(int64_t stride) { syn_index_t i = {0, stride, 2147483647, 0}; return i; }
static syn_index_t IX2(int64_t outer, int64_t inner, int div) { syn_index_t i = {outer, inner, div, 0}; return i; }
static double rnd() { return (double)rand() / RAND_MAX - 0.5; }

int main() {
    printf("syn_version %d, SMs %d\n", syn_version(), syn_device_sm_count());
    // ---- 1. C = A B ------------------------------------------------------------------------------------
    const int M = 70, N = 45, K = 33;
    std::vector<double> A(M * K), B(K * N), C(M * N), R(M * N, 0.0);
    for (auto& v : A) v = rnd();
    for (auto& v : B) v = rnd();
    for (int m = 0; m < M; m++) for (int n = 0; n < N; n++) for (int k = 0; k < K; k++) R[m * N + n] += A[m * K + k] * B[k * N + n];
    double *dA, *dB, *dC;
    CK(cudaMalloc(&dA, A.size() * 8)); CK(cudaMalloc(&dB, B.size() * 8)); CK(cudaMalloc(&dC, C.size() * 8));
    CK(cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice));
    syn_gemm_desc_t d = {M, N, K, 1, IX(K), IX(1), IX(0), IX(N), IX(1), IX(0), IX(N), IX(1), IX(0), 1.0, 0.0};
    SYN(syn_gemm_f64(&d, dA, dB, dC, nullptr));
    CK(cudaMemcpy(C.data(), dC, C.size() * 8, cudaMemcpyDeviceToHost));
    double e1 = 0; for (size_t i = 0; i < C.size(); i++) e1 = fmax(e1, fabs(C[i] - R[i]));
    printf("gemm max err %.2e\n", e1);
    // ---- 2. MPO x MPS site contraction as ONE strided GEMM (matrix_product_operator.py:184-190) ----------
    const int a = 5, i = 3, b = 4, l = 2, o = 3, r = 6;
    std::vector<double> X(a * i * b), W(l * i * o * r), S(a * l * o * b * r), SR(S.size(), 0.0);
    for (auto& v : X) v = rnd();
    for (auto& v : W) v = rnd();
    for (int aa = 0; aa < a; aa++) for (int ll = 0; ll < l; ll++) for (int oo = 0; oo < o; oo++) for (int bb = 0; bb < b; bb++) for (int rr = 0; rr < r; rr++) {
        double s = 0; for (int ii = 0; ii < i; ii++) s += X[(aa * i + ii) * b + bb] * W[((ll * i + ii) * o + oo) * r + rr];
        SR[(((aa * l + ll) * o + oo) * b + bb) * r + rr] = s;
    }
    double *dX, *dW, *dS;
    CK(cudaMalloc(&dX, X.size() * 8)); CK(cudaMalloc(&dW, W.size() * 8)); CK(cudaMalloc(&dS, S.size() * 8));
    CK(cudaMemcpy(dX, X.data(), X.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dW, W.data(), W.size() * 8, cudaMemcpyHostToDevice));
    syn_gemm_desc_t ds = {b, o * r, i, a * l, IX(1), IX(b), IX2(i * b, 0, l), IX(o * r), IX(1), IX2(0, i * o * r, l),
                          IX(r), IX2(b * r, 1, r), IX(o * b * r), 1.0, 0.0};
    SYN(syn_gemm_f64(&ds, dX, dW, dS, nullptr));
    CK(cudaMemcpy(S.data(), dS, S.size() * 8, cudaMemcpyDeviceToHost));
    double e2 = 0; for (size_t k = 0; k < S.size(); k++) e2 = fmax(e2, fabs(S[k] - SR[k]));
    printf("site contraction max err %.2e\n", e2);
    // ---- 3. qrt: Q^T Q = I and Q S = projection of L on its first q columns ----------------------------------
    const int m = 96, n = 40, q = 12;
    std::vector<double> L(m * n), Q(m * q), St(q * n);
    for (auto& v : L) v = rnd();
    double *dL, *dQ, *dSt; void* ws;
    size_t wsb = syn_qrt_workspace_f64(m, n, q, 1);
    CK(cudaMalloc(&dL, L.size() * 8)); CK(cudaMalloc(&dQ, Q.size() * 8)); CK(cudaMalloc(&dSt, St.size() * 8)); CK(cudaMalloc(&ws, wsb));
    CK(cudaMemcpy(dL, L.data(), L.size() * 8, cudaMemcpyHostToDevice));
    int qk = 0;
    SYN(syn_qrt_f64(dL, n, 1, 0, m, n, q, 1, dQ, q, 1, 0, dSt, n, 1, 0, ws, wsb, &qk, nullptr));
    CK(cudaMemcpy(Q.data(), dQ, Q.size() * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(St.data(), dSt, St.size() * 8, cudaMemcpyDeviceToHost));
    double e3 = 0, e4 = 0;
    for (int x = 0; x < q; x++) for (int y = 0; y < q; y++) { double s = 0; for (int k = 0; k < m; k++) s += Q[k * q + x] * Q[k * q + y]; e3 = fmax(e3, fabs(s - (x == y))); }
    for (int k = 0; k < m; k++) for (int c = 0; c < q; c++) { double s = 0; for (int x = 0; x < q; x++) s += Q[k * q + x] * St[x * n + c]; e4 = fmax(e4, fabs(s - L[k * n + c])); }
    printf("qrt kept %d, |QtQ-I| %.2e, first-q-columns reproduced to %.2e\n", qk, e3, e4);
    // ---- 4. Jacobi SVD of a diagonal-dominant matrix: singular values sorted, rows orthonormal -----------------
    const int nj = 48;
    std::vector<double> G(nj * nj), Ut(nj * nj), sig(nj);
    for (int x = 0; x < nj; x++) for (int y = 0; y < nj; y++) G[x * nj + y] = (x == y ? (x + 1.0) : 0.0) + 0.01 * rnd();
    double *dG, *dUt, *dsig, *dw; int* dinfo; void* ctrl;
    size_t cb = syn_jacobi_ctrl_bytes(1, 40);
    CK(cudaMalloc(&dG, G.size() * 8)); CK(cudaMalloc(&dUt, Ut.size() * 8)); CK(cudaMalloc(&dsig, nj * 8)); CK(cudaMalloc(&dw, 16)); CK(cudaMalloc(&dinfo, 8)); CK(cudaMalloc(&ctrl, cb));
    CK(cudaMemcpy(dG, G.data(), G.size() * 8, cudaMemcpyHostToDevice));
    SYN(syn_jacobi_rows_f64(dG, nj, (int64_t)nj * nj, nj, 1, ctrl, cb, 40, 1e-14, 1e-14, nullptr));
    SYN(syn_jacobi_finalize_f64(dG, nj, (int64_t)nj * nj, nj, 1, dUt, nj, (int64_t)nj * nj, dsig, nj, dinfo, dw, 10, 0.0, 1e-14, 0, nullptr, nullptr, 0, nullptr));
    int info[2]; CK(cudaMemcpy(info, dinfo, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(sig.data(), dsig, nj * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(Ut.data(), dUt, Ut.size() * 8, cudaMemcpyDeviceToHost));
    double e5 = 0; bool sorted = true;
    for (int x = 0; x + 1 < nj; x++) sorted = sorted && sig[x] >= sig[x + 1];
    for (int x = 0; x < nj; x++) for (int y = 0; y < nj; y++) { double s = 0; for (int k = 0; k < nj; k++) s += Ut[x * nj + k] * Ut[y * nj + k]; e5 = fmax(e5, fabs(s - (x == y))); }
    printf("jacobi keep %d (chi_max 10), sigma_0 %.4f, sorted %d, |UUt-I| %.2e, launches so far %lld\n", info[0], sig[0], (int)sorted, e5, syn_launch_count());
    bool ok = e1 < 1e-13 && e2 < 1e-13 && e3 < 1e-13 && e4 < 1e-12 && e5 < 1e-12 && sorted && info[0] == 10 && qk == q && fabs(sig[0] - 48.0) < 0.1;
    // ---- 5. errors are reported, not thrown -------------------------------------------------------------------
    int rc = syn_gemm_f64(nullptr, dA, dB, dC, nullptr);
    ok = ok && rc != 0;
    // ---- 6. a C2-shaped apply + round (N = 64, d = 2, chi = 256, chi_W = 16 -> 256) in ONE call: syn_apply_round_chain_f64 ----------
    {
        const int n = 64, dphys = 2, chi = 256, chiw = 16, dim = 256;
        std::vector<int> bx(n + 1, 1), bw(n + 1, 1);
        for (int k = 1; k < n; k++) {
            int e = k < n - k ? k : n - k;
            bx[k] = e >= 8 ? chi : (1 << e);  if (bx[k] > chi) bx[k] = chi;
            bw[k] = e >= 2 ? chiw : (1 << (2 * e));  if (bw[k] > chiw) bw[k] = chiw;
        }
        std::vector<int> xs(3 * n), ws4(4 * n), os(3 * n);
        std::vector<const double*> Xp(n), Wp(n);
        std::vector<double*> Op(n);
        for (int k = 0; k < n; k++) {
            xs[3 * k] = bx[k]; xs[3 * k + 1] = dphys; xs[3 * k + 2] = bx[k + 1];
            ws4[4 * k] = bw[k]; ws4[4 * k + 1] = dphys; ws4[4 * k + 2] = dphys; ws4[4 * k + 3] = bw[k + 1];
        }
        SYN(syn_apply_round_chain_shapes(n, xs.data(), ws4.data(), dim, os.data()));
        for (int k = 0; k < n; k++) {
            size_t nx = (size_t)bx[k] * dphys * bx[k + 1], nw = (size_t)bw[k] * dphys * dphys * bw[k + 1];
            std::vector<double> hx(nx), hw(nw);
            for (auto& v : hx) v = rnd() / sqrt((double)bx[k]);
            for (auto& v : hw) v = rnd() / sqrt((double)bw[k]);
            double *dx, *dw, *dout;
            CK(cudaMalloc(&dx, nx * 8)); CK(cudaMalloc(&dw, nw * 8)); CK(cudaMalloc(&dout, (size_t)os[3 * k] * os[3 * k + 1] * os[3 * k + 2] * 8));
            CK(cudaMemcpy(dx, hx.data(), nx * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dw, hw.data(), nw * 8, cudaMemcpyHostToDevice));
            Xp[k] = dx; Wp[k] = dw; Op[k] = dout;
        }
        size_t cwb = syn_apply_round_chain_workspace_f64(n, xs.data(), ws4.data(), dim);
        void* cws; CK(cudaMalloc(&cws, cwb));
        long long l0 = syn_launch_count();
        cudaEvent_t t0, t1; CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
        SYN(syn_apply_round_chain_f64(n, Xp.data(), xs.data(), Wp.data(), ws4.data(), dim, Op.data(), cws, cwb, nullptr));     // warm-up
        CK(cudaEventRecord(t0));
        SYN(syn_apply_round_chain_f64(n, Xp.data(), xs.data(), Wp.data(), ws4.data(), dim, Op.data(), cws, cwb, nullptr));
        CK(cudaEventRecord(t1)); CK(cudaDeviceSynchronize());
        float ms = 0; CK(cudaEventElapsedTime(&ms, t0, t1));
        long long launches = (syn_launch_count() - l0) / 2;
        // plateau core 30 must be left-orthonormal: Q (512 x 256)
        const int k30 = 30, rows = os[3 * k30] * os[3 * k30 + 1], cols = os[3 * k30 + 2];
        std::vector<double> Q((size_t)rows * cols);
        CK(cudaMemcpy(Q.data(), Op[k30], Q.size() * 8, cudaMemcpyDeviceToHost));
        double e6 = 0;
        for (int x = 0; x < cols; x += 17) for (int y = 0; y < cols; y += 13) { double s2 = 0; for (int r2 = 0; r2 < rows; r2++) s2 += Q[(size_t)r2 * cols + x] * Q[(size_t)r2 * cols + y]; e6 = fmax(e6, fabs(s2 - (x == y))); }
        printf("chain apply+round: core 30 is (%d, %d, %d), |QtQ-I| (sampled) %.2e, %lld kernel launches, %.2f ms per sweep in one call\n",
               os[3 * k30], os[3 * k30 + 1], os[3 * k30 + 2], e6, launches, ms);
        ok = ok && os[3 * k30] == 256 && os[3 * k30 + 2] == 256 && e6 < 1e-10 && launches > 0;
    }
    printf("%s\n", ok ? "ABI SMOKE OK" : "ABI SMOKE FAILED");
    return ok ? 0 : 1;
}

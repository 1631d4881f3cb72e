// Whole-chain entry points of the C ABI: ONE call runs a complete sweep of the hot path over arrays of core pointers.
//
//   syn_apply_round_chain_f64   `W @ X` followed by `>> dim` with the reference's semantics -- the site contraction of
//                               MatrixProductOperator.__matmul__ (matrix_product_operator.py:181-192) fused with the strict QR-truncation
//                               sweep of MatrixProductState.compress (matrix_product_state.py:432-468): per site two strided GEMMs contract
//                               the carry with the MPS and MPO cores (the D = chi * chi_W product core is never formed), the truncation step
//                               keeps the span of the first `dim` columns, and the next carry is Q^T L with its columns re-ordered on the fly.
//   syn_round_chain_f64         the `>>` sweep alone on a chain of cores (MPS, or MPO with its physical legs flattened):
//                               matrix_product_state.py:432-468 / matrix_product_operator.py:544-580.
//
// Both are the C++ form of syngular/tensor/_sweeps.py (apply_round_qr, round_qr), which are thin callers of these.  The truncation step
// tries the fused Newton-Schulz orthonormalisation kernel first (csrc/purify.cu; any orthonormal basis of the span gives the same
// projection as the reference's Q[:, :q]) and falls back to the Householder kernels when it does not reach 1e-12 (rank-deficient leading
// columns) or the shape is not covered: that verdict is 8 doubles read by the host once per site, so a sweep is a fixed sequence of
// launches plus one small device-to-host copy per bond.  Every buffer, workspace included, is owned by the caller.
#include <algorithm>
#include <climits>
#include <cmath>

#include <vector>

#include "common.cuh"

namespace syn {

int gemm_f64(const syn_gemm_desc_t& d, const double* A, const double* B, double* C, cudaStream_t st);
int apply_small_core(const double* X, const double* W, int64_t w_ro, int64_t w_ri, double* Y, int64_t Q, int rin, int rout, int L, int64_t x_q,
                     int64_t x_r, int64_t x_l, int64_t y_q, int64_t y_ro_outer, int64_t y_ro_inner, int y_ro_div, int64_t y_l, cudaStream_t stream);
bool small_core_fits(int rin, int rout);

namespace chain {

static inline syn_index_t IX(int64_t stride) {
    syn_index_t i;
    i.outer = 0; i.inner = stride; i.div = INT_MAX; i._pad = 0;
    return i;
}
static inline syn_index_t IX2(int64_t outer, int64_t inner, int div) {
    syn_index_t i;
    i.outer = outer; i.inner = inner; i.div = div; i._pad = 0;
    return i;
}
static inline size_t up16(size_t bytes) { return (bytes + 255) & ~(size_t)255; }

constexpr int ORTHO_MIN_ROWS = 256;       // truncation steps with at least this many rows try the Newton-Schulz kernel first
constexpr int ORTHO_NS_MAX = 60;

// does the truncation step of an (m x n) unfolding to q columns take the Newton-Schulz route?  (the same rule as _sweeps.qrt_step)
static bool ortho_route(int m, int n, int q) { return m >= ORTHO_MIN_ROWS && q <= n && syn_orthonormalize_columns_fits(m, q) != 0; }

// workspace of one truncation step: max of the two routes, plus the 8 info doubles
static size_t trunc_ws_bytes(int m, int n, int q) {
    size_t a = syn_qrt_workspace_f64(m, n, q, 1);
    size_t b = ortho_route(m, n, q) ? syn_orthonormalize_columns_workspace_f64(m, q, ORTHO_NS_MAX) : 0;
    return up16(std::max(a, b)) + 256;
}

// Q (m x kept, contiguous) = orthonormal basis of span(L[:, :q]) (+ completion when q > n); S (kept x n, contiguous) = Q^T L when asked.
// L: m x n, row stride ldl, unit column stride.  ws: trunc_ws_bytes(m, n, q).
static int truncation_step(const double* L, int64_t ldl, int m, int n, int q, double* Q, double* S, void* ws, size_t ws_bytes, int* kept_out,
                           cudaStream_t st, double* late_info = nullptr, bool* late_used = nullptr) {
    // late_info != nullptr: SPECULATIVE step -- the Newton-Schulz route writes its 8 verdict doubles there and the host does not wait for
    // them (the caller reads the verdicts of all sites once, after the sweep, and repeats the sweep with immediate verdicts if one failed)
    const int kept = std::min(q, m);
    *kept_out = kept;
    if (late_used) *late_used = false;
    double* info = reinterpret_cast<double*>(static_cast<char*>(ws) + ws_bytes - 256);
    if (ortho_route(m, n, q) && (((uintptr_t)L) & 15) == 0 && (ldl & 1) == 0) {
        if (int rc = syn_orthonormalize_columns_f64(L, ldl, m, q, ORTHO_NS_MAX, Q, ws, ws_bytes - 256, late_info ? late_info : info, st)) return rc;
        bool ok = true;
        if (late_info) {
            if (late_used) *late_used = true;
        } else {
            double h[8];
            SYN_CUDA(cudaMemcpyAsync(h, info, sizeof(h), cudaMemcpyDeviceToHost, st));
            SYN_CUDA(cudaStreamSynchronize(st));
            ok = h[4] < 1e-12;
            for (double v : h) ok = ok && std::isfinite(v);
        }
        if (ok) {
            if (S) {
                syn_gemm_desc_t d = {kept, n, m, 1, IX(1), IX(kept), IX(0), IX(ldl), IX(1), IX(0), IX(n), IX(1), IX(0), 1.0, 0.0};
                if (int rc = gemm_f64(d, Q, L, S, st)) return rc;
            }
            return 0;
        }
    }
    int qk = 0;
    if (int rc = syn_qrt_f64(L, ldl, 1, 0, m, n, q, 1, Q, kept, 1, 0, S, S ? n : 0, S ? 1 : 0, 0, ws, ws_bytes - 256, &qk, st)) return rc;
    SYN_REQUIRE(qk == kept, "chain sweep: truncation step kept %d columns, expected %d", qk, kept);
    return 0;
}

// the verdicts of a speculative sweep: one copy and one synchronisation for all sites
static int late_verdicts(const double* infos, const std::vector<char>& used, int n_sites, bool* all_ok, cudaStream_t st) {
    std::vector<double> h((size_t)n_sites * 8);
    SYN_CUDA(cudaMemcpyAsync(h.data(), infos, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, st));
    SYN_CUDA(cudaStreamSynchronize(st));
    *all_ok = true;
    for (int k = 0; k < n_sites; k++) {
        if (!used[k]) continue;
        bool ok = h[(size_t)k * 8 + 4] < 1e-12;
        for (int e = 0; e < 8; e++) ok = ok && std::isfinite(h[(size_t)k * 8 + e]);
        if (!ok) *all_ok = false;
    }
    return 0;
}

// M[s, o, (b, r)] = sum_{a,l,i} T[s,l,a] X[a,i,b] W[l,i,o,r]; T1 (s,l,i,b) is scratch.  Carry stored (s, l, a): MPO bond major, so that both
// GEMMs read unit-stride operands; the output columns are in the reference's product order (b major, r minor).
static int contract_carry(const double* T, int s, const double* X, int a, int i, int b, const double* W, int l, int o, int r, double* T1, double* M,
                          cudaStream_t st) {
    syn_gemm_desc_t d1 = {s, i * b, a, l, IX((int64_t)l * a), IX(1), IX(a), IX((int64_t)i * b), IX(1), IX(0), IX((int64_t)l * i * b), IX(1),
                          IX((int64_t)i * b), 1.0, 0.0};
    if (int rc = gemm_f64(d1, T, X, T1, st)) return rc;
    if (small_core_fits(l * i, o * r) && (int64_t)s * b >= 4096)
        // an MPO core of at most 16 x 16 against s b columns: pure streaming (smallcore.cu), the core read in place as its transpose
        return apply_small_core(T1, W, 1, (int64_t)o * r, M, s, l * i, o * r, b, (int64_t)l * i * b, b, 1, (int64_t)o * b * r, (int64_t)b * r, 1, r, r, st);
    syn_gemm_desc_t d2 = {s * b, o * r, l * i, 1, IX2((int64_t)l * i * b, 1, b), IX(b), IX(0), IX((int64_t)o * r), IX(1), IX(0),
                          IX2((int64_t)o * b * r, r, b), IX2((int64_t)b * r, 1, r), IX(0), 1.0, 0.0};
    return gemm_f64(d2, T1, W, M, st);
}

struct ApplyPlan {
    // per site: kept bond entering (s) and the buffer sizes it needs
    size_t t_doubles = 0, t1_doubles = 0, m_doubles = 0, trunc_bytes = 0;
};

static int plan_apply(int n, const int* xs, const int* wshape, int dim, int* out_shape, ApplyPlan& p) {
    int s = 1;
    for (int k = 0; k < n; k++) {
        const int a = xs[3 * k], i = xs[3 * k + 1], b = xs[3 * k + 2];
        const int l = wshape[4 * k], i2 = wshape[4 * k + 1], o = wshape[4 * k + 2], r = wshape[4 * k + 3];
        SYN_REQUIRE(a >= 1 && i >= 1 && b >= 1 && l >= 1 && o >= 1 && r >= 1 && i == i2, "chain sweep: bad core shapes at site %d", k);
        if (k > 0) SYN_REQUIRE(a == xs[3 * k - 1] && l == wshape[4 * k - 1], "chain sweep: bonds of site %d do not match site %d", k, k - 1);
        const int64_t rows = (int64_t)s * o, cols = (int64_t)b * r;
        SYN_REQUIRE(rows < INT_MAX && cols < INT_MAX && (int64_t)s * b < INT_MAX, "chain sweep: unfolding too large at site %d", k);
        p.t_doubles = std::max(p.t_doubles, (size_t)s * l * a);
        p.t1_doubles = std::max(p.t1_doubles, (size_t)s * l * i * b);
        p.m_doubles = std::max(p.m_doubles, (size_t)rows * cols);
        int kept;
        if (k < n - 1) {
            kept = (int)std::min<int64_t>(dim, rows);
            p.trunc_bytes = std::max(p.trunc_bytes, trunc_ws_bytes((int)rows, (int)cols, dim));
            p.t_doubles = std::max(p.t_doubles, (size_t)kept * cols);
        } else {
            kept = (int)cols;
        }
        if (out_shape) { out_shape[3 * k] = s; out_shape[3 * k + 1] = o; out_shape[3 * k + 2] = kept; }
        s = kept;
    }
    return 0;
}

static size_t apply_ws_bytes(const ApplyPlan& p, int n_sites) {
    return 2 * up16(p.t_doubles * 8) + up16(p.t1_doubles * 8) + up16(p.m_doubles * 8) + up16(p.trunc_bytes) + up16((size_t)n_sites * 64) + 256;
}

}  // namespace chain
}  // namespace syn

using namespace syn;
using namespace syn::chain;

extern "C" int syn_apply_round_chain_shapes(int n_sites, const int* xshape, const int* wshape, int dim, int* out_shape) {
    SYN_REQUIRE(n_sites >= 1 && xshape && wshape && out_shape && dim >= 1, "syn_apply_round_chain_shapes: bad arguments");
    ApplyPlan p;
    return plan_apply(n_sites, xshape, wshape, dim, out_shape, p);
}

extern "C" size_t syn_apply_round_chain_workspace_f64(int n_sites, const int* xshape, const int* wshape, int dim) {
    if (n_sites < 1 || !xshape || !wshape || dim < 1) return 0;
    ApplyPlan p;
    if (plan_apply(n_sites, xshape, wshape, dim, nullptr, p)) return 0;
    return apply_ws_bytes(p, n_sites);
}

// one left-to-right pass; speculative: no host wait per site, verdict slots of all sites in `infos`
static int apply_round_pass(int n_sites, const double* const* X, const int* xshape, const double* const* W, const int* wshape, int dim,
                            double* const* out, double* T, double* Tn, double* T1, double* M, void* tws, size_t tws_bytes, double* infos,
                            bool speculative, std::vector<char>& used, cudaStream_t st) {
    const double one = 1.0;
    SYN_CUDA(cudaMemcpyAsync(T, &one, sizeof(double), cudaMemcpyHostToDevice, st));
    int s = 1;
    for (int k = 0; k < n_sites; k++) {
        const int a = xshape[3 * k], i = xshape[3 * k + 1], b = xshape[3 * k + 2];
        const int l = wshape[4 * k], o = wshape[4 * k + 2], r = wshape[4 * k + 3];
        SYN_REQUIRE(X[k] && W[k] && out[k], "syn_apply_round_chain_f64: null core pointer at site %d", k);
        if (k == n_sites - 1) return contract_carry(T, s, X[k], a, i, b, W[k], l, o, r, T1, out[k], st);     // last core: the carry absorbed, no truncation
        if (int rc = contract_carry(T, s, X[k], a, i, b, W[k], l, o, r, T1, M, st)) return rc;
        const int rows = s * o, cols = b * r;
        int kept = 0;
        bool late = false;
        if (int rc = truncation_step(M, cols, rows, cols, dim, out[k], nullptr, tws, tws_bytes, &kept, st, speculative ? infos + (size_t)k * 8 : nullptr,
                                     &late))
            return rc;
        used[k] = late ? 1 : 0;
        // T_next[kept, r, b] = Q^T L with the columns of L (b major) re-ordered to (r, b) on the fly
        syn_gemm_desc_t dc = {kept, cols, rows, 1, IX(1), IX(kept), IX(0), IX(cols), IX(1), IX(0), IX((int64_t)r * b), IX2(1, b, r), IX(0), 1.0, 0.0};
        if (int rc = gemm_f64(dc, out[k], M, Tn, st)) return rc;
        std::swap(T, Tn);
        s = kept;
    }
    return 0;
}

extern "C" int syn_apply_round_chain_f64(int n_sites, const double* const* X, const int* xshape, const double* const* W, const int* wshape, int dim,
                                         double* const* out, void* ws, size_t ws_bytes, void* stream) {
    SYN_REQUIRE(n_sites >= 1 && X && xshape && W && wshape && out && ws && dim >= 1, "syn_apply_round_chain_f64: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    ApplyPlan p;
    if (int rc = plan_apply(n_sites, xshape, wshape, dim, nullptr, p)) return rc;
    SYN_REQUIRE(ws_bytes >= apply_ws_bytes(p, n_sites), "syn_apply_round_chain_f64: workspace too small (%zu < %zu bytes)", ws_bytes,
                apply_ws_bytes(p, n_sites));
    SYN_REQUIRE((((uintptr_t)ws) & 255) == 0, "syn_apply_round_chain_f64: the workspace must be 256-byte aligned");
    char* base = static_cast<char*>(ws);
    double* T = reinterpret_cast<double*>(base);            base += up16(p.t_doubles * 8);
    double* Tn = reinterpret_cast<double*>(base);           base += up16(p.t_doubles * 8);
    double* T1 = reinterpret_cast<double*>(base);           base += up16(p.t1_doubles * 8);
    double* M = reinterpret_cast<double*>(base);            base += up16(p.m_doubles * 8);
    void* tws = base;
    const size_t tws_bytes = up16(p.trunc_bytes);           base += tws_bytes;
    double* infos = reinterpret_cast<double*>(base);
    // Speculative pass: the Newton-Schulz steps are queued without waiting for their verdicts (the host wait per site left the GPU idle for
    // a launch latency 63 times a sweep); the verdicts of all sites are read once at the end.  A rejected site (dependent columns: the
    // iteration did not reach orthonormality) repeats the sweep with immediate verdicts and the Householder fallback.
    std::vector<char> used((size_t)n_sites, 0);
    if (int rc = apply_round_pass(n_sites, X, xshape, W, wshape, dim, out, T, Tn, T1, M, tws, tws_bytes, infos, true, used, st)) return rc;
    bool any = false;
    for (char u : used) any = any || u;
    if (!any) return 0;
    bool all_ok = true;
    if (int rc = late_verdicts(infos, used, n_sites, &all_ok, st)) return rc;
    if (all_ok) return 0;
    return apply_round_pass(n_sites, X, xshape, W, wshape, dim, out, T, Tn, T1, M, tws, tws_bytes, infos, false, used, st);
}

// ---- `>>` on a chain of cores (l, d, r), physical legs flattened ----------------------------------------------------------------------
static int plan_round(int n, const int* shape, int dim, int* out_shape, size_t& carry_doubles, size_t& trunc_bytes) {
    int s = shape[0];
    carry_doubles = 0;
    trunc_bytes = 0;
    for (int k = 0; k < n; k++) {
        const int l = shape[3 * k], d = shape[3 * k + 1], r = shape[3 * k + 2];
        SYN_REQUIRE(l >= 1 && d >= 1 && r >= 1, "chain sweep: bad core shape at site %d", k);
        if (k > 0) SYN_REQUIRE(l == shape[3 * k - 1], "chain sweep: bonds of site %d do not match site %d", k, k - 1);
        const int64_t rows = (int64_t)s * d;
        SYN_REQUIRE(rows < INT_MAX && (int64_t)d * r < INT_MAX, "chain sweep: unfolding too large at site %d", k);
        int kept = r;
        if (k < n - 1) {
            kept = (int)std::min<int64_t>(dim, rows);
            trunc_bytes = std::max(trunc_bytes, trunc_ws_bytes((int)rows, r, dim));
            carry_doubles = std::max(carry_doubles, (size_t)kept * r);
            const int dn = shape[3 * k + 4], rn = shape[3 * k + 5];
            carry_doubles = std::max(carry_doubles, (size_t)kept * dn * rn);           // the next core with the carry absorbed
        }
        if (out_shape) { out_shape[3 * k] = s; out_shape[3 * k + 1] = d; out_shape[3 * k + 2] = kept; }
        s = kept;
    }
    return 0;
}

extern "C" int syn_round_chain_shapes(int n_sites, const int* shape, int dim, int* out_shape) {
    SYN_REQUIRE(n_sites >= 1 && shape && out_shape && dim >= 1, "syn_round_chain_shapes: bad arguments");
    size_t c, t;
    return plan_round(n_sites, shape, dim, out_shape, c, t);
}

extern "C" size_t syn_round_chain_workspace_f64(int n_sites, const int* shape, int dim) {
    if (n_sites < 1 || !shape || dim < 1) return 0;
    size_t c, t;
    if (plan_round(n_sites, shape, dim, nullptr, c, t)) return 0;
    return 3 * up16(c * 8) + up16(t) + up16((size_t)n_sites * 64) + 256;
}

static int round_pass(int n_sites, const double* const* cores, const int* shape, int dim, double* const* out, double* S, double* Ca, double* Cb,
                      void* tws, size_t tws_bytes, double* infos, bool speculative, std::vector<char>& used, cudaStream_t st) {
    const double* cur = cores[0];                   // (s * d) x r unfolding of the current core, contiguous
    int s = shape[0];
    for (int k = 0; k < n_sites; k++) {
        const int d = shape[3 * k + 1], r = shape[3 * k + 2];
        SYN_REQUIRE(cores[k] && out[k], "syn_round_chain_f64: null core pointer at site %d", k);
        if (k == n_sites - 1) {
            SYN_CUDA(cudaMemcpyAsync(out[k], cur, sizeof(double) * (size_t)s * d * r, cudaMemcpyDeviceToDevice, st));
            return 0;
        }
        const int rows = s * d;
        int kept = 0;
        bool late = false;
        if (int rc = truncation_step(cur, r, rows, r, dim, out[k], S, tws, tws_bytes, &kept, st, speculative ? infos + (size_t)k * 8 : nullptr, &late))
            return rc;
        used[k] = late ? 1 : 0;
        // next core <- S (kept x r) @ next.reshape(r, d' r')
        const int dn = shape[3 * k + 4], rn = shape[3 * k + 5];
        double* nxt = (cur == Ca) ? Cb : Ca;
        syn_gemm_desc_t dg = {kept, dn * rn, r, 1, IX(r), IX(1), IX(0), IX((int64_t)dn * rn), IX(1), IX(0), IX((int64_t)dn * rn), IX(1), IX(0), 1.0, 0.0};
        if (int rc = gemm_f64(dg, S, cores[k + 1], nxt, st)) return rc;
        cur = nxt;
        s = kept;
    }
    return 0;
}

extern "C" int syn_round_chain_f64(int n_sites, const double* const* cores, const int* shape, int dim, double* const* out, void* ws, size_t ws_bytes,
                                   void* stream) {
    SYN_REQUIRE(n_sites >= 1 && cores && shape && out && ws && dim >= 1, "syn_round_chain_f64: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    size_t cd, tb;
    if (int rc = plan_round(n_sites, shape, dim, nullptr, cd, tb)) return rc;
    SYN_REQUIRE(ws_bytes >= 3 * up16(cd * 8) + up16(tb) + up16((size_t)n_sites * 64) + 256, "syn_round_chain_f64: workspace too small");
    SYN_REQUIRE((((uintptr_t)ws) & 255) == 0, "syn_round_chain_f64: the workspace must be 256-byte aligned");
    char* base = static_cast<char*>(ws);
    double* S = reinterpret_cast<double*>(base);            base += up16(cd * 8);
    double* Ca = reinterpret_cast<double*>(base);           base += up16(cd * 8);      // current core with the carry absorbed (ping)
    double* Cb = reinterpret_cast<double*>(base);           base += up16(cd * 8);      // (pong)
    void* tws = base;
    const size_t tws_bytes = up16(tb);                      base += tws_bytes;
    double* infos = reinterpret_cast<double*>(base);
    std::vector<char> used((size_t)n_sites, 0);             // speculative pass first, see syn_apply_round_chain_f64
    if (int rc = round_pass(n_sites, cores, shape, dim, out, S, Ca, Cb, tws, tws_bytes, infos, true, used, st)) return rc;
    bool any = false;
    for (char u : used) any = any || u;
    if (!any) return 0;
    bool all_ok = true;
    if (int rc = late_verdicts(infos, used, n_sites, &all_ok, st)) return rc;
    if (all_ok) return 0;
    return round_pass(n_sites, cores, shape, dim, out, S, Ca, Cb, tws, tws_bytes, infos, false, used, st);
}

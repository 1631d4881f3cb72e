"""GPU: the strided DMMA GEMM against torch.matmul (cuBLAS FP64) on the same inputs.  Tolerance: 1e-13 of
|A||B| (FP64 accumulation-order differences only)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _check(out, ref, a, b, tol=1e-13):
    scale = (a.abs().max() * b.abs().max() * max(a.shape[-1], 1)).item() + 1e-300
    err = (out - ref).abs().max().item() / scale
    assert err < tol, err


@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (7, 5, 3), (64, 64, 16), (128, 128, 64), (130, 70, 33), (512, 4096, 512),
                                   (256, 256, 4096), (1000, 24, 129), (2, 2, 2), (33, 257, 1)])
@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
def test_gemm_layouts(M, N, K, ta, tb):
    from syngular_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M * 131 + N * 17 + K)
    a = torch.randn((K, M) if ta else (M, K), dtype=torch.float64, device="cuda", generator=g)
    b = torch.randn((N, K) if tb else (K, N), dtype=torch.float64, device="cuda", generator=g)
    av = a.t() if ta else a
    bv = b.t() if tb else b
    out = ops.matmul(av, bv)
    _check(out, av @ bv, a, b)
    c0 = torch.randn((M, N), dtype=torch.float64, device="cuda", generator=g)
    out2 = ops.matmul(av, bv, out=c0.clone(), alpha=-0.5, beta=2.0)
    _check(out2, -0.5 * (av @ bv) + 2.0 * c0, a, b, tol=1e-12)


def test_gemm_batched_and_strided_views():
    from syngular_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randn((6, 40, 24), dtype=torch.float64, device="cuda", generator=g)
    b = torch.randn((6, 24, 50), dtype=torch.float64, device="cuda", generator=g)
    _check(ops.matmul(a, b), a @ b, a, b)
    # non-contiguous views: column slices, odd offsets (forces the scalar-load path)
    big = torch.randn((300, 301), dtype=torch.float64, device="cuda", generator=g)
    av = big[3:140, 1:98]
    bv = big[5:102, 7:190]
    _check(ops.matmul(av, bv), av @ bv, av, bv)
    out = torch.zeros((200, 333), dtype=torch.float64, device="cuda")
    ov = out[1:138, 3:186]
    ops.matmul(av, bv, out=ov)
    _check(ov, av @ bv, av, bv)
    assert out[0].abs().max().item() == 0.0 and out[:, :3].abs().max().item() == 0.0


def test_gemm_two_level_indices_site_contraction():
    """MPO x MPS site contraction C[(a,l),o,(b,r)] = sum_i X[a,i,b] W[l,i,o,r] (matrix_product_operator.py:184-190)
    expressed as ONE strided GEMM with two-level indices: m=(a,l)... here done as m=(a), n=(l,o,r) per b."""
    from syngular_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(7)
    a_, i_, b_, l_, o_, r_ = 6, 4, 5, 3, 2, 7
    X = torch.randn((a_, i_, b_), dtype=torch.float64, device="cuda", generator=g)
    W = torch.randn((l_, i_, o_, r_), dtype=torch.float64, device="cuda", generator=g)
    ref = torch.einsum("aib,lior->alobr", X, W).contiguous()
    out = torch.empty_like(ref)
    # GEMM: M = (a, b) two-level rows of X, K = i, N = (l, o, r) -> needs 3 levels; do batch over l.
    # rows m=(a,b): X offset a*(i*b) + b ; out offset a*(l*o*b*r) + b*r
    # cols n=(o,r): W offset o*r_ + r (contiguous) ; out offset o*(b*r) + r
    ops.gemm(X, W, out, M=a_ * b_, N=o_ * r_, K=i_,
             a_m=(i_ * b_, 1, b_), a_k=b_,
             b_k=o_ * r_, b_n=1,
             c_m=(l_ * o_ * b_ * r_, r_, b_), c_n=(b_ * r_, 1, r_),
             batch=l_, a_b=0, b_b=i_ * o_ * r_, c_b=o_ * b_ * r_)
    assert (out - ref).abs().max().item() < 1e-13


@pytest.mark.parametrize("M,N,K", [(512, 4096, 128), (512, 4096, 4096), (256, 8232, 64), (128, 16464, 96)])
def test_gemm_128x112_wave_configuration(M, N, K):
    """Shapes where 128 x 112 tiles fill the 148 SMs in fewer waves than 128 x 128 (csrc/gemm.cu: CfgT, TMA-staged, the last column tile
    hangs over the edge: 4096 = 36 * 112 + 64, 8232 = 73 * 112 + 56): the M.E product of the density-matrix sweep takes this path."""
    from syngular_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn((M, K), dtype=torch.float64, device="cuda", generator=g)
    b = torch.randn((K, N), dtype=torch.float64, device="cuda", generator=g)
    _check(ops.matmul(a, b), a @ b, a, b)
    c0 = torch.randn((M, N), dtype=torch.float64, device="cuda", generator=g)
    _check(ops.matmul(a, b, out=c0.clone(), alpha=0.25, beta=-1.0), 0.25 * (a @ b) - c0, a, b, tol=1e-12)
    # the sweep's own call: output columns permuted on the fly by a two-level C index (gram_with_environment)
    if N == 4096:
        out = torch.empty((M, N), dtype=torch.float64, device="cuda")
        ops.gemm(a, b, out, M=M, N=N, K=K, a_m=K, a_k=1, b_k=N, b_n=1, c_m=N, c_n=(1, 16, 256))
        ref = (a @ b).reshape(M, 16, 256).transpose(1, 2).reshape(M, N)
        _check(out, ref, a, b)


@pytest.mark.parametrize("M,N,K,batch,mr,mc", [(1024, 256, 96, 3, 256, 64), (4096, 256, 64, 2, 1024, 64), (512, 192, 40, 1, 128, 64), (320, 256, 33, 2, 64, 64)])
def test_gemm_block_lower_mask(M, N, K, batch, mr, mc):
    """Block-lower output mask (syn_gemm_desc_t.mask_rows / mask_cols): only n < (m // rows + 1) * cols is computed, the rest of C keeps its
    contents; 128 x 64 tiles when the row blocks allow, 64 x 64 otherwise."""
    from syngular_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn((batch, M, K), dtype=torch.float64, device="cuda", generator=g)
    b = torch.randn((batch, K, N), dtype=torch.float64, device="cuda", generator=g)
    out = torch.full((batch, M, N), 7.0, dtype=torch.float64, device="cuda")
    ops.gemm(a, b, out, M, N, K, K, 1, N, 1, N, 1, batch=batch, a_b=M * K, b_b=K * N, c_b=M * N, mask=(mr, mc))
    rows = torch.arange(M, device="cuda")[:, None]
    cols = torch.arange(N, device="cuda")[None, :]
    keep = cols < (rows // mr + 1) * mc
    ref = torch.where(keep, a @ b, torch.full_like(out, 7.0))
    _check(out, ref, a, b)


def test_gemm_half_width_tiles_for_few_tiles():
    """256 x 4096 x 512 (the carry product U^T M of the sweeps): 64 tiles of 128 x 128 -> 128 tiles of 128 x 64."""
    from syngular_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    q = torch.randn((512, 256), dtype=torch.float64, device="cuda", generator=g)
    l = torch.randn((512, 4096), dtype=torch.float64, device="cuda", generator=g)
    _check(ops.matmul(q.t(), l), q.t() @ l, q, l)
    a = torch.randn((256, 512), dtype=torch.float64, device="cuda", generator=g)
    _check(ops.matmul(a, l), a @ l, a, l)

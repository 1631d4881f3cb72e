"""GPU: TensorDense forward (BASELINE configs[4]) against the numpy restatement of layers/TensorDense.py:103-142.
FP64 on the device; tolerance 1e-10 against the float64 restatement."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ins,outs,bonds,batch", [((4, 5, 3), (3, 2, 4), (2, 3), 7), ((3, 2, 4, 2), (2, 3, 2, 2), (2, 4, 3), 5),
                                                   ((16, 16, 16), (16, 16, 16), (16, 16), 33), ((6,), (5,), (), 4), ((8, 4), (2, 6), (3,), 9)])
def test_tensordense_forward_matches_restatement(ins, outs, bonds, batch):
    from syngular.layers import TensorDense
    from oracle import tensordense_numpy as TD
    rng = np.random.default_rng(len(ins) * 10 + batch)
    layer = TensorDense(ins, outs, bonds)
    cores = [rng.normal(size=s) for s in layer.core_shapes()]
    bias = rng.normal(size=outs)
    layer.build(cores, bias)
    x = rng.normal(size=(batch, int(np.prod(ins))))
    got = layer(x).cpu().numpy()
    ref = TD.forward(x, cores, bias, "relu")
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)) < 1e-10 * max(1.0, np.max(np.abs(ref)))
    lin = TensorDense(ins, outs, bonds, activation=None, use_bias=False).build(cores)
    ref2 = TD.forward(x, cores, None, None)
    assert np.max(np.abs(lin(x, chunk=3).cpu().numpy() - ref2)) < 1e-10 * max(1.0, np.max(np.abs(ref2)))     # chunked path


# ---- TF32 path: the fused tcgen05 kernel (csrc/ttdense.cu) for BASELINE configs[4]'s layer -------------------------------------------
TF32_TOL = 4e-3        # of the output's largest magnitude: TF32 keeps 10 mantissa bits of every operand, three chained contractions


@pytest.mark.parametrize("batch", [1, 2, 75, 300, 1111])
def test_tensordense_tf32_fused_kernel_matches_restatement(batch):
    """float32 in / out on tcgen05.mma.kind::tf32 against the float64 restatement of the reference's einsum (TensorDense.py:103-142) on the
    SAME float32 weights and inputs; batch sizes cover one sample per CTA, ragged sample counts and several samples per CTA."""
    import torch
    from syngular.layers import TensorDense
    from oracle import tensordense_numpy as TD
    rng = np.random.default_rng(batch)
    shape = (16, 16, 16)
    layer = TensorDense(shape, shape, (16, 16), precision="tf32")
    cores = [rng.normal(scale=0.05, size=s).astype(np.float32) for s in layer.core_shapes()]
    bias = (0.01 * rng.normal(size=shape)).astype(np.float32)
    layer.build(cores, bias)
    x = rng.normal(size=(batch, 4096)).astype(np.float32)
    got = layer(x)
    assert got.dtype == torch.float32 and tuple(got.shape) == (batch, 4096)
    ref = TD.forward(x.astype(np.float64), [c.astype(np.float64) for c in cores], bias.astype(np.float64), "relu")
    assert np.max(np.abs(got.cpu().numpy() - ref)) < TF32_TOL * np.max(np.abs(ref))
    lin = TensorDense(shape, shape, (16, 16), activation=None, use_bias=False, precision="tf32").build(cores)
    ref2 = TD.forward(x.astype(np.float64), [c.astype(np.float64) for c in cores], None, None)
    assert np.max(np.abs(lin(x).cpu().numpy() - ref2)) < TF32_TOL * np.max(np.abs(ref2))


def test_tensordense_tf32_is_exact_on_exactly_representable_data():
    """Identity-like cores and small-integer inputs are exact in TF32: any index permutation inside the kernel's re-layouts would show."""
    from syngular.layers import TensorDense
    shape = (16, 16, 16)
    I1 = np.zeros((16, 16, 16)); I2 = np.zeros((16, 16, 16, 16)); I3 = np.zeros((16, 16, 16))
    for i in range(16):
        I1[i, i, 0] = 1; I2[i, i, 0, 0] = 1; I3[i, i, 0] = 1
    layer = TensorDense(shape, shape, (16, 16), activation=None, use_bias=False, precision="tf32").build([I1, I2, I3])
    x = (np.arange(5 * 4096) % 1021).astype(np.float32).reshape(5, 4096)
    assert np.array_equal(layer(x).cpu().numpy(), x)
    # a permutation layer: out (o1,o2,o3) = in (i1,i2,i3) with every mode index reversed
    P1 = np.zeros((16, 16, 16)); P2 = np.zeros((16, 16, 16, 16)); P3 = np.zeros((16, 16, 16))
    for i in range(16):
        P1[i, 15 - i, 3] = 1; P2[i, 15 - i, 3, 7] = 1; P3[i, 15 - i, 7] = 1
    layer = TensorDense(shape, shape, (16, 16), activation=None, use_bias=False, precision="tf32").build([P1, P2, P3])
    want = x.reshape(5, 16, 16, 16)[:, ::-1, ::-1, ::-1].reshape(5, 4096)
    assert np.array_equal(layer(x).cpu().numpy(), want)


@pytest.mark.parametrize("batch", [1, 3, 75, 1111])
def test_tensordense_tf32_cta_pair_variant_matches_restatement_and_the_default_kernel(batch):
    """syn_tt_dense3_tf32_pair (tcgen05 cta_group::2; two CTAs share a sample) against the float64 restatement, and against the single-CTA
    kernel: both accumulate the same TF32 products in FP32, in different orders."""
    import torch
    from syngular_b200 import ops
    from oracle import tensordense_numpy as TD
    rng = np.random.default_rng(100 + batch)
    cores = [rng.normal(scale=0.05, size=s).astype(np.float32) for s in ((16, 16, 16), (16, 16, 16, 16), (16, 16, 16))]
    bias = (0.01 * rng.normal(size=4096)).astype(np.float32)
    x = rng.normal(size=(batch, 4096)).astype(np.float32)
    dev = torch.device("cuda")
    packed = ops.tt_dense3_pack(*[torch.from_numpy(c).to(dev) for c in cores])
    xd, bd = torch.from_numpy(x).to(dev), torch.from_numpy(bias).to(dev)
    got = ops.tt_dense3_tf32(xd, packed, bd, relu=True, pair=True).cpu().numpy()
    one = ops.tt_dense3_tf32(xd, packed, bd, relu=True).cpu().numpy()
    ref = TD.forward(x.astype(np.float64), [c.astype(np.float64) for c in cores], bias.astype(np.float64).reshape(16, 16, 16), "relu")
    assert np.max(np.abs(got - ref)) < TF32_TOL * np.max(np.abs(ref))
    assert np.max(np.abs(got - one)) < 1e-5 * np.max(np.abs(ref))
    # exactly representable data: a permutation layer must come out bit-exact
    P1 = np.zeros((16, 16, 16), np.float32); P2 = np.zeros((16, 16, 16, 16), np.float32); P3 = np.zeros((16, 16, 16), np.float32)
    for i in range(16):
        P1[i, 15 - i, 3] = 1; P2[i, (i + 5) % 16, 3, 7] = 1; P3[i, 15 - i, 7] = 1
    pk = ops.tt_dense3_pack(*[torch.from_numpy(c).to(dev) for c in (P1, P2, P3)])
    xi = (np.arange(batch * 4096) % 1021).astype(np.float32).reshape(batch, 4096)
    want = np.roll(xi.reshape(batch, 16, 16, 16)[:, ::-1, :, ::-1], 5, axis=2).reshape(batch, 4096)
    assert np.array_equal(ops.tt_dense3_tf32(torch.from_numpy(xi).to(dev), pk, None, relu=False, pair=True).cpu().numpy(), want)
    # run-to-run bit identity: the two re-layout writers share the operand ring through mbarrier / tcgen05.commit hand-overs that
    # compute-sanitizer's racecheck does not model (it reports write-after-write hazards there); a real race would show up here
    for _ in range(10):
        assert np.array_equal(ops.tt_dense3_tf32(xd, packed, bd, relu=True, pair=True).cpu().numpy(), got)


def test_tensordense_tf32_rejects_uncovered_shapes():
    from syngular.layers import TensorDense
    with pytest.raises(NotImplementedError):
        TensorDense((8, 8, 8), (8, 8, 8), (4, 4), precision="tf32")


def test_tensordense_tf32_streamed_host_path_equals_the_device_path():
    """layer(x_host, out=y_host) with pinned host tensors: pieces are uploaded, computed and downloaded on three overlapping streams; the
    result is bit-identical to the one-shot device call (same kernel, same per-sample arithmetic), for a ragged last piece too."""
    import torch
    from syngular.layers import TensorDense
    shape = (16, 16, 16)
    layer = TensorDense(shape, shape, (16, 16), seed=11, precision="tf32").build()
    g = torch.Generator().manual_seed(3)
    x = torch.randn((1000, 4096), dtype=torch.float32, generator=g).pin_memory()
    y = torch.empty_like(x).pin_memory()
    want = layer(x.cuda()).cpu()
    for chunk in (128, 300, 4096):                    # several pieces, a ragged tail, one piece
        y.zero_()
        got = layer(x, chunk=chunk, out=y)
        torch.cuda.synchronize()
        assert got is y and torch.equal(y, want)
    with pytest.raises(ValueError):
        layer(torch.randn(4, 4096), out=torch.empty(4, 4096))        # not pinned

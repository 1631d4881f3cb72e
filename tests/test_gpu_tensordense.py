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

"""GPU: the reference's README script (readme.md:42-74) run verbatim through the drop-in package, FROM THE DENSE TENSORS
(decompose / random included), against the values the unmodified reference produced with np.random.seed(0)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_readme_example_end_to_end():
    g = np.load(os.path.join(ROOT, "tests", "golden", "readme_chain.npz"))
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "readme_example.py")], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.strip()]
    xu, zx = float(lines[0]), float(lines[1])
    assert abs(xu - float(g["X_U"])) < 1e-9 * abs(float(g["X_U"]))
    assert abs(zx - float(g["Z_X"])) < 1e-8 * abs(float(g["Z_X"]))
    diag = np.array([float(t) for t in " ".join(lines[2:]).replace("[", " ").replace("]", " ").split()])
    assert diag.shape == (32,) and np.max(np.abs(diag - 1.0)) < 1e-10


def test_error_behaviour_matches_the_reference():
    """Same exception messages as the reference for misuse of the operators."""
    from syngular.tensor import MatrixProductOperator as MPO, MatrixProductState as MPS
    import syngular as syn
    x = np.arange(8, dtype=float).reshape(2, 2, 2)
    X = MPS(x, bond_shape=(2, 2)).decompose()
    with pytest.raises(Exception, match="dimension should be an integer"):
        X >> 2.0                                                            # MPS:139-143
    with pytest.raises(Exception, match="right-hand site must be a MatrixProductState"):
        X | 3                                                               # MPS:129
    with pytest.raises(Exception, match="input indices do not match the number of sites"):
        X[(0, 1)]                                                           # MPS:157-159
    with pytest.raises(Exception, match="dimensions of bond indices do not match order - 1"):
        MPS(x, bond_shape=(2,))                                             # MPS:50-51
    with pytest.raises(Exception, match="canonical form"):
        X + MPS(x, bond_shape=(2, 2))                                       # MPS:102 (other not decomposed)
    w = np.arange(16, dtype=float).reshape(2, 2, 2, 2)
    with pytest.raises(Exception, match="input_shape and output_shape of the tensor must have the same length"):
        MPO(np.zeros((2, 2, 2)), bond_shape=(2,))                           # MPO:43-44
    with pytest.raises(Exception, match="input_shape and output_shape of the tensor must have the same length"):
        MPO(w, bond_shape=(2, 2))                                           # MPO:43-44 fires first, exactly as in the reference
    Wm = MPO(w, bond_shape=(2,)).decompose()
    with pytest.raises(Exception, match="output indices do not match the number of sites"):
        Wm[(0, 1), (0,)]                                                    # MPO:339-340
    with pytest.raises(Exception, match="same number of sites"):
        syn.mul(Wm, X)                                                      # tensor/utils.py:15-16
    # reference metadata attributes exist with the reference's meaning
    Y = MPS(np.arange(64, dtype=float).reshape(4, 4, 4), bond_shape=(4, 4)).decompose()
    assert Y.decomposed and Y.sites_number == 3 and Y.input_shape == (4, 4, 4) and Y.real_parameters_number == 64
    assert Y.parameters_number == sum(int(np.prod(s)) for s in Y.shape) and Y.orthonormalized is None
    Z = Y >> 2
    assert Z is not Y and Z.bond_shape == (4, 4) and [s[2] for s in Z.shape[:-1]] == [2, 2]        # stale bond_shape, fresh shape

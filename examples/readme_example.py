"""The reference's README example (readme.md:42-74), unchanged except for the import path: runs on the GPU."""
import numpy as np

from syngular.tensor import MatrixProductState
from syngular.tensor import MatrixProductOperator

np.random.seed(0)
tensor_W = np.arange(16**6).reshape((16, 16, 16, 16, 16, 16))
tensor_X = np.arange(16**3).reshape((16, 16, 16))

W = MatrixProductOperator(tensor_W, bond_shape=(16, 16,))
W.decompose()

X = MatrixProductState(tensor_X, bond_shape=(4, 4,))
X.decompose()

T = MatrixProductOperator.random((16, 16, 16), (16, 16, 16), (8, 8,))
O = MatrixProductOperator.zeros((16, 16, 16), (16, 16, 16), (4, 4,))
U = MatrixProductState.random((16, 16, 16), (8, 8,))

W = W >> 4
T = T >> 2

Z = ((T + W) @ T) @ X

print(X | U)
print(Z | X)

Z = Z >> 16
Z.left_orthonormalization()

print(np.diag(Z.left_orthogonality(0)))
print(np.diag(Z.left_orthogonality(1)))

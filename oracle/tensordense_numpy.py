"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the TensorDense forward pass (layers/TensorDense.py:103-142).

TensorFlow is not installed, so the Keras layer cannot run here: PARITY UNPINNED by the reference for this row; the
restatement is the one-line einsum the reference hands to opt_einsum (for 3 cores, the only case its index labels are valid for,
TensorDense.py:106-118) generalised to N cores, plus `+ bias` (:139) and the activation (:142)."""
import numpy as np


def forward(x, cores, bias=None, activation="relu", tt_input_shape=None):
    """x: (batch, prod(inputs)); cores in the reference layouts: (i0,o0,b0), (i,o,bl,br)..., (i,o,bl)."""
    n = len(cores)
    ins = tuple(c.shape[0] for c in cores) if tt_input_shape is None else tuple(tt_input_shape)
    full = []
    for k, c in enumerate(cores):
        if n == 1:
            full.append(c.reshape(c.shape[0], c.shape[1], 1, 1))
        elif k == 0:
            full.append(c.reshape(c.shape[0], c.shape[1], 1, c.shape[2]))
        elif k == n - 1:
            full.append(c.reshape(c.shape[0], c.shape[1], c.shape[2], 1))
        else:
            full.append(c)
    T = x.reshape((x.shape[0], 1) + ins)                       # (batch, b, i_1..i_n)
    for k, c in enumerate(full):
        # contract b and i_k (axes 1 and 2+k-k... always the first input axis left) -> append o_k, new b
        T = np.tensordot(T, c, axes=([1, 2], [2, 0]))           # (batch, rest inputs..., [outs so far...], o_k, b_r)
        T = np.moveaxis(T, -1, 1)                               # b to axis 1
    # axes now: (batch, 1, o_1..o_n) because each step removed the first remaining input axis and appended o_k
    y = T.reshape(x.shape[0], -1)
    if bias is not None:
        y = y + bias.reshape(-1)
    if activation == "relu":
        y = np.maximum(y, 0)
    return y

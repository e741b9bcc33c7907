"""oracle/mlp_oracle.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

numpy float64 forward/backward of the encoder stack ``Linear -> LeakyReLU(slope) -> ... -> Linear``
(reference /root/reference/encoders.py:36-58: no activation after the last Linear; ``nn.LeakyReLU()``
default slope 0.01).  ``y = x W^T + b``; LeakyReLU'(z) = 1 for z > 0 else slope (torch's
``leaky_relu_backward`` uses ``x > 0``).
"""
import numpy as np


def mlp_forward(x, weights, biases, slope=0.01):
    """Returns (output, list of layer inputs a_0..a_{L-1}, list of pre-activations z_0..z_{L-1})."""
    a = np.asarray(x, dtype=np.float64)
    acts, pre = [], []
    L = len(weights)
    for li, (W, b) in enumerate(zip(weights, biases)):
        acts.append(a)
        z = a @ np.asarray(W, dtype=np.float64).T
        if b is not None:
            z = z + np.asarray(b, dtype=np.float64)
        pre.append(z)
        a = z if li == L - 1 else np.where(z > 0, z, slope * z)
    return a, acts, pre


def mlp_backward(gy, weights, acts, pre, slope=0.01, need_dx=False):
    """Gradients (dW list, db list, dx or None) for upstream gradient ``gy`` of the output."""
    L = len(weights)
    dz = np.asarray(gy, dtype=np.float64)
    dWs, dbs = [None] * L, [None] * L
    dx = None
    for li in range(L - 1, -1, -1):
        dWs[li] = dz.T @ acts[li]
        dbs[li] = dz.sum(axis=0)
        if li > 0 or need_dx:
            da = dz @ np.asarray(weights[li], dtype=np.float64)
            if li > 0:
                dz = da * np.where(pre[li - 1] > 0, 1.0, slope)
            else:
                dx = da
    return dWs, dbs, dx

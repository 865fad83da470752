"""Drop-in for ``simkit.lbs_jacobian`` (lbs_jacobian.py:12-47): ``d x / d W`` of linear blend skinning in stacked form,
``J[(v, i), (k, a, i)] = W[v, k] * [V_v, 1]_a``.  A dense Kronecker layout product on the host (set-up code; the basis it
returns is what the reduced tier keeps resident on the device with ``MeshPlan.set_basis``)."""

import numpy as np


def lbs_jacobian(V: np.ndarray, W: np.ndarray) -> np.ndarray:
    V = np.asarray(V, dtype=np.float64)
    W = np.asarray(W, dtype=np.float64)
    n, d = V.shape
    k = W.shape[1]
    V1 = np.hstack((V, np.ones((n, 1))))                       # homogeneous rest positions
    # column (bone, a): weight of the bone times homogeneous coordinate a
    cols = (W[:, :, None] * V1[:, None, :]).reshape(n, k * (d + 1))
    return np.kron(cols, np.identity(d))

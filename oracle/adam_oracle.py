"""ORACLE (test infrastructure, NOT product code): numpy restatement of the reference's sparse Adam step,
LoG/model/sparse_optimizer.py:41-78 (_single_tensor_adam) applied to gathered rows and scattered back (:163-196).
PARITY: pinned -- checked against tests/golden/reference_sparse_adam.npz, produced by running the reference's own
function (tests/golden/make_golden.py)."""
import math

import numpy as np


def sparse_adam_step(param, grad, exp_avg, exp_avg_sq, index, step, lr, max_exp_avg_sq=None, beta1=0.9, beta2=0.999, eps=1e-15):
    f = np.float32
    param, exp_avg, exp_avg_sq = param.copy(), exp_avg.copy(), exp_avg_sq.copy()
    vmax = None if max_exp_avg_sq is None else max_exp_avg_sq.copy()
    g = grad.astype(f)
    m = exp_avg[index] * f(beta1) + g * f(1 - beta1)                     # :53
    v = exp_avg_sq[index] * f(beta2) + (f(1 - beta2) * g) * g            # :54
    bc1 = 1 - beta1 ** int(step)                                         # :64
    bc2 = 1 - beta2 ** int(step)
    step_size = lr / bc1
    bc2_sqrt = math.sqrt(bc2)
    vv = v
    if vmax is not None:
        vv = np.maximum(vmax[index], v)                                  # :72
        vmax[index] = vv
    denom = np.sqrt(vv) / f(bc2_sqrt) + f(eps)                           # :73,75
    param[index] = param[index] + f(-step_size) * (m / denom)            # :76
    exp_avg[index], exp_avg_sq[index] = m, v
    return param, exp_avg, exp_avg_sq, vmax

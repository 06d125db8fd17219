"""Fused sparse Adam step -- host side of `lgr_sparse_adam`.

Mirrors what `SparseOptimizer.step` does per parameter in the reference (LoG/model/sparse_optimizer.py:163-196):
gather the optimiser state of the visible rows, `_single_tensor_adam` (:41-78), scatter parameter and state back --
as one in-place kernel, without the `index.cpu()` synchronisation (:168)."""
import ctypes

import torch

from . import _capi


def sparse_adam_step_(param, grad, exp_avg, exp_avg_sq, index, step, lr, max_exp_avg_sq=None, beta1=0.9, beta2=0.999,
                      eps=1e-15):
    """In place.  param / exp_avg / exp_avg_sq [/ max_exp_avg_sq]: (N, ...) float32 CUDA, contiguous;
    index: (K,) int64 unique rows; grad: (K, ...) gradient of the gathered rows param[index]."""
    lib = _capi.load()
    for name, t in (('param', param), ('grad', grad), ('exp_avg', exp_avg), ('exp_avg_sq', exp_avg_sq)):
        _capi.require_cuda(t, name)
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise TypeError(f'{name} must be contiguous float32')
    if index.dtype != torch.int64 or not index.is_contiguous():
        raise TypeError('index must be contiguous int64')
    k = int(index.shape[0])
    c = int(param[0].numel()) if param.shape[0] else 1
    if tuple(grad.shape) != (k,) + tuple(param.shape[1:]):
        raise ValueError(f'grad shape {tuple(grad.shape)} does not match ({k},) + {tuple(param.shape[1:])}')
    p = lambda t: None if t is None or t.numel() == 0 else ctypes.c_void_p(t.data_ptr())
    _capi.check(lib.lgr_sparse_adam(k, c, p(index), p(grad), p(param), p(exp_avg), p(exp_avg_sq), p(max_exp_avg_sq),
                                    int(step), float(lr), float(beta1), float(beta2), float(eps),
                                    _capi.current_stream()), 'lgr_sparse_adam')
    return param

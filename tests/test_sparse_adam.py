"""Sparse Adam: oracle and kernel against golden vectors produced by RUNNING the reference's
`_single_tensor_adam` (LoG/model/sparse_optimizer.py:41-78; generator: tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import adam_oracle

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_sparse_adam.npz'))
CASES = sorted({k.rsplit('_', 2)[0] + '_' for k in G.files if k.endswith('_param_in')})


def close(got, want, scale):
    """|got - want| <= 4 ulp of the operands' magnitude (the reference may or may not fuse a*b+c into one rounding) or
    2e-6 relative."""
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=4 * 6e-8 * float(scale))


def case(prefix):
    step, lr, b1, b2, eps, ams = G[prefix + 'hyper']
    return dict(step=int(step), lr=float(lr), beta1=float(b1), beta2=float(b2), eps=float(eps)), bool(ams)


@pytest.mark.parametrize('prefix', CASES)
def test_oracle_matches_reference(prefix):
    hyper, ams = case(prefix)
    p, m, v, vm = adam_oracle.sparse_adam_step(G[prefix + 'param_in'], G[prefix + 'grad'], G[prefix + 'm_in'], G[prefix + 'v_in'],
                                               G[prefix + 'index'], max_exp_avg_sq=G[prefix + 'vmax_in'] if ams else None, **hyper)
    gmax = np.abs(G[prefix + 'grad']).max()
    close(p, G[prefix + 'param_out'], np.abs(G[prefix + 'param_in']).max())
    close(m, G[prefix + 'm_out'], max(np.abs(G[prefix + 'm_in']).max(), gmax))
    close(v, G[prefix + 'v_out'], max(np.abs(G[prefix + 'v_in']).max(), gmax ** 2))
    if ams:
        close(vm, G[prefix + 'vmax_out'], max(np.abs(G[prefix + 'vmax_in']).max(), gmax ** 2))


@pytest.mark.gpu
@pytest.mark.parametrize('prefix', CASES)
def test_kernel_matches_reference(built, prefix):
    from log_b200.optim import sparse_adam_step_
    hyper, ams = case(prefix)
    from util import device
    dev = device()
    t = lambda k: torch.from_numpy(G[prefix + k]).to(dev)
    p, m, v, vm = t('param_in'), t('m_in'), t('v_in'), t('vmax_in')
    sparse_adam_step_(p, t('grad'), m, v, t('index'), hyper['step'], hyper['lr'], max_exp_avg_sq=vm if ams else None,
                      beta1=hyper['beta1'], beta2=hyper['beta2'], eps=hyper['eps'])
    gmax = np.abs(G[prefix + 'grad']).max()
    close(p.cpu().numpy(), G[prefix + 'param_out'], np.abs(G[prefix + 'param_in']).max())
    close(m.cpu().numpy(), G[prefix + 'm_out'], max(np.abs(G[prefix + 'm_in']).max(), gmax))
    close(v.cpu().numpy(), G[prefix + 'v_out'], max(np.abs(G[prefix + 'v_in']).max(), gmax ** 2))
    close(vm.cpu().numpy(), G[prefix + ('vmax_out' if ams else 'vmax_in')], max(np.abs(G[prefix + 'vmax_in']).max(), gmax ** 2))
    # rows that were not listed are untouched
    mask = np.ones(p.shape[0], bool)
    mask[G[prefix + 'index']] = False
    np.testing.assert_array_equal(p.cpu().numpy()[mask], G[prefix + 'param_in'][mask])


@pytest.mark.gpu
def test_large_sparse_adam_against_oracle(built):
    from log_b200.optim import sparse_adam_step_
    g = torch.Generator().manual_seed(5)
    N, K, C = 200_000, 120_000, 4
    p, m, v = torch.randn(N, C, generator=g), torch.randn(N, C, generator=g) * 0.01, torch.rand(N, C, generator=g) * 1e-3
    idx = torch.randperm(N, generator=g)[:K]
    grad = torch.randn(K, C, generator=g) * 0.1
    want = adam_oracle.sparse_adam_step(p.numpy(), grad.numpy(), m.numpy(), v.numpy(), idx.numpy(), step=123, lr=2e-3)
    dev = torch.device('cuda:0')
    pd, md, vd = p.to(dev), m.to(dev), v.to(dev)
    sparse_adam_step_(pd, grad.to(dev), md, vd, idx.to(dev), 123, 2e-3)
    close(pd.cpu().numpy(), want[0], 5.0)
    close(md.cpu().numpy(), want[1], 0.5)
    close(vd.cpu().numpy(), want[2], 0.3)


def test_cpu_tensors_raise(built):
    from log_b200._capi import LgrError
    from log_b200.optim import sparse_adam_step_
    z = torch.zeros(4, 3)
    with pytest.raises(LgrError):
        sparse_adam_step_(z, z[:2], z.clone(), z.clone(), torch.tensor([0, 1]), 1, 1e-3)

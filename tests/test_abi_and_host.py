"""CPU-side checks: the C-ABI library loads and exports what include/ssnode.h declares,
its host scalar helpers agree with the oracle, and the Python mirror behaves like the
reference modules.  No GPU compute here."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'ssnode.h')).read()
    return re.findall(r'^SSN_API\s+[\w\s\*]+?\b(\w+)\s*\(', text, flags=re.M)


def test_library_exports_every_declared_symbol(built_library):
    import ctypes
    lib = ctypes.CDLL(built_library)
    names = declared_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), name
    from tc_gan_b200 import clib
    assert set(names) == set(clib.EXPORTED_SYMBOLS)


def test_reference_abi_symbols_and_signatures(built_library):
    """tc_gan/clib.py:16-33: three solver symbols named solve_dynamics_{io}_{solver} and
    the four scalar helpers."""
    from tc_gan_b200 import clib
    for io_type in ('asym_power', 'asym_linear', 'asym_tanh'):
        fun = getattr(clib.libssnode, 'solve_dynamics_{}_{}'.format(io_type, 'euler'))
        assert len(fun.argtypes) == 14 and fun.restype is clib.ctypes.c_int
    assert clib.libssnode.ssn_kernel_launches() >= 0


def test_scalar_helpers_match_numpy(built_library, oracle):
    """tc_gan/tests/test_ssn.py:8-63, atol 1e-12."""
    from tc_gan_b200 import ssnode
    from tc_gan_b200.clib import libssnode
    k, n, r0, r1 = 0.01, 2.2, 200., 1000.
    v0 = ssnode.rate_to_volt(r0, k=k, n=n)
    xs = np.linspace(-0.1, v0 * 3, 1000)
    for io_type, cfun in (('asym_tanh', libssnode.io_atanh), ('asym_linear', libssnode.io_alin),
                          ('asym_power', libssnode.io_pow)):
        io_fun = ssnode.make_io_fun(k=k, n=n, rate_soft_bound=r0, rate_hard_bound=r1, io_type=io_type)
        ys_c = np.array([cfun(x, r0, r1, v0, k, n) for x in xs])
        np.testing.assert_allclose(io_fun(xs), ys_c, rtol=0, atol=1e-12)
        np.testing.assert_allclose(oracle.io_fun(xs, io_type, k, n, r0, r1), ys_c, rtol=0, atol=1e-12)
    rates = np.linspace(0, 1000, 1000)
    np.testing.assert_allclose(ssnode.rate_to_volt(rates, k=k, n=n),
                               [libssnode.rate_to_volt(x, k, n) for x in rates], rtol=0, atol=1e-12)
    a, b = np.arange(5.), np.arange(5.)[::-1].copy()
    from tc_gan_b200.clib import double_ptr
    assert libssnode.dot(5, a.ctypes.data_as(double_ptr), b.ctypes.data_as(double_ptr)) == a @ b


def test_io_sub_bound(built_library):
    """tc_gan/tests/test_ssn.py:35-55."""
    from tc_gan_b200 import ssnode
    kw = dict(k=0.01, n=2.2, rate_soft_bound=200, rate_hard_bound=1000)
    v0 = ssnode.rate_to_volt(200, k=0.01, n=2.2)
    io_pow = ssnode.make_io_fun(io_type='asym_power', **kw)
    io_atanh = ssnode.make_io_fun(io_type='asym_tanh', **kw)
    io_alin = ssnode.make_io_fun(io_type='asym_linear', **kw)
    assert 0 == io_pow(np.float64(0)) == io_atanh(np.float64(0)) == io_alin(np.float64(0))
    assert io_pow(np.float64(v0 + 10)) != io_atanh(np.float64(v0 + 10))
    assert io_pow(np.float64(v0 + 10)) != io_alin(np.float64(v0 + 10))
    xs = np.linspace(-10, v0, 500)
    np.testing.assert_allclose(io_pow(xs), io_atanh(xs), rtol=0, atol=1e-12)
    np.testing.assert_allclose(io_pow(xs), io_alin(xs), rtol=0, atol=1e-12)
    with pytest.raises(ValueError):
        ssnode.make_io_fun(0.01, 2.2, io_type='nope')


def test_weight_and_stimuli_mirrors(built_library):
    from tc_gan_b200 import stimuli, weight_gen, ssnode
    g = golden('weights_stimuli.npz')
    np.testing.assert_allclose(weight_gen.generate_weight(7, g['J_new'], g['D_new'], g['S_new'], g['z7']),
                               g['W7'], rtol=0, atol=1e-15)
    P = ssnode.DEFAULT_PARAMS
    x = np.linspace(-.5, .5, 51)
    np.testing.assert_allclose(stimuli.input(P['bandwidths'], x, P['smoothness'], P['contrast']), g['stim8'],
                               rtol=0, atol=1e-14)
    np.testing.assert_allclose(stimuli.input(np.linspace(0, 1, 10), x, P['smoothness'], [5, 10, 20, 30, 40]),
                               g['stim50'], rtol=0, atol=1e-13)
    np.testing.assert_allclose(stimuli.input([0.25, 0.5], x, P['smoothness'], [20, 10], [-0.25, 0.0, 0.25]),
                               g['stim_off'], rtol=0, atol=1e-13)
    j = ssnode.new_JDS()
    np.testing.assert_allclose(j['J'], g['J_new'])
    np.testing.assert_allclose(j['D'], g['D_new'])
    W, z = weight_gen.generate_parameter(5, j['J'], j['D'], j['S'], seed=1)
    assert W.shape == z.shape == (10, 10) and (W[:, :5] >= 0).all() and (W[:, 5:] <= 0).all()


def test_subsample_neurons_doctest_cases(built_library):
    """tc_gan/gradient_expressions/utils.py:74-126."""
    from tc_gan_b200.gradient_expressions.utils import subsample_neurons, sample_sites_from_stim_space
    N, NZ, NB = 7, 5, 2
    rate_vector = np.tile(np.arange(2 * N), (NZ, NB, 1))
    red0 = subsample_neurons(rate_vector, [2, 3, 4], False)
    assert red0.shape == (NZ * 3, NB) and red0[:3].tolist() == [[2, 2], [3, 3], [4, 4]]
    red1 = subsample_neurons(rate_vector, [2, 3, 4], True)
    assert red1.shape == (NZ, NB * 3) and red1[0].tolist() == [2, 3, 4, 2, 3, 4]
    red2 = subsample_neurons(rate_vector, [2], include_inhibitory_neurons=True)
    assert red2[0].tolist() == [2, 2] and red2[1].tolist() == [9, 9]
    assert sample_sites_from_stim_space([0, 0.5, 1], 101) == [50, 75, 100]
    with pytest.raises(ValueError):
        sample_sites_from_stim_space([0, 0.0001], 101)
    import torch
    t = subsample_neurons(torch.as_tensor(rate_vector), [2, 3, 4], False)
    assert t.shape == (NZ * 3, NB) and t.numpy().tolist() == red0.tolist()


def test_argument_validation_without_gpu(built_library):
    from tc_gan_b200 import ssnode, clib
    with pytest.raises(ValueError):
        ssnode.fixed_point(np.eye(2), np.zeros(2), k=1, n=1, io_type='bogus')
    with pytest.raises(ValueError):
        ssnode.fixed_point(np.eye(2), np.zeros(2), k=1, n=1, solver='rk4')
    with pytest.raises(ValueError):
        ssnode.find_fixed_points(1, iter([]), np.zeros((1, 2)), method='threads')
    with pytest.raises(ValueError):
        clib.make_solver(io_type='bogus')
    sv = clib.make_solver(io_type='asym_power', rate_stop_at=200.)
    assert sv.rate_hard_bound == 200. and sv.io_type == 0          # ssnode.py:241-242
    sv = clib.make_solver(io_type='asym_tanh', rate_stop_at=200.)
    assert sv.rate_hard_bound == 1000. and sv.io_type == 2


def test_no_cpu_fallback(built_library):
    """Without a GPU the solver must fail loudly, never compute on the CPU."""
    from tc_gan_b200 import ssnode, clib
    if clib.libssnode.ssn_device_count() > 0:
        pytest.skip('a GPU is visible')
    with pytest.raises(clib.SSNLibraryError):
        ssnode.fixed_point(np.eye(2) * 0.1, np.ones(2), k=1, n=1, io_type='asym_linear')
    with pytest.raises(clib.SSNLibraryError):
        ssnode.fixed_points_batch(np.zeros((1, 2, 2)), np.ones((1, 2)), k=1, n=1)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'tc_gan_b200')
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(base, f)).read()
                assert 'ssn_oracle' not in text and 'oracle/' not in text, os.path.join(base, f)


def test_hetero_input_and_critic_loss_on_cpu(built_library):
    """Host-side torch helpers of the widened rows (no GPU needed): heterogeneous stimulus
    (tc_gan/networks/ssn.py:645-727) and the WGAN-GP critic loss (networks/wgan.py:194-215)."""
    import torch
    from tc_gan_b200 import gan, torch_ops
    ext = torch.arange(12., dtype=torch.float64).reshape(2, 6)          # nb=2, 2N=6
    zs = torch.tensor([[1., -1., 1., -1., 1., -1.], [0.5, 0.5, 0.5, -0.5, -0.5, -0.5]], dtype=torch.float64)
    V = torch.tensor([0.2, 0.4], dtype=torch.float64, requires_grad=True)
    out = torch_ops.hetero_input(ext, zs, V)
    assert out.shape == (2, 2, 6)
    want = (1 + torch.tensor([0.2] * 3 + [0.4] * 3, dtype=torch.float64) * zs[:, None, :]) * ext[None]
    assert torch.allclose(out, want)
    out.sum().backward()
    assert torch.allclose(V.grad, torch.stack([(zs[:, None, :3] * ext[None, :, :3]).sum(),
                                               (zs[:, None, 3:] * ext[None, :, 3:]).sum()]))
    deg = torch_ops.hetero_input(ext, zs, torch.tensor(0.3, dtype=torch.float64))
    assert torch.allclose(deg, (1 + 0.3 * zs[:, None, :]) * ext[None])
    torch.manual_seed(0)
    critic = gan.Critic(6, layers=(8,), layer_norm=True).double()
    fake, real = torch.randn(5, 6, dtype=torch.float64), torch.randn(7, 6, dtype=torch.float64)
    loss, acc = gan.critic_loss(critic, fake, real, lipschitz_cost=10.0)
    assert loss.requires_grad and not acc.requires_grad and torch.isfinite(loss)
    assert abs(float(acc) - float(critic(fake).mean() - critic(real).mean())) < 1e-12
    loss0, _ = gan.critic_loss(critic, fake, real, lipschitz_cost=0.0)
    assert abs(float(loss0) - float(acc)) < 1e-12                     # penalty-free loss is the accuracy
    with pytest.raises(ValueError):
        gan.SSNWassersteinGAN(np.zeros((4, 8), np.float32), mode='nope', device='cpu')

"""the WGAN-GP learners that call the SSN hot path (torch MLP critic around the CUDA
generator).  One generator step of each flavour, as tc_gan/networks/tests/test_wgan.py:50-64 does, plus a
finite-difference check of the generator gradient through prober + critic + implicit gradient."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def make_gan(mode, **kw):
    import torch
    from tc_gan_b200 import gan
    if not torch.cuda.is_available():
        pytest.fail('GPU tests need a CUDA device')
    rs = np.random.RandomState(0)
    data = np.abs(rs.randn(64, 8 * 3)).astype(np.float32) * 5
    args = dict(num_sites=21, mode=mode, sample_sites=(-0.5, 0.0, 0.5), num_models=6,
                critic_layers=(32, 32), critic_iters_init=2, critic_iters=1,
                seqlen=60, skip_steps=40, seed=1)
    args.update(kw)
    return gan.SSNWassersteinGAN(data, **args)


@pytest.mark.parametrize('mode', ['fixed_point', 'bptt'])
def test_single_generator_step(mode):
    g = make_gan(mode)
    J0 = g.J.detach().clone()
    infos = []
    for info in g.learning():
        infos.append(info)
        if not info['is_discriminator']:
            break
    assert [i['is_discriminator'] for i in infos] == [True, True, False]
    assert all(np.isfinite(i['disc_loss']) for i in infos[:2]) and np.isfinite(infos[2]['gen_loss'])
    assert not np.allclose(g.J.detach().cpu().numpy(), J0.cpu().numpy())       # Adam moved the parameters
    assert (g.J >= g.param_min).all() and (g.S <= g.param_max).all()
    assert g.sample_sites == [5, 10, 15]


@pytest.mark.parametrize('mode,ssn_type', [('bptt', 'deg-heteroin'), ('fixed_point', 'heteroin')])
def test_heteroin_generator_step(mode, ssn_type):
    """The paper's runs use ssn_type 'deg-heteroin' (scripts/fig4/gan/run.json): V is learned too."""
    g = make_gan(mode, ssn_type=ssn_type, V=0.5, dist_in='bernoulli' if mode == 'bptt' else 'uniform')
    V0 = g.V.detach().clone()
    for info in g.learning():
        if not info['is_discriminator']:
            break
    assert np.isfinite(info['gen_loss'])
    assert g.V.shape == (V0.shape) and not np.allclose(g.V.detach().cpu().numpy(), V0.cpu().numpy())
    assert (g.V >= 0).all() and (g.V <= 1).all()


def test_generator_gradient_matches_finite_differences():
    import torch
    g = make_gan('fixed_point', num_models=4, solver_kwargs=dict(atol=1e-9, max_iter=200000))
    z = g.sample_z()

    def loss_at(J):
        with torch.no_grad():
            R, status, _ = __import__('tc_gan_b200.torch_ops', fromlist=['x']).fixed_points(
                z, J, g.D, g.S, g.exts, solver=g.solver)
            assert (status == 0).all()
            return float(-g.critic(g.tuning_curves(R).float()).double().mean())

    tc, _, _ = g.generate(z, differentiable=True)
    loss = -g.critic(tc.float()).mean()
    loss.backward()
    for (a, b) in ((0, 0), (1, 1), (0, 1)):
        h = 2e-3 * float(g.J[a, b])
        Jp, Jm = g.J.detach().clone(), g.J.detach().clone()
        Jp[a, b] += h
        Jm[a, b] -= h
        fd = (loss_at(Jp) - loss_at(Jm)) / (2 * h)
        assert abs(fd - float(g.J.grad[a, b])) <= 2e-2 * max(abs(fd), 1e-3), (a, b, fd, float(g.J.grad[a, b]))

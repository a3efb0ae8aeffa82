"""
GPU tests of the learners around the SSN path with the reference's names: probe kernels,
the BPTT WGAN (tc_gan/networks/wgan.py) and the conditional WGAN behind tc_gan.run.bptt_cwgan
(tc_gan/networks/cwgan.py), the truth-dataset provider and the driver + recorders.
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def torch_cuda(built_library):
    import torch
    from tc_gan_b200 import clib
    if clib.libssnode.ssn_device_count() < 1 or not torch.cuda.is_available():
        pytest.fail('GPU tests need a CUDA device: the library has no CPU fallback')
    return torch


def test_probe_gather_and_scatter_kernels(torch_cuda):
    """ssn_probe_gather == time_avg[model_ids, :, probes] (cwgan.py:96-99); its backward == index_put accumulate;
    FixedProber == time_avg[:, :, probes].reshape (networks/ssn.py:838-851)."""
    torch = torch_cuda
    from tc_gan_b200 import torch_ops
    from tc_gan_b200.networks.ssn import FixedProber
    rs = np.random.RandomState(0)
    nz, nb, dim, batch = 5, 8, 42, 17
    rates = torch.tensor(rs.rand(nz, nb, dim), dtype=torch.float32, device='cuda:0', requires_grad=True)
    ids = torch.tensor(rs.randint(0, nz, batch), device='cuda:0')
    probes = torch.tensor(rs.randint(0, dim, batch), device='cuda:0')
    probes[3], ids[3] = probes[2], ids[2]                       # a duplicate: the scatter must accumulate
    out = torch_ops.probe_rates(rates, ids, probes)
    want = rates.detach()[ids, :, probes]
    assert out.shape == (batch, nb)
    torch.testing.assert_close(out, want, rtol=0, atol=0)
    go = torch.tensor(rs.randn(batch, nb), dtype=torch.float32, device='cuda:0')
    out.backward(go)
    ref = torch.zeros_like(rates)
    ref.index_put_((ids[:, None].expand(batch, nb), torch.arange(nb, device='cuda:0')[None].expand(batch, nb),
                    probes[:, None].expand(batch, nb)), go, accumulate=True)
    torch.testing.assert_close(rates.grad, ref, rtol=1e-6, atol=1e-6)
    fp = FixedProber(None, [3, 20, 41])
    tc = fp.tuning_curve(rates.detach())
    np.testing.assert_array_equal(tc.cpu().numpy(), fp.probe_numpy(rates.detach().cpu().numpy()))


def small_config(**over):
    from tc_gan_b200 import ssnode
    jds = ssnode.new_JDS()
    cfg = dict(J0=jds['J'], D0=jds['D'], S0=jds['S'], num_sites=21, seqlen=60, skip_steps=40,
               bandwidths=[0, 0.25, 0.5, 1], contrasts=[5, 20], critic_iters_init=2, critic_iters=1,
               disc=dict(layers=[16, 16]), gen=dict(rate_cost=0.01, dynamics_cost=1.0), seed=3)
    cfg.update(over)
    return cfg


def test_bptt_wgan_make_gan_and_learning(torch_cuda):
    """tc_gan/networks/tests/test_wgan.py:50-64: make_gan + a few steps of learning(); rmsprop + L2 decay on the
    critic, the rate-penalty bound skipping critic updates (networks/wgan.py:395-400)."""
    from tc_gan_b200.networks import wgan
    cfg = small_config(batchsize=6, sample_sites=[-0.5, 0, 0.5], include_inhibitory_neurons=True)
    cfg['disc'].update(update_name='rmsprop', reg_l2_decay=1e-3, learning_rate=1e-3)
    gan, rest = wgan.make_gan(cfg)
    assert not rest, rest
    assert gan.gen.output_shape == (6, 8 * 6) and list(gan.gen.prober.probes) == [5, 10, 15, 26, 31, 36]
    assert gan.sample_sites == [5, 10, 15]
    data = np.abs(np.random.RandomState(0).randn(40, 48)) * 5
    gan.set_dataset(data)
    J0 = gan.get_gen_param()[0].copy()
    infos = []
    for info in gan.learning():
        infos.append(info)
        if len(infos) == 5:
            break
    assert [i.is_discriminator for i in infos] == [True, True, False, True, False]
    assert all(np.isfinite(i.disc_loss) for i in infos if i.is_discriminator)
    assert np.isfinite(infos[2].gen_loss) and infos[2].gen_time > 0
    assert not np.allclose(gan.get_gen_param()[0], J0)
    out = gan.gen_forward()
    assert out.prober_tuning_curve.shape == (6, 48) and out.model_dynamics_penalty >= 0
    # a bound below the current rate penalty skips the critic update and reports NaN
    gan.disc_rate_penalty_bound, gan.rate_penalty_threshold = 1e-9, 0.0
    gan.gen_forward_watch = gan.disc_train_watch = wgan.StopWatch()
    skipped = gan.train_discriminator(wgan.Namespace(is_discriminator=True, gen_step=9, disc_step=0))
    assert np.isnan(skipped.disc_loss) and np.isnan(skipped.accuracy)


def test_conditional_generator_gradient_matches_float64_autograd(torch_cuda, oracle):
    """The cWGAN generator loss (cwgan.py:96-120, networks/wgan.py:236-241) differentiated through the CUDA path
    (ConditionalProber gather -> BPTT kernels) against torch float64 autograd through the oracle's Euler unroll,
    the same indexing and the same critic weights: dL/d(J, D, S) at rtol 1e-4."""
    torch = torch_cuda
    from tc_gan_b200.networks import cwgan
    cfg = small_config(num_models=3, probes_per_model=2, norm_probes=[-0.5, 0, 0.5], num_sites=15,
                       include_inhibitory_neurons=True, contrasts=[5, 20])
    gan, rest = cwgan.make_gan(cfg)
    assert not rest, rest
    rs = np.random.RandomState(5)
    n_sites, nb = 15, 4
    kw = dict(stimulator_bandwidths=np.tile(np.array(cfg['bandwidths'], dtype='float32'), (3, 1)),
              stimulator_contrasts=np.array([[5.] * nb, [20.] * nb, [5.] * nb], dtype='float32'),
              prober_norm_probes=np.array([-0.5, 0.5, 0, -0.5, 0.5, 0], dtype='float32'),
              prober_cell_types=np.array([0, 1, 1, 0, 0, 1], dtype='uint16'),
              prober_model_ids=np.array([0, 0, 1, 1, 2, 2], dtype='uint16'),
              model_zs=rs.rand(3, 2 * n_sites, 2 * n_sites).astype('float32'),
              model_rate_penalty_threshold=0.5)
    trainer = gan.gen_trainer
    trainer.dynamics_cost, trainer.rate_cost = 3.0, 2.0
    loss = trainer.loss(**kw)
    loss.backward()
    m = gan.gen.model
    # ---- float64 restatement ----
    t64 = lambda a, g=False: torch.tensor(np.asarray(a, dtype=float), dtype=torch.float64, requires_grad=g)
    J, D, S = (t64(p.detach().cpu().numpy(), True) for p in (m.J, m.D, m.S))
    x = np.linspace(-.5, .5, n_sites)
    sig = lambda u: 1 / (1 + np.exp(-u / (0.25 / 8)))
    b, c = kw['stimulator_bandwidths'].astype(float)[..., None], kw['stimulator_contrasts'].astype(float)[..., None]
    stim = c * sig(x + b / 2) * sig(b / 2 - x)
    ext = np.concatenate([stim, stim], axis=-1)
    avg, dyn, rate = oracle.euler_unroll_torch(t64(kw['model_zs']), J, D, S, t64(ext), 60, 40, 0.01, 0.1,
                                               rate_penalty_threshold=0.5)
    probes = ((kw['prober_norm_probes'].astype(float) + 1) * (n_sites - 1) / 2).astype(int) + \
        kw['prober_cell_types'].astype(int) * n_sites
    tc = avg[torch.tensor(kw['prober_model_ids'].astype(int)), :, torch.tensor(probes)]
    cond = t64(np.array([kw['stimulator_contrasts'][kw['prober_model_ids'].astype(int), 0],
                         np.abs(kw['prober_norm_probes']), kw['prober_cell_types'].astype(float)]).T)
    critic64 = [(t64(l.weight.detach().cpu().numpy()), t64(l.bias.detach().cpu().numpy()))
                for l in gan.disc.l_out if hasattr(l, 'weight')]
    h = torch.cat([tc, cond], dim=1)
    for i, (w, bb) in enumerate(critic64):
        h = h @ w.T + bb
        if i < len(critic64) - 1:
            h = torch.relu(h)
    loss64 = -h.mean() + 3.0 * dyn + 2.0 * rate
    loss64.backward()
    # the critic runs in float32 on the device: compare the loss loosely, the generator gradients at 1e-4
    np.testing.assert_allclose(float(loss.detach()), float(loss64.detach()), rtol=1e-4)
    for got, want in ((m.J.grad, J.grad), (m.D.grad, D.grad), (m.S.grad, S.grad)):
        np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-4 * float(want.abs().max()))


def test_cwgan_learning_on_ssnode_dataset_with_driver(torch_cuda, tmp_path):
    """run.bptt_cwgan end to end at toy size: truth data from dataset_by_ssnode (networks/dataset.py:28-71),
    RandomChoiceSampler minibatches, ConditionalBPTTWassersteinGAN.learning() driven by BPTTcWGANDriver, tables
    written by the recorders (tc_gan/recorders.py), exit.json = end_of_iteration."""
    from tc_gan_b200 import drivers, execution
    from tc_gan_b200.networks import cwgan, dataset
    cfg = small_config(num_models=4, probes_per_model=2, norm_probes=[-0.5, 0, 0.5], ssn_type='deg-heteroin', V0=0.3,
                       include_inhibitory_neurons=True)
    cfg['gen'].update(V_min=0, V_max=1)
    gan, rest = cwgan.make_gan(cfg)
    assert not rest, rest
    data = dataset.dataset_by_ssnode(
        num_sites=21, bandwidths=cfg['bandwidths'], contrasts=cfg['contrasts'], truth_size=12, truth_seed=1,
        sample_sites=gan.sample_sites, include_inhibitory_neurons=True,
        true_ssn_options=dict(max_iter=5000))
    assert data.shape == (12, 2 * 4 * 2 * 3) and np.isfinite(data).all() and data.max() < 200
    gan.set_dataset(data)
    run = execution.pre_learn(datastore=str(tmp_path / 'run'), **{k: v for k, v in cfg.items() if k != 'seed'})
    with execution.DataStore(run['datastore'], table_format='csv') as ds:
        driver = drivers.BPTTcWGANDriver(gan, ds, iterations=3, quiet=True, tc_stats_record_interval=1)
        driver.run(gan)
    d = run['datastore']
    assert json.load(open(os.path.join(d, 'exit.json'))) == dict(reason='end_of_iteration', good=True)
    import pandas
    learning = pandas.read_csv(os.path.join(d, 'learning.csv'))
    assert list(learning['gen_step']) == [0, 1, 2] and np.isfinite(learning['Gloss']).all()
    gen = pandas.read_csv(os.path.join(d, 'generator.csv'))
    assert list(gen.columns)[-1] == 'V' and len(gen) == 3 and (gen['V'] >= 0).all() and (gen['V'] <= 1).all()
    disc = pandas.read_csv(os.path.join(d, 'disc_learning.csv'))
    assert len(disc) == 2 + 1 + 1                                   # critic_iters_init, then critic_iters per step
    assert os.path.exists(os.path.join(d, 'tc_stats.csv')) and os.path.exists(os.path.join(d, 'TC_mean.csv'))
    assert os.path.exists(os.path.join(d, 'disc_param_stats.csv'))


def test_fixed_point_gan_redraws_rejected_networks(torch_cuda):
    """tc_gan/ssnode.py:468-487 inside the fixed-point GAN: with the original (less stable) J, D and asym_power +
    rate_stop_at some draws are rejected; every generator batch still has exactly num_models tuning curves, the
    rejected draws are counted, and the implicit gradient attached to the kept fixed points is finite."""
    import torch
    from tc_gan_b200 import gan, ssnode
    P = ssnode.DEFAULT_PARAMS
    data = np.abs(np.random.RandomState(0).randn(64, 4)).astype(np.float32) * 5
    g = gan.SSNWassersteinGAN(data, num_sites=51, mode='fixed_point', J=P['J'], D=P['D'], S=P['S'], num_models=12,
                              bandwidths=[0, 0.0625, 0.125, 0.25],           # ~45 % of the draws are rejected
                              io_type='asym_power', solver_kwargs=dict(rate_stop_at=200.0, max_iter=3000),
                              critic_layers=(16,), critic_iters_init=1, critic_iters=1, seed=2)
    tc, _, _ = g.generate(g.sample_z(), differentiable=True)
    assert tc.shape == (12, 4) and g.rejections > 0 and g.draws == 12 + g.rejections + g.unused
    assert 0 < g.rejection_rate() < 1
    (-g.critic(tc.float()).mean()).backward()
    assert all(torch.isfinite(p.grad).all() for p in (g.J, g.D, g.S))
    infos = []
    for info in g.learning():
        infos.append(info)
        if not info['is_discriminator']:
            break
    assert np.isfinite(infos[-1]['gen_loss'])
    g.max_redraw_rounds = 0
    with pytest.raises(RuntimeError):
        g.generate(g.sample_z(), differentiable=False)

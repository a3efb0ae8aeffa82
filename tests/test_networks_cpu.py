"""
CPU tests of the host-side mirrors around the SSN path: stimulator, probes, grid helper,
conditional minibatch sampler, update rules, rejection limiter, datastore / recorders.  No CUDA kernel is called.
"""
import json
import os

import numpy as np
import pytest
import torch

from tc_gan_b200 import execution, recorders, stimuli
from tc_gan_b200.drivers import SSNRejectionLimiter, WGANDiscLossLimiter
from tc_gan_b200.gradient_expressions.utils import subsample_neurons
from tc_gan_b200.networks import cwgan, wgan
from tc_gan_b200.networks.ssn import BandwidthContrastStimulator, FixedProber, make_flat_param_names
from tc_gan_b200.networks.utils import gridify_tc_samples


def test_stimulator_matches_stimuli_input():
    """BandwidthContrastStimulator (networks/ssn.py:177-188) == stimuli.input (stimuli.py:3-10) on the
    contrast-major grid of grid_stimulator_inputs (networks/wgan.py:291-296)."""
    num_sites, bandwidths, contrasts = 21, [0, 0.125, 0.5, 1], [5, 20]
    stim = BandwidthContrastStimulator(num_sites, len(bandwidths) * len(contrasts), 0.25 / 8, device='cpu')
    c, b = wgan.grid_stimulator_inputs(contrasts, bandwidths, batchsize=3)
    assert c.shape == b.shape == (3, 8)
    got = stim.stimulus(b, c).numpy()
    want = stimuli.input(bandwidths, np.linspace(-.5, .5, num_sites), 0.25 / 8, contrasts)
    assert got.shape == (3, 8, 2 * num_sites)
    for z in range(3):
        np.testing.assert_allclose(got[z], want, rtol=1e-6, atol=1e-6)


def test_fixed_prober_layout_and_gridify_roundtrip():
    """FixedProber mixes the probe axis into the tuning-curve domain (networks/ssn.py:838-851) exactly as
    subsample_neurons(track_offset_identity=True); gridify_tc_samples inverts that layout."""
    nz, n_c, n_b, N = 4, 2, 3, 10
    rates = np.random.RandomState(0).rand(nz, n_c * n_b, 2 * N)
    sites = [2, 5, 7]
    probes = sites + [s + N for s in sites]
    tc = FixedProber(None, probes).probe_numpy(rates)
    np.testing.assert_array_equal(tc, subsample_neurons(rates, sites, track_offset_identity=True,
                                                        include_inhibitory_neurons=True))
    grid = gridify_tc_samples(tc, num_contrasts=n_c, num_bandwidths=n_b, num_cell_types=2, num_probes=3)
    assert grid.shape == (nz, 2, 3, n_c, n_b)
    for ct in range(2):
        for p in range(3):
            for c in range(n_c):
                for b in range(n_b):
                    np.testing.assert_array_equal(grid[:, ct, p, c, b], rates[:, c * n_b + b, sites[p] + ct * N])


def test_random_choice_sampler_minibatches():
    """cwgan.py:277-390: shapes, conditions, model ids, uniqueness of cells per model, e_ratio weighting."""
    rs = np.random.RandomState(1)
    bandwidths, contrasts, norm_probes = np.array([0, .25, .5, 1]), np.array([5., 20.]), np.array([-.5, 0, .5])
    n_data = 7
    grid_shape = (n_data, 2, len(norm_probes), len(contrasts), len(bandwidths))
    # data as subsample_neurons lays it out: (sample, contrast, bandwidth, cell_type, probe)
    data = rs.rand(n_data, len(contrasts), len(bandwidths), 2, len(norm_probes))
    sampler = cwgan.RandomChoiceSampler.from_grid_data(data.reshape(n_data, -1), bandwidths, contrasts, norm_probes,
                                                       include_inhibitory_neurons=True, e_ratio=0.8, seed=3)
    assert sampler.nested.shape == grid_shape
    batch = sampler.select_minibatch(num_models=5, probes_per_model=4)
    assert batch.batchsize == 20 and batch.tuning_curves.shape == (20, 4) and batch.conditions.shape == (20, 3)
    kw = batch.gen_kwargs
    assert kw['stimulator_bandwidths'].shape == kw['stimulator_contrasts'].shape == (5, 4)
    np.testing.assert_array_equal(kw['prober_model_ids'], np.repeat(np.arange(5), 4))
    assert set(np.unique(kw['prober_cell_types'])) <= {0, 1}
    # every row of the minibatch is a true tuning curve at the stated condition
    for i in range(batch.batchsize):
        contrast, norm_probe, cell_type = batch.conditions[i]
        ic = list(contrasts).index(contrast); ip = list(norm_probes).index(norm_probe)
        assert any(np.array_equal(batch.tuning_curves[i], data[s, ic, :, int(cell_type), ip]) for s in range(n_data))
        assert kw['stimulator_contrasts'][kw['prober_model_ids'][i], 0] == contrast
    # a cell is chosen at most once per model
    cells = np.stack([kw['prober_cell_types'].reshape(5, 4), kw['prober_norm_probes'].reshape(5, 4)], axis=-1)
    for m in range(5):
        assert len({tuple(c) for c in cells[m]}) == 4
    # e_ratio = 1: excitatory cells only
    s1 = cwgan.RandomChoiceSampler(sampler.nested, sampler.cond_values, e_ratio=1.0, seed=0)
    assert (s1.select_minibatch(6, 3).gen_kwargs['prober_cell_types'] == 0).all()
    b2 = next(cwgan.NaiveRandomChoiceSampler(sampler.nested, sampler.cond_values, e_ratio=0.5).random_minibatches(3, 2))
    assert b2.tuning_curves.shape == (6, 4)


@pytest.mark.parametrize('name,cfg', [('adam-wgan', {}), ('rmsprop', {'rho': 0.8}), ('sgd', {}), ('momentum', {})])
def test_updater_rules_and_regularisation(name, cfg):
    """networks/wgan.py:106-166: named rules; L2/L1 penalties enter the loss, decays act on the update."""
    p = torch.tensor([1.0, -2.0, 3.0], dtype=torch.float64, requires_grad=True)
    q = p.detach().clone().requires_grad_()
    lr = 0.01
    plain = wgan.Updater(lr, name, cfg)
    reg = wgan.Updater(lr, name, cfg, reg_l2_penalty=0.1, reg_l1_penalty=0.05, reg_l2_decay=0.5, reg_l1_decay=0.25)
    (p ** 2).sum().backward()
    plain.step([p])
    assert not torch.equal(p.detach(), torch.tensor([1.0, -2.0, 3.0], dtype=torch.float64))
    loss = (q ** 2).sum() + reg.penalty([q])
    q0 = q.detach().clone()
    np.testing.assert_allclose(float(reg.penalty([q]).detach()), 0.1 * 14 + 0.05 * 6)
    loss.backward()
    np.testing.assert_allclose(q.grad.numpy(), 2 * q0.numpy() + 0.2 * q0.numpy() + 0.05 * np.sign(q0.numpy()))
    ref = q0.clone().requires_grad_()
    ref.grad = q.grad.clone()
    wgan.Updater(lr, name, cfg).step([ref])
    reg.step([q])
    np.testing.assert_allclose(q.detach().numpy(),
                               ref.detach().numpy() - lr * 0.5 * q0.numpy() - lr * 0.25 * np.sign(q0.numpy()))
    if name == 'adam-wgan':
        assert plain._opt.defaults['betas'] == (0.5, 0.9)
    with pytest.raises(ValueError):
        wgan.Updater(lr, 'no-such-rule').step([p])


def test_rejection_and_disc_loss_limiters(tmp_path):
    """tc_gan/drivers.py:214-297."""
    ds = execution.DataStore(str(tmp_path), table_format='csv')
    lim = SSNRejectionLimiter(ds, n_samples=10, rejection_limit=0.6, max_consecutive_exceedings=2)
    for _ in range(2):
        lim(20)                           # 20 / 30 > 0.6
    lim(1)                                # resets
    for _ in range(2):
        lim(20)
    with pytest.raises(execution.KnownError) as err:
        lim(20)
    assert err.value.exit_code == 4
    assert json.load(open(os.path.join(str(tmp_path), 'exit.json'))) == dict(reason='too_many_rejections', good=False)
    dl = WGANDiscLossLimiter(ds, hist_length=5)
    for _ in range(4):
        dl(1e6)
    with pytest.raises(execution.KnownError):
        dl(1e6)


def _fake_updates(n):
    for k in range(n):
        info = wgan.Namespace(gen_loss=0.5 + k, gen_forward_time=0.1, gen_train_time=0.2, disc_time=0.3)
        disc = wgan.Namespace(disc_loss=-1.0 * k, accuracy=0.25, rate_penalty=0.0, dynamics_penalty=1e-9)
        yield k, recorders.UpdateResult(info=info, disc_info=disc)


@pytest.mark.parametrize('table_format', ['csv', 'hdf5'])
def test_datastore_tables_read_back(tmp_path, table_format):
    """learning / generator / disc_learning / tc_stats tables with the reference's column names and dtypes
    (tc_gan/recorders.py:113-361), info.json and exit.json (tc_gan/execution.py:222-346); read back the way
    tc_gan/loaders/datastore_loader.py:58-75 does (<table>.csv with a header line first, else the HDF5 tables)."""
    if table_format == 'hdf5' and not execution.have_h5py():
        pytest.skip('h5py is not installed in this image: the CSV tables are the format written here')
    pandas = pytest.importorskip('pandas')
    run_config = execution.pre_learn(datastore=str(tmp_path / 'run'), iterations=3, layers=[16], J0=np.eye(2))
    info = json.load(open(os.path.join(run_config['datastore'], 'info.json')))
    assert info['run_config']['iterations'] == 3 and info['run_config']['J0'] == [[1.0, 0.0], [0.0, 1.0]]
    assert set(info) == {'run_config', 'extra_info', 'meta_info'}

    class Gen(object):
        def get_flat_param_names(self):
            return make_flat_param_names([('J', torch.zeros(2, 2)), ('D', torch.zeros(2, 2)), ('S', torch.zeros(2, 2)),
                                          ('V', torch.zeros(()))])

        def get_flat_param_values(self):
            return list(np.arange(13.0))

    class Gan(object):
        gen = Gen()

        def get_gen_param(self):
            return [np.ones((2, 2))] * 3

    with execution.DataStore(run_config['datastore'], table_format=table_format) as ds:
        learning = recorders.LearningRecorder.make(ds)
        generator = recorders.FlexGenParamRecorder.make(ds, Gan())
        disc = recorders.DiscLearningRecorder.make(ds)
        tcs = recorders.ConditionalTuningCurveStatsRecorder.make(ds, 2)
        for k, res in _fake_updates(3):
            disc.record(k, 0, res.disc_info.disc_loss, 0.25, 0.01, 0.02, 0, 0)
            learning.record(k, res)
            generator.record(k)
        xd = np.array([[1., 2.], [3., 4.], [5., 6.]])
        cd = np.array([[20., 0., 0.], [20., 0., 0.], [5., .5, 1.]])
        tcs.record(7, wgan.Namespace(xd=xd, cd=cd, xg=xd + 1, cg=cd))
        ds.save_exit_reason(reason='end_of_iteration', good=True)
        ds.flush_all()
    d = run_config['datastore']
    if table_format == 'csv':
        load = lambda name: pandas.read_csv(os.path.join(d, name + '.csv'))
    else:
        import h5py

        def load(name):
            fname = name + '.hdf5' if os.path.exists(os.path.join(d, name + '.hdf5')) else 'store.hdf5'
            with h5py.File(os.path.join(d, fname), 'r') as f:
                return pandas.DataFrame(f[name][...])
    t = load('learning')
    assert list(t.columns) == ['gen_step', 'Gloss', 'Dloss', 'Daccuracy', 'gen_forward_time', 'gen_train_time',
                               'disc_time', 'rate_penalty', 'dynamics_penalty']
    np.testing.assert_allclose(t['Gloss'], [0.5, 1.5, 2.5])
    g = load('generator')
    assert list(g.columns) == ['gen_step', 'J_EE', 'J_EI', 'J_IE', 'J_II', 'D_EE', 'D_EI', 'D_IE', 'D_II',
                               'S_EE', 'S_EI', 'S_IE', 'S_II', 'V']
    assert len(g) == 3 and g['V'].iloc[0] == 12.0
    assert list(load('disc_learning').columns)[:4] == ['gen_step', 'disc_step', 'Dloss', 'Daccuracy']
    s = load('tc_stats')
    assert len(s) == 4 and list(s['count']) == [1, 2, 1, 2] and list(s['is_fake']) == [0, 0, 1, 1]
    np.testing.assert_allclose(s['mean_0'].iloc[1], 2.0)
    assert json.load(open(os.path.join(d, 'exit.json')))['good'] is True


def test_reference_loader_reads_our_csv_tables(tmp_path):
    """When the reference checkout is present (this container, not the GPU box), its own DataStoreLoader must read
    a run written here."""
    ref = '/root/reference/tc_gan/loaders/datastore_loader.py'
    if not os.path.exists(ref):
        pytest.skip('reference checkout not present')
    pytest.importorskip('pandas')
    import importlib.util
    spec = importlib.util.spec_from_file_location('ref_datastore_loader', ref)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    with execution.DataStore(str(tmp_path), table_format='csv') as ds:
        learning = recorders.LearningRecorder.make(ds)
        for k, res in _fake_updates(4):
            learning.record(k, res)
    cls = [getattr(mod, n) for n in dir(mod) if n.startswith('DataStoreLoader')]
    assert cls, 'reference loader class not found'
    loader = cls[0](str(tmp_path))
    table = loader.default_load('learning')
    assert list(table['gen_step']) == [0, 1, 2, 3] and 'Daccuracy' in table.columns

"""
Conditional Wasserstein GAN over the BPTT generator -- mirror of tc_gan/networks/cwgan.py, the learner behind
`tc_gan.run.bptt_cwgan` (BASELINE north_star):

* `ConditionalProber` (:34-104): tuning_curve = time_avg[model_ids, :, probes], one probed neuron per batch
  element -- here the CUDA gather `ssn_probe_gather`, whose backward scatter-adds dL/d tuning_curve into the
  dL/d time_avg array the BPTT kernel consumes;
* `ConditionalTuningCurveGenerator` (:107-120), `ConditionalDiscriminator` / `CellTypeBlindDiscriminator`
  (:123-187), `ConditionalCriticTrainer` (:190-214), `ConditionalMinibatch` (:217-274),
  `RandomChoiceSampler` / `NaiveRandomChoiceSampler` (:277-407), `ConditionalBPTTWassersteinGAN` (:410-552),
  `make_gan` (:555-614).
"""
from logging import getLogger

import numpy as np
import torch
from torch import nn

from ..gradient_expressions.utils import sample_sites_from_stim_space, sample_sites_from_stim_space_impl
from .. import torch_ops
from .ssn import TuningCurveGenerator, make_tuning_curve_generator
from .utils import gridify_tc_samples
from .wgan import (BPTTWassersteinGAN, CriticTrainer, GeneratorTrainer, Namespace, StopWatch, Updater,
                   _split_subdicts, _DISC_NET_KEYS, cartesian_product, emit_generator_trainer, gradient_penalty,
                   make_net)
from .wgan import DEFAULT_PARAMS as _WGAN_DEFAULTS

logger = getLogger(__name__)

DEFAULT_PARAMS = dict(_WGAN_DEFAULTS, num_models=1, probes_per_model=1, e_ratio=0.8, hide_cell_type=False,
                      norm_probes=[0])
del DEFAULT_PARAMS['batchsize']
del DEFAULT_PARAMS['sample_sites']


def as_randomstate(seed):
    return seed if hasattr(seed, 'rand') else np.random.RandomState(seed)


class ConditionalProber(object):
    """
    Probe `time_avg` (num_models, num_tcdom, num_neurons) with a probe that varies over the batch:
    ``tuning_curve[i] = time_avg[model_ids[i], :, probes[i]]`` with
    ``probes = ((norm_probes + 1) (num_sites - 1) / 2).astype(int) + cell_types * num_sites``.
    """

    inputs = ('norm_probes', 'model_ids', 'cell_types')
    outputs = ('tuning_curve',)

    def __init__(self, model):
        self.model = model

    def probe_indices(self, norm_probes, cell_types):
        norm_probes = np.asarray(norm_probes, dtype=float)
        assert np.all(norm_probes >= -1) and np.all(norm_probes <= 1)
        return (sample_sites_from_stim_space_impl(norm_probes, self.model.num_sites)
                + np.asarray(cell_types, dtype=int) * self.model.num_sites)

    def tuning_curve(self, time_avg, norm_probes, model_ids, cell_types):
        dev = time_avg.device
        probes = torch.as_tensor(self.probe_indices(norm_probes, cell_types), dtype=torch.int32, device=dev)
        ids = torch.as_tensor(np.asarray(model_ids, dtype=np.int32), device=dev)
        return torch_ops.probe_rates(time_avg, ids, probes)                 # (batchsize, num_tcdom)


class ConditionalTuningCurveGenerator(TuningCurveGenerator):

    output_shape = property(lambda self: (self.batchsize, self.num_tcdom))
    cond_shape = property(lambda self: (self.batchsize, 3))

    @staticmethod
    def conditions(stimulator_contrasts, prober_model_ids, prober_norm_probes, prober_cell_types, **_):
        """(batchsize, 3): contrast of the probed model, normalized probe, cell type (cwgan.py:113-120)."""
        contrasts = np.asarray(stimulator_contrasts)[np.asarray(prober_model_ids, dtype=int), 0]
        return np.array([contrasts, np.asarray(prober_norm_probes, dtype=float),
                         np.asarray(prober_cell_types, dtype=float)]).T


class ConditionalDiscriminator(nn.Module):
    """MLP critic on [tuning curve, contrast, |norm_probe|, cell_type] (cwgan.py:123-174)."""

    def __init__(self, shape, cond_shape, loss_type='WD', layers=(128, 128), normalization='none',
                 nonlinearity='rectify', net_options=None):
        super(ConditionalDiscriminator, self).__init__()
        assert loss_type == 'WD'
        self.l_out = make_net(shape[-1] + cond_shape[-1], layers, normalization, nonlinearity)

    def preprocess_condition(self, condition):
        # the SSN is symmetric: feed the absolute value of norm_probes
        return torch.stack([condition[:, 0], condition[:, 1].abs(), condition[:, 2]], dim=1)

    def get_output(self, inputs):
        tc, condition = inputs
        return self.l_out(torch.cat([tc, self.preprocess_condition(condition)], dim=1))

    forward = get_output

    def accuracy(self, xg, cg, xd, cd):
        with torch.no_grad():
            return float(self.get_output([xg, cg]).mean() - self.get_output([xd, cd]).mean())

    def get_all_params(self):
        return list(self.parameters())

    def prepare(self):
        pass


class CellTypeBlindDiscriminator(ConditionalDiscriminator):

    def preprocess_condition(self, condition):
        c = super(CellTypeBlindDiscriminator, self).preprocess_condition(condition)
        return torch.stack([c[:, 0], c[:, 1], torch.zeros_like(c[:, 2])], dim=1)


class ConditionalCriticTrainer(CriticTrainer):
    """loss = D(xg, cg) - D(xd, cd) + lmd * mean (||dD/dxp|| - 1)^2, gradient w.r.t. xp only (cwgan.py:190-214)."""

    def loss(self, xg, xd, xp, cg, cd, cp, lmd):
        d = self.disc.get_output
        return (d([xg, cg]).mean() - d([xd, cd]).mean()
                + lmd * gradient_penalty(lambda x: d([x, cp]), xp))

    def train(self, xg, xd, xp, cg, cd, cp, lmd):
        params = self.disc.get_all_params()
        for p in params:
            p.grad = None
        loss = self.loss(xg, xd, xp, cg, cd, cp, lmd)
        (loss + self.updater.penalty(params)).backward()
        from .wgan import _allreduce_mean_grads
        _allreduce_mean_grads(params)
        self.updater.step(params)
        return float(loss.detach())


class ConditionalGeneratorTrainer(GeneratorTrainer):

    def loss(self, rng=None, **kwargs):
        tc, dyn, rate = self.gen.get_output(rng=rng, **kwargs)
        cond = torch.as_tensor(self.gen.conditions(**kwargs), dtype=tc.dtype, device=tc.device)
        return -self.disc.get_output([tc, cond]).mean() + self.dynamics_cost * dyn + self.rate_cost * rate


class ConditionalMinibatch(object):

    def __init__(self, tc_md, conditions_md, bandwidths, contrasts):
        self.tc_md = tc_md
        self.conditions_md = np.asarray(conditions_md)
        self.bandwidths = bandwidths
        self.contrasts = contrasts
        assert self.tc_md.shape[:-1] == self.conditions_md.shape[1:]
        assert self.tc_md.shape[-1] == len(bandwidths)

    num_models = property(lambda self: self.tc_md.shape[0])
    probes_per_model = property(lambda self: self.tc_md.shape[1])
    num_bandwidths = property(lambda self: self.tc_md.shape[2])
    batchsize = property(lambda self: self.num_models * self.probes_per_model)

    @property
    def gen_kwargs(self):
        contrasts, bandwidths = np.broadcast_arrays(np.asarray(self.contrasts).reshape((-1, 1)),
                                                    np.asarray(self.bandwidths).reshape((1, -1)))
        _, norm_probes, cell_types = self._conditions_T
        return dict(stimulator_bandwidths=bandwidths.astype('float32'),
                    stimulator_contrasts=contrasts.astype('float32'),
                    prober_norm_probes=norm_probes.astype('float32'),
                    prober_cell_types=cell_types.astype('uint16'),
                    prober_model_ids=self.model_ids.astype('uint16'))

    @property
    def tuning_curves(self):
        return self.tc_md.reshape((self.batchsize, self.num_bandwidths))

    @property
    def conditions(self):
        return self._conditions_T.T

    @property
    def _conditions_T(self):
        return self.conditions_md.reshape((-1, self.batchsize))

    @property
    def model_ids(self):
        ids = np.arange(self.num_models, dtype='uint16').reshape((-1, 1))
        return np.broadcast_to(ids, self.conditions_md.shape[1:]).flatten()


class RandomChoiceSampler(object):
    """Minibatch sampler based on random choice (cwgan.py:277-390)."""

    @classmethod
    def from_grid_data(cls, data, bandwidths, contrasts, norm_probes, include_inhibitory_neurons, **kwargs):
        """`data`: what `subsample_neurons(..., track_offset_identity=True)` returns for the truth networks."""
        cell_types = [0, 1] if include_inhibitory_neurons else [0]
        nested = gridify_tc_samples(data, num_contrasts=len(contrasts), num_bandwidths=len(bandwidths),
                                    num_cell_types=len(cell_types), num_probes=len(norm_probes))
        cond_values = [cell_types, norm_probes, contrasts, bandwidths]
        assert nested.shape == (len(data),) + tuple(map(len, cond_values))
        return cls(nested, cond_values, **kwargs)

    def __init__(self, nested, cond_values, e_ratio, seed=0):
        self.nested = np.asarray(nested)
        self.cond_values = cond_values = list(map(np.asarray, cond_values))
        self.e_ratio = e_ratio
        self.cell_types, self.norm_probes, self.contrasts, self.bandwidths = cond_values
        self.rng = as_randomstate(seed)
        assert tuple(self.cell_types) in [(0,), (0, 1)]

    def random_cells(self, num_models, probes_per_model):
        """Choose cells in such a way that every cell is chosen at most once per model."""
        cellids = cartesian_product(np.arange(len(self.cell_types)), np.arange(len(self.norm_probes))).T.astype(int)
        if len(self.cell_types) == 2:
            probs = np.zeros(len(cellids))
            probs[:len(self.norm_probes)] = self.e_ratio
            probs[len(self.norm_probes):] = 1 - self.e_ratio
            probs /= probs.sum()
        else:
            probs = None
        ids = np.asarray([cellids[self.rng.choice(len(cellids), probes_per_model, replace=False, p=probs)]
                          for _ in range(num_models)])                   # (num_models, probes_per_model, 2)
        ids_cell_type, ids_norm_probes = ids.transpose((2, 0, 1))
        return ids_cell_type, ids_norm_probes

    def select_minibatch(self, num_models, probes_per_model):
        shape = (num_models, probes_per_model)
        ids_sample = self.rng.choice(len(self.nested), shape)
        ids_cell_type, ids_norm_probes = self.random_cells(*shape)
        ids_flat_contrast = self.rng.choice(len(self.contrasts), num_models)
        ids_contrast = np.broadcast_to(ids_flat_contrast.reshape((-1, 1)), shape)
        tc_md = self.nested[ids_sample, ids_cell_type, ids_norm_probes, ids_contrast]
        assert tc_md.shape == (num_models, probes_per_model, len(self.bandwidths))
        return ConditionalMinibatch(
            tc_md,
            [self.contrasts[ids_contrast], self.norm_probes[ids_norm_probes], self.cell_types[ids_cell_type]],
            self.bandwidths, self.contrasts[ids_flat_contrast])

    def random_minibatches(self, *args, **kwargs):
        while True:
            yield self.select_minibatch(*args, **kwargs)


class NaiveRandomChoiceSampler(RandomChoiceSampler):

    def random_cell_types(self, shape):
        if len(self.cell_types) == 2:
            return self.rng.choice(2, shape, p=[self.e_ratio, 1 - self.e_ratio])
        return np.zeros(shape, dtype='uint16')

    def random_cells(self, num_models, probes_per_model):
        shape = (num_models, probes_per_model)
        return self.random_cell_types(shape), self.rng.choice(len(self.norm_probes), shape)


class ConditionalBPTTWassersteinGAN(BPTTWassersteinGAN):

    def __init__(self, gen, disc, gen_trainer, disc_trainer, bandwidths, contrasts, norm_probes, e_ratio,
                 include_inhibitory_neurons, rate_penalty_threshold, num_models, probes_per_model,
                 critic_iters_init, critic_iters, lipschitz_cost, disc_rate_penalty_bound, seed=0):
        self.gen, self.disc, self.gen_trainer, self.disc_trainer = gen, disc, gen_trainer, disc_trainer
        self.bandwidths, self.contrasts = np.asarray(bandwidths), np.asarray(contrasts)
        self.norm_probes = np.asarray(norm_probes)
        self.e_ratio, self.include_inhibitory_neurons = e_ratio, include_inhibitory_neurons
        self.rate_penalty_threshold = rate_penalty_threshold
        self.num_models, self.probes_per_model = num_models, probes_per_model
        self.critic_iters_init, self.critic_iters = critic_iters_init, critic_iters
        self.lipschitz_cost, self.disc_rate_penalty_bound = lipschitz_cost, disc_rate_penalty_bound
        self.rng = as_randomstate(seed)
        assert gen.batchsize == self.num_models * self.probes_per_model
        assert self.probes_per_model < gen.num_neurons

    num_sites = property(lambda self: self.gen.model.num_sites)

    @property
    def sample_sites(self):
        return sample_sites_from_stim_space(self.norm_probes, self.num_sites)

    def set_dataset(self, data, **kwargs):
        kwargs.setdefault('seed', self.rng)
        self.sampler = RandomChoiceSampler.from_grid_data(
            data, bandwidths=self.bandwidths, contrasts=self.contrasts, norm_probes=self.norm_probes,
            e_ratio=self.e_ratio, include_inhibitory_neurons=self.include_inhibitory_neurons, **kwargs)
        self.dataset = self.sampler.random_minibatches(self.num_models, self.probes_per_model)

    def gen_forward(self, batch):
        return self.gen.forward(rng=self.rng, model_rate_penalty_threshold=self.rate_penalty_threshold,
                                **batch.gen_kwargs)

    def train_discriminator(self, info):
        batch = self.next_minibatch()
        xd, cd = batch.tuning_curves, batch.conditions
        eps = self.rng.rand(batch.batchsize, 1)
        with self.gen_forward_watch:
            gen_out = self.gen_forward(batch)
        xg = gen_out.prober_tuning_curve
        xp = eps * xd + (1 - eps) * xg
        info.gen_out = gen_out
        info.dynamics_penalty, info.rate_penalty = gen_out.model_dynamics_penalty, gen_out.model_rate_penalty
        info.xd, info.xg, info.xp = xd, xg, xp
        info.cd = info.cg = info.cp = cd
        info.batch = batch
        info.gen_time = self.gen_forward_watch.times[-1]
        bound = self.disc_rate_penalty_bound
        if bound > 0 and gen_out.model_rate_penalty > bound:
            info.disc_loss = info.accuracy = info.disc_time = np.nan
            return info
        t = self._t
        with self.disc_train_watch:
            info.disc_loss = self.disc_trainer.train(t(xg), t(xd), t(xp), t(cd), t(cd), t(cd), self.lipschitz_cost)
        info.accuracy = self.disc.accuracy(t(xg), t(cd), t(xd), t(cd))
        info.disc_time = self.disc_train_watch.times[-1]
        return info

    def train_generator(self, info, batch):
        with self.gen_train_watch:
            info.gen_loss = self.gen_trainer.train(rng=self.rng, model_rate_penalty_threshold=self.rate_penalty_threshold,
                                                   **batch.gen_kwargs)
        info.gen_forward_time = self.gen_forward_watch.sum()
        info.gen_train_time = self.gen_train_watch.sum()
        info.gen_time = info.gen_train_time + info.gen_forward_time
        info.disc_time = self.disc_train_watch.sum()
        return info

    def _single_gen_step(self, gen_step, critic_iters):
        self.gen_forward_watch, self.gen_train_watch, self.disc_train_watch = StopWatch(), StopWatch(), StopWatch()
        for disc_step in range(critic_iters):
            info = Namespace(is_discriminator=True, gen_step=gen_step, disc_step=disc_step)
            info = self.train_discriminator(info)
            yield info
        disc_info, batch = info, info.batch
        info = Namespace(is_discriminator=False, gen_step=gen_step)
        info = self.train_generator(info, batch)
        logger.debug('[Loss] Acc: %-9.3g D: %-9.3g G: %-9.3g [Time] Fwd: %.3g D: %.3g G: %.3g',
                     disc_info.accuracy, disc_info.disc_loss, info.gen_loss, self.gen_forward_watch.mean(),
                     self.disc_train_watch.mean(), self.gen_train_watch.mean())
        yield info


def make_gan(config, device=None):
    """``(ConditionalBPTTWassersteinGAN, unused config)`` (cwgan.py:555-614).  Required: J0, D0, S0."""
    kw = _split_subdicts(config, DEFAULT_PARAMS)
    gen_cfg, disc_cfg = kw.pop('gen'), kw.pop('disc')
    bandwidths, contrasts = kw.pop('bandwidths'), kw.pop('contrasts')
    num_models, probes_per_model = kw.pop('num_models'), kw.pop('probes_per_model')
    hide_cell_type = kw.pop('hide_cell_type')
    if 'V0' in kw:
        kw['V'] = kw.pop('V0')
    rate_penalty_threshold = gen_cfg.pop('rate_penalty_threshold')
    disc_rate_penalty_bound = disc_cfg.pop('rate_penalty_bound')
    loop = {k: kw.pop(k) for k in ('critic_iters_init', 'critic_iters', 'lipschitz_cost', 'norm_probes', 'e_ratio',
                                   'include_inhibitory_neurons')}
    seed = kw.pop('seed', 0)
    gen, rest = make_tuning_curve_generator(
        kw, batchsize=num_models * probes_per_model, num_tcdom=len(bandwidths),
        J=kw.pop('J0'), D=kw.pop('D0'), S=kw.pop('S0'),
        emit_prober=ConditionalProber, emit_tcg=ConditionalTuningCurveGenerator, device=device)
    for k in ('J0', 'D0', 'S0'):
        rest.pop(k, None)
    disc_cls = CellTypeBlindDiscriminator if hide_cell_type else ConditionalDiscriminator
    disc = disc_cls(gen.output_shape, gen.cond_shape, 'WD',
                    **{k: disc_cfg.pop(k) for k in _DISC_NET_KEYS if k in disc_cfg}).to(gen.stimulator.device)
    gen_trainer, gen_rest = emit_generator_trainer(gen, disc, **gen_cfg)
    gen_trainer.__class__ = ConditionalGeneratorTrainer
    disc_updater, disc_rest = Updater.consume_kwargs(**disc_cfg)
    disc_trainer = ConditionalCriticTrainer(disc, disc_updater)
    rest.update({'gen': gen_rest, 'disc': disc_rest} if (gen_rest or disc_rest) else {})
    gan = ConditionalBPTTWassersteinGAN(
        gen, disc, gen_trainer, disc_trainer, bandwidths, contrasts, num_models=num_models,
        probes_per_model=probes_per_model, rate_penalty_threshold=rate_penalty_threshold,
        disc_rate_penalty_bound=disc_rate_penalty_bound, seed=seed, **loop)
    return gan, rest

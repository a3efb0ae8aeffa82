"""
Wasserstein GAN (gradient penalty) fitting the BPTT tuning-curve generator -- mirror of
tc_gan/networks/wgan.py: `UnConditionalDiscriminator` (:66-103), `Updater` (:106-166: 'adam-wgan' =
Adam(0.5, 0.9), any other lasagne.updates name such as 'rmsprop', L1/L2 penalties and decoupled decays),
`CriticTrainer` (:194-215), `GeneratorTrainer` (:218-260, parameter clipping), `BPTTWassersteinGAN`
(:299-444, incl. the `disc_rate_penalty_bound` critic skip :395-400) and `make_gan` (:454-509).
The critic is a torch.nn MLP (tc_gan/networks/simple_discriminator.py); all SSN arithmetic is in libssnode.so.
"""
import itertools
import time

import numpy as np
import torch
from torch import nn

from .. import dist as sdist
from .. import ssnode
from ..gradient_expressions.utils import sample_sites_from_stim_space
from .ssn import is_heteroin, make_tuning_curve_generator


DEFAULT_PARAMS = dict(
    bandwidths=ssnode.DEFAULT_PARAMS['bandwidths'],
    contrasts=ssnode.DEFAULT_PARAMS['contrast'],
    smoothness=ssnode.DEFAULT_PARAMS['smoothness'],
    sample_sites=[0],
    num_sites=ssnode.DEFAULT_PARAMS['N'],
    k=ssnode.DEFAULT_PARAMS['k'],
    n=ssnode.DEFAULT_PARAMS['n'],
    io_type='asym_tanh',
    tau_E=10,
    tau_I=1,
    dt=0.1,
    seqlen=1200,
    batchsize=1,
    skip_steps=1000,
    include_inhibitory_neurons=False,
    gen=dict(
        rate_cost=0.01,
        rate_penalty_threshold=200.0,
        dynamics_cost=1.0,
        J_min=1e-3, J_max=10, D_min=1e-3, D_max=10, S_min=1e-3, S_max=10,
    ),
    disc=dict(
        rate_penalty_bound=-1.0,
        layers=[128, 128],
        normalization='none',
        nonlinearity='rectify',
    ),
    critic_iters_init=50,
    critic_iters=5,
    lipschitz_cost=10.0,
)


class Namespace(object):
    def __init__(self, **kwargs):
        self.__dict__.update(kwargs)


class StopWatch(object):
    """Wall-clock stopwatch of tc_gan/utils/misc.py:27-52 (synchronises the device on exit)."""

    def __init__(self):
        self.times = []

    def __enter__(self):
        self._t0 = time.time()
        return self

    def __exit__(self, *exc):
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        self.times.append(time.time() - self._t0)

    def sum(self):
        return float(np.sum(self.times))

    def mean(self):
        return float(np.mean(self.times)) if self.times else float('nan')


_NONLINEARITIES = {'rectify': nn.ReLU, 'relu': nn.ReLU, 'tanh': nn.Tanh, 'sigmoid': nn.Sigmoid,
                   'leaky_rectify': lambda: nn.LeakyReLU(0.01), 'LeakyRectify': lambda: nn.LeakyReLU(0.01)}


def make_net(n_in, layers, normalization='none', nonlinearity='rectify'):
    """Hidden MLP stack + linear output (loss_type 'WD'), simple_discriminator.py:6-75, 104-165."""
    mods, width = [], n_in
    for units in layers:
        mods.append(nn.Linear(width, units))
        if normalization == 'layer':
            mods.append(nn.LayerNorm(units))
        elif normalization not in ('none', None):
            raise ValueError('Unknown normalization: {}'.format(normalization))
        mods.append(_NONLINEARITIES[nonlinearity]())
        width = units
    mods.append(nn.Linear(width, 1))
    return nn.Sequential(*mods)


class UnConditionalDiscriminator(nn.Module):

    def __init__(self, shape, loss_type='WD', layers=(128, 128), normalization='none', nonlinearity='rectify',
                 net_options=None):
        super(UnConditionalDiscriminator, self).__init__()
        assert loss_type == 'WD'
        self.l_out = make_net(shape[-1], layers, normalization, nonlinearity)

    def get_output(self, inputs):
        return self.l_out(inputs)

    forward = get_output

    def accuracy(self, xg, xd):
        with torch.no_grad():
            return float(self.get_output(xg).mean() - self.get_output(xd).mean())

    def get_all_params(self):
        return list(self.parameters())

    def prepare(self):
        pass


class Updater(object):
    """
    Parameter update rule by name (networks/wgan.py:106-166).  `update_name`: 'adam-wgan' (Adam with
    beta1 = 0.5, beta2 = 0.9, Gulrajani et al. 2017), 'adam', 'rmsprop', 'sgd', 'momentum', 'nesterov_momentum',
    'adagrad', 'adadelta' -- the lasagne.updates names with lasagne's defaults; `update_config` overrides
    (lasagne spellings beta1/beta2/rho/epsilon/momentum accepted).  `reg_l2_penalty` / `reg_l1_penalty` add
    sum(p^2) / sum|p| to the loss; `reg_l2_decay` / `reg_l1_decay` subtract lr*decay*p / lr*decay*sign(p) from the
    updated parameter (decoupled weight decay, Loshchilov & Hutter 2017).
    """

    _named_update_configs = {'adam-wgan': ('adam', dict(beta1=0.5, beta2=0.9))}

    def __init__(self, learning_rate=0.001, update_name='adam-wgan', update_config={},
                 reg_l2_penalty=0.0, reg_l2_decay=0.0, reg_l1_penalty=0.0, reg_l1_decay=0.0):
        self.learning_rate, self.update_name, self.update_config = learning_rate, update_name, dict(update_config)
        self.reg_l2_penalty, self.reg_l1_penalty = reg_l2_penalty, reg_l1_penalty
        self.reg_l2_decay, self.reg_l1_decay = reg_l2_decay, reg_l1_decay
        self._opt = None

    def _make(self, params):
        name, default = self._named_update_configs.get(self.update_name, (self.update_name, {}))
        cfg = dict(default, **self.update_config)
        lr = self.learning_rate
        if name == 'adam':
            return torch.optim.Adam(params, lr=lr, betas=(cfg.get('beta1', 0.9), cfg.get('beta2', 0.999)),
                                    eps=cfg.get('epsilon', 1e-8))
        if name == 'rmsprop':
            return torch.optim.RMSprop(params, lr=lr, alpha=cfg.get('rho', 0.9), eps=cfg.get('epsilon', 1e-6))
        if name == 'sgd':
            return torch.optim.SGD(params, lr=lr)
        if name in ('momentum', 'nesterov_momentum'):
            return torch.optim.SGD(params, lr=lr, momentum=cfg.get('momentum', 0.9), nesterov=name.startswith('nest'))
        if name == 'adagrad':
            return torch.optim.Adagrad(params, lr=lr, eps=cfg.get('epsilon', 1e-6))
        if name == 'adadelta':
            return torch.optim.Adadelta(params, lr=lr, rho=cfg.get('rho', 0.95), eps=cfg.get('epsilon', 1e-6))
        raise ValueError('Unknown update rule: {}'.format(self.update_name))

    def penalty(self, params):
        """Regularisation term to add to the loss (0-dim tensor or 0.0)."""
        reg = 0.0
        if self.reg_l2_penalty:
            reg = reg + self.reg_l2_penalty * sum((p ** 2).sum() for p in params)
        if self.reg_l1_penalty:
            reg = reg + self.reg_l1_penalty * sum(p.abs().sum() for p in params)
        return reg

    def step(self, params):
        """Apply the update from the gradients already in `p.grad`, then the decoupled decays."""
        params = list(params)
        if self._opt is None:
            self._opt = self._make(params)
        before = [p.detach().clone() for p in params] if (self.reg_l2_decay or self.reg_l1_decay) else None
        self._opt.step()
        if before is not None:
            with torch.no_grad():
                for p, p0 in zip(params, before):
                    if self.reg_l2_decay:
                        p.sub_(self.learning_rate * self.reg_l2_decay * p0)
                    if self.reg_l1_decay:
                        p.sub_(self.learning_rate * self.reg_l1_decay * torch.sign(p0))

    _KEYS = ('learning_rate', 'update_name', 'update_config', 'reg_l2_penalty', 'reg_l2_decay',
             'reg_l1_penalty', 'reg_l1_decay')

    @classmethod
    def consume_kwargs(cls, **kwargs):
        own = {k: kwargs.pop(k) for k in cls._KEYS if k in kwargs}
        return cls(**own), kwargs


def gradient_penalty(disc_fn, xp):
    """mean (||d D(x_hat) / d x_hat||_2 - 1)^2 (networks/wgan.py:209-212), differentiable in the critic."""
    xp = xp.detach().requires_grad_(True)
    grad, = torch.autograd.grad(disc_fn(xp).sum(), xp, create_graph=True)
    return ((grad.norm(2, dim=1) - 1) ** 2).mean()


def _allreduce_mean_grads(params):
    rank, world = sdist.world()
    if world > 1:
        for p in params:
            if p.grad is not None:
                torch.distributed.all_reduce(p.grad)
                p.grad /= world


class CriticTrainer(object):
    """Discriminator/critic trainer for WGAN-GP (networks/wgan.py:194-215)."""

    def __init__(self, disc, updater):
        self.target = self.disc = disc
        self.updater = updater

    def loss(self, xg, xd, xp, lmd):
        d = self.disc.get_output
        return d(xg).mean() - d(xd).mean() + lmd * gradient_penalty(d, xp)

    def train(self, xg, xd, xp, lmd):
        params = self.disc.get_all_params()
        for p in params:
            p.grad = None
        loss = self.loss(xg, xd, xp, lmd)
        (loss + self.updater.penalty(params)).backward()
        _allreduce_mean_grads(params)
        self.updater.step(params)
        return float(loss.detach())

    def prepare(self):
        pass


class GeneratorTrainer(object):
    """loss = -mean D(G) + dynamics_cost * dynamics_penalty + rate_cost * rate_penalty, then clipping of every
    generator parameter to [<name>_min, <name>_max] (networks/wgan.py:218-260)."""

    def __init__(self, gen, disc, dynamics_cost, rate_cost, J_min, J_max, D_min, D_max, S_min, S_max, updater,
                 V_min=0, V_max=1):
        self.target = self.gen = gen
        self.disc = disc
        self.dynamics_cost, self.rate_cost = dynamics_cost, rate_cost
        self.J_min, self.J_max, self.D_min, self.D_max = J_min, J_max, D_min, D_max
        self.S_min, self.S_max, self.V_min, self.V_max = S_min, S_max, V_min, V_max
        self.updater = updater

    def disc_output(self, gen_output):
        return self.disc.get_output(gen_output)

    def loss(self, rng=None, **kwargs):
        tc, dyn, rate = self.gen.get_output(rng=rng, **kwargs)
        return -self.disc_output(tc).mean() + self.dynamics_cost * dyn + self.rate_cost * rate

    def clip_params(self):
        with torch.no_grad():
            for name, p in self.gen.get_all_params():
                p.clamp_(getattr(self, name + '_min'), getattr(self, name + '_max'))

    def train(self, rng=None, **kwargs):
        named = self.gen.get_all_params()
        params = [p for _, p in named]
        for p in params:
            p.grad = None
        for q in self.disc.get_all_params():
            q.requires_grad_(False)
        try:
            loss = self.loss(rng=rng, **kwargs)
            (loss + self.updater.penalty(params)).backward()
        finally:
            for q in self.disc.get_all_params():
                q.requires_grad_(True)
        # networks are sharded over the ranks: ONE all-reduce of the packed generator gradient
        rank, world = sdist.world()
        if world > 1:
            packed = torch.cat([p.grad.reshape(-1) for p in params])
            torch.distributed.all_reduce(packed)
            packed /= world
            o = 0
            for p in params:
                p.grad = packed[o:o + p.numel()].reshape(p.shape).clone()
                o += p.numel()
        self.updater.step(params)
        self.clip_params()
        return float(loss.detach())

    def prepare(self):
        pass


HeteroInGeneratorTrainer = GeneratorTrainer          # V_min / V_max are always accepted here


def emit_generator_trainer(gen, disc, **kwargs):
    updater, kwargs = Updater.consume_kwargs(**kwargs)
    keys = ('dynamics_cost', 'rate_cost', 'J_min', 'J_max', 'D_min', 'D_max', 'S_min', 'S_max', 'V_min', 'V_max')
    own = {k: kwargs.pop(k) for k in keys if k in kwargs}
    if not is_heteroin(gen):
        own.pop('V_min', None), own.pop('V_max', None)
    return GeneratorTrainer(gen, disc, updater=updater, **own), kwargs


def cartesian_product(*arrays):
    """Rows of the product as columns: shape (len(arrays), prod(lens)) -- tc_gan/utils/numerics.py:37-64."""
    grids = np.meshgrid(*[np.asarray(a) for a in arrays], indexing='ij')
    return np.array([g.reshape(-1) for g in grids])


def grid_stimulator_inputs(contrasts, bandwidths, batchsize):
    """(stimulator_contrasts, stimulator_bandwidths), each (batchsize, n_contrasts * n_bandwidths), contrast-major
    (networks/wgan.py:291-296)."""
    product = cartesian_product(contrasts, bandwidths)
    return np.tile(product.reshape((1,) + product.shape), (batchsize,) + (1,) * product.ndim).swapaxes(0, 1)


def random_minibatches(batchsize, data, strict=False, seed=0):
    """Endless minibatches from shuffled epochs (tc_gan/utils/numerics.py:67-82)."""
    rng = seed if hasattr(seed, 'permutation') else np.random.RandomState(seed)
    num_batches = len(data) // batchsize
    if strict and len(data) % batchsize:
        raise ValueError('len(data) is not a multiple of batchsize')
    assert num_batches >= 1, 'dataset smaller than one minibatch'
    while True:
        idx = rng.permutation(len(data))
        for i in range(num_batches):
            yield data[idx[i * batchsize:(i + 1) * batchsize]]


class BPTTWassersteinGAN(object):
    """The learning loop of networks/wgan.py:299-444 over torch tensors on the generator's device."""

    loss_type = 'WD'

    def __init__(self, gen, disc, gen_trainer, disc_trainer, bandwidths, contrasts, include_inhibitory_neurons,
                 rate_penalty_threshold, critic_iters_init, critic_iters, lipschitz_cost, disc_rate_penalty_bound,
                 seed=0):
        self.gen, self.disc, self.gen_trainer, self.disc_trainer = gen, disc, gen_trainer, disc_trainer
        self.critic_iters_init, self.critic_iters = critic_iters_init, critic_iters
        self.lipschitz_cost, self.disc_rate_penalty_bound = lipschitz_cost, disc_rate_penalty_bound
        self.rng = seed if hasattr(seed, 'rand') else np.random.RandomState(seed)
        self.bandwidths, self.contrasts = bandwidths, contrasts
        self.stimulator_contrasts, self.stimulator_bandwidths = grid_stimulator_inputs(
            contrasts, bandwidths, self.batchsize)
        self.include_inhibitory_neurons = include_inhibitory_neurons
        self.rate_penalty_threshold = rate_penalty_threshold

    batchsize = property(lambda self: self.gen.batchsize)
    num_neurons = property(lambda self: self.gen.num_neurons)
    discriminator = property(lambda self: self.disc.l_out)
    NZ = property(lambda self: self.batchsize)
    device = property(lambda self: self.gen.stimulator.device)

    @property
    def sample_sites(self):
        probes = list(self.gen.prober.probes)
        return probes[:len(probes) // 2] if self.include_inhibitory_neurons else probes

    def get_gen_param(self):
        m = self.gen.model
        return [p.detach().cpu().numpy() for p in (m.J, m.D, m.S)]

    def set_dataset(self, data, **kwargs):
        kwargs.setdefault('seed', self.rng)
        self.dataset = random_minibatches(self.batchsize, np.asarray(data), **kwargs)

    def next_minibatch(self):
        return next(self.dataset)

    def prepare(self):
        pass

    def _t(self, a):
        return torch.as_tensor(np.asarray(a, dtype=np.float32), device=self.device)

    def _gen_kwargs(self):
        return dict(stimulator_bandwidths=self.stimulator_bandwidths, stimulator_contrasts=self.stimulator_contrasts,
                    model_rate_penalty_threshold=self.rate_penalty_threshold)

    def gen_forward(self):
        return self.gen.forward(rng=self.rng, **self._gen_kwargs())

    def train_discriminator(self, info):
        xd = self.next_minibatch()
        eps = self.rng.rand(self.batchsize).reshape((-1, 1))
        with self.gen_forward_watch:
            gen_out = self.gen_forward()
        xg = gen_out.prober_tuning_curve
        xp = eps * xd + (1 - eps) * xg
        info.gen_out = gen_out
        info.dynamics_penalty, info.rate_penalty = gen_out.model_dynamics_penalty, gen_out.model_rate_penalty
        info.xd, info.xg, info.xp = xd, xg, xp
        info.gen_time = self.gen_forward_watch.times[-1]
        bound = self.disc_rate_penalty_bound
        if bound > 0 and gen_out.model_rate_penalty > bound:           # networks/wgan.py:395-400
            info.disc_loss = info.accuracy = info.disc_time = np.nan
            return info
        with self.disc_train_watch:
            info.disc_loss = self.disc_trainer.train(self._t(xg), self._t(xd), self._t(xp), self.lipschitz_cost)
        info.accuracy = self.disc.accuracy(self._t(xg), self._t(xd))
        info.disc_time = self.disc_train_watch.times[-1]
        return info

    def train_generator(self, info):
        with self.gen_train_watch:
            info.gen_loss = self.gen_trainer.train(rng=self.rng, **self._gen_kwargs())
        info.gen_forward_time = self.gen_forward_watch.sum()
        info.gen_train_time = self.gen_train_watch.sum()
        info.gen_time = info.gen_train_time + info.gen_forward_time
        info.disc_time = self.disc_train_watch.sum()
        return info

    def _single_gen_step(self, gen_step, critic_iters):
        self.gen_forward_watch, self.gen_train_watch, self.disc_train_watch = StopWatch(), StopWatch(), StopWatch()
        for disc_step in range(critic_iters):
            info = Namespace(is_discriminator=True, gen_step=gen_step, disc_step=disc_step)
            yield self.train_discriminator(info)
        info = Namespace(is_discriminator=False, gen_step=gen_step)
        yield self.train_generator(info)

    def learning(self):
        for info in self._single_gen_step(0, self.critic_iters_init):
            yield info
        for gen_step in itertools.count(1):
            for info in self._single_gen_step(gen_step, self.critic_iters):
                yield info


def probes_from_stim_space(stim_locs, num_sites, include_inhibitory_neurons):
    probes = sample_sites_from_stim_space(stim_locs, num_sites)
    if include_inhibitory_neurons:
        probes.extend([p + num_sites for p in list(probes)])
    return probes


def _split_subdicts(config, defaults):
    kwargs = dict(defaults, **config)
    kwargs['gen'] = dict(defaults['gen'], **config.get('gen', {}))
    kwargs['disc'] = dict(defaults['disc'], **config.get('disc', {}))
    return kwargs


_DISC_NET_KEYS = ('layers', 'normalization', 'nonlinearity', 'net_options')


def make_gan(config, device=None):
    """``(BPTTWassersteinGAN, unused config)`` from a flat config with `gen` / `disc` sub-dicts
    (networks/wgan.py:454-509).  Required: J0, D0, S0."""
    kw = _split_subdicts(config, DEFAULT_PARAMS)
    gen_cfg, disc_cfg = kw.pop('gen'), kw.pop('disc')
    bandwidths, contrasts = kw.pop('bandwidths'), kw.pop('contrasts')
    num_sites = kw['num_sites']
    include_inh = kw.pop('include_inhibitory_neurons')
    probes = probes_from_stim_space(kw.pop('sample_sites'), num_sites, include_inh)
    if 'V0' in kw:
        kw['V'] = kw.pop('V0')
    rate_penalty_threshold = gen_cfg.pop('rate_penalty_threshold')
    disc_rate_penalty_bound = disc_cfg.pop('rate_penalty_bound')
    loop = {k: kw.pop(k) for k in ('critic_iters_init', 'critic_iters', 'lipschitz_cost')}
    seed = kw.pop('seed', 0)
    gen, rest = make_tuning_curve_generator(
        kw, num_tcdom=len(bandwidths) * len(contrasts), J=kw.pop('J0'), D=kw.pop('D0'), S=kw.pop('S0'),
        probes=probes, device=device)
    for k in ('J0', 'D0', 'S0'):
        rest.pop(k, None)
    disc = UnConditionalDiscriminator(gen.output_shape, 'WD', **{k: disc_cfg.pop(k) for k in _DISC_NET_KEYS
                                                               if k in disc_cfg}).to(gen.stimulator.device)
    gen_trainer, gen_rest = emit_generator_trainer(gen, disc, **gen_cfg)
    disc_updater, disc_rest = Updater.consume_kwargs(**disc_cfg)
    disc_trainer = CriticTrainer(disc, disc_updater)
    rest.update({'gen': gen_rest, 'disc': disc_rest} if (gen_rest or disc_rest) else {})
    gan = BPTTWassersteinGAN(gen, disc, gen_trainer, disc_trainer, bandwidths, contrasts,
                             include_inhibitory_neurons=include_inh, rate_penalty_threshold=rate_penalty_threshold,
                             disc_rate_penalty_bound=disc_rate_penalty_bound, seed=seed, **loop)
    return gan, rest

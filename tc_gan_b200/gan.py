"""
WGAN-GP learners over the CUDA SSN generator -- SURVEY.md section 8f rank 1 (the first
caller of the hot path), kept deliberately small: the critic is a few-hundred-unit MLP
(torch.nn), all of the simulation work stays in libssnode.so.

What it mirrors in the reference (tc_gan/networks/wgan.py unless noted):
  * critic loss  mean D(fake) - mean D(real) + lambda * mean (||grad_x D(x_hat)||_2 - 1)^2
    with x_hat = eps * real + (1 - eps) * fake                      :194-215, :380-395
  * generator loss  -mean D(G(z)) + dynamics_cost * dynamics_penalty + rate_cost * rate_penalty
    and clipping of J, D, S to [min, max] after each update           :218-254
  * updates 'adam-wgan' = Adam(beta1=0.5, beta2=0.9)                  :112-118
  * schedule: critic_iters_init critic steps before the first generator step, then
    critic_iters per generator step                                   :430-444
  * defaults (seqlen 1200, skip_steps 1000, dt 0.1, tau (10, 1), rate_cost 0.01,
    rate_penalty_threshold 200)                                       :39-63
  * the fixed-point ("legacy") GAN of tc_gan/run/gan.py:617-672, 830-941: generator = fixed
    points of sampled networks, implicit gradient; a network that does not converge for every
    stimulus is rejected and a fresh one drawn until `num_models` have converged, exactly as
    `find_fixed_points` does (tc_gan/ssnode.py:468-487).  `max_redraw_rounds` bounds the loop
    (a generator whose draws keep diverging raises instead of spinning), the solver's `max_iter`
    bounds every diverging draw, and `SSNRejectionLimiter` semantics (tc_gan/drivers.py:214-255)
    are available through `rejection_rate()`.
Tuning curves fed to the critic are rates probed at `sample_sites`
(gradient_expressions/utils.py:74-149, track_offset_identity=True layout).

Multi-GPU: networks of a generator batch are sharded over the ranks; the only collective
on the generator path is one all-reduce of the packed (dJ, dD, dS); critic gradients are
all-reduced as usual data-parallel training does.
"""
import numpy as np
import torch
from torch import nn

from . import dist as sdist
from . import ssnode, stimuli, torch_ops
from .gradient_expressions.utils import sample_sites_from_stim_space, subsample_neurons


class Critic(nn.Module):
    """MLP critic, tc_gan/networks/simple_discriminator.py:6-75 (rectifier units, optional
    layer normalisation, linear output for the WGAN loss)."""

    def __init__(self, n_in, layers=(128, 128), layer_norm=False):
        super().__init__()
        mods, width = [], n_in
        for units in layers:
            mods.append(nn.Linear(width, units))
            if layer_norm:
                mods.append(nn.LayerNorm(units))
            mods.append(nn.ReLU())
            width = units
        mods.append(nn.Linear(width, 1))
        self.net = nn.Sequential(*mods)

    def forward(self, x):
        return self.net(x)


def critic_loss(critic, fake, real, lipschitz_cost=10.0, generator=None):
    """WGAN-GP critic loss and the accuracy D(fake) - D(real) (wgan.py:194-215, :90)."""
    eps = torch.rand((real.shape[0], 1), device=real.device, dtype=real.dtype, generator=generator)
    n = min(len(fake), len(real))
    x_hat = (eps[:n] * real[:n] + (1 - eps[:n]) * fake[:n]).requires_grad_(True)
    grad, = torch.autograd.grad(critic(x_hat).sum(), x_hat, create_graph=True)
    penalty = ((grad.norm(2, dim=1) - 1) ** 2).mean()
    d_fake, d_real = critic(fake).mean(), critic(real).mean()
    return d_fake - d_real + lipschitz_cost * penalty, (d_fake - d_real).detach()


class SSNWassersteinGAN(object):
    """
    One object for both generator flavours:
      mode='fixed_point' -- tuning curves from converged fixed points, implicit gradient;
      mode='bptt'        -- tuning curves from the time-averaged unrolled Euler dynamics, BPTT.
    `data` is [n_data, nb * len(sample_sites)] true tuning curves.
    """

    def __init__(self, data, num_sites=ssnode.DEFAULT_PARAMS['N'], mode='fixed_point',
                 J=None, D=None, S=None,
                 bandwidths=ssnode.DEFAULT_PARAMS['bandwidths'], contrasts=(20.,),
                 smoothness=ssnode.DEFAULT_PARAMS['smoothness'], sample_sites=(0,),
                 num_models=128, io_type='asym_tanh', k=0.01, n=2.2,
                 seqlen=1200, skip_steps=1000, dt=0.1, tau_E=10., tau_I=1.,
                 dynamics_cost=0.0, rate_cost=0.01, rate_penalty_threshold=200.0,
                 critic_layers=(128, 128), layer_norm=False,
                 critic_iters_init=50, critic_iters=5, lipschitz_cost=10.0,
                 gen_learning_rate=0.001, disc_learning_rate=0.001,
                 param_min=1e-3, param_max=10.0, seed=0, device='cuda', solver_kwargs=None,
                 ssn_type='default', V=0.5, dist_in='bernoulli', V_min=0.0, V_max=1.0,
                 max_redraw_rounds=20, disc_rate_penalty_bound=-1.0,
                 gen_update='adam-wgan', disc_update='adam-wgan', gen_update_config=None, disc_update_config=None):
        if mode not in ('fixed_point', 'bptt'):
            raise ValueError('Unknown mode: {}'.format(mode))
        if ssn_type not in ('default', 'heteroin', 'deg-heteroin'):
            raise ValueError('Unknown ssn_type: {}'.format(ssn_type))
        if dist_in not in ('bernoulli', 'uniform'):
            raise ValueError('Unknown dist_in: {}'.format(dist_in))
        self.ssn_type, self.dist_in, self.V_min, self.V_max = ssn_type, dist_in, V_min, V_max
        self.mode, self.num_sites, self.num_models = mode, int(num_sites), int(num_models)
        self.device = torch.device(device)
        jds = ssnode.new_JDS()
        mk = lambda a, key: torch.tensor(np.asarray(jds[key] if a is None else a), dtype=torch.float64,
                                         device=self.device, requires_grad=True)
        self.J, self.D, self.S = mk(J, 'J'), mk(D, 'D'), mk(S, 'S')
        x = np.linspace(-0.5, 0.5, self.num_sites)
        self.exts = torch.tensor(stimuli.input(bandwidths, x, smoothness, list(contrasts)),
                                 dtype=torch.float32, device=self.device)
        # probe locations are given in stimulus space [-1, 1] (wgan.py:30, probes_from_stim_space)
        self.sample_sites = sample_sites_from_stim_space(list(sample_sites), self.num_sites)
        self.data = torch.as_tensor(data, dtype=torch.float32, device=self.device)
        n_in = self.exts.shape[0] * len(self.sample_sites)
        assert self.data.shape[1] == n_in, (self.data.shape, n_in)
        torch.manual_seed(seed)
        self.critic = Critic(n_in, critic_layers, layer_norm).to(self.device)
        # heterogeneous input (networks/ssn.py:645-727): V is a generator parameter, 2-vector or scalar
        self.V = None
        if ssn_type != 'default':
            v0 = np.broadcast_to(np.asarray(V, dtype=float), (2,)).copy() if ssn_type == 'heteroin' else float(np.asarray(V))
            self.V = torch.tensor(v0, dtype=torch.float64, device=self.device, requires_grad=True)
        self.gen_params = [self.J, self.D, self.S] + ([self.V] if self.V is not None else [])
        from .networks.wgan import Updater          # update rules by the reference's names (adam-wgan, rmsprop, ...)
        self.gen_updater = Updater(gen_learning_rate, gen_update, gen_update_config or {})
        self.disc_updater = Updater(disc_learning_rate, disc_update, disc_update_config or {})
        self.max_redraw_rounds, self.disc_rate_penalty_bound = int(max_redraw_rounds), disc_rate_penalty_bound
        self.draws = self.unused = 0
        self.rng = torch.Generator(device=self.device)
        rank, world = sdist.world()
        self.rng.manual_seed(seed * 1000 + rank)
        self.local_models = len(sdist.shard_indices(self.num_models, rank, world))
        self.io = dict(io_type=io_type, k=k, n=n)
        self.euler = dict(seqlen=seqlen, skip_steps=skip_steps, dt=dt, tau_E=tau_E, tau_I=tau_I)
        self.costs = dict(dynamics_cost=dynamics_cost, rate_cost=rate_cost,
                          rate_penalty_threshold=rate_penalty_threshold)
        self.critic_iters_init, self.critic_iters = critic_iters_init, critic_iters
        self.lipschitz_cost = lipschitz_cost
        self.param_min, self.param_max = param_min, param_max
        self.solver = torch_ops.make_solver(**dict(self.io, **(solver_kwargs or {})))
        self.rejections = 0

    # ---- generator --------------------------------------------------------------------
    def sample_z(self, n=None):
        dim = 2 * self.num_sites
        n = self.local_models if n is None else n
        return torch.rand((n, dim, dim), generator=self.rng, device=self.device)

    def rejection_rate(self):
        """#rejections / #draws so far (what tc_gan/drivers.py:232 thresholds at 0.6)."""
        return self.rejections / max(self.draws, 1)

    def tuning_curves(self, rates):
        return subsample_neurons(rates, self.sample_sites, track_offset_identity=True)

    def input_noise(self, nz):
        """z_in [nz, 2N] of the heteroin types (+-1 bernoulli or uniform in [-1, 1]); None otherwise."""
        if self.V is None:
            return None
        shape = (nz, 2 * self.num_sites)
        if self.dist_in == 'bernoulli':
            return torch.randint(0, 2, shape, generator=self.rng, device=self.device).float() * 2 - 1
        return torch.rand(shape, generator=self.rng, device=self.device) * 2 - 1

    def stimulus(self, nz, zs_in=None):
        """[nb, 2N], or [nz, nb, 2N] scaled per neuron by 1 + V z_in for the heteroin types."""
        if self.V is None:
            return self.exts
        if zs_in is None:
            zs_in = self.input_noise(nz)
        return torch_ops.hetero_input(self.exts, zs_in, self.V)

    def converged_networks(self, z):
        """
        Rejection sampling as `find_fixed_points` (tc_gan/ssnode.py:468-487): solve the networks `z`; every one
        that fails for some stimulus is replaced by a fresh draw until `len(z)` have converged.  Runs without
        autograd; returns (z_kept, zs_in_kept or None, R_kept), in the order the networks were accepted.
        """
        want = z.shape[0]
        zs, noises, Rs = [], [], []
        missing, rounds = want, 0
        while missing > 0:
            rounds += 1
            if rounds > self.max_redraw_rounds:
                raise RuntimeError('fixed-point generator: %d of %d networks still unconverged after %d rounds of '
                                   're-drawing (rejection rate %.2f)' % (missing, want, self.max_redraw_rounds,
                                                                         self.rejection_rate()))
            zs_in = self.input_noise(z.shape[0])
            with torch.no_grad():
                R, status, _ = torch_ops.fixed_points(z, self.J, self.D, self.S, self.stimulus(z.shape[0], zs_in),
                                                      solver=self.solver)
            ok = (status == 0).all(dim=1)
            n_ok = int(ok.sum())                       # the control flow needs the count (as the reference's does)
            self.draws += z.shape[0]
            self.rejections += z.shape[0] - n_ok
            if n_ok > missing:                         # an over-drawn round: keep the first `missing` successes
                self.unused += n_ok - missing          # (FixedPointsInfo.unused of the reference's pool)
                ok &= torch.cumsum(ok.to(torch.int32), 0) <= missing
                n_ok = missing
            zs.append(z[ok]); Rs.append(R[ok])
            if zs_in is not None:
                noises.append(zs_in[ok])
            missing -= n_ok
            if missing > 0:
                # draw for the acceptance rate seen so far (the reference's pool over-submits likewise,
                # ssnode.py:468-487 resubmit_threshold), at most 4x the batch per round
                accept = max((self.draws - self.rejections) / max(self.draws, 1), 0.05)
                z = self.sample_z(int(min(4 * want, max(missing, np.ceil(missing / accept)))))
        return torch.cat(zs), (torch.cat(noises) if noises else None), torch.cat(Rs)

    def generate(self, z, differentiable):
        """(tuning curves [num local models, nb * n_sites], dynamics_penalty, rate_penalty)."""
        if self.mode == 'fixed_point':
            z, zs_in, R = self.converged_networks(z)
            if differentiable:
                # the implicit gradient needs only the fixed points: attach them to the graph
                R = torch_ops.attach_fixed_point(z, self.J, self.D, self.S, self.stimulus(z.shape[0], zs_in), R,
                                                 solver=self.solver)
            zero = R.new_zeros(())
            return self.tuning_curves(R), zero, zero
        ctx = torch.enable_grad() if differentiable else torch.no_grad()
        with ctx:
            exts = self.stimulus(z.shape[0])
            avg, dyn, rate = torch_ops.euler_ssn(
                z, self.J, self.D, self.S, exts, rate_penalty_threshold=self.costs['rate_penalty_threshold'],
                **dict(self.euler, **self.io))
            return self.tuning_curves(avg), dyn, rate

    # ---- training steps -----------------------------------------------------------------
    def next_minibatch(self, n):
        idx = torch.randint(0, self.data.shape[0], (n,), generator=self.rng, device=self.device)
        return self.data[idx]

    def train_discriminator(self):
        fake, dyn, rate = self.generate(self.sample_z(), differentiable=False)
        bound = self.disc_rate_penalty_bound
        if bound > 0 and float(rate) > bound:               # tc_gan/networks/wgan.py:395-400: skip the critic update
            return dict(is_discriminator=True, disc_loss=float('nan'), accuracy=float('nan'),
                        rate_penalty=float(rate), dynamics_penalty=float(dyn))
        real = self.next_minibatch(len(fake))
        params = list(self.critic.parameters())
        for p in params:
            p.grad = None
        loss, accuracy = critic_loss(self.critic, fake.float(), real, self.lipschitz_cost, self.rng)
        (loss + self.disc_updater.penalty(params)).backward()
        rank, world = sdist.world()
        if world > 1:
            for p in params:
                torch.distributed.all_reduce(p.grad)
                p.grad /= world
        self.disc_updater.step(params)
        return dict(is_discriminator=True, disc_loss=float(loss.detach()), accuracy=float(accuracy),
                    rate_penalty=float(rate), dynamics_penalty=float(dyn))

    def train_generator(self):
        tc, dyn, rate = self.generate(self.sample_z(), differentiable=True)
        loss = -self.critic(tc.float()).mean() + self.costs['dynamics_cost'] * dyn + self.costs['rate_cost'] * rate
        for p in self.gen_params:
            p.grad = None
        for p in self.critic.parameters():
            p.requires_grad_(False)
        loss.backward()
        for p in self.critic.parameters():
            p.requires_grad_(True)
        rank, world = sdist.world()
        dJ, dD, dS = sdist.allreduce_generator_grads(self.J.grad, self.D.grad, self.S.grad)
        self.J.grad, self.D.grad, self.S.grad = dJ / world, dD / world, dS / world
        if self.V is not None and world > 1:
            torch.distributed.all_reduce(self.V.grad)
            self.V.grad /= world
        self.gen_updater.step(self.gen_params)
        with torch.no_grad():
            for p in (self.J, self.D, self.S):
                p.clamp_(self.param_min, self.param_max)
            if self.V is not None:
                self.V.clamp_(self.V_min, self.V_max)
        return dict(is_discriminator=False, gen_loss=float(loss.detach()), dynamics_penalty=float(dyn.detach()),
                    rate_penalty=float(rate.detach()))

    def learning(self):
        """Generator of per-step info dicts, as BPTTWassersteinGAN.learning (wgan.py:439-444)."""
        gen_step, critic_iters = 0, self.critic_iters_init
        while True:
            for disc_step in range(critic_iters):
                yield dict(self.train_discriminator(), gen_step=gen_step, disc_step=disc_step)
            yield dict(self.train_generator(), gen_step=gen_step)
            gen_step, critic_iters = gen_step + 1, self.critic_iters

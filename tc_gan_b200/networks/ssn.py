"""
The BPTT tuning-curve generator -- mirror of tc_gan/networks/ssn.py with the Theano graph replaced by
the CUDA operators of ``tc_gan_b200.torch_ops``:

* `BandwidthContrastStimulator`  (networks/ssn.py:126-198)   stimulus from per-network (bandwidths, contrasts)
* `HeteroInputWrapper` / `DegenerateHeteroInputWrapper` (:645-745)   stimulus x (1 + V z_in), V learned
* `EulerSSNModel` and its heteroin variants (:579-772)   unrolled Euler dynamics -> time_avg + penalties
* `FixedProber` (:779-851)   time_avg[:, :, probes] -> [batch, num_tcdom * n_probes]
* `TuningCurveGenerator` (:862-968), `make_tuning_curve_generator` (:992-1018)

Inputs keep the reference's names (``stimulator_bandwidths``, ``stimulator_contrasts``,
``model_rate_penalty_threshold``, ``model_zs``, ``model_zs_in``, ``prober_*``).  `forward` returns numpy
values in a namedtuple as the compiled Theano function did; `get_output` returns torch tensors that stay in the
autograd graph (what the trainers differentiate).  Generator parameters J, D, S (and V) are float64 CUDA
tensors with ``requires_grad``.
"""
import collections

import numpy as np
import torch

from .. import ssnode, torch_ops


def _dev(device):
    return torch.device(device if device is not None else 'cuda')


def as_device_f32(a, device):
    if torch.is_tensor(a):
        return a.to(device=device, dtype=torch.float32)
    return torch.as_tensor(np.asarray(a, dtype=np.float32), device=device)


def concat_flat(arrays):
    flat = []
    for a in arrays:
        flat.extend(np.asarray(a).flat)
    return flat


def make_flat_param_names(parameters):
    """('J_EE', 'J_EI', ..., 'V_E', 'V_I' | 'V') as networks/ssn.py:105-123; `parameters` = [(name, tensor)]."""
    names = []
    for name, p in parameters:
        if p.dim() == 0:
            names.append(name)
        elif p.dim() == 1:
            names.extend([name + '_E', name + '_I'])
        elif p.dim() == 2:
            names.extend([name + s for s in ('_EE', '_EI', '_IE', '_II')])
        else:
            raise ValueError('Only ndim=1 and ndim=2 are supported. Given parameter {} has ndim={}.'
                             .format(name, p.dim()))
    return tuple(names)


class BandwidthContrastStimulator(object):
    r"""
    I_i(s) = A sigma((s/2 + x_i)/l) sigma((s/2 - x_i)/l), duplicated for the E and I populations
    (networks/ssn.py:126-198 == stimuli.py:3-10).  `bandwidths`, `contrasts`: (batchsize, num_tcdom);
    result (batchsize, num_tcdom, 2 num_sites) float32 on `device`.
    """

    inputs = ('bandwidths', 'contrasts')
    outputs = ()

    def __init__(self, num_sites, num_tcdom, smoothness=ssnode.DEFAULT_PARAMS['smoothness'], device=None):
        self.num_sites, self.num_tcdom, self.smoothness = int(num_sites), int(num_tcdom), float(smoothness)
        self.device = _dev(device)
        self.site_to_band = torch.linspace(-0.5, 0.5, self.num_sites, dtype=torch.float64, device=self.device)

    num_neurons = property(lambda self: self.num_sites * 2)

    def stimulus(self, bandwidths, contrasts):
        b = torch.as_tensor(np.asarray(bandwidths, dtype=np.float64), device=self.device)[..., None]
        c = torch.as_tensor(np.asarray(contrasts, dtype=np.float64), device=self.device)[..., None]
        assert b.shape[-2] == self.num_tcdom, (b.shape, self.num_tcdom)
        x = self.site_to_band.reshape(1, 1, -1)
        sig = lambda u: 1 / (1 + torch.exp(-u / self.smoothness))
        stim = c * sig(x + b / 2) * sig(b / 2 - x)
        return torch.cat([stim, stim], dim=-1).to(torch.float32)

    def get_all_params(self):
        return []

    def gen_noise(self, rng, **_):
        return {}


class HeteroInputWrapper(object):
    """Stimulator wrapper: stimulus x (1 + vpop z_in), V a 2-vector generator parameter (E, I)."""

    dist_in_choices = ('bernoulli', 'uniform')
    inputs = ('zs_in',) + BandwidthContrastStimulator.inputs

    def __init__(self, stimulator, V=0, dist_in='bernoulli'):
        assert dist_in in self.dist_in_choices
        self.stimulator, self.dist_in = stimulator, dist_in
        self.init_variability(V)

    def init_variability(self, V):
        V = np.ascontiguousarray(np.broadcast_to(np.asarray(V, dtype=float), 2))
        self.V = torch.tensor(V, dtype=torch.float64, device=self.stimulator.device, requires_grad=True)

    def __getattr__(self, name):
        return getattr(self.__dict__['stimulator'], name)

    def stimulus(self, bandwidths, contrasts, zs_in):
        base = self.stimulator.stimulus(bandwidths, contrasts)                   # (batch, tcdom, 2N)
        zs = as_device_f32(zs_in, base.device)
        n_sites = self.stimulator.num_sites
        vpop = self.V.expand(2) if self.V.dim() == 0 else self.V
        vs = torch.cat([vpop[0].expand(n_sites), vpop[1].expand(n_sites)]).to(base.dtype)
        return (1 + vs[None, None, :] * zs[:, None, :]) * base

    def get_all_params(self):
        return self.stimulator.get_all_params() + [('V', self.V)]

    def gen_noise(self, rng, stimulator_bandwidths, **_):
        shape = (np.shape(stimulator_bandwidths)[0], self.stimulator.num_neurons)
        if self.dist_in == 'bernoulli':
            return dict(zs_in=rng.choice(2, shape) * 2 - 1)
        return dict(zs_in=rng.rand(*shape) * 2 - 1)


class DegenerateHeteroInputWrapper(HeteroInputWrapper):
    """One scalar V for both populations (ssn_type 'deg-heteroin', the paper's runs)."""

    def init_variability(self, V):
        V = np.asarray(V, dtype=float)
        assert V.ndim == 0
        self.V = torch.tensor(float(V), dtype=torch.float64, device=self.stimulator.device, requires_grad=True)


ModelOut = collections.namedtuple('ModelOut', ['time_avg', 'dynamics_penalty', 'rate_penalty'])


class EulerSSNModel(object):
    """
    r_{t+1} = (1 - eps) r_t + eps f(W(z; J, D, S) r_t + I), r_0 = 0, unrolled `seqlen` steps on the GPU
    (networks/ssn.py:555-576); outputs as :619-633: `time_avg` (batch, num_tcdom, 2N) over t >= skip_steps,
    `dynamics_penalty`, `rate_penalty`.  Backward (BPTT to J, D, S and to the stimulus) is the CUDA adjoint
    recursion of torch_ops.EulerSSN.
    """

    ssn_type = 'default'
    outputs = ('dynamics_penalty', 'rate_penalty')

    def __init__(self, stimulator, J, D, S, k=ssnode.DEFAULT_PARAMS['k'], n=ssnode.DEFAULT_PARAMS['n'],
                 io_type='asym_tanh', tau_E=10, tau_I=1, dt=0.1, seqlen=1200, skip_steps=1000,
                 rate_soft_bound=ssnode.DEFAULT_PARAMS['rate_soft_bound'],
                 rate_hard_bound=ssnode.DEFAULT_PARAMS['rate_hard_bound'],
                 include_rate_penalty=True, include_time_avg=False, unroll_scan=False):
        self.stimulator = stimulator
        dev = stimulator.device
        mk = lambda a: torch.tensor(np.asarray(a, dtype=float), dtype=torch.float64, device=dev, requires_grad=True)
        self.J, self.D, self.S = mk(J), mk(D), mk(S)
        self.io = dict(io_type=io_type, k=k, n=n, rate_soft_bound=rate_soft_bound, rate_hard_bound=rate_hard_bound)
        self.tau_E, self.tau_I, self.dt = tau_E, tau_I, dt
        self.seqlen, self.skip_steps = int(seqlen), int(skip_steps)
        self.include_rate_penalty, self.include_time_avg = include_rate_penalty, include_time_avg
        if include_time_avg:
            self.outputs = self.outputs + ('time_avg',)

    inputs = ('zs', 'rate_penalty_threshold')
    num_sites = property(lambda self: self.stimulator.num_sites)
    num_neurons = property(lambda self: self.stimulator.num_neurons)
    num_tcdom = property(lambda self: self.stimulator.num_tcdom)

    def get_all_params(self):
        return [('J', self.J), ('D', self.D), ('S', self.S)] + self.stimulator.get_all_params()

    def gen_noise(self, rng, stimulator_bandwidths, **kwargs):
        batch = np.shape(stimulator_bandwidths)[0]
        noise = dict(zs=rng.rand(batch, self.num_neurons, self.num_neurons))          # networks/ssn.py:438
        noise.update(self.stimulator.gen_noise(rng, stimulator_bandwidths=stimulator_bandwidths, **kwargs))
        return noise

    def _stimulus(self, bandwidths, contrasts, zs_in=None):
        return self.stimulator.stimulus(bandwidths, contrasts)

    def run(self, zs, bandwidths, contrasts, rate_penalty_threshold, zs_in=None):
        """Differentiable forward: ModelOut of torch tensors."""
        dev = self.stimulator.device
        z = as_device_f32(zs, dev)
        ext = self._stimulus(bandwidths, contrasts, zs_in)
        avg, dyn, rate = torch_ops.euler_ssn(
            z, self.J, self.D, self.S, ext, seqlen=self.seqlen, skip_steps=self.skip_steps, dt=self.dt,
            tau_E=self.tau_E, tau_I=self.tau_I, rate_penalty_threshold=float(rate_penalty_threshold), **self.io)
        return ModelOut(avg, dyn, rate)


class HeteroInEulerSSNModel(EulerSSNModel):
    """SSN with heterogeneous (random) input (networks/ssn.py:748-768)."""

    ssn_type = 'heteroin'
    input_wrapper_class = HeteroInputWrapper
    inputs = ('zs_in',) + EulerSSNModel.inputs

    def __init__(self, stimulator, *args, V=0, dist_in='bernoulli', **kwargs):
        super(HeteroInEulerSSNModel, self).__init__(self.input_wrapper_class(stimulator, V=V, dist_in=dist_in),
                                                    *args, **kwargs)

    def _stimulus(self, bandwidths, contrasts, zs_in=None):
        assert zs_in is not None, 'heteroin models need model_zs_in'
        return self.stimulator.stimulus(bandwidths, contrasts, zs_in)


class DegenerateHeteroInEulerSSNModel(HeteroInEulerSSNModel):
    ssn_type = 'deg-heteroin'
    input_wrapper_class = DegenerateHeteroInputWrapper


def is_heteroin(gen):
    return hasattr(gen.model.stimulator, 'V')


class FixedProber(object):
    """
    Probe `time_avg` with constant `probes` (neuron indices): (batch, num_tcdom * len(probes)), the probe axis
    mixed into the tuning-curve domain, bandwidth-major (networks/ssn.py:779-851).

    >>> import numpy as np
    >>> time_avg = np.arange(3 * 5 * 7).reshape((3, 5, 7))
    >>> FixedProber(None, [0, 5]).probe_numpy(time_avg)[0].tolist()
    [0, 5, 7, 12, 14, 19, 21, 26, 28, 33]
    """

    inputs = ()
    outputs = ('tuning_curve',)

    def __init__(self, model, probes):
        self.model = model
        self.probes = np.asarray(probes, dtype=int)

    def probe_numpy(self, time_avg):
        tc = np.asarray(time_avg)[:, :, self.probes]
        return tc.reshape((tc.shape[0], -1))

    def tuning_curve(self, time_avg, **_):
        """Differentiable probe on the device through the probe kernels (gather fwd / scatter-add bwd)."""
        nz, nb, _ = time_avg.shape
        n_p = len(self.probes)
        dev = time_avg.device
        model_ids = torch.arange(nz, device=dev, dtype=torch.int32).repeat_interleave(n_p)
        probes = torch.as_tensor(self.probes, device=dev, dtype=torch.int32).repeat(nz)
        tc = torch_ops.probe_rates(time_avg, model_ids, probes)                 # (nz * n_p, nb)
        return tc.reshape(nz, n_p, nb).transpose(1, 2).reshape(nz, nb * n_p)


class TuningCurveGenerator(object):
    """Stimulator + model + prober bundled (networks/ssn.py:862-968)."""

    def __init__(self, stimulator, model, prober, batchsize):
        self.stimulator, self.model, self.prober, self.batchsize = stimulator, model, prober, batchsize
        self._input_names = (['stimulator_' + k for k in BandwidthContrastStimulator.inputs] +
                             ['model_' + k for k in self.model.inputs] +
                             ['prober_' + k for k in self.prober.inputs])
        out_names = ['model_' + k for k in self.model.outputs] + ['prober_' + k for k in self.prober.outputs]
        self.OutType = collections.namedtuple('OutType', out_names)

    num_tcdom = property(lambda self: self.model.num_tcdom)
    num_sites = property(lambda self: self.model.num_sites)
    num_neurons = property(lambda self: self.model.num_neurons)
    dt = property(lambda self: self.model.dt)
    probes = property(lambda self: self.prober.probes)
    output_shape = property(lambda self: (self.batchsize, self.num_tcdom * len(self.probes)))

    def gen_noise(self, rng, **kwargs):
        return {'model_' + k: v for k, v in self.model.gen_noise(rng, **kwargs).items()}

    def get_all_params(self):
        return self.model.get_all_params()

    def get_flat_param_names(self):
        return make_flat_param_names(self.get_all_params())

    def get_flat_param_values(self):
        return concat_flat(p.detach().cpu().numpy() for _, p in self.get_all_params())

    def set_params(self, params):
        rest = dict(params)
        with torch.no_grad():
            for name, p in self.get_all_params():
                if name in rest:
                    p.copy_(torch.as_tensor(np.asarray(rest.pop(name), dtype=float)).reshape(p.shape))
        if rest:
            raise ValueError('Unknown parameters: {}'.format(rest))

    def _mixin_noise(self, rng, kwargs):
        if rng is None:
            return kwargs
        noise = self.gen_noise(rng, **kwargs)
        return dict(noise, **kwargs)

    def _run(self, kwargs):
        kwargs = dict(kwargs)
        values = {k: kwargs.pop(k) for k in self._input_names}
        assert not kwargs, 'unknown inputs: {}'.format(sorted(kwargs))
        out = self.model.run(values['model_zs'], values['stimulator_bandwidths'], values['stimulator_contrasts'],
                             values['model_rate_penalty_threshold'], zs_in=values.get('model_zs_in'))
        prober_kw = {k[len('prober_'):]: v for k, v in values.items() if k.startswith('prober_')}
        return out, self.prober.tuning_curve(out.time_avg, **prober_kw)

    def get_output(self, rng=None, **kwargs):
        """Differentiable: (tuning_curve, dynamics_penalty, rate_penalty) as torch tensors."""
        out, tc = self._run(self._mixin_noise(rng, kwargs))
        return tc, out.dynamics_penalty, out.rate_penalty

    def forward(self, rng=None, **kwargs):
        """Numpy outputs in an `OutType` namedtuple (no graph), as the compiled Theano function."""
        with torch.no_grad():
            out, tc = self._run(self._mixin_noise(rng, kwargs))
        vals = {'model_dynamics_penalty': float(out.dynamics_penalty), 'model_rate_penalty': float(out.rate_penalty),
                'model_time_avg': out.time_avg.cpu().numpy(), 'prober_tuning_curve': tc.cpu().numpy()}
        return self.OutType(*[vals[k] for k in self.OutType._fields])

    def prepare(self):
        """Nothing to compile: the CUDA library is built ahead of time."""

    def to_config(self):
        m = self.model
        config = dict(num_sites=self.num_sites, num_tcdom=self.num_tcdom, smoothness=m.stimulator.smoothness,
                      tau_E=m.tau_E, tau_I=m.tau_I, dt=m.dt, seqlen=m.seqlen, skip_steps=m.skip_steps,
                      batchsize=self.batchsize, ssn_type=m.ssn_type, ssn_impl='default', **m.io)
        for name, p in self.get_all_params():
            config[name] = p.detach().cpu().numpy().tolist()
        if hasattr(self.prober, 'probes'):
            config['probes'] = [int(p) for p in self.prober.probes]
        return config


_ssn_classes = {
    'default': EulerSSNModel,
    'heteroin': HeteroInEulerSSNModel,
    'deg-heteroin': DegenerateHeteroInEulerSSNModel,
}
ssn_type_choices = tuple(_ssn_classes)


def ssn_type_of(gen):
    return gen.model.ssn_type


_MODEL_KEYS = ('k', 'n', 'io_type', 'tau_E', 'tau_I', 'dt', 'seqlen', 'skip_steps', 'rate_soft_bound',
               'rate_hard_bound', 'include_rate_penalty', 'include_time_avg', 'unroll_scan')


def make_tuning_curve_generator(config, consume_union=True, emit_prober=None, emit_tcg=None, device=None, **kwargs):
    """
    ``(TuningCurveGenerator, unused part of config)`` as networks/ssn.py:992-1018.  Consumed keys:
    num_sites, num_tcdom, smoothness, J, D, S, ssn_type, V, dist_in, the model keys (k, n, io_type, tau_E,
    tau_I, dt, seqlen, skip_steps, ...), probes, batchsize.
    """
    rest = dict(config, **kwargs)
    rest.pop('ssn_impl', None)
    stimulator = BandwidthContrastStimulator(rest.pop('num_sites'), rest.pop('num_tcdom'),
                                             rest.pop('smoothness', ssnode.DEFAULT_PARAMS['smoothness']), device=device)
    ssn_type = rest.pop('ssn_type', 'default')
    model_kw = {k: rest.pop(k) for k in _MODEL_KEYS if k in rest}
    if ssn_type != 'default':
        for k in ('V', 'dist_in'):
            if k in rest:
                model_kw[k] = rest.pop(k)
    model = _ssn_classes[ssn_type](stimulator, rest.pop('J'), rest.pop('D'), rest.pop('S'), **model_kw)
    prober = emit_prober(model) if emit_prober else FixedProber(model, rest.pop('probes'))
    tcg = (emit_tcg or TuningCurveGenerator)(stimulator, model, prober, rest.pop('batchsize'))
    if consume_union:
        for key in ('V', 'dist_in'):
            rest.pop(key, None)
    return tcg, rest

"""
SSN fixed-point solver API -- mirror of tc_gan/ssnode.py over the CUDA library.

Same names, arguments, return types and error behaviour as the reference
(``fixed_point``, ``solve_dynamics``, ``find_fixed_points``, ``make_io_fun``,
``sample_fixed_points``, ``sample_tuning_curves``, ``DEFAULT_PARAMS``,
``FixedPointResult`` / ``FixedPointError`` / ``FixedPointsInfo``), but:

* ``fixed_point`` crosses the same ctypes boundary (ssnode.py:244-255) into a
  float64 CUDA kernel instead of the C loop;
* ``find_fixed_points`` replaces the thread pool over (network, stimulus) solves
  (ssnode.py:390-510) by batched launches: every drawn network is solved for all
  stimuli at once by persistent thread-block clusters.

There is no CPU solver in this module.
"""
from __future__ import print_function, division

import collections
import ctypes
import itertools

import numpy as np

from . import clib
from .clib import libssnode, double_ptr
from .gradient_expressions.utils import subsample_neurons


_J0 = np.array([[.0957, .0638], [.1197, .0479]])
_D0 = np.array([[.7660, .5106], [.9575, .3830]])
_S0 = np.array([[.6667, .2], [1.333, .2]]) / 8

# Same keys and values as tc_gan.ssnode.DEFAULT_PARAMS (ssnode.py:27-41).
DEFAULT_PARAMS = {
    'N': 102, 'J': _J0, 'D': _D0, 'S': _S0,
    'bandwidths': [0, 0.0625, 0.125, 0.1875, 0.25, 0.5, 0.75, 1],
    'smoothness': 0.25 / 8, 'contrast': [20], 'offset': [0],
    'io_type': 'asym_tanh', 'k': 0.01, 'n': 2.2,
    'rate_soft_bound': 200, 'rate_hard_bound': 1000,
    'tau': (0.01589, 0.002),
}

_IO_TYPES = ('asym_linear', 'asym_tanh', 'asym_power')
_MESSAGES = {0: "Converged", 1: "SSN Convergence Failed", 2: "Reached to rate_stop_at"}


def new_JDS():
    """More stable generator parameters (networks/fixed_time_sampler.py:12-23): D/2, J + D/4."""
    return dict(J=_J0 + _D0 / 4, D=_D0 / 2, S=_S0.copy())


class FixedPointResult(object):
    """Outcome of one solve: state `x`, reference error code `error`, `message`, `success`;
    `iterations` is filled by the batched path (the reference does not report it)."""

    __slots__ = ('x', 'error', 'message', 'iterations')

    def __init__(self, x, error, iterations=None):
        self.x, self.error, self.iterations, self.message = x, error, iterations, None

    success = property(lambda self: self.error == 0)

    def to_exception(self):
        return FixedPointError(self.message, self)


class FixedPointError(Exception):
    """Raised with ``check=True``; carries the failed `FixedPointResult` as ``.result``."""

    def __init__(self, message, result):
        Exception.__init__(self, message)
        self.result = result


def take(n, iterable):
    return list(itertools.islice(iterable, n))


def make_neu_vec(N, E, I):
    """2N-vector from population-level values."""
    return np.repeat(np.array([E, I]), N)


def any_to_neu_vec(N, vec):
    vec = np.asarray(vec)
    return make_neu_vec(N, *vec) if len(vec) == 2 else vec


def thlin(x):
    return x * (x > 0)


def rate_to_volt(rate, k, n):
    return (rate / k) ** (1 / n)


def _xp(a):
    """numpy, or torch for torch tensors (the reference dispatches numpy / Theano the same way)."""
    try:
        import torch
        if isinstance(a, torch.Tensor):
            return torch
    except ImportError:
        pass
    return np


def io_power(v, k, n):
    return k * thlin(v) ** n


def io_alin(v, volt_max, k, n):
    """Power law up to volt_max, then its tangent line (ssnode.py:129-134)."""
    xp = _xp(v)
    below = k * xp.clip(v, 0, volt_max) ** n
    slope = k * n * volt_max ** (n - 1)
    return xp.where(v <= volt_max, below, below + slope * (v - volt_max))


def io_atanh(v, r0, r1, v0, k, n):
    """Power law up to v0 (rate r0), then a tanh saturating at r1 (ssnode.py:142-149)."""
    xp = _xp(v)
    below = k * xp.clip(v, 0, v0) ** n
    span = r1 - r0
    above = r0 + span * xp.tanh(n * r0 / span * (v - v0) / v0)
    return xp.where(v <= v0, below, above)


def make_io_fun(k, n,
                rate_soft_bound=DEFAULT_PARAMS['rate_soft_bound'],
                rate_hard_bound=DEFAULT_PARAMS['rate_hard_bound'],
                io_type=DEFAULT_PARAMS['io_type']):
    """Elementwise transfer function on numpy arrays / torch tensors (ssnode.py:276-292)."""
    v0 = rate_to_volt(rate_soft_bound, k, n)
    table = {
        'asym_linear': lambda v: io_alin(v, v0, k, n),
        'asym_tanh': lambda v: io_atanh(v, rate_soft_bound, rate_hard_bound, v0, k, n),
        'asym_power': lambda v: io_power(v, k, n),
    }
    if io_type not in table:
        raise ValueError("Unknown I/O type: {}".format(io_type))
    return table[io_type]


def solve_dynamics(*args, **kwds):
    """`fixed_point(...).x`, printing the message of a failed solve (ssnode.py:152-156)."""
    sol = fixed_point(*args, **kwds)
    if not sol.success:
        print(sol.message)
    return sol.x


def _set_message(sol):
    """Error code -> message exactly as ssnode.py:256-270 (a converged but non-finite state is error 1)."""
    code = sol.error
    if code == 0 and not np.isfinite(sol.x).all():
        sol.error, sol.message = 1, "Converged to non-finite value"
    elif code in _MESSAGES:
        sol.message = _MESSAGES[code]
    elif code > 900:
        sol.message = "CUDA error {}: {}".format(
            code - 1000, libssnode.ssn_last_error().decode('utf-8', 'replace'))
    else:
        sol.message = "Unknown error: code={}".format(code)
    return sol


def fixed_point(
        W, ext, k, n, r0=None, tau=DEFAULT_PARAMS['tau'],
        max_iter=10000, atol=1e-5, dt=.0008, solver='euler',
        rate_soft_bound=DEFAULT_PARAMS['rate_soft_bound'],
        rate_hard_bound=DEFAULT_PARAMS['rate_hard_bound'],
        rate_stop_at=np.inf,
        io_type='asym_tanh', check=False):
    """
    Solve the SSN ODE for one (W, ext) until it reaches a fixed point.

    Same signature and result as tc_gan.ssnode.fixed_point (ssnode.py:159-273): `W` is
    (2N, 2N), `ext` and `r0` (2N,), the result a `FixedPointResult`.  The Euler loop runs on
    the GPU in float64 behind the reference's own C symbol ``solve_dynamics_{io_type}_{solver}``.
    Unlike the reference, non-contiguous inputs are honoured (it passes the raw buffer).  A
    failure of the GPU call itself raises `clib.SSNLibraryError`: there is no CPU fallback.
    """
    if io_type not in _IO_TYPES:
        raise ValueError("Unknown I/O type: {}".format(io_type))
    if solver != 'euler':
        raise ValueError("Unknown solver: {}".format(solver))

    Wc = np.ascontiguousarray(W, dtype=np.float64)
    ec = np.ascontiguousarray(ext, dtype=np.float64)
    dim = Wc.shape[0]
    if Wc.ndim != 2 or Wc.shape != (dim, dim) or dim % 2 or ec.shape != (dim,):
        raise AssertionError('W must be (2N, 2N) and ext (2N,)')
    state = np.zeros(dim) if r0 is None else np.array(r0, dtype=np.float64)   # copy: overwritten by the solver
    if state.shape != (dim,):
        raise AssertionError('r0 must be (2N,)')
    scratch = np.empty(dim)
    # power / linear transfer functions have no saturation: the hard bound is the stop criterion
    bound = rate_hard_bound if io_type == 'asym_tanh' else rate_stop_at

    symbol = getattr(libssnode, 'solve_dynamics_{}_{}'.format(io_type, solver))
    code = symbol(dim // 2, Wc.ctypes.data_as(double_ptr), ec.ctypes.data_as(double_ptr),
                  float(k), float(n), state.ctypes.data_as(double_ptr), scratch.ctypes.data_as(double_ptr),
                  tau[0], tau[1], dt, max_iter, atol, rate_soft_bound, bound)
    if code > 900:
        clib.check_call(code, symbol.__name__)
    sol = _set_message(FixedPointResult(state, code))
    if check and not sol.success:
        raise sol.to_exception()
    return sol


FixedPointsInfo = collections.namedtuple('FixedPointsInfo', [
    'solutions', 'counter', 'rejections', 'unused',
])
null_FixedPointsInfo = FixedPointsInfo(None, None, 0, 0)


def fixed_points_batch(Ws, exts, k, n, r0=None, tau=DEFAULT_PARAMS['tau'],
                       max_iter=10000, atol=1e-5, dt=.0008, solver='euler',
                       rate_soft_bound=DEFAULT_PARAMS['rate_soft_bound'],
                       rate_hard_bound=DEFAULT_PARAMS['rate_hard_bound'],
                       rate_stop_at=np.inf, io_type='asym_tanh', precise=False):
    """
    All (network, stimulus) solves of ``Ws`` [nz, 2N, 2N] x ``exts`` [nb, 2N] in one
    batched GPU call.  Returns ``(Rs [nz, nb, 2N] float64, errors [nz, nb], iters [nz, nb])``
    with the reference's error codes.  ``precise=True`` runs the float64 kernel;
    the default contracts in FP32 with a float64 state update.
    """
    if solver not in ('euler',):
        raise ValueError("Unknown solver: {}".format(solver))
    Ws = np.ascontiguousarray(Ws, dtype='double')
    exts = np.ascontiguousarray(exts, dtype='double')
    nz, dim = Ws.shape[0], Ws.shape[1]
    nb = exts.shape[0]
    assert Ws.shape == (nz, dim, dim) and exts.shape == (nb, dim) and dim % 2 == 0
    sv = clib.make_solver(io_type=io_type, k=k, n=n, tau=tau, dt=dt, max_iter=max_iter, atol=atol,
                          rate_soft_bound=rate_soft_bound, rate_hard_bound=rate_hard_bound,
                          rate_stop_at=rate_stop_at)
    r_init = None
    if r0 is not None:
        r0 = np.asarray(r0, dtype='double')
        if r0.any():
            r_init = np.ascontiguousarray(np.broadcast_to(r0, (nz, nb, dim)))
    Rs = np.empty((nz, nb, dim))
    errors = np.empty((nz, nb), dtype=np.int32)
    iters = np.empty((nz, nb), dtype=np.int32)
    clib.check_call(libssnode.ssn_fixed_point_batch_f64(
        sv, nz, nb, dim // 2, Ws.ctypes.data, exts.ctypes.data,
        None if r_init is None else r_init.ctypes.data,
        Rs.ctypes.data, errors.ctypes.data, iters.ctypes.data, int(bool(precise))),
        'ssn_fixed_point_batch_f64')
    return Rs, errors, iters


class _Solutions(object):
    """`FixedPointsInfo.solutions` of the batched finder: behaves as the reference's tuple (one list of `nb`
    converged `FixedPointResult` per kept network, ssnode.py:503-510) but builds the result objects on access --
    8192 Python objects per 1024-network call would cost more than the solves."""

    def __init__(self, xs, its):
        self._xs, self._its = xs, its

    def __len__(self):
        return len(self._xs)

    def __getitem__(self, k):
        if isinstance(k, slice):
            return tuple(self[i] for i in range(*k.indices(len(self))))
        return [_set_message(FixedPointResult(x, 0, int(i))) for x, i in zip(self._xs[k], self._its[k])]

    def __iter__(self):
        return (self[i] for i in range(len(self)))


def _solve_drawn(drawn, exts, jds, precise, host_threads, **kwargs):
    """One batched GPU call for a list of drawn (Z, W): the streamed pointer-list entry point for the default
    FP32-contraction kernel (no concatenation of the per-network arrays), `fixed_points_batch` for `precise`."""
    if precise:
        Ws = np.array([np.asarray(W, dtype='double') for _, W in drawn]) if jds is None else np.array(
            [_generate_weight_from_z(Z, jds) for Z, _ in drawn])
        return fixed_points_batch(Ws, exts, precise=True, **kwargs)
    solver = kwargs.pop('solver', 'euler')
    if solver != 'euler':
        raise ValueError("Unknown solver: {}".format(solver))
    r0 = kwargs.pop('r0', None)
    sv = clib.make_solver(**kwargs)
    items = [np.asarray(Z if jds is not None else W) for Z, W in drawn]
    f32 = all(it.dtype == np.float32 for it in items)                 # float32 z / W is staged without conversion
    items = [np.ascontiguousarray(it, dtype=np.float32 if f32 else np.float64) for it in items]
    nz, dim, nb = len(items), items[0].shape[0], len(exts)
    for it in items:
        assert it.shape == (dim, dim), 'every network must be (2N, 2N)'
    assert exts.shape == (nb, dim) and dim % 2 == 0
    ptrs = (ctypes.c_void_p * nz)(*[it.ctypes.data for it in items])
    r_init = None
    if r0 is not None:
        r0 = np.asarray(r0, dtype='double')
        if r0.any():
            r_init = np.ascontiguousarray(np.broadcast_to(r0, (nz, nb, dim)))
    Rs = np.empty((nz, nb, dim))
    errors = np.empty((nz, nb), dtype=np.int32)
    iters = np.empty((nz, nb), dtype=np.int32)
    clib.check_call(libssnode.ssn_fixed_point_batch_ptrs(
        sv, nz, nb, dim // 2, clib.W_FROM_Z if jds is not None else clib.W_DENSE, ptrs, int(f32),
        None if jds is None else clib.make_jds(jds['J'], jds['D'], jds['S']), exts.ctypes.data,
        None if r_init is None else r_init.ctypes.data, Rs.ctypes.data, errors.ctypes.data, iters.ctypes.data,
        int(host_threads)), 'ssn_fixed_point_batch_ptrs')
    return Rs, errors, iters


def _stack(arrays, host_threads=0):
    """np.array(arrays) -- with the copy spread over host threads when the elements are equal-shaped contiguous
    ndarrays (the kept Z of a 1024-network call are 1.3 GB)."""
    first = arrays[0]
    if not (isinstance(first, np.ndarray) and first.nbytes >= 1 << 16 and len(arrays) > 1 and all(
            isinstance(a, np.ndarray) and a.shape == first.shape and a.dtype == first.dtype and
            a.flags['C_CONTIGUOUS'] for a in arrays)):
        return np.array(arrays)
    out = np.empty((len(arrays),) + first.shape, dtype=first.dtype)
    ptrs = (ctypes.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])
    clib.check_call(libssnode.ssn_host_gather(ptrs, len(arrays), first.nbytes, out.ctypes.data, int(host_threads)),
                    'ssn_host_gather')
    return out


class _Gather(object):
    """Copy of the Z of one round of draws into rows [row0, row0 + n) of the output array, on a host thread that runs
    beside the GPU call (both release the GIL)."""

    def __init__(self, zs, out, row0, host_threads):
        import threading
        self.out = out
        ptrs = (ctypes.c_void_p * len(zs))(*[z.ctypes.data for z in zs])
        dst = out.ctypes.data + row0 * zs[0].nbytes
        self._keep = (zs, ptrs)
        self._rc = None

        def run():
            self._rc = libssnode.ssn_host_gather(ptrs, len(zs), zs[0].nbytes, dst, int(host_threads))
        self._thread = threading.Thread(target=run)
        self._thread.start()

    def join(self):
        self._thread.join()
        clib.check_call(self._rc, 'ssn_host_gather')
        return self.out


def _start_gather(drawn, out, row0, num, host_threads):
    """Start gathering the Z of `drawn` into the result array (allocated on the first round) when they are
    equal-shaped contiguous ndarrays worth a threaded copy; None otherwise (the caller then stacks at the end)."""
    zs = [Z for Z, _ in drawn]
    first = zs[0]
    if not (isinstance(first, np.ndarray) and first.nbytes >= 1 << 16 and all(
            isinstance(z, np.ndarray) and z.shape == first.shape and z.dtype == first.dtype and
            z.flags['C_CONTIGUOUS'] for z in zs)):
        return None
    if out is None:
        if row0:
            return None
        out = np.empty((num,) + first.shape, dtype=first.dtype)
    elif out.shape[1:] != first.shape or out.dtype != first.dtype:
        return None
    # half of the threads: the solver's own staging threads run at the same time (measured at 1024 networks x 1.3 MB:
    # 4 threads need 110 ms for float64 Z, longer than the 85 ms solve they are meant to hide under)
    return _Gather(zs, out, row0, max(1, (host_threads or 16) // 2))


def _generate_weight_from_z(Z, jds):
    from .weight_gen import generate_weight
    Z = np.asarray(Z, dtype='double')
    return generate_weight(Z.shape[0] // 2, jds['J'], jds['D'], jds['S'], Z)


def find_fixed_points(num, Z_W_gen, exts, method='parallel', **common_kwargs):
    """
    Find `num` sets of fixed points using weight matrices from `Z_W_gen`.

    Same contract as tc_gan.ssnode.find_fixed_points (ssnode.py:332-387): a
    network is kept only if every stimulus in `exts` converges; returns
    ``(Zs [num, ...], Rs [num, NB, 2N], FixedPointsInfo)``.

    Both ``method='parallel'`` and ``'serial'`` run the same batched GPU path
    (``method`` is validated for compatibility).  Networks are drawn from
    `Z_W_gen` in rounds of exactly as many as are still missing, so the generator
    is consumed precisely as far as the reference's serial finder would consume
    it, the kept networks are the first `num` successes in generator order
    (ssnode.py:489-495) and ``info.unused`` is 0.

    Extra keyword arguments understood here: those of `fixed_point`, plus

    * ``precise`` -- float64 kernel instead of the FP32-contraction / float64-state one;
    * ``jds=dict(J=, D=, S=)`` -- the generator's ``Z`` is the uniform noise z itself and W is built on
      the GPU from (J, D, S): the ``W`` element of each pair is ignored (may be ``None``), so a caller
      that owns the generator parameters skips its own weight construction and the float64 W never
      crosses the bus;
    * ``host_threads`` -- threads staging the matrices into pinned memory (default: up to 16);
    * the thread-pool knobs ``no_pool``, ``resubmit_threshold``, ``deterministic``, which are accepted and
      have no effect (there is no pool: results are always deterministic, nothing is over-submitted).
    """
    if method not in ('parallel', 'serial'):
        raise ValueError('Unknown method: {}'.format(method))
    kwargs = dict(common_kwargs)
    for ignored in ('no_pool', 'resubmit_threshold', 'deterministic'):
        kwargs.pop(ignored, None)
    check = kwargs.pop('check', False)
    jds = kwargs.pop('jds', None)
    precise = kwargs.pop('precise', False)
    host_threads = kwargs.pop('host_threads', 0)
    if 'io_type' in kwargs and kwargs['io_type'] not in _IO_TYPES:
        raise ValueError("Unknown I/O type: {}".format(kwargs['io_type']))
    exts = np.ascontiguousarray(exts, dtype='double')
    Z_W_gen = iter(Z_W_gen)

    kept_z, kept_R, kept_it = [], [], []
    counter = collections.Counter()
    # The stacked `Zs` the reference returns (np.array(zs), ssnode.py:503-510) is a 1.3 GB copy at the benchmark
    # size.  It is written WHILE the GPU solves the round: a host thread gathers the drawn Z into their final rows
    # (most draws converge), and only the rows of rejected networks are closed up afterwards.
    zs_out, zs_rows = None, 0
    while zs_rows + len(kept_z) < num:
        drawn = take(num - zs_rows - len(kept_z), Z_W_gen)
        if not drawn:
            break
        gather = _start_gather(drawn, zs_out, zs_rows, num, host_threads) if not kept_z else None
        if gather is None and zs_out is not None:           # draws that no longer fit the array: back to a list
            kept_z, zs_out, zs_rows = [z for z in zs_out[:zs_rows]], None, 0
        try:
            Rs, errors, iters = _solve_drawn(drawn, exts, jds, precise, host_threads, **kwargs)
        finally:
            if gather is not None:
                zs_out = gather.join()
        # a NaN or Inf anywhere in a network's rates survives the sum (rates are bounded by rate_hard_bound, no overflow):
        # one pass over Rs instead of a boolean copy of it
        ok = (errors == 0).all(axis=1) & np.isfinite(Rs.sum(axis=(1, 2)))
        for i in np.flatnonzero(~ok):
            err = errors[i].copy()
            err[(err == 0) & ~np.isfinite(Rs[i]).all(axis=1)] = 1          # ssnode.py:257-262
            # the reference visits stimuli last to first and records the first failure
            b = max(np.flatnonzero(err != 0))
            sol = _set_message(FixedPointResult(Rs[i, b], int(errors[i, b]), int(iters[i, b])))
            counter[sol.error] += 1
            if check:
                raise sol.to_exception()
        all_ok = bool(ok.all())
        if gather is not None:
            good = np.flatnonzero(ok)
            if len(good) < len(drawn):                     # close up the rows of the rejected networks
                for dst, src in enumerate(good):
                    if dst != src:
                        zs_out[zs_rows + dst] = zs_out[zs_rows + src]
            zs_rows += len(good)
        else:
            kept_z.extend(drawn[i][0] for i in np.flatnonzero(ok))
        kept_R.append(Rs if all_ok else Rs[ok])
        kept_it.append(iters if all_ok else iters[ok])

    if zs_out is not None:
        if zs_rows == 0:
            raise ValueError('find_fixed_points: Z_W_gen was exhausted before any network converged')
        xs = kept_R[0] if len(kept_R) == 1 else np.concatenate(kept_R)
        its = kept_it[0] if len(kept_it) == 1 else np.concatenate(kept_it)
        Zs = zs_out if zs_rows == len(zs_out) else zs_out[:zs_rows].copy()
        return Zs, xs, FixedPointsInfo(_Solutions(xs, its), counter, sum(counter.values()), 0)
    if not kept_z:
        raise ValueError('find_fixed_points: Z_W_gen was exhausted before any network converged')
    xs, its = np.concatenate(kept_R), np.concatenate(kept_it)
    return _stack(kept_z, host_threads), xs, FixedPointsInfo(_Solutions(xs, its), counter, sum(counter.values()), 0)


def find_fixed_points_serial(num, Z_W_gen, exts, **common_kwargs):
    return find_fixed_points(num, Z_W_gen, exts, method='serial', **common_kwargs)


def find_fixed_points_parallel(num, Z_W_gen, exts, **common_kwargs):
    return find_fixed_points(num, Z_W_gen, exts, method='parallel', **common_kwargs)


def make_solver_params(
        N=DEFAULT_PARAMS['N'],
        J=DEFAULT_PARAMS['J'],
        D=DEFAULT_PARAMS['D'],
        S=DEFAULT_PARAMS['S'],
        io_type=DEFAULT_PARAMS['io_type'],
        seed=65,
        bandwidth=1,
        smoothness=DEFAULT_PARAMS['smoothness'],
        contrast=DEFAULT_PARAMS['contrast'],
        k=DEFAULT_PARAMS['k'],
        n=DEFAULT_PARAMS['n'],
        ):
    """One seeded (W, ext) problem, as ssnode.py:526-558."""
    from . import stimuli
    from .weight_gen import generate_weight

    rs = np.random.RandomState(seed) if isinstance(seed, int) else seed
    Z = rs.rand(1, 2*N, 2*N)
    W = generate_weight(N, J, D, S, Z[0])
    X = np.linspace(-0.5, 0.5, N)
    ext, = stimuli.input([bandwidth], X, smoothness, contrast)
    return dict(W=W, ext=ext, r0=np.zeros(W.shape[0]), k=k, n=n, io_type=io_type)


def sample_fixed_points(
        NZ=30, seed=0,
        N=DEFAULT_PARAMS['N'],
        J=DEFAULT_PARAMS['J'],
        D=DEFAULT_PARAMS['D'],
        S=DEFAULT_PARAMS['S'],
        bandwidths=DEFAULT_PARAMS['bandwidths'],
        smoothness=DEFAULT_PARAMS['smoothness'],
        contrast=DEFAULT_PARAMS['contrast'],
        offset=DEFAULT_PARAMS['offset'],
        io_type=DEFAULT_PARAMS['io_type'],
        k=DEFAULT_PARAMS['k'],
        n=DEFAULT_PARAMS['n'],
        **solver_kwargs):
    """Seeded rejection sampling of NZ networks, as ssnode.py:561-590."""
    from . import stimuli
    from .weight_gen import generate_weight

    X = np.linspace(-0.5, 0.5, N)
    exts = stimuli.input(bandwidths, X, smoothness, contrast, offset)
    rs = np.random.RandomState(seed)

    def Z_W_gen():
        while True:
            z = rs.rand(1, 2*N, 2*N)
            yield z[0], generate_weight(N, J, D, S, z[0])

    solver_kwargs.setdefault('r0', np.zeros(2 * N))
    solver_kwargs.update(k=k, n=n, io_type=io_type)
    return find_fixed_points(NZ, Z_W_gen(), exts, **solver_kwargs)


def sample_tuning_curves(sample_sites=[0], track_offset_identity=False,
                         include_inhibitory_neurons=False,
                         **kwargs):
    """Tuning curves of sampled networks, as ssnode.py:593-602."""
    _, rates, _ = sample = sample_fixed_points(**kwargs)
    rates = np.array(rates)
    tunings = subsample_neurons(
        rates, sample_sites,
        include_inhibitory_neurons=include_inhibitory_neurons,
        track_offset_identity=track_offset_identity).T
    return tunings, sample

"""
PyTorch-facing operators over libssnode.so (device pointers through the C ABI).

These replace the Theano ops of the reference on the SSN path:

* `fixed_points` / `SSNFixedPoint`  -- batched fixed points of W(z; J, D, S) with the
  implicit-function-theorem gradient w.r.t. (J, D, S) in backward
  (tc_gan/ssnode.py:332-510 + gradient_expressions/SS_grad.py:17-76 +
  make_w_batch.py:36-121 + run/gan.py:902-911).
* `euler_ssn` / `EulerSSN` -- fixed-length unrolled Euler dynamics with
  time-average / dynamics-penalty / rate-penalty outputs and BPTT in backward
  (tc_gan/networks/ssn.py:555-576, 598-633).

torch is used for device memory, streams and autograd bookkeeping only; all
arithmetic on the path is in the CUDA library.  CPU tensors are rejected.
"""
import weakref

import torch

from . import clib
from .clib import libssnode


_jds_cache = {'refs': None, 'versions': None, 'struct': None}


def _jds_struct(J, D, S):
    """Host copy of (J, D, S) for the kernels' by-value constants.  The device->host read drains the stream, so
    it is done once per parameter VERSION: the cache is keyed on the identity of the three tensor objects (weak
    references, so a recycled address can never alias) and their in-place version counters -- one read after
    each optimizer step, none for repeated calls with unchanged parameters -- and uses a single copy.
    (Writing through ``.data`` does not bump the version counter; use in-place ops under ``no_grad``.)"""
    params = (J, D, S)
    refs, versions = _jds_cache['refs'], tuple(t._version for t in params)
    if refs is None or versions != _jds_cache['versions'] or any(r() is not t for r, t in zip(refs, params)):
        flat = torch.cat([t.detach().reshape(-1).double() for t in params]).cpu().numpy()
        _jds_cache['struct'] = clib.make_jds(flat[0:4], flat[4:8], flat[8:12])
        _jds_cache['refs'] = tuple(weakref.ref(t) for t in params)
        _jds_cache['versions'] = versions
    return _jds_cache['struct']


def _check_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise clib.SSNLibraryError('tc_gan_b200.torch_ops needs CUDA tensors (there is no CPU fallback)')


def _f32c(t):
    return t.detach().to(torch.float32).contiguous()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def make_solver(**kw):
    """Solver knobs with the defaults of tc_gan.ssnode.fixed_point (ssnode.py:159-165)."""
    return clib.make_solver(**kw)


def fixed_points(z, J, D, S, ext, solver=None, r_init=None, precise=False):
    """
    Fixed points of the networks W(z; J, D, S) for every stimulus.

    z [nz, 2N, 2N] float32 CUDA, ext [nb, 2N] or [nz, nb, 2N]; returns
    (R [nz, nb, 2N] float32, status [nz, nb] int32, iters [nz, nb] int32) with the
    reference's error codes (0 converged, 1 max_iter / non-finite, 2 rate_stop_at).
    """
    _check_cuda(z, ext, r_init)
    solver = solver or make_solver()
    z32, e32 = _f32c(z), _f32c(ext)
    nz, dim = z32.shape[0], z32.shape[1]
    nb = e32.shape[-2]
    R = torch.empty((nz, nb, dim), dtype=torch.float32, device=z.device)
    status = torch.empty((nz, nb), dtype=torch.int32, device=z.device)
    iters = torch.empty((nz, nb), dtype=torch.int32, device=z.device)
    r0 = _f32c(r_init) if r_init is not None else None
    with torch.cuda.device(z.device):
        clib.check_call(libssnode.ssn_fixed_point_batch(
            solver, nz, nb, dim // 2, clib.W_FROM_Z, z32.data_ptr(), _jds_struct(J, D, S), e32.data_ptr(),
            int(e32.dim() == 3), None if r0 is None else r0.data_ptr(), R.data_ptr(), status.data_ptr(),
            iters.data_ptr(), int(bool(precise)), clib.MEM_DEVICE, _stream()), 'ssn_fixed_point_batch')
    return R, status, iters


def ift_gradient(z, J, D, S, ext, R, grad_R, solver=None, rtol=1e-6, return_mu=False, return_grad_ext=False):
    """dL/d(J, D, S) (three float64 [2, 2] CUDA tensors) from dL/dR at the fixed points R;
    with return_grad_ext also dL/d ext [nz, nb, 2N] (= Phi mu, for heterogeneous-input generators)."""
    _check_cuda(z, ext, R, grad_R)
    solver = solver or make_solver()
    z32, e32, R32, g32 = _f32c(z), _f32c(ext), _f32c(R), _f32c(grad_R)
    nz, dim = z32.shape[0], z32.shape[1]
    nb = R32.shape[1]
    grad = torch.empty(12, dtype=torch.float64, device=z.device)
    mu = torch.empty_like(R32) if return_mu else None
    gext = torch.zeros_like(R32) if return_grad_ext else None
    status = torch.empty((nz, nb), dtype=torch.int32, device=z.device)
    iters = torch.empty((nz, nb), dtype=torch.int32, device=z.device)
    with torch.cuda.device(z.device):
        clib.check_call(libssnode.ssn_ift_gradient_batch(
            solver, nz, nb, dim // 2, z32.data_ptr(), _jds_struct(J, D, S), e32.data_ptr(), int(e32.dim() == 3),
            R32.data_ptr(), g32.data_ptr(), float(rtol), grad.data_ptr(), None if mu is None else mu.data_ptr(),
            status.data_ptr(), iters.data_ptr(), None if gext is None else gext.data_ptr(),
            clib.MEM_DEVICE, _stream()), 'ssn_ift_gradient_batch')
    dJ, dD, dS = grad[0:4].reshape(2, 2), grad[4:8].reshape(2, 2), grad[8:12].reshape(2, 2)
    out = (dJ, dD, dS)
    if return_mu:
        out = out + (mu, status, iters)
    if return_grad_ext:
        out = out + (gext,)
    return out


# (status, iters) of the adjoint solves of the most recent implicit-gradient backward, device tensors
# [nz, nb]: status 1 = the adjoint solve did not reach rtol (marginally stable fixed point).  Reading
# them synchronises; `adjoint_failures()` does so on demand.
last_adjoint = {'status': None, 'iters': None}


def adjoint_failures():
    st = last_adjoint['status']
    return 0 if st is None else int((st != 0).sum())


def _ift_backward(ctx, z, J, D, S, ext, R, grad_R):
    need_ext = ctx.needs_input_grad[4]
    res = ift_gradient(z, J, D, S, ext, R, grad_R, solver=ctx.solver, return_mu=True, return_grad_ext=need_ext)
    dJ, dD, dS = res[:3]
    last_adjoint['status'], last_adjoint['iters'] = res[4], res[5]
    g_ext = None
    if need_ext:
        g_ext = res[6] if ext.dim() == 3 else res[6].sum(dim=0)
        g_ext = g_ext.to(ext.dtype)
    return (None, dJ.to(J.dtype).to(J.device), dD.to(D.dtype).to(D.device), dS.to(S.dtype).to(S.device), g_ext)


class SSNFixedPoint(torch.autograd.Function):
    """R = fixed_point(W(z; J, D, S), ext); backward = implicit gradient w.r.t. J, D, S."""

    @staticmethod
    def forward(ctx, z, J, D, S, ext, solver, precise):
        R, status, iters = fixed_points(z, J, D, S, ext, solver=solver, precise=precise)
        ctx.save_for_backward(z, J, D, S, ext, R)
        ctx.solver = solver
        ctx.mark_non_differentiable(status, iters)
        return R, status, iters

    @staticmethod
    def backward(ctx, grad_R, _gs, _gi):
        z, J, D, S, ext, R = ctx.saved_tensors
        return _ift_backward(ctx, z, J, D, S, ext, R, grad_R) + (None, None)


class SSNAttachFixedPoint(torch.autograd.Function):
    """Identity on fixed points R that were solved beforehand (e.g. by a rejection-sampling loop run without
    autograd): puts them into the graph so that backward is the implicit gradient w.r.t. J, D, S (and ext).
    The implicit-function theorem needs only the fixed point itself, not the path that led to it."""

    @staticmethod
    def forward(ctx, z, J, D, S, ext, R, solver):
        ctx.save_for_backward(z, J, D, S, ext, R)
        ctx.solver = solver
        return R.clone()

    @staticmethod
    def backward(ctx, grad_R):
        z, J, D, S, ext, R = ctx.saved_tensors
        return _ift_backward(ctx, z, J, D, S, ext, R, grad_R) + (None, None)


def attach_fixed_point(z, J, D, S, ext, R, solver=None):
    """Differentiable view of already-solved fixed points `R` of the networks `z` (see SSNAttachFixedPoint)."""
    return SSNAttachFixedPoint.apply(z, J, D, S, ext, R.detach(), solver or make_solver())


def ssn_fixed_point(z, J, D, S, ext, solver=None, precise=False):
    """Differentiable (w.r.t. J, D, S) batched fixed-point solve; returns (R, status, iters)."""
    return SSNFixedPoint.apply(z, J, D, S, ext, solver or make_solver(), precise)


# ---- unrolled Euler dynamics ------------------------------------------------------------

def euler_forward(z, J, D, S, ext, seqlen, skip_steps, solver, rate_penalty_threshold=200.0,
                  store=True):
    _check_cuda(z, ext)
    z32, e32 = _f32c(z), _f32c(ext)
    nz, dim = z32.shape[0], z32.shape[1]
    nb = e32.shape[-2]
    dev = z.device
    time_avg = torch.empty((nz, nb, dim), dtype=torch.float32, device=dev)
    pen = torch.zeros(2, dtype=torch.float64, device=dev)
    # rows of the scratch arrays are padded to 16 bytes (ssn_traj_pitch): the backward pass loads them by TMA
    pitch = int(libssnode.ssn_traj_pitch(dim // 2))
    traj = torch.empty((nz, seqlen, nb, pitch), dtype=torch.float32, device=dev) if store else None
    gain = torch.empty((nz, seqlen, nb, pitch), dtype=torch.float32, device=dev) if store else None
    with torch.cuda.device(dev):
        clib.check_call(libssnode.ssn_euler_forward(
            solver, nz, nb, dim // 2, z32.data_ptr(), _jds_struct(J, D, S), e32.data_ptr(), int(e32.dim() == 3),
            int(seqlen), int(skip_steps), float(rate_penalty_threshold), time_avg.data_ptr(), pen.data_ptr(),
            None if traj is None else traj.data_ptr(), None if gain is None else gain.data_ptr(), _stream()),
            'ssn_euler_forward')
    return time_avg, pen, traj, gain


class EulerSSN(torch.autograd.Function):
    """(time_avg, dynamics_penalty, rate_penalty) of the unrolled Euler SSN; BPTT backward."""

    @staticmethod
    def forward(ctx, z, J, D, S, ext, seqlen, skip_steps, solver, threshold, grad_enabled):
        # needs_input_grad only mirrors requires_grad; under torch.no_grad() (critic updates) nothing will ever
        # call backward, and storing the trajectory (2 x nz*seqlen*nb*2N floats) would be pure waste.  Grad mode
        # is always off inside Function.forward, so the caller passes it in.
        need_grad = grad_enabled and any(ctx.needs_input_grad[1:5])
        time_avg, pen, traj, gain = euler_forward(z, J, D, S, ext, seqlen, skip_steps, solver, threshold,
                                                  store=need_grad)
        nz, nb, dim = time_avg.shape
        T = seqlen - skip_steps
        n_dyn = max(nz * (T - 1) * nb * dim, 1)
        n_rate = nz * T * nb * dim
        dyn = (pen[0] / n_dyn).to(torch.float32)
        rate = (pen[1] / n_rate).to(torch.float32)
        ctx.save_for_backward(z, J, D, S, traj, gain)
        ctx.meta = (seqlen, skip_steps, solver, threshold, n_dyn, n_rate, ext.dim(), ext.dtype)
        return time_avg, dyn, rate

    @staticmethod
    def backward(ctx, g_avg, g_dyn, g_rate):
        z, J, D, S, traj, gain = ctx.saved_tensors
        seqlen, skip_steps, solver, threshold, n_dyn, n_rate, ext_dim, ext_dtype = ctx.meta
        need_ext = ctx.needs_input_grad[4]
        nz, _, nb, _pitch = traj.shape
        dim = z.shape[1]
        dev = traj.device
        g32 = _f32c(g_avg) if g_avg is not None else torch.zeros((nz, nb, dim), dtype=torch.float32, device=dev)
        # the upstream scalar gradients stay on the device (no float(): that would drain the stream)
        w_dev = torch.zeros(2, dtype=torch.float32, device=dev)
        if g_dyn is not None:
            w_dev[0] = g_dyn
        if g_rate is not None:
            w_dev[1] = g_rate
        w_dyn, w_rate = 1.0 / n_dyn, 1.0 / n_rate
        adj = torch.empty_like(traj)
        grad = torch.empty(12, dtype=torch.float64, device=dev)
        gext = torch.zeros((nz, nb, dim), dtype=torch.float32, device=dev) if need_ext else None
        with torch.cuda.device(dev):
            clib.check_call(libssnode.ssn_euler_backward(
                solver, nz, nb, dim // 2, _f32c(z).data_ptr(), _jds_struct(J, D, S), int(seqlen), int(skip_steps),
                float(threshold), g32.data_ptr(), w_dyn, w_rate, w_dev.data_ptr(), traj.data_ptr(), gain.data_ptr(),
                adj.data_ptr(),
                grad.data_ptr(), None if gext is None else gext.data_ptr(), _stream()), 'ssn_euler_backward')
        dJ, dD, dS = grad[0:4].reshape(2, 2), grad[4:8].reshape(2, 2), grad[8:12].reshape(2, 2)
        g_ext = None
        if need_ext:
            g_ext = (gext if ext_dim == 3 else gext.sum(dim=0)).to(ext_dtype)
        return (None, dJ.to(J.dtype).to(J.device), dD.to(D.dtype).to(D.device), dS.to(S.dtype).to(S.device),
                g_ext, None, None, None, None, None)


def euler_ssn(z, J, D, S, ext, seqlen=1200, skip_steps=1000, dt=0.1, tau_E=10.0, tau_I=1.0,
              io_type='asym_tanh', k=0.01, n=2.2, rate_soft_bound=200., rate_hard_bound=1000.,
              rate_penalty_threshold=200.0):
    """
    Unrolled Euler SSN as tc_gan.networks.ssn.EulerSSNModel: r_0 = 0,
    r_{t+1} = (1 - dt/tau) r_t + dt/tau f(W r_t + I); returns
    (time_avg [nz, nb, 2N], dynamics_penalty, rate_penalty), differentiable w.r.t. J, D, S.
    Defaults are those of tc_gan/networks/wgan.py:39-63.
    """
    solver = clib.make_solver(io_type=io_type, k=k, n=n, tau=(tau_E, tau_I), dt=dt,
                              rate_soft_bound=rate_soft_bound, rate_hard_bound=rate_hard_bound,
                              rate_stop_at=rate_hard_bound)
    return EulerSSN.apply(z, J, D, S, ext, int(seqlen), int(skip_steps), solver, float(rate_penalty_threshold),
                          torch.is_grad_enabled())


class ProbeRates(torch.autograd.Function):
    """tuning_curve[i, b] = rates[model_ids[i], b, probes[i]] (tc_gan/networks/cwgan.py:96-99); backward is the
    scatter-add into dL/d rates that the BPTT / implicit-gradient kernels consume."""

    @staticmethod
    def forward(ctx, rates, model_ids, probes):
        _check_cuda(rates, model_ids, probes)
        r32 = _f32c(rates)
        nz, nb, dim = r32.shape
        ids = model_ids.to(torch.int32).contiguous()
        prb = probes.to(torch.int32).contiguous()
        batch = ids.numel()
        assert prb.numel() == batch
        out = torch.empty((batch, nb), dtype=torch.float32, device=rates.device)
        with torch.cuda.device(rates.device):
            clib.check_call(libssnode.ssn_probe_gather(r32.data_ptr(), ids.data_ptr(), prb.data_ptr(), batch, nz, nb,
                                                       dim // 2, out.data_ptr(), _stream()), 'ssn_probe_gather')
        ctx.save_for_backward(ids, prb)
        ctx.shape = (nz, nb, dim)
        ctx.dtype = rates.dtype
        return out

    @staticmethod
    def backward(ctx, grad_out):
        ids, prb = ctx.saved_tensors
        nz, nb, dim = ctx.shape
        g = _f32c(grad_out)
        grad_rates = torch.empty((nz, nb, dim), dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            clib.check_call(libssnode.ssn_probe_scatter(g.data_ptr(), ids.data_ptr(), prb.data_ptr(), ids.numel(), nz,
                                                        nb, dim // 2, grad_rates.data_ptr(), _stream()),
                            'ssn_probe_scatter')
        return grad_rates.to(ctx.dtype), None, None


def probe_rates(rates, model_ids, probes):
    """Differentiable gather of one probed neuron per batch element: [batch, nb] from rates [nz, nb, 2N]."""
    return ProbeRates.apply(rates, model_ids, probes)


def hetero_input(ext, zs_in, V):
    """
    Heterogeneous stimulus of tc_gan.networks.ssn.HeteroInputWrapper (networks/ssn.py:645-727):
    stimulus[z, b, i] = (1 + V_pop(i) * zs_in[z, i]) * ext[b, i].  `V` is a 2-vector (E, I;
    ssn_type 'heteroin') or a scalar ('deg-heteroin'); `zs_in` [nz, 2N] is +-1 (bernoulli) or
    uniform in [-1, 1].  Differentiable w.r.t. V (the kernels return dL/d ext).
    """
    n_sites = ext.shape[-1] // 2
    V = torch.as_tensor(V, device=ext.device) if not torch.is_tensor(V) else V
    vpop = V.expand(2) if V.dim() == 0 else V
    vs = torch.cat([vpop[0].expand(n_sites), vpop[1].expand(n_sites)]).to(ext.dtype)
    return (1 + vs[None, None, :] * zs_in.to(ext.dtype)[:, None, :]) * ext[None]


def smoke(oracle):
    """Tiny forward + backward of both generator paths on cuda:0, checked against the oracle."""
    import numpy as np
    dev = torch.device('cuda:0')
    n_sites, nz = 12, 2
    jds = oracle.new_JDS()
    exts = oracle.stimulus_input([0.125, 0.5, 1.0], n_sites)
    rs = np.random.RandomState(5)
    z = rs.rand(nz, 2 * n_sites, 2 * n_sites)
    gR = rs.randn(nz, len(exts), 2 * n_sites)
    W = oracle.generate_weight(n_sites, jds['J'], jds['D'], jds['S'], z)
    Ro, st, _ = oracle.fixed_point_batch(W, exts)
    t = lambda a, dt=torch.float32: torch.tensor(np.asarray(a), dtype=dt, device=dev)
    J, D, S = (t(jds[k], torch.float64).requires_grad_() for k in 'JDS')
    R, status, iters = ssn_fixed_point(t(z), J, D, S, t(exts))
    assert (status == 0).all()
    np.testing.assert_allclose(R.detach().cpu().numpy(), Ro, rtol=1e-4, atol=3e-4)
    (R * t(gR)).sum().backward()
    dJ, dD, dS, _ = oracle.ift_param_gradient(R.detach().double().cpu().numpy(), W, z, exts,
                                               jds['J'], jds['D'], jds['S'], gR)
    for got, want in ((J.grad, dJ), (D.grad, dD), (S.grad, dS)):
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-4, atol=1e-4 * np.abs(want).max())

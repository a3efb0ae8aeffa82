"""
CPU oracle for the SSN hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.  The product
package ``tc_gan_b200`` never does: it fails loudly without its CUDA library.

Everything here is float64 numpy (or float64 torch for the BPTT oracle) and
restates the reference's algorithm; citations are into /root/reference/tc_gan.

Pinning (see tests/test_oracle.py):
  * fixed points  -- against ``oracle/_ref/libssnode.so`` (the unmodified
    reference C file) and the MATLAB golden vectors ``assets/*.mat`` that the
    reference's own tests use (tests/test_dynamics.py:43-126), committed as
    ``tests/golden/*.npz`` by ``oracle/make_golden.py``.
  * W(z), stimuli -- against arrays produced by importing the reference's
    ``weight_gen.py`` / ``stimuli.py`` (same script).
  * gradients     -- PARITY UNPINNED BY THE REFERENCE: its tests only print them
    (tests/test_dynamics.py:140-276).  They are pinned here by central finite
    differences through the reference C solver (IFT) and by torch float64
    autograd (BPTT).
"""
import ctypes
import os
import subprocess
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
IO_TYPES = {'asym_power': 0, 'asym_linear': 1, 'asym_tanh': 2}

# ssnode.py:27-41
DEFAULT_J = np.array([[.0957, .0638], [.1197, .0479]])
DEFAULT_D = np.array([[.7660, .5106], [.9575, .3830]])
DEFAULT_S = np.array([[.6667, .2], [1.333, .2]]) / 8
DEFAULT_BANDWIDTHS = [0, 0.0625, 0.125, 0.1875, 0.25, 0.5, 0.75, 1]


def new_JDS():
    """More stable parameters, networks/fixed_time_sampler.py:12-23."""
    D_new = DEFAULT_D / 2
    return dict(J=DEFAULT_J + DEFAULT_D / 2 - D_new / 2, D=D_new, S=DEFAULT_S.copy())


_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)
_lock = threading.Lock()
_libs = {}


def build(ref=True):
    """Compile the C restatement (and oracle/_ref when the reference is here)."""
    subprocess.check_call(['make', '-s', '-C', HERE, 'all' if ref else
                           os.path.join(HERE, '_build', 'libssn_oracle.so')])


def lib():
    """Our C restatement (oracle/ssn_oracle.c)."""
    with _lock:
        if 'port' not in _libs:
            path = os.path.join(HERE, '_build', 'libssn_oracle.so')
            if not os.path.exists(path):
                build(ref=os.path.isdir('/root/reference'))
            L = ctypes.CDLL(path)
            L.oracle_rate_to_volt.argtypes = [ctypes.c_double] * 3
            L.oracle_rate_to_volt.restype = ctypes.c_double
            for f in (L.oracle_io, L.oracle_io_gain):
                f.argtypes = [ctypes.c_int] + [ctypes.c_double] * 6
                f.restype = ctypes.c_double
            L.oracle_fixed_point.argtypes = [
                ctypes.c_int, ctypes.c_int, _dp, _dp, ctypes.c_double, ctypes.c_double,
                _dp, _dp, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, _ip]
            L.oracle_fixed_point.restype = ctypes.c_int
            L.oracle_fixed_point_batch.argtypes = [
                ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _dp, _dp,
                ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                ctypes.c_double, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                ctypes.c_double, ctypes.c_int, ctypes.c_int, _dp, _ip, _ip]
            L.oracle_fixed_point_batch.restype = None
            _libs['port'] = L
        return _libs['port']


def ref_lib():
    """The unmodified reference solver (oracle/_ref/libssnode.so) or None."""
    with _lock:
        if 'ref' not in _libs:
            path = os.path.join(HERE, '_ref', 'libssnode.so')
            if not os.path.exists(path):
                _libs['ref'] = None
            else:
                L = ctypes.CDLL(path)
                # clib.py:16-33
                for name in ('power', 'linear', 'tanh'):
                    f = getattr(L, 'solve_dynamics_asym_%s_euler' % name)
                    f.argtypes = [ctypes.c_int, _dp, _dp, ctypes.c_double, ctypes.c_double,
                                  _dp, _dp, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                  ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double]
                    f.restype = ctypes.c_int
                for f in (L.io_pow, L.io_alin, L.io_atanh):
                    f.argtypes = [ctypes.c_double] * 6
                    f.restype = ctypes.c_double
                L.rate_to_volt.argtypes = [ctypes.c_double] * 3
                L.rate_to_volt.restype = ctypes.c_double
                _libs['ref'] = L
        return _libs['ref']


# --------------------------------------------------------------------------
# W(z; J, D, S), its parameter derivatives, and the stimulus
# --------------------------------------------------------------------------

def _blocks(n_sites, S):
    """g[a, b, i, j] = exp(-(x_i - x_j)^2 / (2 S_ab^2)), d2[i, j] = (x_i - x_j)^2."""
    x = np.linspace(-0.5, 0.5, n_sites)
    d2 = (x[:, None] - x[None, :]) ** 2
    S = np.asarray(S, dtype=float)
    g = np.exp(-d2[None, None] / (2 * S[:, :, None, None] ** 2))
    return g, d2


def generate_weight(n_sites, J, D, S, z):
    """
    W[aN+i, bN+j] = s_b g_ab(i,j) (J_ab + D_ab z[aN+i, bN+j]), s_E=+1, s_I=-1.
    weight_gen.py:6-26 == gradient_expressions/make_w_batch.py:8-34.
    `z` may be [2N, 2N] or [nz, 2N, 2N].
    """
    z = np.asarray(z, dtype=float)
    single = z.ndim == 2
    z = z.reshape((-1, 2 * n_sites, 2 * n_sites))
    J, D = np.asarray(J, float), np.asarray(D, float)
    g, _ = _blocks(n_sites, S)
    W = np.empty_like(z)
    N = n_sites
    for a in range(2):
        for b in range(2):
            sgn = 1.0 if b == 0 else -1.0
            zz = z[:, a * N:(a + 1) * N, b * N:(b + 1) * N]
            W[:, a * N:(a + 1) * N, b * N:(b + 1) * N] = sgn * g[a, b] * (J[a, b] + D[a, b] * zz)
    return W[0] if single else W


def weight_param_contraction(n_sites, J, D, S, z, G):
    """
    <G, dW/dtheta_ab> for theta in (J, D, S): three [2, 2] arrays, where
    G[nz, 2N, 2N] is dL/dW.  Restates make_w_batch.py:36-121 contracted on the
    fly (the reference materialises [nz, 2N, 2N, 2, 2] tensors).
      dW/dJ_ab = s_b g_ab;  dW/dD_ab = s_b g_ab z;  dW/dS_ab = s_b g_ab d2/S_ab^3 (J_ab + D_ab z)
    """
    z = np.asarray(z, float).reshape((-1, 2 * n_sites, 2 * n_sites))
    G = np.asarray(G, float).reshape(z.shape)
    J, D, S = (np.asarray(a, float) for a in (J, D, S))
    g, d2 = _blocks(n_sites, S)
    N = n_sites
    dJ, dD, dS = np.zeros((2, 2)), np.zeros((2, 2)), np.zeros((2, 2))
    for a in range(2):
        for b in range(2):
            sgn = 1.0 if b == 0 else -1.0
            sl = (slice(None), slice(a * N, (a + 1) * N), slice(b * N, (b + 1) * N))
            zz, GG = z[sl], G[sl]
            dJ[a, b] = np.sum(GG * sgn * g[a, b])
            dD[a, b] = np.sum(GG * sgn * g[a, b] * zz)
            dS[a, b] = np.sum(GG * sgn * g[a, b] * d2 / S[a, b] ** 3 * (J[a, b] + D[a, b] * zz))
    return dJ, dD, dS


def stimulus_input(bandwidths, n_sites, smoothness=0.25 / 8, contrasts=(20.,), offsets=(0.,)):
    """
    I[c, o, b][i] = c * sig((x_i - o + b/2)/l) * sig((b/2 - (x_i - o))/l), tiled for E and I.
    Order: contrast-major, then offset, then bandwidth.  stimuli.py:3-10.
    """
    x = np.linspace(-0.5, 0.5, n_sites)

    def sig(u):
        return 1.0 / (1.0 + np.exp(-u / smoothness))

    rows = []
    for c in contrasts:
        for o in offsets:
            for b in bandwidths:
                band = sig((x - o) + b / 2) * sig(b / 2 - (x - o))
                rows.append(c * np.concatenate([band, band]))
    return np.array(rows)


# --------------------------------------------------------------------------
# transfer function and gain
# --------------------------------------------------------------------------

def rate_to_volt(rate, k, n):
    return (np.asarray(rate, float) / k) ** (1.0 / n)      # ssnode.py:125-126


def io_fun(v, io_type='asym_tanh', k=0.01, n=2.2, r_soft=200., r_hard=1000.):
    """ssnode.py:129-149 / ext/ssnode.c:25-53, vectorised."""
    v = np.asarray(v, float)
    v0 = rate_to_volt(r_soft, k, n)
    if io_type == 'asym_power':
        return k * np.clip(v, 0, None) ** n
    low = k * np.clip(v, 0, v0) ** n
    if io_type == 'asym_linear':
        return np.where(v <= v0, low, r_soft + k * v0 ** (n - 1) * n * (v - v0))
    span = r_hard - r_soft
    return np.where(v <= v0, low, r_soft + span * np.tanh(n * r_soft / span * (v - v0) / v0))


def io_gain(v, io_type='asym_tanh', k=0.01, n=2.2, r_soft=200., r_hard=1000.):
    """f'(v) as the reference's implicit gradient defines it, SS_grad.py:78-99."""
    v = np.asarray(v, float)
    v0 = rate_to_volt(r_soft, k, n)
    vc = np.clip(v, 0, None)
    if io_type == 'asym_power':
        return n * k * vc ** (n - 1.)
    if io_type == 'asym_linear':
        return n * k * np.clip(v, 0, v0) ** (n - 1.)
    arg = (n * r_soft / v0) * (vc - v0) / (r_hard - r_soft)
    return np.where(vc <= v0, n * k * vc ** (n - 1.), (n * r_soft / v0) * np.cosh(arg) ** -2.)


# --------------------------------------------------------------------------
# fixed points
# --------------------------------------------------------------------------

def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def fixed_point(W, ext, k=0.01, n=2.2, r0=None, tau=(0.01589, 0.002), max_iter=10000,
                atol=1e-5, dt=.0008, rate_soft_bound=200., rate_hard_bound=1000.,
                rate_stop_at=np.inf, io_type='asym_tanh', impl='port'):
    """
    One solve, same meaning as ssnode.fixed_point (ssnode.py:159-273).
    impl='port' runs our C restatement, impl='ref' the reference's own C file.
    Returns (x, error_code, iterations or None).
    """
    W, ext = _c(W), _c(ext)
    dim = W.shape[0]
    r = np.zeros(dim) if r0 is None else np.array(r0, dtype=np.float64)
    scratch = np.empty(dim)
    if io_type in ('asym_power', 'asym_linear'):          # ssnode.py:241-242
        rate_hard_bound = rate_stop_at
    args = [W.ctypes.data_as(_dp), ext.ctypes.data_as(_dp), float(k), float(n),
            r.ctypes.data_as(_dp), scratch.ctypes.data_as(_dp),
            float(tau[0]), float(tau[1]), float(dt), int(max_iter), float(atol),
            float(rate_soft_bound), float(rate_hard_bound)]
    if impl == 'ref':
        L = ref_lib()
        if L is None:
            raise RuntimeError('oracle/_ref/libssnode.so is not built')
        f = getattr(L, 'solve_dynamics_%s_euler' % io_type)
        code = f(dim // 2, *args)
        return r, code, None
    it = ctypes.c_int(0)
    code = lib().oracle_fixed_point(IO_TYPES[io_type], dim // 2, *args, ctypes.byref(it))
    return r, code, it.value


def fixed_point_batch(W, exts, threads=1, stop_at_first_failure=False, k=0.01, n=2.2,
                      tau=(0.01589, 0.002), max_iter=10000, atol=1e-5, dt=.0008,
                      rate_soft_bound=200., rate_hard_bound=1000., rate_stop_at=np.inf,
                      io_type='asym_tanh'):
    """All (network, stimulus) pairs from r=0 with our C restatement (OpenMP)."""
    W, exts = _c(W), _c(exts)
    nz, dim = W.shape[0], W.shape[1]
    nb = exts.shape[0]
    if io_type in ('asym_power', 'asym_linear'):
        rate_hard_bound = rate_stop_at
    R = np.zeros((nz, nb, dim))
    status = np.zeros((nz, nb), dtype=np.int32)
    iters = np.zeros((nz, nb), dtype=np.int32)
    lib().oracle_fixed_point_batch(
        IO_TYPES[io_type], nz, nb, dim // 2, W.ctypes.data_as(_dp), exts.ctypes.data_as(_dp),
        float(k), float(n), float(tau[0]), float(tau[1]), float(dt), int(max_iter),
        float(atol), float(rate_soft_bound), float(rate_hard_bound),
        int(stop_at_first_failure), int(threads),
        R.ctypes.data_as(_dp), status.ctypes.data_as(_ip), iters.ctypes.data_as(_ip))
    return R, status, iters


def ref_fixed_point_batch(W, exts, threads=1, **kw):
    """
    The reference arm: the UNMODIFIED reference C solver driven the way
    ssnode.find_fixed_points_parallel drives it (ssnode.py:423-510): a pool of
    Python threads, one network per job, stimuli visited last to first, a
    network abandoned at its first failure.  ctypes drops the GIL in the call.
    Returns R[nz, nb, 2N], status[nz, nb] (-1 = skipped).
    """
    W, exts = _c(W), _c(exts)
    nz, dim = W.shape[0], W.shape[1]
    nb = exts.shape[0]
    R = np.zeros((nz, nb, dim))
    status = -np.ones((nz, nb), dtype=np.int32)

    def job(z):
        for b in range(nb - 1, -1, -1):
            x, code, _ = fixed_point(W[z], exts[b], impl='ref', **kw)
            if code == 0 and not np.isfinite(x).all():     # ssnode.py:257-262
                code = 1
            R[z, b], status[z, b] = x, code
            if code != 0:
                break

    if threads <= 1:
        for z in range(nz):
            job(z)
    else:
        with ThreadPoolExecutor(threads) as pool:
            list(pool.map(job, range(nz)))
    return R, status


# --------------------------------------------------------------------------
# implicit-function-theorem gradient (fixed-point GAN generator gradient)
# --------------------------------------------------------------------------

def ift_rate_jacobians(R, W, z, exts, J, D, S, io_type='asym_tanh', k=0.01, n=2.2,
                       r_soft=200., r_hard=1000.):
    """
    dr/dtheta for theta in (J, D, S): three arrays [nz, nb, 2N, 2, 2], the direct
    form of SS_grad.WRgrad_batch (SS_grad.py:17-76) with the dW/dtheta tensors of
    make_w_batch.py:36-121:   dr/dtheta = (I - Phi W)^-1 Phi (dW/dtheta r).
    Small sizes only (dense solves, O(nz nb (2N)^3)).
    """
    R, W, z, exts = (np.asarray(a, float) for a in (R, W, z, exts))
    nz, nb, dim = R.shape
    N = dim // 2
    J, D, S = (np.asarray(a, float) for a in (J, D, S))
    g, d2 = _blocks(N, S)
    out = [np.zeros((nz, nb, dim, 2, 2)) for _ in range(3)]
    for iz in range(nz):
        for ib in range(nb):
            r = R[iz, ib]
            phi = io_gain(W[iz] @ r + exts[ib], io_type, k, n, r_soft, r_hard)
            A = np.eye(dim) - phi[:, None] * W[iz]
            for a in range(2):
                for b in range(2):
                    sgn = 1.0 if b == 0 else -1.0
                    rows, cols = slice(a * N, (a + 1) * N), slice(b * N, (b + 1) * N)
                    zz = z[iz, rows, cols]
                    dWs = (sgn * g[a, b],
                           sgn * g[a, b] * zz,
                           sgn * g[a, b] * d2 / S[a, b] ** 3 * (J[a, b] + D[a, b] * zz))
                    for t, dW in enumerate(dWs):
                        rhs = np.zeros(dim)
                        rhs[rows] = dW @ r[cols]
                        out[t][iz, ib, :, a, b] = np.linalg.solve(A, phi * rhs)
    return out


def ift_param_gradient(R, W, z, exts, J, D, S, grad_R, io_type='asym_tanh', k=0.01, n=2.2,
                       r_soft=200., r_hard=1000.):
    """
    dL/d(J, D, S) = sum_{z,b,i} dL/dr * dr/dtheta (run/gan.py:902-911), computed
    by the adjoint:  (I - W^T Phi) mu = g,  dL/dW = (Phi mu) r^T,  then the
    on-the-fly contraction with dW/dtheta.  Returns (dJ, dD, dS, mu).
    """
    R, W, grad_R, exts = (np.asarray(a, float) for a in (R, W, grad_R, exts))
    nz, nb, dim = R.shape
    G = np.zeros((nz, dim, dim))
    mu_all = np.zeros_like(R)
    for iz in range(nz):
        for ib in range(nb):
            r = R[iz, ib]
            phi = io_gain(W[iz] @ r + exts[ib], io_type, k, n, r_soft, r_hard)
            A_T = np.eye(dim) - W[iz].T * phi[None, :]
            mu = np.linalg.solve(A_T, grad_R[iz, ib])
            mu_all[iz, ib] = mu
            G[iz] += np.outer(phi * mu, r)
    dJ, dD, dS = weight_param_contraction(dim // 2, J, D, S, z, G)
    return dJ, dD, dS, mu_all


# --------------------------------------------------------------------------
# unrolled Euler dynamics + BPTT (torch float64 autograd as the backward oracle)
# --------------------------------------------------------------------------

def euler_unroll_torch(z, J, D, S, exts, seqlen, skip_steps, eps_E, eps_I,
                       io_type='asym_tanh', k=0.01, n=2.2, r_soft=200., r_hard=1000.,
                       rate_penalty_threshold=200.0, return_trajectory=False):
    """
    r_{t+1} = (1-eps) r_t + eps f(W r_t + I), r_0 = 0, trajectory = (r_1 .. r_seqlen)
    (Lasagne CustomRecurrentLayer returns the hidden state AFTER each step),
    networks/ssn.py:555-576;  outputs as networks/ssn.py:619-633:
      time_avg         = mean_{t >= skip} r_t                       [nz, nb, 2N]
      dynamics_penalty = mean (r_{t+1} - r_t)^2 over the kept steps  scalar
      rate_penalty     = mean relu(r_t - threshold) over kept steps  scalar
    All arguments are torch float64 tensors (J, D, S may require grad); exts is
    [nb, 2N] or [nz, nb, 2N].
    """
    import torch
    nz, dim = z.shape[0], z.shape[1]
    N = dim // 2
    x = torch.linspace(-0.5, 0.5, N, dtype=torch.float64)
    d2 = (x[:, None] - x[None, :]) ** 2
    blocks = []
    for a in range(2):
        row = []
        for b in range(2):
            sgn = 1.0 if b == 0 else -1.0
            g = torch.exp(-d2 / (2 * S[a, b] ** 2))
            row.append(sgn * g * (J[a, b] + D[a, b] * z[:, a * N:(a + 1) * N, b * N:(b + 1) * N]))
        blocks.append(torch.cat(row, dim=2))
    W = torch.cat(blocks, dim=1)                                  # [nz, 2N, 2N]
    eps = torch.cat([torch.full((N,), float(eps_E), dtype=torch.float64),
                     torch.full((N,), float(eps_I), dtype=torch.float64)])
    v0 = (r_soft / k) ** (1.0 / n)

    def f(v):
        if io_type == 'asym_power':
            return k * torch.clamp(v, min=0) ** n
        low = k * torch.clamp(v, 0, v0) ** n
        if io_type == 'asym_linear':
            return torch.where(v <= v0, low, r_soft + k * v0 ** (n - 1) * n * (v - v0))
        span = r_hard - r_soft
        return torch.where(v <= v0, low,
                           r_soft + span * torch.tanh(n * r_soft / span * (v - v0) / v0))

    I = exts if exts.dim() == 3 else exts[None]
    r = torch.zeros((nz, I.shape[1], dim), dtype=torch.float64)
    traj = []
    for _ in range(seqlen):
        r = (1 - eps) * r + eps * f(torch.einsum('zij,zbj->zbi', W, r) + I)
        traj.append(r)
    rs = torch.stack(traj[skip_steps:], dim=1)                    # [nz, T, nb, 2N]
    time_avg = rs.mean(dim=1)
    dyn = ((rs[:, 1:] - rs[:, :-1]) ** 2).mean() if rs.shape[1] > 1 else rs.sum() * 0
    rate = torch.clamp(rs - rate_penalty_threshold, min=0).mean()
    if return_trajectory:
        return time_avg, dyn, rate, torch.stack(traj, dim=1)
    return time_avg, dyn, rate

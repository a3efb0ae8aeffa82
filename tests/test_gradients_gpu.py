"""
GPU parity tests of the generator gradients through the fixed point (K2) and through the
unrolled Euler dynamics (K3/K4).  The reference's tests leave gradients unpinned (they only
print them, tc_gan/tests/test_dynamics.py:140-276); the oracle's restatements are pinned by
finite differences / torch float64 autograd in tests/test_oracle.py.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops(built_library):
    import torch
    from tc_gan_b200 import clib, torch_ops
    if clib.libssnode.ssn_device_count() < 1 or not torch.cuda.is_available():
        pytest.fail('GPU tests need a CUDA device: the library has no CPU fallback')
    return torch_ops


def tens(a, dtype=None, grad=False):
    import torch
    t = torch.tensor(np.asarray(a), dtype=dtype or torch.float32, device='cuda:0')
    return t.requires_grad_() if grad else t


@pytest.mark.parametrize('n_sites,nz,nb,io_type', [(12, 2, 3, 'asym_tanh'), (51, 3, 8, 'asym_tanh'),
                                                   (51, 2, 8, 'asym_power'), (40, 2, 11, 'asym_linear'),
                                                   (201, 2, 8, 'asym_tanh'),
                                                   # cluster widths 2 and 8 of the shared-memory core (4 is 2N = 402)
                                                   (125, 2, 8, 'asym_tanh'), (280, 2, 5, 'asym_tanh')])
def test_ift_gradient_matches_oracle(ops, oracle, n_sites, nz, nb, io_type):
    """Same R and dL/dR into the CUDA adjoint path and the float64 restatement of
    SS_grad.WRgrad_batch + make_w_batch + run/gan.py:902-911."""
    import torch
    jds = oracle.new_JDS()
    bw = oracle.DEFAULT_BANDWIDTHS if nb == 8 else np.linspace(0.05, 1, nb)
    exts = oracle.stimulus_input(bw, n_sites)
    rs = np.random.RandomState(n_sites + nb)
    z = rs.rand(nz, 2 * n_sites, 2 * n_sites).astype(np.float32).astype(np.float64)
    W = oracle.generate_weight(n_sites, jds['J'], jds['D'], jds['S'], z)
    R, st, _ = oracle.fixed_point_batch(W, exts, io_type=io_type, threads=8)
    assert (st == 0).all()
    R = R.astype(np.float32).astype(np.float64)
    gR = rs.randn(nz, nb, 2 * n_sites).astype(np.float32).astype(np.float64)
    dJ, dD, dS, mu = oracle.ift_param_gradient(R, W, z, exts, jds['J'], jds['D'], jds['S'], gR, io_type=io_type)
    solver = ops.make_solver(io_type=io_type)
    J, D, S = (tens(jds[k], torch.float64) for k in 'JDS')
    gJ, gD, gS, mu_gpu, status, iters = ops.ift_gradient(tens(z), J, D, S, tens(exts), tens(R), tens(gR),
                                                         solver=solver, return_mu=True)
    assert (status.cpu().numpy() == 0).all()
    np.testing.assert_allclose(mu_gpu.cpu().numpy(), mu, rtol=1e-3, atol=2e-4 * np.abs(mu).max())
    for got, want in ((gJ, dJ), (gD, dD), (gS, dS)):      # north_star: generator gradients within rtol 1e-4
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-4, atol=1e-4 * np.abs(want).max())


@pytest.mark.parametrize('n_sites,nz,nb', [(1, 3, 2), (3, 2, 9), (17, 4, 8), (64, 2, 16), (201, 3, 8)])
def test_ift_gmres_residual_and_damped_agreement(ops, oracle, monkeypatch, n_sites, nz, nb):
    """The restarted GMRES of K2 (default) against (a) the defining linear system, checked in float64 on the host:
    |g - (I - W^T Phi) mu|_2 <= ~rtol |g|_2 per solve, and (b) the damped adjoint iteration of round 1
    (SSN_IFT=damped), which it replaces: same gradients to 2e-4, in at least 4x fewer contractions at 2N >= 34.
    Sizes include 2N = 2 (Krylov space exhausted after two steps: the breakdown branch), a ragged stimulus count
    with a partly empty second panel, and BASELINE's 2N = 402."""
    import torch
    jds = oracle.new_JDS()
    dim = 2 * n_sites
    exts = oracle.stimulus_input(np.linspace(0.05, 1, nb), n_sites) if n_sites > 1 else np.full((nb, 2), 5.0) * (1 + np.arange(nb))[:, None]
    rs = np.random.RandomState(100 + n_sites)
    z = rs.rand(nz, dim, dim).astype(np.float32).astype(np.float64)
    W = oracle.generate_weight(n_sites, jds['J'], jds['D'], jds['S'], z)
    R, st, _ = oracle.fixed_point_batch(W, exts, threads=8)
    assert (st == 0).all()
    R = R.astype(np.float32).astype(np.float64)
    gR = rs.randn(nz, nb, dim).astype(np.float32).astype(np.float64)
    J, D, S = (tens(jds[k], torch.float64) for k in 'JDS')
    args = (tens(z), J, D, S, tens(exts), tens(R), tens(gR))
    gm = ops.ift_gradient(*args, rtol=1e-6, return_mu=True)
    mu, status, iters = (t.cpu().numpy() for t in gm[3:6])
    assert (status == 0).all()
    for iz in range(nz):
        for ib in range(nb):
            phi = oracle.io_gain(W[iz] @ R[iz, ib] + exts[ib])
            res = gR[iz, ib] - mu[iz, ib] + W[iz].T @ (phi * mu[iz, ib])
            # 1e-6 asked; float32 storage of mu adds ~6e-8 |A mu|
            assert np.linalg.norm(res) <= 4e-6 * np.linalg.norm(gR[iz, ib]) + 2e-7 * np.linalg.norm(mu[iz, ib])
    monkeypatch.setenv('SSN_IFT', 'damped')
    dm = ops.ift_gradient(*args, rtol=1e-6, return_mu=True)
    monkeypatch.delenv('SSN_IFT')
    assert (dm[4].cpu().numpy() == 0).all()
    for a, b in zip(gm[:3], dm[:3]):
        np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=2e-4, atol=2e-4 * float(b.abs().max()))
    if n_sites >= 17:
        assert iters.mean() * 4 < dm[5].cpu().numpy().mean()
    # the fallback of a stalled GMRES (forced here after the first cycle): damped steps from the GMRES iterate
    monkeypatch.setenv('SSN_IFT', 'stall')
    fb = ops.ift_gradient(*args, rtol=1e-6, return_mu=True)
    monkeypatch.delenv('SSN_IFT')
    assert (fb[4].cpu().numpy() == 0).all()
    for a, b in zip(fb[:3], gm[:3]):
        np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=2e-4, atol=2e-4 * float(b.abs().max()))


def test_fixed_point_autograd_function(ops, oracle):
    """loss = <R, G> through SSNFixedPoint.apply; J.grad etc. against the oracle, and the
    zero-gradient / ragged-panel edge cases."""
    import torch
    n_sites, nz = 30, 5
    jds = oracle.new_JDS()
    exts = oracle.stimulus_input(np.linspace(0, 1, 9), n_sites)
    rs = np.random.RandomState(2)
    z = rs.rand(nz, 2 * n_sites, 2 * n_sites).astype(np.float32).astype(np.float64)
    G = rs.randn(nz, len(exts), 2 * n_sites)
    G[1] = 0.0                                   # a network whose loss gradient vanishes
    J, D, S = (tens(jds[k], torch.float64, grad=True) for k in 'JDS')
    R, status, iters = ops.ssn_fixed_point(tens(z), J, D, S, tens(exts))
    assert (status == 0).all() and not status.requires_grad
    (R * tens(G)).sum().backward()
    W = oracle.generate_weight(n_sites, jds['J'], jds['D'], jds['S'], z)
    dJ, dD, dS, _ = oracle.ift_param_gradient(R.detach().double().cpu().numpy(), W, z, exts,
                                               jds['J'], jds['D'], jds['S'], G)
    for got, want in ((J.grad, dJ), (D.grad, dD), (S.grad, dS)):
        assert got.shape == (2, 2) and got.dtype == torch.float64
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-4, atol=1e-4 * np.abs(want).max())
    with pytest.raises(Exception):
        ops.fixed_points(tens(z).cpu(), J, D, S, tens(exts).cpu())


def euler_oracle(oracle, z, jds, exts, seqlen, skip, eps, io_type, thr, G, c_dyn, c_rate):
    import torch
    t64 = lambda a, g=False: torch.tensor(np.asarray(a), dtype=torch.float64, requires_grad=g)
    J, D, S = (t64(jds[k], True) for k in 'JDS')
    avg, dyn, rate = oracle.euler_unroll_torch(t64(z), J, D, S, t64(exts), seqlen, skip, eps[0], eps[1],
                                               io_type=io_type, rate_penalty_threshold=thr)
    loss = (avg * t64(G)).sum() + c_dyn * dyn + c_rate * rate
    loss.backward()
    return avg.detach().numpy(), float(dyn.detach()), float(rate.detach()), J.grad.numpy(), D.grad.numpy(), S.grad.numpy()


@pytest.mark.parametrize('n_sites,nz,nb,seqlen,skip,io_type', [
    (10, 2, 8, 60, 40, 'asym_tanh'), (25, 3, 5, 40, 25, 'asym_tanh'), (51, 2, 8, 30, 20, 'asym_power'),
    (30, 2, 11, 24, 0, 'asym_linear'), (201, 1, 8, 12, 6, 'asym_tanh'),
    # every cluster width of the shared-memory core: 1 CTA (2N = 2, 20), 2 (2N = 250), 4 (402 above), 8 (2N = 560);
    # more networks than resident clusters of 8 (17 > 15), a ragged second panel
    (1, 3, 2, 9, 3, 'asym_tanh'), (125, 3, 8, 16, 8, 'asym_tanh'), (280, 2, 8, 10, 4, 'asym_tanh'),
    (280, 17, 3, 6, 2, 'asym_power'), (64, 40, 9, 14, 5, 'asym_tanh')])
def test_euler_unroll_forward_backward(ops, oracle, n_sites, nz, nb, seqlen, skip, io_type):
    """K3/K4 against torch float64 autograd through the restated Euler unroll
    (tc_gan/networks/ssn.py:555-576, 619-633): outputs and gradients to rtol 1e-4 (BASELINE north_star).
    A low rate threshold makes the rate penalty and its gradient non-trivial."""
    import torch
    jds = oracle.new_JDS()
    bw = oracle.DEFAULT_BANDWIDTHS if nb == 8 else np.linspace(0.05, 1, nb)
    exts = oracle.stimulus_input(bw, n_sites)
    rs = np.random.RandomState(seqlen + n_sites)
    z = rs.rand(nz, 2 * n_sites, 2 * n_sites).astype(np.float32).astype(np.float64)
    G = rs.randn(nz, nb, 2 * n_sites)
    eps, thr, c_dyn, c_rate = (0.01, 0.1), 0.5, 3.0, 2.0
    avg_o, dyn_o, rate_o, dJ, dD, dS = euler_oracle(oracle, z, jds, exts, seqlen, skip, eps, io_type, thr,
                                                    G, c_dyn, c_rate)
    J, D, S = (tens(jds[k], torch.float64, grad=True) for k in 'JDS')
    avg, dyn, rate = ops.euler_ssn(tens(z), J, D, S, tens(exts), seqlen=seqlen, skip_steps=skip, dt=0.1,
                                   tau_E=10.0, tau_I=1.0, io_type=io_type, rate_penalty_threshold=thr)
    np.testing.assert_allclose(avg.detach().cpu().numpy(), avg_o, rtol=1e-4, atol=1e-5)
    if seqlen - skip > 1:
        np.testing.assert_allclose(float(dyn), dyn_o, rtol=1e-4)
    np.testing.assert_allclose(float(rate), rate_o, rtol=1e-4)
    loss = (avg * tens(G)).sum() + c_dyn * dyn + c_rate * rate
    loss.backward()
    for got, want in ((J.grad, dJ), (D.grad, dD), (S.grad, dS)):
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-4, atol=1e-4 * np.abs(want).max())


@pytest.mark.parametrize('nz,seqlen,skip,c_dyn', [(2, 1200, 1000, 0.0), (3, 400, 200, 3.0)])
def test_bptt_config3_size_parity(ops, oracle, nz, seqlen, skip, c_dyn):
    """BASELINE configs[2] at full size per network: 2N=402, 8 stimuli, seqlen 1200 / skip 1000 (and a 400-step
    case whose kept window still moves, so the dynamics penalty and its gradient are not rounding noise):
    time_avg, both penalties and dL/d(J, D, S) against torch float64 autograd at rtol 1e-4.  1200 FP32 adjoint
    steps and the K = seqlen * nb = 9600 accumulation of the parameter-gradient contraction are where drift
    would show."""
    import torch
    n_sites, nb, io_type = 201, 8, 'asym_tanh'
    jds = oracle.new_JDS()
    exts = oracle.stimulus_input(oracle.DEFAULT_BANDWIDTHS, n_sites)
    rs = np.random.RandomState(seqlen)
    z = rs.rand(nz, 2 * n_sites, 2 * n_sites).astype(np.float32).astype(np.float64)
    G = rs.randn(nz, nb, 2 * n_sites)
    eps, thr, c_rate = (0.01, 0.1), 5.0, 2.0
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    avg_o, dyn_o, rate_o, dJ, dD, dS = euler_oracle(oracle, z, jds, exts, seqlen, skip, eps, io_type, thr,
                                                    G, c_dyn, c_rate)
    J, D, S = (tens(jds[k], torch.float64, grad=True) for k in 'JDS')
    avg, dyn, rate = ops.euler_ssn(tens(z), J, D, S, tens(exts), seqlen=seqlen, skip_steps=skip, dt=0.1,
                                   tau_E=10.0, tau_I=1.0, io_type=io_type, rate_penalty_threshold=thr)
    np.testing.assert_allclose(avg.detach().cpu().numpy(), avg_o, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(float(rate), rate_o, rtol=1e-4)
    if c_dyn:
        np.testing.assert_allclose(float(dyn), dyn_o, rtol=1e-4)
    else:
        # after 1000 steps the kept window moves by ~1e-5 per step: the penalty (~1e-10) is below the FP32
        # resolution of the contraction, in the reference (Theano floatX) as here; bound it absolutely
        assert abs(float(dyn) - dyn_o) <= 1e-4 * dyn_o + 1e-9 * float(np.abs(avg_o).max()) ** 2
    loss = (avg * tens(G)).sum() + c_dyn * dyn + c_rate * rate
    loss.backward()
    for got, want in ((J.grad, dJ), (D.grad, dD), (S.grad, dS)):
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-4, atol=1e-4 * np.abs(want).max())


def test_euler_long_unroll_reaches_fixed_point(ops, oracle):
    """tc_gan/networks/tests/test_euler_ssn.py:29-72 with our two CUDA paths: the last
    state of a long Euler unroll equals the fixed-point solver's answer (rtol=atol=5e-4)."""
    import torch
    n_sites, nz = 10, 2
    jds = oracle.new_JDS()
    exts = oracle.stimulus_input(oracle.DEFAULT_BANDWIDTHS, n_sites)
    z = np.random.RandomState(n_sites * nz).rand(nz, 2 * n_sites, 2 * n_sites)
    J, D, S = (tens(jds[k], torch.float64) for k in 'JDS')
    avg, _, _ = ops.euler_ssn(tens(z), J, D, S, tens(exts), seqlen=4000, skip_steps=3999)
    solver = ops.make_solver(atol=1e-10, max_iter=100000)       # as the reference test; float64 kernel
    R, status, _ = ops.fixed_points(tens(z), J, D, S, tens(exts), solver=solver, precise=True)
    assert (status == 0).all()
    np.testing.assert_allclose(avg.cpu().numpy(), R.cpu().numpy(), rtol=5e-4, atol=5e-4)


def test_bptt_of_a_long_unroll_equals_the_implicit_gradient(ops, oracle):
    """Cross-check of the two gradient paths at BASELINE size (2N=402, 8 stimuli): the gradient of a loss on the final
    state of a LONG Euler unroll (K3 + K4 + K4b: 2000 steps with the fixed-point solver's dt / tau, by which the
    dynamics has converged) equals the implicit-function-theorem gradient at the fixed point (K1 + K2: GMRES on the
    adjoint system).  Two different algorithms, five kernels, no oracle in between: dL/d(J, D, S) to 2e-4."""
    import torch
    n_sites, nz = 201, 3
    jds = oracle.new_JDS()
    exts = oracle.stimulus_input(oracle.DEFAULT_BANDWIDTHS, n_sites)
    rs = np.random.RandomState(77)
    z = tens(rs.rand(nz, 2 * n_sites, 2 * n_sites))
    G = tens(rs.randn(nz, len(exts), 2 * n_sites))
    grads = []
    for path in ('bptt', 'ift'):
        J, D, S = (tens(jds[k], torch.float64, grad=True) for k in 'JDS')
        if path == 'bptt':
            R, _, _ = ops.euler_ssn(z, J, D, S, tens(exts), seqlen=2000, skip_steps=1999, dt=0.0008,
                                    tau_E=0.01589, tau_I=0.002)
        else:
            R, status, _ = ops.ssn_fixed_point(z, J, D, S, tens(exts), solver=ops.make_solver(atol=1e-9))
            assert (status == 0).all()
        (R * G).sum().backward()
        grads.append([p.grad.cpu().numpy() for p in (J, D, S)] + [R.detach().cpu().numpy()])
    np.testing.assert_allclose(grads[0][3], grads[1][3], rtol=1e-5, atol=1e-5)        # the states agree
    for a, b in zip(grads[0][:3], grads[1][:3]):
        np.testing.assert_allclose(a, b, rtol=2e-4, atol=2e-4 * np.abs(b).max())


def test_heterogeneous_input_gradients(ops, oracle):
    """heteroin / deg-heteroin generators (networks/ssn.py:645-727): stimulus scaled per
    neuron by 1 + V z_in; dL/dV through both gradient paths against float64 (torch autograd for BPTT,
    the adjoint formula dL/d ext = Phi mu for the fixed point)."""
    import torch
    n_sites, nz, nb, seqlen, skip = 20, 3, 8, 40, 25
    jds = oracle.new_JDS()
    exts = oracle.stimulus_input(oracle.DEFAULT_BANDWIDTHS, n_sites)
    rs = np.random.RandomState(11)
    z = rs.rand(nz, 2 * n_sites, 2 * n_sites).astype(np.float32).astype(np.float64)
    zs_in = rs.choice(2, (nz, 2 * n_sites)) * 2.0 - 1.0
    G = rs.randn(nz, nb, 2 * n_sites)
    V0 = np.array([0.3, 0.15])
    # ---- BPTT path, ssn_type 'heteroin' (two-component V) ----
    t64 = lambda a, g=False: torch.tensor(np.asarray(a), dtype=torch.float64, requires_grad=g)
    Vo = t64(V0, True)
    J, D, S = (t64(jds[k], True) for k in 'JDS')
    vs = torch.cat([Vo[0].expand(n_sites), Vo[1].expand(n_sites)])
    ext_o = (1 + vs[None, None, :] * t64(zs_in)[:, None, :]) * t64(exts)[None]
    avg_o, dyn_o, rate_o = oracle.euler_unroll_torch(t64(z), J, D, S, ext_o, seqlen, skip, 0.01, 0.1,
                                                     rate_penalty_threshold=0.5)
    ((avg_o * t64(G)).sum() + 3.0 * dyn_o + 2.0 * rate_o).backward()
    Vg = tens(V0, torch.float64, grad=True)
    Jg, Dg, Sg = (tens(jds[k], torch.float64, grad=True) for k in 'JDS')
    ext_g = ops.hetero_input(tens(exts), tens(zs_in), Vg)
    avg, dyn, rate = ops.euler_ssn(tens(z), Jg, Dg, Sg, ext_g, seqlen=seqlen, skip_steps=skip,
                                   rate_penalty_threshold=0.5)
    np.testing.assert_allclose(avg.detach().cpu().numpy(), avg_o.detach().numpy(), rtol=1e-4, atol=1e-5)
    ((avg * tens(G)).sum() + 3.0 * dyn + 2.0 * rate).backward()
    np.testing.assert_allclose(Vg.grad.cpu().numpy(), Vo.grad.numpy(), rtol=1e-4, atol=1e-4 * np.abs(Vo.grad.numpy()).max())
    np.testing.assert_allclose(Jg.grad.cpu().numpy(), J.grad.numpy(), rtol=1e-4, atol=1e-4 * np.abs(J.grad.numpy()).max())
    # ---- fixed-point path, 'deg-heteroin' (scalar V) ----
    Vs = tens(0.25, torch.float64, grad=True)
    Jg, Dg, Sg = (tens(jds[k], torch.float64, grad=True) for k in 'JDS')
    ext_g = ops.hetero_input(tens(exts), tens(zs_in), Vs)
    R, status, _ = ops.ssn_fixed_point(tens(z), Jg, Dg, Sg, ext_g)
    assert (status == 0).all()
    (R * tens(G)).sum().backward()
    ext_np = (1 + 0.25 * zs_in[:, None, :]) * exts[None]
    W = oracle.generate_weight(n_sites, jds['J'], jds['D'], jds['S'], z)
    Rn = R.detach().double().cpu().numpy()
    dV = 0.0
    for iz in range(nz):
        for ib in range(nb):
            phi = oracle.io_gain(W[iz] @ Rn[iz, ib] + ext_np[iz, ib])
            mu = np.linalg.solve(np.eye(2 * n_sites) - W[iz].T * phi[None, :], G[iz, ib])
            dV += np.sum(phi * mu * zs_in[iz] * exts[ib])           # dL/d ext * d ext/dV
    assert abs(float(Vs.grad) - dV) <= 2e-4 * abs(dV), (float(Vs.grad), dV)


def test_rejected_network_does_not_poison_the_implicit_gradient(ops, oracle):
    """A network masked out by the caller (dL/dr = 0) whose state is non-finite -- e.g. a diverged asym_power solve --
    must contribute exactly nothing: the 12 accumulators equal those of the batch without it, and the adjoint
    status of its solves is 0 with 0 sweeps."""
    import torch
    n_sites, nz, nb = 30, 3, 8
    jds = oracle.new_JDS()
    exts = oracle.stimulus_input(oracle.DEFAULT_BANDWIDTHS, n_sites)
    rs = np.random.RandomState(3)
    z = rs.rand(nz, 2 * n_sites, 2 * n_sites).astype(np.float32).astype(np.float64)
    W = oracle.generate_weight(n_sites, jds['J'], jds['D'], jds['S'], z)
    R, st, _ = oracle.fixed_point_batch(W, exts)
    assert (st == 0).all()
    gR = rs.randn(nz, nb, 2 * n_sites)
    gR[1] = 0.0
    R_bad = R.copy()
    R_bad[1] = np.nan
    R_bad[1, 3] = np.inf
    solver = ops.make_solver()
    J, D, S = (tens(jds[k], torch.float64) for k in 'JDS')
    keep = [0, 2]
    want = ops.ift_gradient(tens(z[keep]), J, D, S, tens(exts), tens(R[keep]), tens(gR[keep]), solver=solver)
    got = ops.ift_gradient(tens(z), J, D, S, tens(exts), tens(R_bad), tens(gR), solver=solver, return_mu=True)
    for a, b in zip(got[:3], want[:3]):
        assert torch.isfinite(a).all()
        np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-9, atol=1e-9 * float(b.abs().max()))
    mu, status, iters = got[3:6]
    assert (status.cpu().numpy() == 0).all() and (iters[1].cpu().numpy() == 0).all()
    assert (mu[1] == 0).all()


def test_no_grad_forward_stores_no_trajectory(ops, monkeypatch):
    """Critic updates run the generator under torch.no_grad(): J, D, S still have requires_grad, but the forward
    must not store the trajectory and gain arrays (2 x nz * seqlen * nb * 2N floats) that only backward reads."""
    import torch
    from tc_gan_b200 import ssnode
    seen = []
    real = ops.euler_forward

    def spy(*args, **kwargs):
        seen.append(kwargs.get('store'))
        return real(*args, **kwargs)

    monkeypatch.setattr(ops, 'euler_forward', spy)
    jds = ssnode.new_JDS()
    J, D, S = (tens(jds[k], torch.float64, grad=True) for k in 'JDS')
    z = tens(np.random.RandomState(0).rand(2, 20, 20))
    ext = tens(np.random.RandomState(1).rand(3, 20))
    with torch.no_grad():
        avg, _, _ = ops.euler_ssn(z, J, D, S, ext, seqlen=20, skip_steps=10)
    avg2, _, _ = ops.euler_ssn(z, J, D, S, ext, seqlen=20, skip_steps=10)
    assert seen == [False, True] and not avg.requires_grad and avg2.requires_grad
    torch.testing.assert_close(avg, avg2.detach())


@pytest.mark.parametrize('seed', [21, 22])
def test_randomised_gradient_stress(ops, seed):
    """24 random cases per seed of tools/dev_stress_grad.py: K2 (GMRES) and K3/K4/K4b against the float64 oracle
    over sizes from 2N = 2 to 560 (every cluster width), ragged stimulus counts, all transfer functions, unrolls of
    1..41 steps, to rtol 1e-4.  (460 cases of it ran clean on a B200 after the round-2 kernel changes.)"""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CASES='24', SEED=str(seed))
    out = subprocess.run([sys.executable, os.path.join(root, 'tools', 'dev_stress_grad.py')], env=env,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert 'done: 24 cases, 0 mismatches' in out.stdout, out.stdout[-3000:]

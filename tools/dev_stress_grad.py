"""Randomised parity stress of the gradient kernels (K2 GMRES, K3/K4 Euler + BPTT, K4b) against the float64 oracle
(development): random sizes across every cluster width, stimulus counts, transfer functions, unroll lengths."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'oracle'))
import numpy as np, torch
import ssn_oracle as so
from tc_gan_b200 import torch_ops as ops

dev = torch.device('cuda:0')
rs = np.random.RandomState(int(os.environ.get('SEED', 0)))
n_cases = int(os.environ.get('CASES', 40))
t = lambda a, dt=torch.float32, g=False: (lambda x: x.requires_grad_() if g else x)(torch.tensor(np.asarray(a), dtype=dt, device=dev))
bad = 0
t0 = time.time()
for case in range(n_cases):
    n_sites = int(rs.choice([1, 2, 5, 13, 28, 33, 57, 64, 101, 125, 201, 280]))
    nb = int(rs.choice([1, 3, 4, 5, 8, 8, 9, 12, 17]))
    nz = int(rs.randint(1, 4)) if n_sites > 64 else int(rs.choice([1, 2, 5, 19, 41]))
    io_type = str(rs.choice(['asym_tanh', 'asym_tanh', 'asym_linear', 'asym_power']))
    jds = so.new_JDS()
    dim = 2 * n_sites
    bw = np.sort(rs.rand(nb))
    exts = so.stimulus_input(bw, n_sites, contrasts=(float(rs.choice([5., 20., 40.])),))[:nb]
    z = rs.rand(nz, dim, dim).astype(np.float32).astype(np.float64)
    W = so.generate_weight(n_sites, jds['J'], jds['D'], jds['S'], z)
    G = rs.randn(nz, nb, dim).astype(np.float32).astype(np.float64)
    what = str(rs.choice(['ift', 'bptt']))
    ok = True
    if what == 'ift':
        R, st, _ = so.fixed_point_batch(W, exts, io_type=io_type, threads=8)
        keep = (st == 0).all(axis=1)
        if not keep.any():
            print('case %d: no converged network, skipped' % case); continue
        R = R.astype(np.float32).astype(np.float64)
        G[~keep] = 0.0                                   # rejected networks are masked out by the caller
        R[~keep] = 0.0
        dJ, dD, dS, mu = so.ift_param_gradient(R, W, z, exts, jds['J'], jds['D'], jds['S'], G, io_type=io_type)
        J, D, S = (t(jds[k], torch.float64) for k in 'JDS')
        got = ops.ift_gradient(t(z), J, D, S, t(exts), t(R), t(G), solver=ops.make_solver(io_type=io_type), return_mu=True)
        detail = 'sweeps mean %.1f max %d status!=0 %d' % (got[5].float().mean(), int(got[5].max()), int((got[4] != 0).sum()))
        want = (dJ, dD, dS)
        got = got[:3]
    else:
        seqlen = int(rs.choice([1, 2, 7, 20, 41])); skip = int(rs.randint(0, seqlen))
        eps = (0.01, 0.1); thr = 0.5; c_dyn, c_rate = 3.0, 2.0
        t64 = lambda a, g=False: (lambda x: x.requires_grad_() if g else x)(torch.tensor(np.asarray(a), dtype=torch.float64))
        Jo, Do, So = (t64(jds[k], True) for k in 'JDS')
        avg_o, dyn_o, rate_o = so.euler_unroll_torch(t64(z), Jo, Do, So, t64(exts), seqlen, skip, eps[0], eps[1],
                                                     io_type=io_type, rate_penalty_threshold=thr)
        ((avg_o * t64(G)).sum() + c_dyn * dyn_o + c_rate * rate_o).backward()
        want = tuple(p.grad.numpy() for p in (Jo, Do, So))
        J, D, S = (t(jds[k], torch.float64, True) for k in 'JDS')
        avg, dyn, rate = ops.euler_ssn(t(z), J, D, S, t(exts), seqlen=seqlen, skip_steps=skip, dt=0.1, tau_E=10.0, tau_I=1.0,
                                       io_type=io_type, rate_penalty_threshold=thr)
        ((avg * t(G)).sum() + c_dyn * dyn + c_rate * rate).backward()
        got = (J.grad, D.grad, S.grad)
        ok = np.allclose(avg.detach().cpu().numpy(), avg_o.detach().numpy(), rtol=1e-4, atol=1e-5)
        detail = 'seqlen %d skip %d avg %s' % (seqlen, skip, ok)
    errs = []
    for g_, w_ in zip(got, want):
        g_ = g_.detach().cpu().numpy()
        errs.append(np.abs(g_ - w_).max() / max(np.abs(w_).max(), 1e-300))
    ok = ok and max(errs) < 1e-4
    bad += not ok
    print('case %2d: %-4s n_sites %3d nb %2d nz %2d %-11s | %s | max rel err %.1e %s' % (
        case, what, n_sites, nb, nz, io_type, detail, max(errs), 'ok' if ok else 'MISMATCH'), flush=True)
print('done: %d cases, %d mismatches, %.1f s' % (n_cases, bad, time.time() - t0))

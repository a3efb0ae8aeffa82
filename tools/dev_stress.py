"""Randomised parity stress of the default fixed-point kernel against the float64 oracle (development)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'oracle'))
import numpy as np
import ssn_oracle as so
from tc_gan_b200 import ssnode
from tc_gan_b200.weight_gen import generate_weight

rs = np.random.RandomState(int(os.environ.get('SEED', 0)))
n_cases = int(os.environ.get('CASES', 40))
bad = 0
t0 = time.time()
for case in range(n_cases):
    n_sites = int(rs.choice([1, 2, 5, 13, 28, 29, 40, 56, 57, 84, 101, 130, 168, 201, 224]))
    nb = int(rs.choice([1, 2, 3, 4, 5, 7, 8, 9, 12, 13, 16, 17, 23]))
    nz = int(rs.randint(1, 4))
    io_type = str(rs.choice(['asym_tanh', 'asym_tanh', 'asym_linear', 'asym_power']))
    use_r0 = rs.rand() < 0.3
    atol = float(rs.choice([1e-5, 1e-5, 1e-7, 3e-4]))
    max_iter = int(rs.choice([10000, 10000, 300, 57]))
    jds = so.new_JDS()
    scale = float(rs.choice([1.0, 1.0, 1.3, 0.7]))
    J = jds['J'] * scale
    zs = rs.rand(nz, 2 * n_sites, 2 * n_sites)
    W = np.array([generate_weight(n_sites, J, jds['D'], jds['S'], z) for z in zs])
    bw = np.sort(rs.rand(nb))
    exts = so.stimulus_input(bw, n_sites, contrasts=(float(rs.choice([5., 20., 40.])),))[:nb]
    r0 = rs.rand(2 * n_sites) * 5 if use_r0 else None
    kw = dict(io_type=io_type, atol=atol, max_iter=max_iter)
    if io_type != 'asym_tanh':
        kw['rate_stop_at'] = 200.0
    if r0 is None:
        Ro, st_o, it_o = so.fixed_point_batch(W, exts, threads=16, **kw)
    else:
        Ro = np.empty((nz, nb, 2 * n_sites)); st_o = np.empty((nz, nb), int); it_o = np.empty((nz, nb), int)
        for z in range(nz):
            for b in range(nb):
                Ro[z, b], st_o[z, b], it_o[z, b] = so.fixed_point(W[z], exts[b], r0=r0, **kw)
    R, err, its = ssnode.fixed_points_batch(W, exts, k=0.01, n=2.2, r0=r0, **kw)
    ok_status = (err == st_o).all()
    conv = st_o == 0
    d_it = np.abs(its - it_o)
    tol = 1e-5 * np.maximum(1, it_o[..., None] / 1000.0) + 1e-5 * np.abs(Ro)
    ok_val = (np.abs(R - Ro)[conv] <= tol[conv] * max(1.0, atol / 1e-5)).all() if conv.any() else True
    ok_it = (d_it[conv] <= 1 + it_o[conv] // 1000).all() if conv.any() else True
    flag = '' if (ok_status and ok_val and ok_it) else '   <<<<<< MISMATCH'
    bad += bool(flag)
    print('case %2d: n_sites %3d nb %2d nz %d %-11s r0 %d atol %.0e max_iter %5d scale %.1f | status %s values %s sweeps %s (max d_it %d, codes %s)%s' % (
        case, n_sites, nb, nz, io_type, use_r0, atol, max_iter, scale, ok_status, ok_val, ok_it, int(d_it[conv].max()) if conv.any() else -1,
        np.bincount(st_o.ravel(), minlength=3).tolist(), flag), flush=True)
print('done: %d cases, %d mismatches, %.1f s' % (n_cases, bad, time.time() - t0))

"""Development check of K1 on a GPU box: parity vs the oracle + rough timing."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'oracle'))
import ssn_oracle as so
from tc_gan_b200 import ssnode, stimuli, clib
from tc_gan_b200.weight_gen import generate_weight

P = ssnode.DEFAULT_PARAMS
jds = ssnode.new_JDS()
cs, rc = clib.c_int(), clib.c_int()
for n_sites, nz in ((51, 6), (101, 4), (201, 8)):
    clib.libssnode.ssn_fixed_point_occupancy(n_sites, cs, rc)
    print('n_sites', n_sites, 'cluster', cs.value, 'resident clusters', rc.value, flush=True)
    rs = np.random.RandomState(0)
    zs = rs.rand(nz, 2 * n_sites, 2 * n_sites)
    Ws = np.array([generate_weight(n_sites, jds['J'], jds['D'], jds['S'], z) for z in zs])
    exts = stimuli.input(P['bandwidths'], np.linspace(-.5, .5, n_sites), P['smoothness'], P['contrast'])
    Ro, so_status, so_iters = so.fixed_point_batch(Ws, exts, threads=8)
    for precise in (False, True):
        t0 = time.time()
        R, err, its = ssnode.fixed_points_batch(Ws, exts, k=P['k'], n=P['n'], precise=precise)
        dt = time.time() - t0
        rel = np.abs(R - Ro) / (1e-4 + np.abs(Ro))
        print('  precise', precise, 'status ok', (err == so_status).all(), 'max rel', rel.max(),
              'max abs', np.abs(R - Ro).max(), 'iters diff', np.abs(its - so_iters).max(),
              'iters', its.min(), its.max(), 'time %.3f' % dt, flush=True)
    sol = ssnode.fixed_point(Ws[0], exts[3], k=P['k'], n=P['n'])
    print('  legacy:', sol.message, np.abs(sol.x - Ro[0, 3]).max(), flush=True)

# rough throughput at 2N=402 with z on the host (float32) through the batched C ABI
n_sites, nz, nb = 201, 256, 8
dim = 2 * n_sites
z = np.random.RandomState(1).rand(nz, dim, dim).astype(np.float32)
exts = stimuli.input(P['bandwidths'], np.linspace(-.5, .5, n_sites), P['smoothness'], P['contrast']).astype(np.float32)
R = np.empty((nz, nb, dim), np.float32); st = np.empty((nz, nb), np.int32); it = np.empty((nz, nb), np.int32)
sv = clib.make_solver(k=P['k'], n=P['n'])
j = clib.make_jds(jds['J'], jds['D'], jds['S'])
for rep in range(3):
    t0 = time.time()
    clib.check_call(clib.libssnode.ssn_fixed_point_batch(sv, nz, nb, n_sites, clib.W_FROM_Z, z.ctypes.data, j,
        exts.ctypes.data, 0, None, R.ctypes.data, st.ctypes.data, it.ctypes.data, 0, clib.MEM_HOST, None), 'batch')
    dt = time.time() - t0
    print('batch %d x %d: %.3f s -> %.0f solves/s (converged %d, mean iters %.1f, max %d)' % (
        nz, nb, dt, (st == 0).sum() / dt, (st == 0).sum(), it.mean(), it.max()), flush=True)
Wchk = np.array([generate_weight(n_sites, jds['J'], jds['D'], jds['S'], zz.astype(np.float64)) for zz in z[:4]])
Ro, so_status, so_iters = so.fixed_point_batch(Wchk, exts.astype(np.float64), threads=8)
print('z-path parity: max rel', (np.abs(R[:4] - Ro) / (1e-4 + np.abs(Ro))).max(), 'iters diff', np.abs(it[:4] - so_iters).max())
print('launches', clib.kernel_launches())

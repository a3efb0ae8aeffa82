"""Cycles per sweep of K1: every solve runs exactly MAXIT sweeps (atol = 0)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
from tc_gan_b200 import clib, ssnode, stimuli
n_sites = int(os.environ.get('NSITES', 201)); dim = 2 * n_sites
maxit = int(os.environ.get('MAXIT', 300))
P = ssnode.DEFAULT_PARAMS; jds = ssnode.new_JDS()
exts = stimuli.input(P['bandwidths'], np.linspace(-.5, .5, n_sites), P['smoothness'], P['contrast'])
nb = len(exts); dev = torch.device('cuda:0')
cs, rc = clib.c_int(), clib.c_int()
clib.libssnode.ssn_fixed_point_occupancy(n_sites, cs, rc)
nz = rc.value * int(os.environ.get('ROUNDS', 2))
g = torch.Generator(device=dev); g.manual_seed(0)
z = torch.rand((nz, dim, dim), generator=g, device=dev)
e = torch.tensor(exts, dtype=torch.float32, device=dev)
R = torch.empty((nz, nb, dim), device=dev); st = torch.empty((nz, nb), dtype=torch.int32, device=dev); it = torch.empty_like(st)
sv = clib.make_solver(k=P['k'], n=P['n'], atol=0.0, max_iter=maxit); jd = clib.make_jds(jds['J'], jds['D'], jds['S'])
def run():
    clib.check_call(clib.libssnode.ssn_fixed_point_batch(sv, nz, nb, n_sites, clib.W_FROM_Z, z.data_ptr(), jd, e.data_ptr(), 0, None,
        R.data_ptr(), st.data_ptr(), it.data_ptr(), 0, clib.MEM_DEVICE, torch.cuda.current_stream().cuda_stream), 'k1')
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
rounds = nz / rc.value
print('  mean sweeps reported %.1f (lib %s)' % (float(it.float().mean()), os.environ.get('SSN_LIBNAME', 'libssnode')))
print('dbg=%s cluster=%d resident=%d nz=%d: %.3f ms -> %.2f us/sweep = %.0f cycles @1.965GHz (incl. W load amortised over %d sweeps)' % (
    os.environ.get('SSN_DBG', '0'), cs.value, rc.value, nz, ms, ms * 1e3 / (rounds * maxit), ms * 1e3 / (rounds * maxit) * 1965, maxit))

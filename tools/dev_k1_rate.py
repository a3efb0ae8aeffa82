"""Converged solves/s of K1 at configs[1] (device-resident), for comparing build variants / shapes (development)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
from tc_gan_b200 import clib, ssnode, stimuli
n_sites, nz = 201, int(os.environ.get('NZ', 1024)); dim = 2 * n_sites
P = ssnode.DEFAULT_PARAMS; jds = ssnode.new_JDS()
exts = stimuli.input(P['bandwidths'], np.linspace(-.5, .5, n_sites), P['smoothness'], P['contrast'])
nb = len(exts); dev = torch.device('cuda:0')
g = torch.Generator(device=dev); g.manual_seed(0)
z = torch.rand((nz, dim, dim), generator=g, device=dev)
e = torch.tensor(exts, dtype=torch.float32, device=dev)
R = torch.empty((nz, nb, dim), device=dev); st = torch.empty((nz, nb), dtype=torch.int32, device=dev); it = torch.empty_like(st)
sv = clib.make_solver(k=P['k'], n=P['n']); jd = clib.make_jds(jds['J'], jds['D'], jds['S'])
def run():
    clib.check_call(clib.libssnode.ssn_fixed_point_batch(sv, nz, nb, n_sites, clib.W_FROM_Z, z.data_ptr(), jd, e.data_ptr(), 0, None,
        R.data_ptr(), st.data_ptr(), it.data_ptr(), 0, clib.MEM_DEVICE, torch.cuda.current_stream().cuda_stream), 'k1')
run(); run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print('%s shape=%s %s: %.2f ms -> %.1f k solves/s (converged %d, mean sweeps %.1f, checksum %.6f)' % (
    os.environ.get('SSN_LIBNAME', 'libssnode'), os.environ.get('SSN_WS_SHAPE', '-'), clib.fixed_point_kernel_tag(n_sites), ms,
    int((st == 0).sum()) / ms, int((st == 0).sum()), float(it.float().mean()), float(R.double().sum())))

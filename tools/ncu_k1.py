"""Small K1 run for ncu: 66 networks x 8 stimuli at 2N=402 on the device path."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
from tc_gan_b200 import clib, ssnode, stimuli
n_sites, nz = 201, int(os.environ.get('NZ', 66)); dim = 2 * n_sites
P = ssnode.DEFAULT_PARAMS; jds = ssnode.new_JDS()
exts = stimuli.input(P['bandwidths'], np.linspace(-.5, .5, n_sites), P['smoothness'], P['contrast'])
nb = len(exts); dev = torch.device('cuda:0')
g = torch.Generator(device=dev); g.manual_seed(0)
z = torch.rand((nz, dim, dim), generator=g, device=dev)
e = torch.tensor(exts, dtype=torch.float32, device=dev)
R = torch.empty((nz, nb, dim), device=dev); st = torch.empty((nz, nb), dtype=torch.int32, device=dev); it = torch.empty_like(st)
sv = clib.make_solver(k=P['k'], n=P['n']); jd = clib.make_jds(jds['J'], jds['D'], jds['S'])
for rep in range(int(os.environ.get('REPS', 2))):
    clib.check_call(clib.libssnode.ssn_fixed_point_batch(sv, nz, nb, n_sites, clib.W_FROM_Z, z.data_ptr(), jd, e.data_ptr(), 0, None,
        R.data_ptr(), st.data_ptr(), it.data_ptr(), 0, clib.MEM_DEVICE, torch.cuda.current_stream().cuda_stream), 'k1')
torch.cuda.synchronize()
print('ok', int((st == 0).sum()), float(it.float().mean()))

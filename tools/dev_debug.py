import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'oracle'))
import numpy as np, ssn_oracle as so
from tc_gan_b200 import ssnode
n_sites = int(sys.argv[1]) if len(sys.argv) > 1 else 51
jds = so.new_JDS(); rs = np.random.RandomState(0)
zs = rs.rand(2, 2*n_sites, 2*n_sites)
W = so.generate_weight(n_sites, jds['J'], jds['D'], jds['S'], zs)
exts = so.stimulus_input(so.DEFAULT_BANDWIDTHS, n_sites)
Ro, st, it = so.fixed_point_batch(W, exts)
R, err, its = ssnode.fixed_points_batch(W, exts, k=0.01, n=2.2)
print('status', err.tolist(), 'iters', its.tolist(), 'ref iters', it.tolist())
d = np.abs(R - Ro)
print('err by stim', d.max(axis=(0, 2)))

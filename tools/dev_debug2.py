import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'oracle'))
import numpy as np, ssn_oracle as so
from tc_gan_b200 import ssnode
n_sites, nz, nb = 33, 5, 8
jds = so.new_JDS(); rs = np.random.RandomState(n_sites + nb)
zs = np.array([rs.rand(2*n_sites, 2*n_sites) for _ in range(nz)])
W = so.generate_weight(n_sites, jds['J'], jds['D'], jds['S'], zs)
exts = so.stimulus_input(np.linspace(0, 1, nb), n_sites)
Ro, st, it = so.fixed_point_batch(W, exts)
R, err, its = ssnode.fixed_points_batch(W, exts, k=0.01, n=2.2)
print('its', its.tolist()); print('ref', it.tolist()); print('err', np.abs(R-Ro).max(axis=2).round(7).tolist())
# trajectory of max|dr| near the end for the worst case
d = np.abs(its - it); z, b = np.unravel_index(d.argmax(), d.shape)
print('worst', z, b, its[z, b], it[z, b])
r = np.zeros(2*n_sites); eps = np.r_[np.full(n_sites, 8e-4/0.01589), np.full(n_sites, 8e-4/0.002)]
for k in range(1, it[z, b] + 3):
    rn = r + (so.io_fun(W[z] @ r + exts[b]) - r) * eps
    if k > min(its[z, b], it[z, b]) - 3: print(k, np.abs(rn - r).max())
    r = rn

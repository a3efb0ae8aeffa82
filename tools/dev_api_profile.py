"""Where the time of ssnode.find_fixed_points goes at configs[1] size (development): cProfile of one call with float64 W
per network and one with float32 z + jds."""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
from tc_gan_b200 import ssnode, stimuli
from tc_gan_b200.weight_gen import generate_weight
n_sites, nz = 201, int(os.environ.get('NZ', 1024)); dim = 2 * n_sites
P = ssnode.DEFAULT_PARAMS; jds = ssnode.new_JDS()
exts = stimuli.input(P['bandwidths'], np.linspace(-.5, .5, n_sites), P['smoothness'], P['contrast'])
rs = np.random.RandomState(0)
zs = [rs.rand(dim, dim) for _ in range(nz)]
Ws = [generate_weight(n_sites, jds['J'], jds['D'], jds['S'], z) for z in zs]
z32 = [z.astype(np.float32) for z in zs]
kw = dict(k=P['k'], n=P['n'])
for name, pairs, extra in (('float64 W', list(zip(zs, Ws)), {}), ('float32 z + jds', [(z, None) for z in z32], {'jds': jds})):
    ssnode.find_fixed_points(nz, iter(pairs), exts, **kw, **extra)
    t0 = time.time()
    for _ in range(3):
        ssnode.find_fixed_points(nz, iter(pairs), exts, **kw, **extra)
    dt = (time.time() - t0) / 3
    print('%s: %.1f ms per call -> %.1f k solves/s' % (name, dt * 1e3, nz * 8 / dt / 1e3))
    pr = cProfile.Profile(); pr.enable()
    ssnode.find_fixed_points(nz, iter(pairs), exts, **kw, **extra)
    pr.disable()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(12)

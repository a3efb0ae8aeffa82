"""Throughput of the reference-ABI symbols (one network x one stimulus per call) from a thread pool,
as tc_gan.ssnode.find_fixed_points_parallel drives them (ssnode.py:423-510)."""
import os, sys, time
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
from tc_gan_b200 import ssnode, stimuli
from tc_gan_b200.weight_gen import generate_weight
n_sites = int(os.environ.get('NSITES', 201)); nz = int(os.environ.get('NZ', 96))
P = ssnode.DEFAULT_PARAMS; jds = ssnode.new_JDS()
exts = stimuli.input(P['bandwidths'], np.linspace(-.5, .5, n_sites), P['smoothness'], P['contrast'])
rs = np.random.RandomState(0)
Ws = [generate_weight(n_sites, jds['J'], jds['D'], jds['S'], rs.rand(2 * n_sites, 2 * n_sites)) for _ in range(nz)]
jobs = [(W, e) for W in Ws for e in exts]
def solve(job):
    return ssnode.fixed_point(job[0], job[1], k=P['k'], n=P['n']).success
solve(jobs[0])
for threads in (1, 4, 8, 16, 32):
    # a warm pool, as in the reference (its pool lives for a whole find_fixed_points call): the first call of a thread
    # creates its stream and pinned staging buffer
    with ThreadPoolExecutor(threads) as ex:
        sum(ex.map(solve, jobs[:4 * threads]))
        t0 = time.time()
        ok = sum(ex.map(solve, jobs))
        dt = time.time() - t0
    print('legacy symbols, %2d host threads: %d solves in %.2f s -> %.0f solves/s (%d converged)' % (threads, len(jobs), dt, len(jobs) / dt, ok), flush=True)

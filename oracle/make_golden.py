"""
Generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF in this container.

Run once here (``python oracle/make_golden.py``); the fixtures are committed
because /root/reference does not exist on the GPU box.  What is imported from
the reference, unmodified:
  * tc_gan/weight_gen.py, tc_gan/stimuli.py                (numpy only)
  * tc_gan/ssnode.py -> fixed_point / find_fixed_points    (through tc_gan/clib.py
    and the reference C file compiled with the reference's flags)
  * tc_gan/assets/*.mat                                    (the MATLAB golden vectors
    used by the reference's tests/test_dynamics.py:43-126)
tc_gan's package __init__ chain imports theano (utils/numerics.py:4,
utils/theanoutils.py:4), which is not installed; a stub module that provides
only the attributes touched at import time is injected.  /root/reference is
read-only, so the package is used from a scratch copy that also receives the
compiled libssnode.so (clib.py:7-15 looks for it under tc_gan/ext/).
"""
import os
import shutil
import subprocess
import sys
import tempfile
import types
import warnings

import numpy as np
import scipy.io

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '..', 'tests', 'golden')
REF = '/root/reference'


def import_reference():
    subprocess.check_call(['make', '-s', '-C', HERE, 'ref'])
    scratch = tempfile.mkdtemp(prefix='tcgan_ref_')
    shutil.copytree(os.path.join(REF, 'tc_gan'), os.path.join(scratch, 'tc_gan'))
    shutil.copy(os.path.join(HERE, '_ref', 'libssnode.so'),
                os.path.join(scratch, 'tc_gan', 'ext', 'libssnode.so'))
    theano = types.ModuleType('theano')
    theano.config = types.SimpleNamespace(floatX='float32')
    theano.Variable = type('Variable', (), {})
    theano.tensor = types.ModuleType('theano.tensor')
    theano.function = None
    sys.modules['theano'] = theano
    sys.modules['theano.tensor'] = theano.tensor
    sys.path.insert(0, scratch)
    import tc_gan.ssnode as ssnode          # noqa: E402
    import tc_gan.stimuli as stimuli        # noqa: E402
    import tc_gan.weight_gen as weight_gen  # noqa: E402
    return ssnode, stimuli, weight_gen


def new_JDS(ssnode):
    # networks/fixed_time_sampler.py:12-23 (module itself needs lasagne)
    J, D, S = (ssnode.DEFAULT_PARAMS[k] for k in 'JDS')
    D_new = D / 2
    return J + D / 2 - D_new / 2, D_new, S


def main():
    warnings.simplefilter('ignore')
    os.makedirs(OUT, exist_ok=True)
    ssnode, stimuli, weight_gen = import_reference()
    P = ssnode.DEFAULT_PARAMS

    # ---- 1. MATLAB golden vectors (tests/test_dynamics.py:43-126) -------------
    conn = scipy.io.loadmat(os.path.join(REF, 'tc_gan/assets/target_parameters_GAN-SSN_Ne51-Zs.mat'))
    data = scipy.io.loadmat(os.path.join(REF, 'tc_gan/assets/training_data_TCs_Ne51-Zs.mat'))
    mz = conn['Zs']
    N = mz.shape[0]
    Z = np.zeros((2 * N, 2 * N))
    Z[:N, :N], Z[N:, :N] = mz[:, :, 0, 0], mz[:, :, 1, 0]     # test_dynamics.py:58-62
    Z[:N, N:], Z[N:, N:] = mz[:, :, 0, 1], mz[:, :, 1, 1]
    tp = conn['Targetparams']
    mp = data['Modelparams'][0, 0]
    L = mp['L'][0, 0]
    # modern scipy hands back a Fortran-ordered array; the reference passes the raw
    # buffer to C (ssnode.py:227,247), so force the row-major layout it assumes.
    W_mat = np.ascontiguousarray(conn['W'].toarray(), dtype=float)
    bandwidths = mp['bandwidths'][0] / L
    smoothness = mp['l_margin'][0, 0] / L
    contrast = float(mp['c'][0, 0])
    k, n = float(mp['k'][0, 0]), float(mp['n'][0, 0])
    exts = stimuli.input(bandwidths, np.linspace(-.5, .5, N), smoothness, [contrast])
    fps = {}
    for io_type in ('asym_linear', 'asym_power', 'asym_tanh'):
        _, (x,), _ = ssnode.find_fixed_points(
            1, iter([(None, W_mat)]), exts, k=k, n=n, r0=np.zeros(2 * N),
            io_type=io_type, method='serial', check=True)
        fps[io_type] = np.array(x)
    np.savez_compressed(
        os.path.join(OUT, 'matlab_ne51.npz'),
        W=W_mat, Z=Z, J=tp['Jlow'][0, 0], D=tp['dJ'][0, 0], S=tp['sigmas'][0, 0] / 8,
        E_Tuning=data['E_Tuning'], bandwidths=bandwidths, smoothness=smoothness,
        contrast=contrast, k=k, n=n, n_sites=N, exts=exts,
        fp_asym_linear=fps['asym_linear'], fp_asym_power=fps['asym_power'],
        fp_asym_tanh=fps['asym_tanh'])

    # ---- 2. W(z) and stimuli from the reference's numpy modules ----------------
    J1, D1, S1 = new_JDS(ssnode)
    z7 = np.random.RandomState(7).rand(2 * 7, 2 * 7)
    W7 = weight_gen.generate_weight(7, J1, D1, S1, z7)
    W7_default = weight_gen.generate_weight(7, P['J'], P['D'], P['S'], z7)
    x51 = np.linspace(-.5, .5, 51)
    stim8 = stimuli.input(P['bandwidths'], x51, P['smoothness'], P['contrast'])
    stim50 = stimuli.input(np.linspace(0, 1, 10), x51, P['smoothness'], [5, 10, 20, 30, 40])
    stim_off = stimuli.input([0.25, 0.5], x51, P['smoothness'], [20, 10], [-0.25, 0.0, 0.25])
    np.savez_compressed(os.path.join(OUT, 'weights_stimuli.npz'),
                        z7=z7, W7=W7, W7_default=W7_default, stim8=stim8, stim50=stim50,
                        stim_off=stim_off, J_new=J1, D_new=D1, S_new=S1)

    # ---- 3. batched fixed points, configs 1 and 2 in miniature -----------------
    out = {}
    for n_sites, nz in ((51, 6), (201, 2)):
        x = np.linspace(-.5, .5, n_sites)
        exts = stimuli.input(P['bandwidths'], x, P['smoothness'], P['contrast'])
        rs = np.random.RandomState(0)

        def gen():
            while True:
                z = rs.rand(2 * n_sites, 2 * n_sites)
                yield z, weight_gen.generate_weight(n_sites, J1, D1, S1, z)

        for io_type in (('asym_tanh', 'asym_linear', 'asym_power') if n_sites == 51 else ('asym_tanh',)):
            rs.seed(0)
            zs, Rs, info = ssnode.find_fixed_points(
                nz, gen(), exts, method='serial', k=P['k'], n=P['n'], io_type=io_type,
                r0=np.zeros(2 * n_sites))
            assert info.rejections == 0
            out['R_%d_%s' % (n_sites, io_type)] = Rs
    np.savez_compressed(os.path.join(OUT, 'fixed_points.npz'), **out)

    # ---- 4. tests/test_ssn.py:66-74 inputs (default J,D,S, bandwidth 1, atol 1e-10)
    xs, zs_seed = [], []
    Nn = P['N']
    for seed in range(10):
        Zr = np.random.RandomState(seed).rand(1, 2 * Nn, 2 * Nn)
        Wn = weight_gen.generate_weight(Nn, P['J'], P['D'], P['S'], Zr[0])
        ext, = stimuli.input([1], np.linspace(-.5, .5, Nn), P['smoothness'], P['contrast'])
        sol = ssnode.fixed_point(Wn, ext, r0=np.zeros(2 * Nn), k=P['k'], n=P['n'],
                                 io_type='asym_tanh', atol=1e-10, tau=(.016, .002))
        assert sol.success
        xs.append(sol.x)
    np.savez_compressed(os.path.join(OUT, 'ssn_seeds_atol1e-10.npz'), x=np.array(xs))

    # ---- 5. failure codes: divergence (test_dynamics.py:129-137), rejections ----
    sol = ssnode.fixed_point(W=[[2, 0], [0, 0]], ext=[10, 10], k=1, n=1, r0=[0, 0],
                             max_iter=10000000, io_type='asym_linear')
    codes = dict(inf_code=sol.error, inf_message=sol.message)
    # asym_power + rate_stop_at=200 with the ORIGINAL (less stable) J, D: the
    # configuration dataset generation uses (networks/dataset.py:28-71).
    n_sites = 51
    exts = stimuli.input(P['bandwidths'], np.linspace(-.5, .5, n_sites), P['smoothness'], P['contrast'])
    status = np.zeros((24, len(exts)), dtype=int)
    Rrej = np.zeros((24, len(exts), 2 * n_sites))
    rs = np.random.RandomState(3)
    for i in range(24):
        z = rs.rand(2 * n_sites, 2 * n_sites)
        Wn = weight_gen.generate_weight(n_sites, P['J'], P['D'], P['S'], z)
        for b, ext in enumerate(exts):
            sol = ssnode.fixed_point(Wn, ext, r0=np.zeros(2 * n_sites), k=P['k'], n=P['n'],
                                     io_type='asym_power', rate_stop_at=200, max_iter=3000)
            status[i, b] = sol.error
            if sol.error == 0:
                Rrej[i, b] = sol.x
    np.savez_compressed(os.path.join(OUT, 'failure_codes.npz'), status_power=status,
                        R_power=Rrej, **codes)
    print('status histogram (asym_power, stop_at 200):', np.bincount(status.ravel()))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == '__main__':
    main()

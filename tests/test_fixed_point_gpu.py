"""
GPU parity tests of the fixed-point path (K1) through the C ABI.

Tolerances.  The reference stops as soon as one Euler step moves every rate by
< atol = 1e-5, i.e. ~1e-3 away from the true fixed point, so a solver that stops k
sweeps earlier or later differs by k x 1e-5: matching the reference to better than
1e-4 means stopping at the SAME sweep.
* default kernel (register-resident W, FP32 FFMA on r - r_ref, float64 f and state):
  rtol = atol = 1e-5 (BASELINE north_star asks 1e-4) and the sweep count of every
  solve equal to the float64 oracle's (|difference| <= 1 tolerated, >= 95 % equal);
* float64 kernel (`precise=True`, reference-ABI symbols): 1e-10, identical sweep counts;
* shared-memory fallback kernel (sizes beyond 2N = 448; forced here with
  SSN_FORCE_SMEM_KERNEL): rtol 1e-4 + atol 3e-4 as the reference's own cross-solver
  tests (tc_gan/networks/tests/test_euler_ssn.py:36,86).
"""
import os
import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-5
RTOL_SMEM, ATOL_SMEM = 1e-4, 3e-4


def check_sweeps(its, it_ref):
    its, it_ref = np.asarray(its, dtype=np.int64), np.asarray(it_ref, dtype=np.int64)
    d = np.abs(its - it_ref)
    # a marginally stable network (thousands of sweeps, |dr| shrinking by 0.1 % per sweep) may cross atol
    # a few sweeps apart: allow 1 + 0.1 % of the reference count
    assert (d <= 1 + it_ref // 1000).all(), (int(d.max()), int(it_ref.ravel()[d.argmax()]))
    assert (d == 0).mean() >= 0.95, (d == 0).mean()


@pytest.fixture(scope='module')
def ssn(built_library):
    from tc_gan_b200 import clib, ssnode
    if clib.libssnode.ssn_device_count() < 1:
        pytest.fail('GPU tests need a CUDA device: the library has no CPU fallback')
    return ssnode


def seeded_problem(oracle, n_sites, nz, seed=0, jds=None, bandwidths=None, contrasts=(20.,)):
    jds = jds or oracle.new_JDS()
    rs = np.random.RandomState(seed)
    zs = np.array([rs.rand(2 * n_sites, 2 * n_sites) for _ in range(nz)])
    W = oracle.generate_weight(n_sites, jds['J'], jds['D'], jds['S'], zs)
    exts = oracle.stimulus_input(bandwidths if bandwidths is not None else oracle.DEFAULT_BANDWIDTHS,
                                 n_sites, contrasts=contrasts)
    return zs, W, exts


@pytest.mark.parametrize('n_sites,io_type', [(51, 'asym_tanh'), (51, 'asym_linear'), (51, 'asym_power'),
                                             (201, 'asym_tanh')])
def test_golden_fixed_points(ssn, oracle, n_sites, io_type):
    """Fixtures produced by the reference's own find_fixed_points (oracle/make_golden.py)."""
    g = golden('fixed_points.npz')['R_%d_%s' % (n_sites, io_type)]
    _, W, exts = seeded_problem(oracle, n_sites, len(g))
    R, err, its = ssn.fixed_points_batch(W, exts, k=0.01, n=2.2, io_type=io_type)
    assert (err == 0).all()
    np.testing.assert_allclose(R, g, rtol=RTOL, atol=ATOL)
    Rp, errp, itsp = ssn.fixed_points_batch(W, exts, k=0.01, n=2.2, io_type=io_type, precise=True)
    assert (errp == 0).all()
    np.testing.assert_allclose(Rp, g, rtol=0, atol=1e-10)
    _, _, it_ref = oracle.fixed_point_batch(W, exts, io_type=io_type, threads=4)
    np.testing.assert_array_equal(itsp, it_ref)
    check_sweeps(its, it_ref)


@pytest.mark.parametrize('io_type', ['asym_linear', 'asym_power', 'asym_tanh'])
def test_matlab_golden(ssn, io_type):
    """tc_gan/tests/test_dynamics.py:76-126 through find_fixed_points(method='parallel')."""
    g = golden('matlab_ne51.npz')
    N = int(g['n_sites'])
    (z,), (fps,), info = ssn.find_fixed_points(
        1, iter([('dummy', g['W'])]), g['exts'], k=float(g['k']), n=float(g['n']),
        r0=np.zeros(2 * N), io_type=io_type, method='parallel', check=True)
    assert z == 'dummy' and info.rejections == 0 and info.unused == 0
    center, ofs = N // 2, len(g['E_Tuning']) // 2
    np.testing.assert_allclose(fps[:, center - ofs:center + ofs + 1].T, g['E_Tuning'], rtol=0.1)
    np.testing.assert_allclose(fps, g['fp_' + io_type], rtol=RTOL, atol=ATOL)
    assert all(s.success and s.message == 'Converged' for s in info.solutions[0])


@pytest.mark.parametrize('seed', range(10))
def test_fixed_point_c_abi_seeds(ssn, seed):
    """tc_gan/tests/test_ssn.py:66-74 (atol=1e-10) through the reference ABI symbol."""
    g = golden('ssn_seeds_atol1e-10.npz')['x']
    kwargs = ssn.make_solver_params(seed=seed, io_type='asym_tanh')
    kwargs.update(atol=1e-10, tau=(.016, .002))
    sol = ssn.fixed_point(**kwargs)
    assert sol.success and sol.message == 'Converged'
    np.testing.assert_allclose(sol.x, g[seed], rtol=0, atol=1e-9)


def test_divergence_codes(ssn):
    """tc_gan/tests/test_dynamics.py:129-137 and the check=True exception."""
    sol = ssn.fixed_point(W=[[2, 0], [0, 0]], ext=[10, 10], k=1, n=1, r0=[0, 0],
                          max_iter=10000000, io_type='asym_linear')
    assert sol.message == "Reached to rate_stop_at" and sol.error == 2 and not sol.success
    sol = ssn.fixed_point(W=[[2, 0], [0, 0]], ext=[10, 10], k=1, n=1, r0=[0, 0], max_iter=50,
                          io_type='asym_tanh')
    assert sol.error == 1 and sol.message == "SSN Convergence Failed"
    with pytest.raises(ssn.FixedPointError):
        ssn.fixed_point(W=[[2, 0], [0, 0]], ext=[10, 10], k=1, n=1, max_iter=50, io_type='asym_tanh',
                        check=True)
    # Fortran-ordered input must be honoured (the reference silently transposes it)
    W = np.asfortranarray(np.array([[0., 0.5], [0., 0.]]))
    sol = ssn.fixed_point(W=W, ext=[1., 2.], k=1, n=1, io_type='asym_linear', atol=1e-12, max_iter=100000)
    np.testing.assert_allclose(sol.x, [2., 2.], atol=1e-8)


def test_status_codes_match_reference(ssn, oracle):
    """asym_power + rate_stop_at=200 with the original J, D: a mix of 0 and 2."""
    g = golden('failure_codes.npz')
    nz = len(g['status_power'])
    jds = dict(J=oracle.DEFAULT_J, D=oracle.DEFAULT_D, S=oracle.DEFAULT_S)
    _, W, exts = seeded_problem(oracle, 51, nz, seed=3, jds=jds)
    kw = dict(k=0.01, n=2.2, io_type='asym_power', rate_stop_at=200, max_iter=3000)
    for precise in (False, True):
        R, err, _ = ssn.fixed_points_batch(W, exts, precise=precise, **kw)
        np.testing.assert_array_equal(err, g['status_power'])
        ok = err == 0
        np.testing.assert_allclose(R[ok], g['R_power'][ok], rtol=RTOL, atol=ATOL if not precise else 1e-9)
    assert set(np.unique(g['status_power'])) == {0, 2}


def test_find_fixed_points_rejection_bookkeeping(ssn, oracle):
    """Kept networks = first `num` successes in generator order; counter keyed by the
    error of the first failing stimulus when visited last to first (ssnode.py:390-420)."""
    jds = dict(J=oracle.DEFAULT_J, D=oracle.DEFAULT_D, S=oracle.DEFAULT_S)
    n_sites, num = 51, 6
    exts = oracle.stimulus_input(oracle.DEFAULT_BANDWIDTHS, n_sites)
    kw = dict(k=0.01, n=2.2, io_type='asym_power', rate_stop_at=200, max_iter=3000)

    def gen(rs):
        idx = 0
        while True:
            z = rs.rand(2 * n_sites, 2 * n_sites)
            yield z, oracle.generate_weight(n_sites, jds['J'], jds['D'], jds['S'], z)
            idx += 1

    rs = np.random.RandomState(3)
    zs, Rs, info = ssn.find_fixed_points(num, gen(rs), exts, r0=np.zeros(2 * n_sites), **kw)
    # serial emulation with the oracle
    rs2 = np.random.RandomState(3)
    kept, counter, consumed = [], {}, 0
    for z, W in gen(rs2):
        consumed += 1
        R, st, _ = oracle.fixed_point_batch(W[None], exts, stop_at_first_failure=True,
                                            io_type='asym_power', rate_stop_at=200, max_iter=3000)
        bad = [b for b in range(len(exts) - 1, -1, -1) if st[0, b] > 0]
        if bad:
            counter[int(st[0, bad[0]])] = counter.get(int(st[0, bad[0]]), 0) + 1
        else:
            kept.append((z, R[0]))
        if len(kept) == num:
            break
    np.testing.assert_array_equal(np.array([zz for zz, _ in kept]), zs)
    assert dict(info.counter) == counter and info.rejections == sum(counter.values()) > 0
    assert info.unused == 0
    np.testing.assert_allclose(Rs, np.array([r for _, r in kept]), rtol=RTOL, atol=ATOL)
    # the generator was consumed exactly as far as the serial finder would
    assert rs.rand() == rs2.rand()
    assert Rs.shape == (num, len(exts), 2 * n_sites) and len(info.solutions) == num


@pytest.mark.parametrize('n_sites,nz,nb', [(1, 1, 1), (7, 3, 3), (10, 2, 11), (33, 5, 8), (101, 3, 9),
                                           (128, 2, 2), (201, 2, 50)])
def test_ragged_shapes(ssn, oracle, n_sites, nz, nb):
    """Odd sizes, partial stimulus panels, several panels per network, every cluster width."""
    bw = np.linspace(0, 1, nb) if nb > 1 else [0.5]
    contrasts = (20.,) if nb != 50 else (5, 10, 20, 30, 40)
    if nb == 50:
        bw = np.linspace(0, 1, 10)
    _, W, exts = seeded_problem(oracle, n_sites, nz, seed=n_sites + nb, bandwidths=bw, contrasts=contrasts)
    assert len(exts) == nb
    Ro, st_o, it_o = oracle.fixed_point_batch(W, exts, threads=8)
    R, err, its = ssn.fixed_points_batch(W, exts, k=0.01, n=2.2)
    np.testing.assert_array_equal(err, st_o)
    # (one seeded network here needs 9325 sweeps: a 3-sweep difference is 3e-5 in the rates)
    tol = ATOL * np.maximum(1, it_o[..., None] / 1000.0) + RTOL * np.abs(Ro)
    assert (np.abs(R - Ro) <= tol).all(), float((np.abs(R - Ro) / tol).max())
    check_sweeps(its, it_o)
    Rp, errp, itsp = ssn.fixed_points_batch(W, exts, k=0.01, n=2.2, precise=True)
    np.testing.assert_array_equal(itsp, it_o)
    np.testing.assert_allclose(Rp, Ro, rtol=0, atol=1e-10)


def test_fifty_stimuli_sixteen_networks(ssn, oracle):
    """BASELINE configs[4] shape per network (2N=402, 5 contrasts x 10 bandwidths = 50 stimuli): every network
    runs 13 half-panels through the two streams, i.e. 11 refills of the half-panel queue per network; 16 networks
    so that more clusters than one wave are busy and the queue is exercised with differing convergence orders."""
    n_sites, nz = 201, 16
    _, W, exts = seeded_problem(oracle, n_sites, nz, seed=50, bandwidths=np.linspace(0, 1, 10),
                                contrasts=(5, 10, 20, 30, 40))
    assert len(exts) == 50
    Ro, st_o, it_o = oracle.fixed_point_batch(W, exts, threads=16)
    R, err, its = ssn.fixed_points_batch(W, exts, k=0.01, n=2.2)
    np.testing.assert_array_equal(err, st_o)
    tol = ATOL * np.maximum(1, it_o[..., None] / 1000.0) + RTOL * np.abs(Ro)
    assert (np.abs(R - Ro) <= tol).all(), float((np.abs(R - Ro) / tol).max())
    check_sweeps(its, it_o)


def test_find_fixed_points_streamed_variants(ssn, oracle):
    """The streamed pointer-list entry point behind find_fixed_points: float64 W per network (the reference's
    calling convention, run/gan.py:622-634), the z-based variant (jds=..., W built on the GPU, float64 or float32
    z), more networks than one staging slab, a non-zero r0, and `info.solutions` behaving like the reference's."""
    n_sites, nz = 40, 70
    jds = oracle.new_JDS()
    zs, W, exts = seeded_problem(oracle, n_sites, nz, seed=77)
    Ro, st_o, it_o = oracle.fixed_point_batch(W, exts, threads=8)
    assert (st_o == 0).all()
    kw = dict(k=0.01, n=2.2, r0=np.zeros(2 * n_sites))
    Zs, Rs, info = ssn.find_fixed_points(nz, ((zs[i], W[i]) for i in range(nz)), exts, **kw)
    np.testing.assert_array_equal(Zs, zs)
    np.testing.assert_allclose(Rs, Ro, rtol=RTOL, atol=ATOL)
    assert len(info.solutions) == nz and len(info.solutions[3]) == len(exts)
    assert all(s.success and s.message == 'Converged' for s in info.solutions[-1])
    np.testing.assert_array_equal(info.solutions[5][2].x, Rs[5, 2])
    assert info.solutions[5][2].iterations == it_o[5, 2] or abs(info.solutions[5][2].iterations - it_o[5, 2]) <= 1
    for ztype in (np.float64, np.float32):
        z_in = zs.astype(ztype)
        Wz = oracle.generate_weight(n_sites, jds['J'], jds['D'], jds['S'], z_in.astype(np.float64))
        Rz_o, _, _ = oracle.fixed_point_batch(Wz, exts, threads=8)
        Zs2, Rs2, _ = ssn.find_fixed_points(nz, ((z_in[i], None) for i in range(nz)), exts, jds=jds, host_threads=3, **kw)
        assert Zs2.dtype == ztype
        np.testing.assert_allclose(Rs2, Rz_o, rtol=RTOL, atol=ATOL)
    r0 = np.full(2 * n_sites, 2.5)
    x, code, _ = oracle.fixed_point(W[1], exts[6], r0=r0)
    _, Rs3, _ = ssn.find_fixed_points(3, ((zs[i], W[i]) for i in range(3)), exts, k=0.01, n=2.2, r0=r0)
    np.testing.assert_allclose(Rs3[1, 6], x, rtol=RTOL, atol=ATOL)
    # the float64 kernel through the same API
    _, Rs4, _ = ssn.find_fixed_points(4, ((zs[i], W[i]) for i in range(4)), exts, precise=True, **kw)
    np.testing.assert_allclose(Rs4, Ro[:4], rtol=0, atol=1e-10)


def test_initial_state_and_max_iter(ssn, oracle):
    n_sites = 20
    _, W, exts = seeded_problem(oracle, n_sites, 2, seed=1)
    r0 = np.full(2 * n_sites, 3.0)
    for precise in (False, True):
        R, err, its = ssn.fixed_points_batch(W, exts, k=0.01, n=2.2, r0=r0, max_iter=25, precise=precise)
        assert (err == 1).all() and (its == 25).all()
        # 25 Euler sweeps from r0, float64
        ref = np.empty_like(R)
        for z in range(2):
            for b in range(len(exts)):
                x, code, it = oracle.fixed_point(W[z], exts[b], r0=r0, max_iter=25)
                assert code == 1
                ref[z, b] = x
        np.testing.assert_allclose(R, ref, rtol=1e-5, atol=1e-5 if not precise else 1e-12)
    # a non-zero initial state that does converge: same fixed point, same sweep count
    Ro, st_o, it_o = oracle.fixed_point_batch(W, exts)
    x, code, it1 = oracle.fixed_point(W[1], exts[4], r0=r0)
    R, err, its = ssn.fixed_points_batch(W, exts, k=0.01, n=2.2, r0=r0)
    assert (err == 0).all() and code == 0
    np.testing.assert_allclose(R[1, 4], x, rtol=RTOL, atol=ATOL)
    assert abs(int(its[1, 4]) - it1) <= 1


def test_from_z_device_path_and_weight_kernel(ssn, oracle):
    """W built on chip from z (SSN_W_FROM_Z) equals W supplied densely; ssn_generate_weight
    equals weight_gen.generate_weight."""
    from tc_gan_b200 import clib
    from tc_gan_b200.weight_gen import generate_weight_batch_gpu
    n_sites, nz = 201, 3
    jds = oracle.new_JDS()
    zs, W, exts = seeded_problem(oracle, n_sites, nz, seed=2)
    Wg = generate_weight_batch_gpu(n_sites, jds['J'], jds['D'], jds['S'], zs)
    np.testing.assert_allclose(Wg, W, rtol=2e-6, atol=1e-7)
    dim, nb = 2 * n_sites, len(exts)
    z32 = np.ascontiguousarray(zs, np.float32)
    e32 = np.ascontiguousarray(exts, np.float32)
    R = np.empty((nz, nb, dim), np.float32)
    st = np.empty((nz, nb), np.int32)
    it = np.empty((nz, nb), np.int32)
    sv = clib.make_solver(k=0.01, n=2.2)
    clib.check_call(clib.libssnode.ssn_fixed_point_batch(
        sv, nz, nb, n_sites, clib.W_FROM_Z, z32.ctypes.data, clib.make_jds(jds['J'], jds['D'], jds['S']),
        e32.ctypes.data, 0, None, R.ctypes.data, st.ctypes.data, it.ctypes.data, 0, clib.MEM_HOST, None),
        'ssn_fixed_point_batch')
    Ro, st_o, it_o = oracle.fixed_point_batch(W, exts, threads=8)
    np.testing.assert_array_equal(st, st_o)
    np.testing.assert_allclose(R, Ro, rtol=RTOL, atol=ATOL)
    check_sweeps(it, it_o)


def test_tight_atol_on_the_fast_path(ssn, oracle):
    """The reference-point iteration converges to atol far below FP32 resolution of the rates
    (the reference's own tests use atol=1e-10, tc_gan/tests/test_ssn.py:66-74)."""
    _, W, exts = seeded_problem(oracle, 51, 2, seed=4)
    Ro, st_o, it_o = oracle.fixed_point_batch(W, exts, atol=1e-10, max_iter=100000)
    R, err, its = ssn.fixed_points_batch(W, exts, k=0.01, n=2.2, atol=1e-10, max_iter=100000)
    assert (err == 0).all() and (st_o == 0).all()
    np.testing.assert_allclose(R, Ro, rtol=1e-6, atol=1e-6)
    assert np.abs(its - it_o).max() <= 2


def test_reference_abi_from_a_thread_pool(ssn, oracle):
    """The reference drives solve_dynamics_*_euler from cpu_count() threads with the GIL released
    (tc_gan/ssnode.py:455-460): concurrent calls must not disturb each other (per-thread stream, device and
    pinned staging), and every result must equal the float64 oracle."""
    from concurrent.futures import ThreadPoolExecutor
    n_sites = 51
    _, W, exts = seeded_problem(oracle, n_sites, 4, seed=31)
    jobs = [(z, b) for z in range(len(W)) for b in range(len(exts))]
    ref = {}
    for z, b in jobs:
        x, code, it = oracle.fixed_point(W[z], exts[b])
        ref[(z, b)] = (x, code)

    def solve(job):
        z, b = job
        return job, ssn.fixed_point(W[z], exts[b], k=0.01, n=2.2)

    for rounds in range(2):
        with ThreadPoolExecutor(8) as pool:
            for job, sol in pool.map(solve, jobs):
                x, code = ref[job]
                assert sol.error == code
                np.testing.assert_allclose(sol.x, x, rtol=0, atol=1e-10)


def test_float64_streaming_fallback(ssn, oracle):
    """The float64 kernel that streams W from L2 (used when the slice does not fit in cluster shared memory;
    forced here with SSN_F64=streamed) gives the same results as the cluster-resident one."""
    _, W, exts = seeded_problem(oracle, 33, 3, seed=32)
    Ro, st_o, it_o = oracle.fixed_point_batch(W, exts, threads=8)
    os.environ['SSN_F64'] = 'streamed'
    try:
        Rp, errp, itsp = ssn.fixed_points_batch(W, exts, k=0.01, n=2.2, precise=True)
    finally:
        del os.environ['SSN_F64']
    np.testing.assert_array_equal(errp, st_o)
    np.testing.assert_array_equal(itsp, it_o)
    np.testing.assert_allclose(Rp, Ro, rtol=0, atol=1e-10)
    # a size whose double-precision slice exceeds shared memory with an 8-stimulus panel takes that path by itself
    n_sites = 220
    _, W, exts = seeded_problem(oracle, n_sites, 1, seed=33)
    Ro, st_o, it_o = oracle.fixed_point_batch(W, exts, threads=8)
    Rp, errp, itsp = ssn.fixed_points_batch(W, exts, k=0.01, n=2.2, precise=True)
    np.testing.assert_array_equal(errp, st_o)
    np.testing.assert_array_equal(itsp, it_o)
    np.testing.assert_allclose(Rp, Ro, rtol=0, atol=1e-10)


@pytest.mark.parametrize('seed', [11, 12])
def test_randomised_shapes_and_solver_settings(ssn, oracle, seed):
    """Random sizes (every cluster width), stimulus counts (partial and refilled half-panels), transfer functions,
    tolerances, iteration caps, initial states and weight scales against the float64 oracle: same codes, same
    sweep counts (to the tolerance of check_sweeps), same rates."""
    from tc_gan_b200.weight_gen import generate_weight
    rs = np.random.RandomState(seed)
    for case in range(20):
        n_sites = int(rs.choice([1, 2, 5, 13, 28, 29, 40, 56, 57, 84, 101, 130, 168, 201, 224]))
        nb = int(rs.choice([1, 2, 3, 4, 5, 7, 8, 9, 12, 13, 16, 17, 23]))
        nz = int(rs.randint(1, 4))
        io_type = str(rs.choice(['asym_tanh', 'asym_tanh', 'asym_linear', 'asym_power']))
        atol = float(rs.choice([1e-5, 1e-5, 1e-7, 3e-4]))
        max_iter = int(rs.choice([10000, 10000, 300, 57]))
        jds = oracle.new_JDS()
        J = jds['J'] * float(rs.choice([1.0, 1.0, 1.3, 0.7]))
        zs = rs.rand(nz, 2 * n_sites, 2 * n_sites)
        W = np.array([generate_weight(n_sites, J, jds['D'], jds['S'], z) for z in zs])
        exts = oracle.stimulus_input(np.sort(rs.rand(nb)), n_sites, contrasts=(float(rs.choice([5., 20., 40.])),))
        r0 = rs.rand(2 * n_sites) * 5 if rs.rand() < 0.3 else None
        kw = dict(io_type=io_type, atol=atol, max_iter=max_iter)
        if io_type != 'asym_tanh':
            kw['rate_stop_at'] = 200.0
        if r0 is None:
            Ro, st_o, it_o = oracle.fixed_point_batch(W, exts, threads=8, **kw)
        else:
            Ro = np.empty((nz, nb, 2 * n_sites)); st_o = np.empty((nz, nb), int); it_o = np.empty((nz, nb), int)
            for z in range(nz):
                for b in range(nb):
                    Ro[z, b], st_o[z, b], it_o[z, b] = oracle.fixed_point(W[z], exts[b], r0=r0, **kw)
        R, err, its = ssn.fixed_points_batch(W, exts, k=0.01, n=2.2, r0=r0, **kw)
        label = (case, n_sites, nb, nz, io_type, atol, max_iter, r0 is not None)
        np.testing.assert_array_equal(err, st_o, err_msg=str(label))
        conv = st_o == 0
        if conv.any():
            d_it = np.abs(its - it_o)[conv]
            assert (d_it <= 1 + it_o[conv] // 1000).all(), (label, int(d_it.max()))
            tol = (ATOL * np.maximum(1, it_o[..., None] / 1000.0) + RTOL * np.abs(Ro)) * max(1.0, atol / 1e-5)
            assert (np.abs(R - Ro)[conv] <= tol[conv]).all(), (label, float((np.abs(R - Ro) / tol)[conv].max()))


def test_shared_memory_fallback_kernel(ssn, oracle):
    """The cluster/DSMEM kernel that keeps W in shared memory (used beyond 2N = 448)."""
    os.environ['SSN_FORCE_SMEM_KERNEL'] = '1'
    try:
        for n_sites, nz in ((51, 3), (201, 2)):
            _, W, exts = seeded_problem(oracle, n_sites, nz, seed=9)
            Ro, st_o, it_o = oracle.fixed_point_batch(W, exts, threads=8)
            R, err, its = ssn.fixed_points_batch(W, exts, k=0.01, n=2.2)
            np.testing.assert_array_equal(err, st_o)
            np.testing.assert_allclose(R, Ro, rtol=RTOL_SMEM, atol=ATOL_SMEM)
            assert np.abs(its - it_o).max() <= 60
    finally:
        del os.environ['SSN_FORCE_SMEM_KERNEL']
    # a size only the fallback covers
    n_sites = 240
    _, W, exts = seeded_problem(oracle, n_sites, 1, seed=10)
    Ro, st_o, it_o = oracle.fixed_point_batch(W, exts, threads=8)
    R, err, its = ssn.fixed_points_batch(W, exts, k=0.01, n=2.2)
    np.testing.assert_array_equal(err, st_o)
    np.testing.assert_allclose(R, Ro, rtol=RTOL_SMEM, atol=ATOL_SMEM)


def test_full_size_fixed_point_property(ssn, oracle):
    """BASELINE config 2 at full size (1024 networks x 8 stimuli, 2N=402), through the
    device-pointer C ABI: every solve converges and every returned state is a fixed point
    of one further float64 Euler step to a few atol (size-independent property; the oracle
    needs ~35 ms per solve, so it checks a sample)."""
    import torch
    from tc_gan_b200 import clib
    n_sites, nz = 201, 1024
    dim = 2 * n_sites
    jds = oracle.new_JDS()
    exts = oracle.stimulus_input(oracle.DEFAULT_BANDWIDTHS, n_sites)
    nb = len(exts)
    dev = torch.device('cuda:0')
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    z = torch.rand((nz, dim, dim), generator=g, device=dev, dtype=torch.float32)
    e = torch.tensor(exts, dtype=torch.float32, device=dev)
    R = torch.empty((nz, nb, dim), dtype=torch.float32, device=dev)
    st = torch.empty((nz, nb), dtype=torch.int32, device=dev)
    it = torch.empty((nz, nb), dtype=torch.int32, device=dev)
    sv = clib.make_solver(k=0.01, n=2.2)
    clib.check_call(clib.libssnode.ssn_fixed_point_batch(
        sv, nz, nb, n_sites, clib.W_FROM_Z, z.data_ptr(), clib.make_jds(jds['J'], jds['D'], jds['S']),
        e.data_ptr(), 0, None, R.data_ptr(), st.data_ptr(), it.data_ptr(), 0, clib.MEM_DEVICE,
        torch.cuda.current_stream().cuda_stream), 'ssn_fixed_point_batch')
    torch.cuda.synchronize()
    st, it = st.cpu().numpy(), it.cpu().numpy()
    assert (st == 0).all() and it.min() > 50 and it.max() < 5000
    sample = [0, 1, 511, 1023]
    zs = z[sample].double().cpu().numpy()
    W = oracle.generate_weight(n_sites, jds['J'], jds['D'], jds['S'], zs)
    Rs = R[sample].double().cpu().numpy()
    eps = np.r_[np.full(n_sites, 8e-4 / 0.01589), np.full(n_sites, 8e-4 / 0.002)]
    step = (oracle.io_fun(np.einsum('zij,zbj->zbi', W, Rs) + exts[None]) - Rs) * eps
    assert np.abs(step).max() < 5e-5
    Ro, st_o, it_o = oracle.fixed_point_batch(W, exts, threads=8)
    np.testing.assert_allclose(Rs, Ro, rtol=RTOL, atol=ATOL)
    check_sweeps(it[sample], it_o)
    assert np.isfinite(R.cpu().numpy()).all()


def test_sample_tuning_curves_dataset_path(ssn, oracle):
    """truth-dataset generation through the solver, as networks/dataset.py:28-71 calls
    ssnode.sample_tuning_curves (asym_power + rate_stop_at: rejected networks are re-drawn)."""
    kw = dict(NZ=5, seed=1, N=51, io_type='asym_power', rate_stop_at=200, max_iter=3000,
              sample_sites=[12, 25, 38])
    tunings, (zs, rates, info) = ssn.sample_tuning_curves(track_offset_identity=True, **kw)
    assert tunings.shape == (3 * 8, 5) and rates.shape == (5, 8, 102) and zs.shape == (5, 102, 102)
    np.testing.assert_array_equal(tunings.T.reshape(5, 8, 3), rates[:, :, [12, 25, 38]])
    # every kept network is a genuine fixed point of the reference dynamics and was drawn in order
    rs = np.random.RandomState(1)
    P = ssn.DEFAULT_PARAMS
    exts = oracle.stimulus_input(P['bandwidths'], 51)
    kept = 0
    for _ in range(5 + info.rejections):
        z = rs.rand(1, 102, 102)[0]
        W = oracle.generate_weight(51, P['J'], P['D'], P['S'], z)
        R, st, _ = oracle.fixed_point_batch(W[None], exts, io_type='asym_power', rate_stop_at=200, max_iter=3000)
        if (st == 0).all():
            np.testing.assert_array_equal(z, zs[kept])
            np.testing.assert_allclose(rates[kept], R[0], rtol=RTOL, atol=ATOL)
            kept += 1
    assert kept == 5 and info.unused == 0

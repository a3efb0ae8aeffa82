"""
Connectivity matrices -- mirror of tc_gan/weight_gen.py (host-side numpy helper;
inside the solver W is built on chip from z, see csrc/ssn_common.cuh).
"""
import numpy


def weight(x, J, delta, sigma, z):
    """One N x N block: exp(-(x_i - x_j)^2 / (2 sigma^2)) * (J + delta z).  weight_gen.py:6-10."""
    diff = x.reshape((-1, 1)) - x.reshape((1, -1))
    return numpy.exp(-diff ** 2 / (2 * sigma ** 2)) * (J + delta * z)


def generate_weight(N, J, delta, sigma, z):
    """2N x 2N matrix; excitatory columns positive, inhibitory negative.  weight_gen.py:13-26."""
    J, delta, sigma = (numpy.asarray(a, dtype=float) for a in (J, delta, sigma))
    z = numpy.asarray(z)
    x = numpy.linspace(-0.5, 0.5, N)
    W = numpy.empty((2 * N, 2 * N))
    for a in range(2):
        for b in range(2):
            sgn = 1.0 if b == 0 else -1.0
            rows, cols = slice(a * N, (a + 1) * N), slice(b * N, (b + 1) * N)
            W[rows, cols] = weight(x, sgn * J[a, b], sgn * delta[a, b], sigma[a, b], z[rows, cols])
    return W


def generate_parameter(N, J, delta, sigma, seed=None):
    """W and its latent z ~ U[0,1)^{2N x 2N}.  weight_gen.py:29-35."""
    rs = numpy.random.RandomState(seed)
    z = rs.uniform(size=(2 * N, 2 * N))
    return generate_weight(N, J, delta, sigma, z), z


def generate_weight_batch_gpu(N, J, delta, sigma, zs):
    """W [nz, 2N, 2N] (float32) from z on the GPU (ssn_generate_weight)."""
    from . import clib
    zs = numpy.ascontiguousarray(zs, dtype=numpy.float32).reshape((-1, 2 * N, 2 * N))
    W = numpy.empty_like(zs)
    jds = clib.make_jds(J, delta, sigma)
    clib.check_call(clib.libssnode.ssn_generate_weight(
        zs.shape[0], N, zs.ctypes.data, jds, W.ctypes.data, clib.MEM_HOST, None), 'ssn_generate_weight')
    return W

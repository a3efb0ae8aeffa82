"""
ctypes binding of ``ext/libssnode.so`` -- mirror of tc_gan/clib.py.

The first block is the reference binding verbatim in meaning (clib.py:10-33):
same library name and location (``ext/`` next to this file), same three solver
symbols with the same argtypes, same four scalar helpers.  The second block
binds the batched entry points declared in include/ssnode.h.
"""
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_void_p
import ctypes
import os

import numpy

libdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ext')


def load_library(name):
    try:
        return numpy.ctypeslib.load_library(name, libdir)
    except OSError as err:
        raise ImportError(
            "tc_gan_b200: CUDA library {}/{}.so is missing or unloadable ({}). "
            "Build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C tc_gan_b200/csrc`.  There is no CPU fallback.".format(libdir, name, err))


double_ptr = ctypes.POINTER(ctypes.c_double)
float_ptr = ctypes.POINTER(ctypes.c_float)
int_ptr = ctypes.POINTER(ctypes.c_int)

# SSN_LIBNAME selects an instrumented build of the same library (`make -C tc_gan_b200/csrc prof`)
libssnode = load_library(os.environ.get('SSN_LIBNAME', 'libssnode'))

# ---- reference ABI (tc_gan/clib.py:16-33) ---------------------------------------
for fun in [libssnode.solve_dynamics_asym_power_euler,
            libssnode.solve_dynamics_asym_linear_euler,
            libssnode.solve_dynamics_asym_tanh_euler]:
    fun.argtypes = [
        c_int, double_ptr, double_ptr, c_double, c_double,
        double_ptr, double_ptr,
        c_double, c_double,
        c_double, c_int, c_double,
        c_double, c_double,
    ]
    fun.restype = ctypes.c_int

for fun in [libssnode.io_pow, libssnode.io_alin, libssnode.io_atanh]:
    fun.argtypes = [c_double] * 6
    fun.restype = c_double

libssnode.rate_to_volt.argtypes = [c_double] * 3
libssnode.rate_to_volt.restype = c_double

libssnode.dot.argtypes = [c_int, double_ptr, double_ptr]
libssnode.dot.restype = c_double

# ---- batched entry points (include/ssnode.h section 2) -------------------------------
IO_TYPES = {'asym_power': 0, 'asym_linear': 1, 'asym_tanh': 2}
MEM_HOST, MEM_DEVICE = 0, 1
W_DENSE, W_FROM_Z = 0, 1


class SolverStruct(Structure):
    _fields_ = [('io_type', c_int), ('max_iter', c_int),
                ('k', c_double), ('n', c_double),
                ('tau_E', c_double), ('tau_I', c_double), ('dt', c_double),
                ('atol', c_double),
                ('rate_soft_bound', c_double), ('rate_hard_bound', c_double)]


class JDSStruct(Structure):
    _fields_ = [('J', c_double * 4), ('D', c_double * 4), ('S', c_double * 4)]


def make_solver(io_type='asym_tanh', k=0.01, n=2.2, tau=(0.01589, 0.002), dt=.0008,
                max_iter=10000, atol=1e-5, rate_soft_bound=200., rate_hard_bound=1000.,
                rate_stop_at=float('inf')):
    if io_type not in IO_TYPES:
        raise ValueError("Unknown I/O type: {}".format(io_type))
    if io_type in ('asym_power', 'asym_linear'):        # tc_gan/ssnode.py:241-242
        rate_hard_bound = rate_stop_at
    return SolverStruct(IO_TYPES[io_type], int(max_iter), float(k), float(n),
                        float(tau[0]), float(tau[1]), float(dt), float(atol),
                        float(rate_soft_bound), float(rate_hard_bound))


def make_jds(J, D, S):
    s = JDSStruct()
    for name, arr in (('J', J), ('D', D), ('S', S)):
        flat = numpy.asarray(arr, dtype=float).reshape(4)
        setattr(s, name, (c_double * 4)(*flat))
    return s


libssnode.ssn_fixed_point_batch.argtypes = [
    POINTER(SolverStruct), c_int, c_int, c_int, c_int, c_void_p, POINTER(JDSStruct),
    c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]
libssnode.ssn_fixed_point_batch.restype = c_int

libssnode.ssn_fixed_point_batch_f64.argtypes = [
    POINTER(SolverStruct), c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
    c_void_p, c_void_p, c_void_p, c_int]
libssnode.ssn_fixed_point_batch_f64.restype = c_int

libssnode.ssn_fixed_point_batch_ptrs.argtypes = [
    POINTER(SolverStruct), c_int, c_int, c_int, c_int, c_void_p, c_int, POINTER(JDSStruct), c_void_p, c_void_p,
    c_void_p, c_void_p, c_void_p, c_int]
libssnode.ssn_fixed_point_batch_ptrs.restype = c_int
libssnode.ssn_host_gather.argtypes = [c_void_p, c_int, ctypes.c_size_t, c_void_p, c_int]
libssnode.ssn_host_gather.restype = c_int

libssnode.ssn_ift_gradient_batch.argtypes = [
    POINTER(SolverStruct), c_int, c_int, c_int, c_void_p, POINTER(JDSStruct), c_void_p, c_int,
    c_void_p, c_void_p, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]
libssnode.ssn_ift_gradient_batch.restype = c_int

libssnode.ssn_euler_forward.argtypes = [
    POINTER(SolverStruct), c_int, c_int, c_int, c_void_p, POINTER(JDSStruct), c_void_p, c_int,
    c_int, c_int, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
libssnode.ssn_euler_forward.restype = c_int

libssnode.ssn_euler_backward.argtypes = [
    POINTER(SolverStruct), c_int, c_int, c_int, c_void_p, POINTER(JDSStruct),
    c_int, c_int, c_double, c_void_p, c_double, c_double, c_void_p, c_void_p, c_void_p, c_void_p,
    c_void_p, c_void_p, c_void_p]
libssnode.ssn_euler_backward.restype = c_int

libssnode.ssn_generate_weight.argtypes = [c_int, c_int, c_void_p, POINTER(JDSStruct), c_void_p,
                                          c_int, c_void_p]
libssnode.ssn_generate_weight.restype = c_int

libssnode.ssn_probe_gather.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]
libssnode.ssn_probe_gather.restype = c_int
libssnode.ssn_probe_scatter.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]
libssnode.ssn_probe_scatter.restype = c_int

libssnode.ssn_device_count.argtypes = []
libssnode.ssn_device_count.restype = c_int
libssnode.ssn_last_error.argtypes = []
libssnode.ssn_last_error.restype = c_char_p
libssnode.ssn_kernel_launches.argtypes = []
libssnode.ssn_kernel_launches.restype = c_int
libssnode.ssn_fixed_point_occupancy.argtypes = [c_int, int_ptr, int_ptr]
libssnode.ssn_fixed_point_occupancy.restype = c_int

libssnode.ssn_bptt_param_grad.argtypes = [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, POINTER(JDSStruct),
                                          c_void_p, c_void_p]
libssnode.ssn_bptt_param_grad.restype = c_int
libssnode.ssn_traj_pitch.argtypes = [c_int]
libssnode.ssn_traj_pitch.restype = c_int
libssnode.ssn_fixed_point_kernel_name.argtypes = [c_int, ctypes.c_char_p, c_int]
libssnode.ssn_fixed_point_kernel_name.restype = c_int
libssnode.ssn_profile_enable.argtypes = [c_int]
libssnode.ssn_profile_enable.restype = c_int
libssnode.ssn_profile_read.argtypes = [ctypes.c_char_p, c_int]
libssnode.ssn_profile_read.restype = c_int

libssnode.ssn_measure_fp32_peak.argtypes = [double_ptr]
libssnode.ssn_measure_fp32_peak.restype = c_int

EXPORTED_SYMBOLS = (
    'solve_dynamics_asym_power_euler', 'solve_dynamics_asym_linear_euler',
    'solve_dynamics_asym_tanh_euler', 'dot', 'rate_to_volt', 'io_pow', 'io_alin', 'io_atanh',
    'ssn_fixed_point_batch', 'ssn_fixed_point_batch_f64', 'ssn_ift_gradient_batch',
    'ssn_euler_forward', 'ssn_euler_backward', 'ssn_generate_weight', 'ssn_device_count',
    'ssn_last_error', 'ssn_kernel_launches', 'ssn_fixed_point_occupancy',
    'ssn_measure_fp32_peak', 'ssn_profile_enable', 'ssn_profile_read', 'ssn_probe_gather', 'ssn_probe_scatter', 'ssn_fixed_point_batch_ptrs', 'ssn_host_gather', 'ssn_fixed_point_kernel_name', 'ssn_traj_pitch', 'ssn_bptt_param_grad')


class SSNLibraryError(RuntimeError):
    """A call into libssnode.so failed (CUDA error, unsupported shape, no GPU)."""


def check_call(code, what):
    if code != 0:
        msg = libssnode.ssn_last_error().decode('utf-8', 'replace')
        raise SSNLibraryError('{} failed with code {}: {}'.format(what, code, msg))


def kernel_launches():
    return int(libssnode.ssn_kernel_launches())


def fixed_point_kernel_tag(n_sites):
    buf = ctypes.create_string_buffer(256)
    check_call(libssnode.ssn_fixed_point_kernel_name(int(n_sites), buf, len(buf)), 'ssn_fixed_point_kernel_name')
    return buf.value.decode()


def profile_enable(on=True):
    """Bracket every kernel launch of the library with CUDA events on its stream (include/ssnode.h)."""
    return int(libssnode.ssn_profile_enable(int(bool(on))))


def profile_read():
    """{kernel name: (total ms, launches)} since the previous read; waits for those launches."""
    buf = ctypes.create_string_buffer(8192)
    libssnode.ssn_profile_read(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, ms, n = line.rsplit(' ', 2)
        out[name] = (float(ms), int(n))
    return out

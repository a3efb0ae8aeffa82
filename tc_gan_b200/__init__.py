"""
tc_gan_b200 -- B200-native (sm_100a) SSN simulation hot path of tc-gan.

Host-side mirror of the reference modules on that path (``tc_gan.clib``,
``tc_gan.ssnode``, ``tc_gan.weight_gen``, ``tc_gan.stimuli``,
``tc_gan.gradient_expressions``) over the CUDA library ``ext/libssnode.so``
(C ABI in ``include/ssnode.h``).  There is no CPU fallback: importing
``tc_gan_b200.clib`` fails if the library is not built, and every solver call
fails if no GPU is usable.
"""
__version__ = '0.1.0'

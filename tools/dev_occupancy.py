"""Resident clusters of the fixed-point kernel per cluster width (development)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from tc_gan_b200 import clib
for n_sites in (28, 56, 84, 112, 140, 168, 196, 201):
    cs, rc = clib.c_int(), clib.c_int()
    r = clib.libssnode.ssn_fixed_point_occupancy(n_sites, cs, rc)
    print('n_sites %d: rc %d cluster %d resident clusters %d -> %d SMs' % (n_sites, r, cs.value, rc.value, cs.value * rc.value))

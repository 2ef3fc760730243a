#!/usr/bin/env python
"""Batch smooth-particle run in the shape of pyticles' run_scripts/bspana.py (:19-62), on the B200
backend: the only change from the reference script is where the modules are imported from (and
SpamForce instead of the Fortran-only SpamComplete viscous terms).

    python examples/bspana.py [steps] [side]
"""
import os
import sys
from time import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))      # run from a checkout

import numpy as np

from pyticles_b200 import forces, neighbour_list, particles
from pyticles_b200.properties import spam_properties
from pyticles_b200.spam_nc import create_sph_ncfile, write_step

MAX_STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 50
S = int(sys.argv[2]) if len(sys.argv) > 2 else 10
NDIM = 3
XMAX = YMAX = ZMAX = S + 2
VMAX = 0.0
dt = 0.05
SPACING = 1.0
SIDE = (S, S, S)
NP = SIDE[0] * SIDE[1] * SIDE[2]
TEMPERATURE = 1.5
HLONG = 5.0
HSHORT = 2.5
ofname = 'output.nc'

particles.SPROPS = True
particles.FUSED = True
print("Initialising")
p = particles.SmoothParticleSystem(NP, maxn=NP, d=3, rinit='grid', vmax=VMAX, side=SIDE, spacing=SPACING,
                                   xmax=XMAX, ymax=YMAX, zmax=ZMAX, temperature=TEMPERATURE, hlong=HLONG,
                                   hshort=HSHORT, thermostat_temp=TEMPERATURE, thermostat=True)
nl = neighbour_list.VerletList(p, cutoff=5.0)
p.nlists.append(nl)
p.nl_default = nl
p.forces.append(forces.SpamForce(p, nl))
nl.build()
nl.separations()
spam_properties(p, nl)
create_sph_ncfile(ofname, {'name': 'Andrew', 'age': 33}, NP, NDIM)
print("STEP   seconds   pairs   mean rho   mean T")
for i in range(MAX_STEPS):
    tstart = time()
    nl.compress()                       # prune the list; sets rebuild_list when particles moved too far
    p.update(dt)
    if bool(p.r.isnan().any()):
        print('stopping due to nan')
        break
    if i % 10 == 0:
        write_step(ofname, p)
        print("%4d  %8.4f  %6d  %.5f  %.5f" % (i, time() - tstart, nl.nip, float(p.rho[:NP].mean()), float(p.t[:NP].mean())))
print('Completed', i + 1, 'steps')

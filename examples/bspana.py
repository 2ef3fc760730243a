#!/usr/bin/env python
"""pyticles' run_scripts/bspana.py (:19-62) on the B200 backend.  From `print("Initialising")` on this IS the
reference script -- SpamComplete with its default arguments, the collision force, spam_properties, the NetCDF
output, p.update(dt) with the NaN poll -- only the import lines (and Python 3 print calls) differ, plus a
command line for the number of steps and the lattice side so that tests can run it briefly.

    python examples/bspana.py [steps] [side]
"""
import os
import sys
from time import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))      # run from a checkout

import numpy as np

from pyticles_b200 import forces, neighbour_list, particles, spam_complete_force
from pyticles_b200.properties import spam_properties
from pyticles_b200.spam_nc import create_sph_ncfile, write_step

MAX_STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 50
S = int(sys.argv[2]) if len(sys.argv) > 2 else 10
NDIM = 3
XMAX = YMAX = ZMAX = S + 2                      # bspana.py:21-23 (12 for its 10^3 lattice)
VMAX = 0.0
dt = 0.05
SPACING = 1.0
LIVE_VIEW = False
SIDE = (S, S, S)
NP = SIDE[0] * SIDE[1] * SIDE[2]
TEMPERATURE = 1.5
HLONG = 5.0
HSHORT = 2.5

ofname = 'output.nc'


def initialise():                                # bspana.py's module-level hook; nothing to set up here
    pass


print("Initialising")
p = particles.SmoothParticleSystem(NP,maxn=NP,d=3,rinit='grid',vmax=VMAX
,side=SIDE,spacing=SPACING,xmax=XMAX,ymax=YMAX,zmax=ZMAX
,temperature=TEMPERATURE,hlong=HLONG,hshort=HSHORT,
thermostat_temp=TEMPERATURE,thermostat=True)
nl = neighbour_list.VerletList(p,cutoff=5.0)
p.nlists.append(nl)
p.nl_default = nl
p.forces.append(spam_complete_force.SpamComplete(p,nl))
p.forces.append(forces.FortranCollisionForce(p,nl,cutoff=0.5))
nl.build()
nl.separations()
spam_properties(p,nl)
cnt = 0
attribs = {'name':'Andrew', 'age':33}
create_sph_ncfile(ofname,attribs,NP,NDIM)
initialise()
print("STEP   INT  DERIV =  PAIR + SPAM +  FORCE   ")
for i in range(MAX_STEPS):
    tstart = time()
    p.update(dt)
    if np.isnan(p.r.cpu().numpy()).any():
        print('stopping due to nan')
        break
    if i % 10 == 0:
        write_step(ofname,p)
print('Completed',i,'steps')
print("rho %.6f .. %.6f  mean T %.6f  pairs %d" % (float(p.rho[:NP].min()), float(p.rho[:NP].max()),
                                                   float(p.t[:NP].mean()), nl.nip))

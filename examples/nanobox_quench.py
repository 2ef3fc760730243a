#!/usr/bin/env python
"""Liquid-vapour quench in the shape of pyticles' nanobox_quench.py (:57-101), as a batch run on the
B200 backend: scaled van der Waals constants (Bedaux et al. water, :72-77), short and long smoothing
lengths, the SpamComplete force with cgrad = eta = zeta = 0 as the reference script sets them (:92-93),
periodic box, thermostat.  The reference drives `p.update(dt)` from a pyglet clock and draws; this one
prints.  The only change to the set-up statements is where the modules are imported from.

    python examples/nanobox_quench.py [steps] [side] [dt]

The force arithmetic that is pinned here is the one of the reference's Python twins: the acceleration carries no mass
factor (forces.py:353-368) and the self density is W(0) whatever the mass (properties.py:76-77).  With the script's
particle mass of 0.1386 the surface particles of the lattice block then start with accelerations of 1.5e6, and the
script's dt = 0.01 (kept as the default) throws them across the box in one step -- the reference's Fortran routine,
which is not available, presumably weights by mass.  Pass a smaller dt (1e-4) for a run that stays physical.
"""
import os
import sys
from time import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))      # run from a checkout

from pyticles_b200 import box, neighbour_list, particles, spam_complete_force
from pyticles_b200.properties import spam_properties

MAX_STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 100
S = int(sys.argv[2]) if len(sys.argv) > 2 else 10

SIDE = (S, S, S)
SPACING = 0.5
XMAX = YMAX = ZMAX = 8 * S / 10.0          # nanobox_quench.py:59-61 for SIDE = (10, 10, 10)
VMAX = 0.0
dt = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01          # nanobox_quench.py:64
NP = SIDE[0] * SIDE[1] * SIDE[2]
TEMPERATURE = 0.8
HLONG = 3.0
HSHORT = 1.5
RINIT = 'grid'
ascl = 7.45e+04
bscl = 5.84e-01
kbscl = 3.29e+04
pmass = 1.386e-01


def main():
    simbox = box.PeriodicBox(xmax=XMAX, ymax=YMAX, zmax=ZMAX)
    p = particles.SmoothParticleSystem(NP, maxn=NP, d=3, rinit=RINIT, vmax=VMAX, side=SIDE, spacing=SPACING,
                                       xmax=XMAX, ymax=YMAX, zmax=ZMAX, temperature=TEMPERATURE, hlong=HLONG,
                                       hshort=HSHORT, thermostat_temp=TEMPERATURE, thermostat=True, mass=pmass,
                                       simbox=simbox)
    nl = neighbour_list.VerletList(p, cutoff=4.0)
    p.nlists.append(nl)
    p.nl_default = nl
    p.forces.append(spam_complete_force.SpamComplete(p, nl, adash=ascl, bdash=bscl, kbdash=kbscl, cgrad=0.0,
                                                     eta=0.0, zeta=0.0))
    nl.build()
    nl.separations()
    spam_properties(p, nl)
    print('initial mean temperature', float(p.t[:NP].mean()))
    print('initial mean density', float(p.rho[:NP].mean()))
    print("STEP   seconds   pairs   mean rho   min rho   max rho   mean T")
    tstart = time()
    for i in range(MAX_STEPS):
        p.update(dt)
        if bool(p.r.isnan().any()):
            print('stopping due to nan')
            return p, i
        if i % 10 == 0 or i == MAX_STEPS - 1:
            rho = p.rho[:NP]
            print("%4d  %8.3f  %7d  %.5f  %.5f  %.5f  %.5f" % (i, time() - tstart, nl.nip, float(rho.mean()),
                                                               float(rho.min()), float(rho.max()),
                                                               float(p.t[:NP].mean())))
    print('Completed', MAX_STEPS, 'steps')
    return p, MAX_STEPS


if __name__ == "__main__":
    main()

"""pyticles_b200 -- B200-native backend for the SPH step hot path of pyticles.

The modules mirror pyticles' flat module names, so a script switches backends by import:

    from pyticles_b200 import particles, neighbour_list, forces, properties
    p  = particles.SmoothParticleSystem(...)
    nl = neighbour_list.VerletList(p, cutoff=2.0)
    nl.build(); nl.separations(); properties.spam_properties(p, nl)

Compute lives in libpyticles_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/pyticles_b200.h); there is no CPU fallback and no second backend.
"""
__version__ = "0.1.0"

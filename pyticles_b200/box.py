"""Simulation boxes -- pyticles `box` surface (box.py:16-73).

The hot path only reads xmax / ymax / zmax (minimum image).  `apply(p)` runs as a CUDA
kernel (sph_box_apply) instead of a Python loop over particles.
"""
from . import backend as _backend


class Box(object):
    def __init__(self, p='none', xmax=64, ymax=48, zmax=100):
        self.p = p
        self.xmax = xmax
        self.ymax = ymax
        self.zmax = zmax

    def apply(self, p=None):
        print("Do nothing")


class PeriodicBox(Box):
    """box.py:28-47: a coordinate past a face is reset to the opposite face (not wrapped) -- the reference's
    behaviour and the default.  `wrap=True` is the true periodic image instead (x - L * floor(x / L): a particle
    that leaves through a face re-enters at the distance it overshot by), which is what conserves the pair
    geometry of a periodic run; it is NOT what the reference computes, hence a flag."""

    def __init__(self, p='none', xmax=64, ymax=48, zmax=100, wrap=False):
        Box.__init__(self, p=p, xmax=xmax, ymax=ymax, zmax=zmax)
        self.wrap = bool(wrap)

    def apply(self, p):
        _backend.box_apply((self.xmax, self.ymax, self.zmax), 2 if self.wrap else 1, p.r, p.v, p.n)


class MirrorBox(Box):
    """box.py:49-73: clamp to the face and reverse that velocity component."""

    def apply(self, p):
        _backend.box_apply((self.xmax, self.ymax, self.zmax), 0, p.r, p.v, p.n)

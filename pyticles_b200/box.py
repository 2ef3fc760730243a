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
    """box.py:28-47: a coordinate past a face is reset to the opposite face (not wrapped)."""

    def apply(self, p):
        _backend.box_apply((self.xmax, self.ymax, self.zmax), 1, p.r, p.v, p.n)


class MirrorBox(Box):
    """box.py:49-73: clamp to the face and reverse that velocity component."""

    def apply(self, p):
        _backend.box_apply((self.xmax, self.ymax, self.zmax), 0, p.r, p.v, p.n)

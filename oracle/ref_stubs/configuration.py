"""Stub for the missing `configuration` module (reference particles.py:31,134-137).
The real lattice generators are not in the reference repo, so this is a builder-defined
restatement of what their call sites imply: `n` sites of a side[0] x side[1] x side[2]
simple-cubic lattice with the given spacing, centred on `centre`, x fastest."""
import numpy as np


def grid3d(n, side, centre, spacing=1.0):
    sx, sy, sz = int(side[0]), int(side[1]), int(side[2])
    idx = np.arange(sx * sy * sz)
    i = idx % sx
    j = (idx // sx) % sy
    k = idx // (sx * sy)
    r = np.empty((idx.size, 3))
    r[:, 0] = centre[0] + (i - (sx - 1) / 2.0) * spacing
    r[:, 1] = centre[1] + (j - (sy - 1) / 2.0) * spacing
    r[:, 2] = centre[2] + (k - (sz - 1) / 2.0) * spacing
    return r[:n]


def fcc3d(n, side, centre, spacing=1.0):
    raise RuntimeError("configuration.fcc3d is not available (module absent from the reference)")

"""Initial configurations.  The reference imports a `configuration` module that is not in
its repository (particles.py:31,134-137); this is the builder-defined stand-in its call
sites imply: `n` sites of a side[0] x side[1] x side[2] simple-cubic lattice with the given
spacing, centred on `centre`, x fastest."""
import numpy as np


def grid3d(n, side, centre, spacing=1.0):
    sx, sy, sz = int(side[0]), int(side[1]), int(side[2])
    idx = np.arange(sx * sy * sz)
    r = np.empty((idx.size, 3))
    r[:, 0] = centre[0] + (idx % sx - (sx - 1) / 2.0) * spacing
    r[:, 1] = centre[1] + ((idx // sx) % sy - (sy - 1) / 2.0) * spacing
    r[:, 2] = centre[2] + (idx // (sx * sy) - (sz - 1) / 2.0) * spacing
    return r[:n]

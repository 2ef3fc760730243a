"""Smoothing kernels -- pyticles `spkernel` surface.

`lucy_kernel(r, dx, h)` is the per-pair scalar function of the reference (spkernel.py:86-118),
kept for scripts and tests that call it directly; bulk evaluation over a neighbour list is
the CUDA kernel behind `nl.wij / nl.dwij` (sph_pair_kernels) and the fused density / force
passes.  Only the Lucy kernel is on the SPH path (every reference force uses it;
spam_complete_force.py:59 pins kernel_type = 2), so `kernel()` dispatches to it alone.
"""
from math import pi

ktable = {1: 'gaussian', 2: 'lucy', 3: 'debrun'}          # spkernel.py:20


def lucy_kernel(r, dx, h):
    """The Lucy kernel: returns (w, dwdx) for 1, 2 or 3 dimensions (len(dx))."""
    try:
        dx = [float(x) for x in dx]
    except TypeError:
        dx = [float(dx)]
    dim = len(dx)
    r = float(r)
    h = float(h)
    if dim == 1:
        q = 5. / (4. * h)
    elif dim == 2:
        q = 5. / (pi * h ** 2)
    elif dim == 3:
        q = 105. / (pi * 16. * (h ** 3))
    else:
        raise ValueError("lucy_kernel: dx must have 1, 2 or 3 components")
    if r < 0:
        r = abs(r)
    if r < h:
        w = q * (1 + 3. * r / h) * ((1. - r / h)) ** 3
        if r == 0:
            dwdx = 0.0
        else:
            f = q * ((-12. / (h ** 4)) * (r ** 3) + (24. / (h ** 3)) * (r ** 2) - (12. * r / (h ** 2)))
            dwdx = [f * dx[i] / r for i in range(dim)]
    else:
        w = 0
        dwdx = [0 for _ in range(dim)]
    return w, dwdx


def lucy_w(r, h):
    """2-D Lucy kernel value only (spkernel.py:54-63)."""
    return lucy_kernel(r, (r, 0.0), h)[0]


def lucy_w3d(r, h):
    """3-D Lucy kernel value only (spkernel.py:65-83)."""
    return lucy_kernel(r, (r, 0.0, 0.0), h)[0]


def kernel(r, dx, h, type):
    """spkernel.py:22-37."""
    if type == 'lucy':
        return lucy_kernel(r, dx, h)
    raise NotImplementedError("only the Lucy kernel is on the SPH hot path (got %r)" % (type,))

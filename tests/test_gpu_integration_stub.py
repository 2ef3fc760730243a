"""GPU suite: INTEGRATION.md section 2 made executable -- the REFERENCE's own classes (oracle/_ref: pyticles'
particles.SmoothParticleSystem, neighbour_list.VerletList, forces.SpamForce, built unmodified by oracle/make_ref.py)
with the ctypes stub examples/b200_backend.py behind them, against the reference's pure-Python results on the same
inputs: pair list bit-exact, rho / p / vdot / udot within 1e-10."""
import importlib.util
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def _ref_modules():
    if not (os.path.isdir(REF) and any(f.startswith("particles.") for f in os.listdir(REF))):
        pytest.skip("oracle/_ref not built")
    sys.path.insert(0, REF)
    try:
        import forces
        import neighbour_list
        import particles
        import properties
    finally:
        sys.path.remove(REF)
    return particles, neighbour_list, properties, forces


def _system(particles, neighbour_list, r, v, m, h, t, box, cutoff, tol):
    n = r.shape[0]
    p = particles.SmoothParticleSystem(n, d=3, maxn=n, xmax=box[0], ymax=box[1], zmax=box[2], hshort=2.0, hlong=4.0)
    p.r[:, :], p.v[:, :] = r, v
    p.m[:], p.h[:], p.t[:] = m, h, t
    nl = neighbour_list.VerletList(p, cutoff=cutoff, tolerance=tol)
    return p, nl


def test_reference_classes_bound_to_the_c_abi():
    from oracle import oracle as O
    particles, neighbour_list, properties, forces = _ref_modules()
    spec = importlib.util.spec_from_file_location("b200_backend", os.path.join(ROOT, "examples", "b200_backend.py"))
    stub = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(stub)
    r, v, box = O.lattice_workload(8, 8, 8, seed=91, jitter=0.3)
    n = r.shape[0]
    rng = np.random.default_rng(4)
    m, h, t = rng.uniform(0.8, 1.2, n), np.full(n, 2.0), rng.uniform(0.8, 1.3, n)
    cutoff, tol = 2.0, 0.5
    # the reference, pure Python (fp64 base-class separations: SURVEY.md facts 5-6)
    p, nl = _system(particles, neighbour_list, r, v, m, h, t, box, cutoff, tol)
    nl.build()
    neighbour_list.NeighbourList.separations(nl)
    properties.spam_properties(p, nl)
    p.vdot[:, :] = 0.0
    p.udot[:] = 0.0
    forces.SpamForce(p, nl).apply()
    # the same classes with the C ABI behind them
    q, ql = _system(particles, neighbour_list, r, v, m, h, t, box, cutoff, tol)
    b = stub.B200(q, cutoff, tol)
    b.build(ql)
    b.spam_properties()
    q.vdot[:, :] = 0.0
    q.udot[:] = 0.0
    b.spam_force()
    assert ql.nip == nl.nip and np.array_equal(ql.iap[:ql.nip], nl.iap[:nl.nip])
    for k in ("rho", "p", "pco", "u", "vdot", "udot"):
        a, c = getattr(q, k)[:n], getattr(p, k)[:n]
        scale = np.maximum(np.abs(c), 1e-3 * np.max(np.abs(c)))
        assert np.max(np.abs(a - c) / scale) < 1e-10, k

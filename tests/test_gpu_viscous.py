"""Velocity gradient and Newtonian viscous pair force (sph_gradv / sph_viscous_force): the eta / zeta terms
of SpamComplete.  PARITY UNPINNED: the reference computes them in the Fortran sphforce3d it does not ship,
so the checks are (i) the CUDA path against the numpy statement of the same builder-defined formulas
(oracle.gradv_two_pass / newtonian_stress / viscous_force) <= 1e-10, and (ii) invariants."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu

from test_gpu_parity import _np, make_system, rel_err  # noqa: E402

RTOL = 1e-10


def _system(shape=(10, 10, 10), seed=51, jitter=0.25, vel=None, uniform_m=False):
    r, v, box = O.lattice_workload(*shape, seed=seed, jitter=jitter)
    n = r.shape[0]
    rng = np.random.default_rng(seed)
    if vel is not None:
        v = vel(r)
    else:
        v = rng.uniform(-0.5, 0.5, size=r.shape)
    m = np.ones(n) if uniform_m else rng.uniform(0.8, 1.2, n)
    h, t = np.full(n, 2.0), rng.uniform(0.8, 1.2, n)
    return r, v, m, h, t, box


def _reference(r, v, m, h, t, box, eta, zeta, cutoff=2.0, tol=0.0, fcut=5.0):
    n = r.shape[0]
    iap = O.verlet_build(r, v, box, cutoff, tol)["iap"]
    drij, rij, rsq, dv = O.separations(iap, r, v, box)
    pr = O.spam_properties(n, m, h, t, iap, rij, drij)
    g = O.gradv_two_pass(n, m, pr["rho"], iap, dv, pr["dwij"])
    pi = O.newtonian_stress(g, eta, zeta)
    vd, ud = O.viscous_force(n, m, pi, pr["rho"], iap, rij, pr["dwij"], dv, cutoff=fcut)
    return pr, g, pi, vd, ud, (iap, rij, dv)


def _gpu(r, v, m, h, t, box, cutoff=2.0, tol=0.0):
    from pyticles_b200 import neighbour_list, properties
    p = make_system(r, v, m, h, t, box)
    nl = neighbour_list.VerletList(p, cutoff=cutoff, tolerance=tol)
    nl.build()
    nl.separations()
    properties.spam_properties(p, nl)
    return p, nl


@pytest.mark.parametrize("shape,tol", [((10, 10, 10), 0.0), ((12, 8, 6), 1.0)])
def test_gradv_and_viscous_force_match_the_numpy_statement(shape, tol):
    from pyticles_b200 import properties
    r, v, m, h, t, box = _system(shape)
    n = r.shape[0]
    eta, zeta = 0.7, 0.3
    pr, g, pi, vd, ud, _ = _reference(r, v, m, h, t, box, eta, zeta, tol=tol)
    p, nl = _gpu(r, v, m, h, t, box, tol=tol)
    properties.spam_gradv(p, nl)
    assert rel_err(_np(p.gradv)[:n], g) < RTOL
    p.vdot[:, :] = 0.0
    p.udot[:] = 0.0
    nl.backend.viscous_force(p.gradv, p.rho, eta, zeta, p.h, True, 5.0, p.vdot, p.udot)
    assert rel_err(_np(p.vdot)[:n], vd) < RTOL
    assert rel_err(_np(p.udot)[:n], ud) < RTOL


def test_non_uniform_smoothing_length():
    from pyticles_b200 import properties
    r, v, m, h, t, box = _system((9, 9, 9), seed=53)
    n = r.shape[0]
    h = np.random.default_rng(3).uniform(1.6, 2.0, n)
    pr, g, pi, vd, ud, _ = _reference(r, v, m, h, t, box, 1.0, 0.1)
    p, nl = _gpu(r, v, m, h, t, box)
    properties.spam_gradv(p, nl)
    assert rel_err(_np(p.gradv)[:n], g) < RTOL
    p.vdot[:, :] = 0.0
    p.udot[:] = 0.0
    nl.backend.viscous_force(p.gradv, p.rho, 1.0, 0.1, p.h, False, 5.0, p.vdot, p.udot)
    assert rel_err(_np(p.vdot)[:n], vd) < RTOL and rel_err(_np(p.udot)[:n], ud) < RTOL


def test_uniform_translation_has_no_velocity_gradient():
    from pyticles_b200 import properties
    r, v, m, h, t, box = _system(vel=lambda r: np.tile(np.array([0.3, -0.2, 0.1]), (r.shape[0], 1)))
    p, nl = _gpu(r, v, m, h, t, box)
    properties.spam_gradv(p, nl)
    assert float(p.gradv.abs().max()) == 0.0


def test_linear_velocity_field_is_recovered_in_the_bulk():
    """v = A r on an unjittered lattice: sum_j (m/rho)(v_j - v_i) x dW = A . M with M = sum (m/rho) dr x dW, the
    same symmetric matrix for every bulk particle (-> -I in the continuum limit).  So gradv_i = A . M exactly,
    and a rigid rotation (A antisymmetric, M = c I on the cubic lattice) gives no Newtonian stress."""
    from pyticles_b200 import properties
    A = np.array([[0.0, 0.02, -0.01], [-0.02, 0.0, 0.03], [0.01, -0.03, 0.0]])       # rigid rotation
    r, v, m, h, t, box = _system((12, 12, 12), jitter=0.0, vel=lambda r: r @ A.T, uniform_m=True)
    n = r.shape[0]
    p, nl = _gpu(r, v, m, h, t, box)
    properties.spam_gradv(p, nl)
    bulk = np.all((r > 3.0) & (r < np.array(box) - 3.0), axis=1)     # away from the periodic seam of v
    g = _np(p.gradv)[:n][bulk]
    c = g[0, 0, 1] / A[0, 1]
    assert -1.2 < c < -0.8                                            # M = c I, c -> -1
    assert np.allclose(g, c * A, rtol=0, atol=1e-12)
    pi = O.newtonian_stress(g, 1.0, 0.1)
    assert np.abs(pi).max() < 1e-12


def test_viscous_force_is_pairwise_antisymmetric_and_dissipates():
    """Equal masses: sum_i vdot_i = 0 (each pair adds +a and -a).  A shear wave heats the fluid:
    sum_i m_i udot_i > 0 from the viscous term alone."""
    box_l = 12
    r, v, m, h, t, box = _system((box_l,) * 3, jitter=0.05, uniform_m=True,
                                 vel=lambda r: np.stack([0.1 * np.sin(2 * np.pi * r[:, 1] / box_l),
                                                         np.zeros(r.shape[0]), np.zeros(r.shape[0])], axis=1))
    n = r.shape[0]
    from pyticles_b200 import properties
    p, nl = _gpu(r, v, m, h, t, box)
    properties.spam_gradv(p, nl)
    p.vdot[:, :] = 0.0
    p.udot[:] = 0.0
    nl.backend.viscous_force(p.gradv, p.rho, 1.0, 0.0, p.h, True, 5.0, p.vdot, p.udot)
    vd, ud = _np(p.vdot)[:n], _np(p.udot)[:n]
    assert np.abs(vd.sum(axis=0)).max() < 1e-12 * np.abs(vd).sum()
    assert ud.sum() > 0.0
    # the force opposes the shear: it decelerates the wave
    assert (vd[:, 0] * v[:, 0]).sum() < 0.0


def test_spam_complete_with_viscosity():
    """SpamComplete(eta, zeta) = its pinned pressure part + the viscous part above; P = (p + pco) I + pi.  The
    class takes T from the integrated internal energy (spam_complete_force.py:134,151-152), so u is set to the
    vdW energy of the test temperature first -- what spam_properties leaves behind in the reference's scripts."""
    from pyticles_b200 import neighbour_list, spam_complete_force
    r, v, m, h, t, box = _system((10, 10, 10), seed=57)
    n = r.shape[0]
    p = make_system(r, v, m, h, t, box)
    p.hlr[:] = 3.0
    nl = neighbour_list.VerletList(p, cutoff=3.0, tolerance=0.0)
    nl.build()
    nl.separations()
    pr, g, pi, vd, ud, _ = _reference(r, v, m, h, t, box, 0.7, 0.3, cutoff=3.0, fcut=10.0)
    p.u[0:n] = t * 1.0 - 2.0 * pr["rho"]                               # properties.py:46
    f0 = spam_complete_force.SpamComplete(p, nl, cgrad=0.0, eta=0.0, zeta=0.0, cutoff=10.0)
    f0.apply()
    base_v, base_u = _np(p.vdot)[:n].copy(), _np(p.udot)[:n].copy()
    f1 = spam_complete_force.SpamComplete(p, nl, cgrad=0.0, eta=0.7, zeta=0.3, cutoff=10.0)
    f1.apply()
    assert rel_err(_np(p.vdot)[:n] - base_v, vd) < 1e-9
    assert rel_err(_np(p.udot)[:n] - base_u, ud) < 1e-9
    P = _np(p.P)[:n]
    want = (_np(p.p)[:n] + _np(p.pco)[:n])[:, None, None] * np.eye(3) + pi
    assert rel_err(P, want) < RTOL


def test_nanobox_quench_example_runs(capsys):
    """examples/nanobox_quench.py (the reference's showcase set-up, nanobox_quench.py:57-101) for a few
    steps: finite state, positive densities, thermostat holds the temperature."""
    import importlib.util
    import os
    import sys
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "nanobox_quench.py")
    spec = importlib.util.spec_from_file_location("nanobox_quench_example", path)
    argv = sys.argv
    sys.argv = [path, "12", "8", "1e-4"]          # see the example's docstring for the time step
    try:
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        p, steps = mod.main()
    finally:
        sys.argv = argv
    assert steps == 12
    n = p.n
    assert bool(torch.isfinite(p.r[:n]).all()) and bool(torch.isfinite(p.v[:n]).all())
    assert float(p.rho[:n].min()) > 0.0
    assert abs(float(p.t[:n].mean()) - 0.8) < 0.05

"""GPU parity suite: the CUDA path (through the C ABI, driven by the pyticles-shaped Python
API) against the golden vectors of the reference and against the CPU oracle.

Bars (BASELINE.json north_star): neighbour-pair set bit-exact; density, pressure, forces
within 1e-10 relative (fp64).  Force components cancel to ~0 on near-lattice inputs, so
vdot/udot are normalised by max(|x_i|, eps*max|x|) (SURVEY.md section 8a note).
"""
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle as C
from oracle import oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-10
CASES = ["sheet_400", "cube_216", "cube_729", "gas_500"]


def _np(t):
    return t.detach().cpu().numpy()


def rel_err(a, b):
    """max |a-b| / max(|b_i|, 1e-3 * max|b|)."""
    a, b = np.asarray(a), np.asarray(b)
    if b.size == 0:
        return 0.0
    scale = np.maximum(np.abs(b), 1e-3 * max(np.max(np.abs(b)), 1e-300))
    return float(np.max(np.abs(a - b) / scale))


def make_system(r, v, m, h, t, box, maxn=None):
    from pyticles_b200 import particles
    n = r.shape[0]
    p = particles.SmoothParticleSystem(n, d=3, maxn=maxn or n, xmax=box[0], ymax=box[1], zmax=box[2])
    p.r[0:n, :] = r
    p.v[0:n, :] = v
    p.m[0:n] = m
    p.h[0:n] = h
    p.t[0:n] = t
    return p


def run_step(p, cutoff, tolerance, fcutoff):
    from pyticles_b200 import forces, neighbour_list, properties
    nl = neighbour_list.VerletList(p, cutoff=cutoff, tolerance=tolerance)
    nl.build()
    nl.separations()
    properties.spam_properties(p, nl)
    p.vdot[:, :] = 0.0
    p.udot[:] = 0.0
    f = forces.SpamForce(p, nl, cutoff=fcutoff)
    f.apply()
    return nl


def check_against(p, nl, ref, n, pair_arrays=True):
    iap = _np(nl.iap).astype(np.int64)
    assert iap.shape == ref["iap"].shape, (iap.shape, ref["iap"].shape)
    assert np.array_equal(iap, ref["iap"].astype(np.int64))          # bit-exact set and order
    assert nl.nip == ref["iap"].shape[0]
    if pair_arrays:
        assert np.array_equal(_np(nl.drij), ref["drij"])
        assert np.array_equal(_np(nl.dv), ref["dv"])
        assert rel_err(_np(nl.rij), ref["rij"]) < 1e-14
        assert rel_err(_np(nl.wij), ref["wij"]) < RTOL
        assert rel_err(_np(nl.dwij), ref["dwij"]) < RTOL
    for name in ("rho", "p", "pco", "u"):
        assert rel_err(_np(getattr(p, name))[:n], ref[name][:n]) < RTOL, name
    assert rel_err(_np(p.vdot)[:n], ref["vdot"][:n]) < RTOL
    assert rel_err(_np(p.udot)[:n], ref["udot"][:n]) < RTOL


@pytest.mark.parametrize("name", CASES)
def test_golden_reference_vectors(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    box = tuple(float(x) for x in g["box"])
    p = make_system(g["r"], g["v"], g["m"], g["h"], g["t_in"], box)
    nl = run_step(p, float(g["cutoff"]), float(g["tolerance"]), float(g["fcutoff"]))
    ref = {k: g[k] for k in g.files}
    check_against(p, nl, ref, g["r"].shape[0])
    assert rel_err(_np(p.t), g["t_out"]) < RTOL
    assert rel_err(_np(nl.rsq), g["rsq_build"]) < 1e-14


@pytest.mark.parametrize("shape,cutoff,tol,jitter", [
    ((32, 32, 32), 2.0, 0.0, 0.1),       # C3-like cube (fast fp32 pre-filter path, interior skip)
    ((24, 24, 24), 2.0, 1.0, 0.45),      # heavy jitter, default Verlet tolerance
    ((128, 128, 1), 2.0, 0.0, 0.1),      # C2-like sheet in a deep box (coarsened z cells)
    ((40, 9, 5), 2.5, 0.5, 0.3),         # anisotropic, 2 cells across z
    ((96, 96, 96), 2.0, 0.0, 0.1),       # 47 cell layers = 6 Morton blocks per dimension: interior block faces,
                                         # multiply-high block-coordinate division (bench grid: 16 per dimension)
    ((104, 72, 56), 2.0, 1.0, 0.3),      # 46 x 32 x 25 layers = 6 x 4 x 4 blocks, default Verlet tolerance
])
def test_lattice_against_c_oracle(shape, cutoff, tol, jitter):
    r, v, box = O.lattice_workload(*shape, seed=11, jitter=jitter)
    if shape[2] == 1:
        box = (box[0], box[1], float(shape[0]))
    n = r.shape[0]
    m, h, t = np.ones(n), np.full(n, 2.0), np.ones(n)
    ref = C.sph_step(r, v, m, h, t, np.array(box), cutoff, tol, 5.0)
    p = make_system(r, v, m, h, t, box)
    nl = run_step(p, cutoff, tol, 5.0)
    check_against(p, nl, ref, n)


def test_positions_outside_box_take_exact_path():
    rng = np.random.default_rng(3)
    n = 3000
    box = (14.0, 12.0, 10.0)
    r = rng.random((n, 3)) * np.array(box)
    r[:200] += rng.uniform(-30.0, 30.0, size=(200, 3))      # far outside: single-shift semantics matter
    v = rng.normal(size=(n, 3))
    m, h, t = rng.uniform(0.5, 1.5, n), rng.uniform(1.5, 2.2, n), rng.uniform(0.5, 1.5, n)
    ref = C.sph_step(r, v, m, h, t, np.array(box), 2.0, 1.0, 5.0)
    p = make_system(r, v, m, h, t, box)
    nl = run_step(p, 2.0, 1.0, 5.0)
    from pyticles_b200 import _lib
    assert nl.backend.status().flags & _lib.SPH_F_OUT_OF_RANGE
    check_against(p, nl, ref, n)


def test_reference_neighbour_list_tests():
    """test/neighbour_list_test.py:8-53, same statements against this backend."""
    from pyticles_b200 import neighbour_list, particles
    n = 3
    p = particles.ParticleSystem(n, d=3, maxn=5)
    p.r[0, :] = (0.0, 0.0, 0.0)
    p.r[1, :] = (1.0, 0.0, 0.0)
    p.r[2, :] = (0.0, 0.0, 1.0)
    nl = neighbour_list.NeighbourList(p)
    nl.build()
    nl.separations()
    k = nl.find_pair(0, 1)
    assert float(nl.rij[k]) == 1.0
    assert nl.nip == 3

    nl = neighbour_list.VerletList(p, cutoff=10, tolerance=2)
    nl.build()
    nl.compress()
    nl.separations()
    assert float(nl.rij[0]) == 1.0
    nl.ponder_rebuild()
    assert nl.rebuild_list is False
    p.r[0, :] = (100.0, 100.0, 100.0)
    nl.ponder_rebuild()
    assert nl.rebuild_list is True


def test_reference_force_test_known_answer():
    """test/force_test.py:10-25 + the two-particle values of SURVEY.md section 8c."""
    from pyticles_b200 import forces, neighbour_list, particles, properties
    p = particles.SmoothParticleSystem(2, d=3, maxn=5)
    p.r[0, :] = (0.0, 0.0, 0.0)
    p.r[1, :] = (1.0, 0.0, 0.0)
    nl = neighbour_list.VerletList(p, cutoff=10, tolerance=2)
    nl.build()
    nl.compress()
    nl.separations()
    properties.spam_properties(p, nl)
    f = forces.SpamForce(p, nl)
    f.apply()
    assert float(p.rho[0]) == pytest.approx(2.088908628081126, rel=1e-13)
    assert float(p.p[1]) == pytest.approx(-46.990009252534314, rel=1e-12)
    assert float(p.pco[0]) == pytest.approx(-8.727078512943546, rel=1e-12)
    assert float(p.u[0]) == pytest.approx(-3.1778172561622524, rel=1e-12)
    assert float(p.vdot[:2].abs().max()) == 0.0


def test_list_maintenance(golden_dir):
    from pyticles_b200 import neighbour_list
    g = np.load(os.path.join(golden_dir, "maintain_343.npz"))
    box = tuple(float(x) for x in g["box"])
    n = g["r0"].shape[0]
    p = make_system(g["r0"], g["v"], np.ones(n), np.full(n, 2.0), np.ones(n), box)
    nl = neighbour_list.VerletList(p, cutoff=float(g["cutoff"]), tolerance=float(g["tolerance"]))
    nl.build()
    assert np.array_equal(_np(nl.iap), g["iap_build"])
    p.r[0:n, :] = g["r1"]
    nl.compress()
    assert np.array_equal(_np(nl.iap), g["iap_compress"])
    assert nl.rebuild_list == bool(g["rebuild"])
    # separations of the kept list at the moved positions
    d = O.separations(g["iap_compress"].astype(np.int64), g["r1"], g["v"], box)
    assert np.array_equal(_np(nl.drij), d[0])
    # density on the kept (stale) list == oracle on the same list
    from pyticles_b200 import properties
    properties.spam_properties(p, nl)
    pr = O.spam_properties(n, np.ones(n), np.full(n, 2.0), np.ones(n), g["iap_compress"].astype(np.int64), d[1], d[0])
    assert rel_err(_np(p.rho), pr["rho"]) < RTOL


def test_empty_single_and_spare_capacity():
    from pyticles_b200 import forces, neighbour_list, particles, properties
    for n in (0, 1):
        p = particles.SmoothParticleSystem(n, d=3, maxn=4, xmax=5., ymax=5., zmax=5.)
        nl = neighbour_list.VerletList(p, cutoff=2.0, tolerance=1.0)
        nl.build()
        nl.separations()
        assert nl.nip == 0 and tuple(nl.iap.shape) == (0, 2)
        properties.spam_properties(p, nl)
        forces.SpamForce(p, nl).apply()
        if n == 1:
            assert float(p.rho[0]) == pytest.approx(O.lucy_kernel(0.0, (0., 0., 0.), 1.0)[0], rel=1e-14)
    # maxn > n: spare slots follow properties.py:119-120 (u = t*kb, t unchanged) and stay out of the list
    r, v, box = O.lattice_workload(6, 6, 6, seed=5)
    n = r.shape[0]
    p = make_system(r, v, np.ones(n), np.full(n, 2.0), np.ones(n), box, maxn=n + 50)
    nl = run_step(p, 2.0, 1.0, 5.0)
    ref = O.sph_step(r, v, np.ones(n), np.full(n, 2.0), np.ones(n), box, 2.0, 1.0, 5.0)
    check_against(p, nl, ref, n)
    assert float(p.rho[n:].abs().max()) == 0.0
    assert float(p.vdot[n:].abs().max()) == 0.0


def test_neighbour_capacity_overflow_grows():
    from pyticles_b200 import neighbour_list
    r, v, box = O.lattice_workload(12, 12, 12, seed=9)
    n = r.shape[0]
    p = make_system(r, v, np.ones(n), np.full(n, 2.0), np.ones(n), box)
    nl = neighbour_list.VerletList(p, cutoff=2.0, tolerance=1.0, max_nbrs=8)     # far too small
    nl.build()
    ref = C.build_pairs(r, np.array(box), 2.0, 1.0)
    assert nl.backend.K > 8
    assert np.array_equal(_np(nl.iap), ref)


def test_cohesive_and_2d_force_variants():
    from pyticles_b200 import forces, neighbour_list, properties
    r, v, box = O.lattice_workload(10, 10, 10, seed=13, jitter=0.3)
    n = r.shape[0]
    rng = np.random.default_rng(1)
    m, h, t = rng.uniform(0.8, 1.2, n), np.full(n, 2.0), rng.uniform(0.8, 1.2, n)
    p = make_system(r, v, m, h, t, box)
    p.hlr[:] = 3.0
    nl = neighbour_list.VerletList(p, cutoff=3.0, tolerance=0.0)
    nl.build()
    nl.separations()
    properties.spam_properties_ls(p, nl)
    iap = _np(nl.iap).astype(np.int64)
    drij, rij, rsq, dv = O.separations(iap, r, v, box)
    pr = O.spam_properties(n, m, h, t, iap, rij, drij)
    hl = np.full(n, 3.0)
    w_lr, dw_lr = O.lucy_kernel_pairs(rij, drij, hl[iap[:, 0]])
    rho_lr = np.full(n, O.lucy_kernel(0.0, (0., 0., 0.), 3.0)[0])
    O._scatter_pairs(rho_lr, iap, w_lr * m[iap[:, 1]], w_lr * m[iap[:, 0]])
    assert rel_err(_np(p.rho_lr), rho_lr) < RTOL
    assert rel_err(_np(nl.dwij_lr), dw_lr) < RTOL
    # cohesive force (forces.py:371-405)
    p.vdot[:, :] = 0.0
    p.udot[:] = 0.0
    forces.CohesiveSpamForce(p, nl, cutoff=10.0).apply()
    vd, ud = O.spam_force(n, m, pr["pco"], rho_lr, iap, rij, dw_lr, dv, cutoff=10.0)
    assert rel_err(_np(p.vdot), vd) < RTOL and rel_err(_np(p.udot), ud) < RTOL
    # 2-D repulsive force (forces.py:246-274), stacked on top: accumulation semantics
    forces.SpamForce2d(p, nl, cutoff=5.0).apply()
    vd2, ud2 = O.spam_force(n, m, pr["p"], pr["rho"], iap, rij, pr["dwij"], dv, cutoff=5.0, dim=2,
                            vdot=vd.copy(), udot=ud.copy())
    assert rel_err(_np(p.vdot), vd2) < RTOL and rel_err(_np(p.udot), ud2) < RTOL


def test_sorted_verlet_list():
    from pyticles_b200 import neighbour_list
    r, v, box = O.lattice_workload(8, 8, 8, seed=17, jitter=0.3)
    n = r.shape[0]
    p = make_system(r, v, np.ones(n), np.full(n, 2.0), np.ones(n), box)
    nl = neighbour_list.SortedVerletList(p, cutoff=2.0, tolerance=1.0)
    nl.build()
    rsq = _np(nl.rsq)
    assert np.all(np.diff(rsq) <= 0)                      # descending (neighbour_list.py:276-277)
    ref = O.verlet_build(r, v, box, 2.0, 1.0)["iap"]
    got = _np(nl.iap).astype(np.int64)
    assert np.array_equal(got[np.lexsort((got[:, 1], got[:, 0]))], ref)
    d = O.separations(got, r, v, box)
    assert np.array_equal(_np(nl.drij), d[0])


def test_update_step_matches_oracle_integrator():
    """One imp_euler update (particles.py:459-494, integrator.py:44-59) with the list rebuilt each
    evaluation, against the same two-stage update done with the oracle."""
    from pyticles_b200 import forces, neighbour_list, particles, properties
    r, v, box = O.lattice_workload(10, 10, 10, seed=21, jitter=0.1)
    n = r.shape[0]
    m, h, t = np.ones(n), np.full(n, 2.0), np.ones(n)
    particles.SPROPS = True
    try:
        from pyticles_b200 import box as boxmod
        p = particles.SmoothParticleSystem(n, d=3, maxn=n, xmax=box[0], ymax=box[1], zmax=box[2],
                                           integrator='ieuler', simbox=boxmod.PeriodicBox(xmax=box[0], ymax=box[1], zmax=box[2]))
        p.r[0:n, :] = r
        p.v[0:n, :] = v
        p.m[:] = 1.0
        p.h[:] = 2.0
        p.t[:] = 1.0
        nl = neighbour_list.VerletList(p, cutoff=2.0, tolerance=1.0)
        p.nlists.append(nl)
        p.nl_default = nl
        p.forces.append(forces.SpamForce(p, nl))
        nl.build()
        nl.separations()
        properties.spam_properties(p, nl)
        dt = 0.01
        p.update(dt)
    finally:
        particles.SPROPS = False
    # oracle: same list (built once, reference behaviour), two derivative evaluations
    iap = O.verlet_build(r, v, box, 2.0, 1.0)["iap"]

    def deriv(rr, vv, tt):
        o = O.sph_step(rr, vv, m, h, tt, box, 2.0, 1.0, 5.0, pairs=iap)
        return o

    o1 = deriv(r, v, t)
    o1b = deriv(r, v, o1["t"])                       # update() calls derivatives() once before step()
    r1 = r + v * dt
    v1 = v + o1b["vdot"] * dt
    o2 = deriv(r1, v1, o1b["t"])
    r2 = r + (v * dt + v1 * dt) / 2
    v2 = v + (o1b["vdot"] * dt + o2["vdot"] * dt) / 2
    assert rel_err(_np(p.v)[:n], v2) < 1e-9
    assert rel_err(_np(p.r)[:n], np.where(r2 > np.array(box), 0.0, np.where(r2 < 0, np.array(box), r2))) < 1e-12


def test_c1_trajectory_matches_reference(golden_dir):
    """BASELINE config 1: the reference's own 20x20x1 sheet run -- list built
    once, SPROPS properties + SpamForce, imp_euler, MirrorBox -- against this backend driven by the
    same statements.  <= 1e-10 after step 1; the bound is relaxed as the trajectory diverges
    (<= 1e-8 at step 5, <= 1e-6 at step 20)."""
    from pyticles_b200 import forces, neighbour_list, particles, properties
    g = np.load(os.path.join(golden_dir, "c1_trajectory.npz"))
    r, v = g["r0"], g["v0"]
    n = r.shape[0]
    particles.SPROPS = True
    try:
        p = particles.SmoothParticleSystem(n, d=3, maxn=n, xmax=20., ymax=20., zmax=20., hshort=2.0, hlong=4.0,
                                           integrator='ieuler')
        p.r[:, :] = r
        p.v[:, :] = v
        nl = neighbour_list.VerletList(p, cutoff=2.0, tolerance=1.0)
        p.nlists.append(nl)
        p.nl_default = nl
        p.forces.append(forces.SpamForce(p, nl))
        nl.build()
        nl.separations()
        properties.spam_properties(p, nl)
        tol = {1: 1e-10, 5: 1e-8, 20: 1e-6}
        for step in range(1, 21):
            p.update(float(g["dt"]))
            if step in tol:
                for name in ("r", "v", "rho", "u"):
                    assert rel_err(_np(getattr(p, name)), g["%s%d" % (name, step)]) < tol[step], (name, step)
    finally:
        particles.SPROPS = False


def test_spam_conduction(golden_dir):
    """c_forces.SpamConduction (c_forces.pyx:196-239): <= 1e-10 against the fp64 oracle, and within the
    float32 temporaries of the compiled Cython original (golden from the reference itself)."""
    from pyticles_b200 import forces, neighbour_list, properties
    g = np.load(os.path.join(golden_dir, "cube_729.npz"))
    c = np.load(os.path.join(golden_dir, "conduction_729.npz"))
    box = tuple(float(x) for x in g["box"])
    n = g["r"].shape[0]
    p = make_system(g["r"], g["v"], g["m"], g["h"], g["t_in"], box)
    nl = neighbour_list.VerletList(p, cutoff=float(g["cutoff"]), tolerance=float(g["tolerance"]))
    nl.build()
    nl.separations()
    properties.spam_properties(p, nl)
    p.jq[0:n, :] = c["jq"]
    p.udot[:] = 0.0
    forces.SpamConduction(p, nl).apply()
    ref = O.spam_conduction(n, g["m"], c["jq"], g["rho"], g["iap"].astype(np.int64), g["dwij"])
    assert rel_err(_np(p.udot), ref) < RTOL
    assert rel_err(_np(p.udot), c["udot"]) < 1e-5


@pytest.mark.parametrize("stepper", ["ieuler", "rk4", "euler"])
def test_fused_steppers_equal_generic_path(stepper):
    """particles.FUSED: the device-resident improved Euler / RK4 / Euler must leave the same state as the generic
    callback-driven integrators (integrator.py:14-95) after several updates."""
    from pyticles_b200 import forces, neighbour_list, particles, properties
    r, v, box = O.lattice_workload(12, 12, 12, seed=23, jitter=0.2)
    n = r.shape[0]
    states = []
    particles.SPROPS = True
    try:
        for fused in (False, True):
            particles.FUSED = fused
            p = particles.SmoothParticleSystem(n, d=3, maxn=n + 7, xmax=box[0], ymax=box[1], zmax=box[2], hshort=2.0,
                                               integrator=stepper)
            p.r[0:n, :] = r
            p.v[0:n, :] = v
            nl = neighbour_list.VerletList(p, cutoff=2.0, tolerance=1.0)
            p.nlists.append(nl)
            p.nl_default = nl
            p.forces.append(forces.SpamForce(p, nl))
            nl.build()
            nl.separations()
            properties.spam_properties(p, nl)
            for _ in range(4):
                p.update(0.02)
            states.append({k: _np(getattr(p, k))[:n].copy() for k in ("r", "v", "u", "rho", "p", "pco")})
    finally:
        particles.FUSED = False
        particles.SPROPS = False
    for k in states[0]:
        assert rel_err(states[1][k], states[0][k]) < 1e-12, k

"""CPU suite: the oracle (numpy + C restatements) against the reference's own answers.

* tests/golden/*.npz were produced by the reference itself (tests/golden/make_golden.py).
* The four assertions of the reference's test/neighbour_list_test.py:24,46,49,53 are re-run
  against the oracle.
"""
import os

import numpy as np
import pytest

from oracle import c_oracle as C
from oracle import oracle as O

CASES = ["sheet_400", "cube_216", "cube_729", "gas_500"]
FIELDS = ["drij", "rij", "dv", "wij", "dwij", "rho", "p", "pco", "u", "vdot", "udot"]


def _load(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    return g, tuple(float(x) for x in g["box"])


def _close(a, b, rtol=1e-12):
    scale = max(np.max(np.abs(b)), 1e-300) if b.size else 1.0
    return np.max(np.abs(a - b)) <= rtol * scale if b.size else True


@pytest.mark.parametrize("name", CASES)
def test_numpy_oracle_matches_reference(golden_dir, name):
    g, box = _load(golden_dir, name)
    out = O.sph_step(g["r"], g["v"], g["m"], g["h"], g["t_in"], box, float(g["cutoff"]),
                     float(g["tolerance"]), float(g["fcutoff"]))
    assert np.array_equal(out["iap"], g["iap"])            # pair set AND order, bit-exact
    assert np.array_equal(out["drij"], g["drij"])
    assert np.array_equal(out["dv"], g["dv"])
    for k in FIELDS:
        assert _close(out[k], g[k]), k
    assert _close(out["t"], g["t_out"])
    assert _close(out["rsq"], g["rsq_build"])


@pytest.mark.parametrize("name", CASES)
def test_c_oracle_matches_reference(golden_dir, name):
    g, box = _load(golden_dir, name)
    out = C.sph_step(g["r"], g["v"], g["m"], g["h"], g["t_in"], np.array(box), float(g["cutoff"]),
                     float(g["tolerance"]), float(g["fcutoff"]))
    assert np.array_equal(out["iap"], g["iap"])
    assert np.array_equal(out["drij"], g["drij"])
    for k in FIELDS:
        assert _close(out[k], g[k]), k


def test_kernel_known_answers(golden_dir):
    k = np.load(os.path.join(golden_dir, "kat.npz"))
    for name in k.files:
        if not name.startswith("lucy_"):
            continue
        r, h, dim, dx0, dx1, dx2, w, g0, g1, g2 = k[name]
        dx = (dx0, dx1, dx2)[:int(dim)]
        ow, og = O.lucy_kernel(r, dx, h)
        assert ow == pytest.approx(w, rel=1e-14, abs=1e-300)
        assert np.allclose(og, (g0, g1, g2)[:int(dim)], rtol=1e-14, atol=1e-300)
    # the survey's pinned numbers
    assert O.lucy_kernel(0.0, (0., 0., 0.), 2.0)[0] == pytest.approx(0.2611135785101408, rel=1e-15)
    assert O.lucy_kernel(1.0, (1., 0., 0.), 2.0)[0] == pytest.approx(0.081597993284419, rel=1e-15)
    assert np.array_equal(np.array(O.vdw(1.0, 1.0)), k["vdw_1_1"])
    assert np.array_equal(np.array(O.vdw(0.5, 1.5)), k["vdw_05_15"])
    assert O.vdw_energy(1.0, 5.0) == float(k["vdw_energy_1_5"]) == 3.0
    assert O.vdw_temp(1.0, 3.0) == float(k["vdw_temp_1_3"]) == 5.0


def test_reference_neighbour_list_assertions():
    """test/neighbour_list_test.py:8-53 restated on the oracle."""
    r = np.array([[0., 0., 0.], [1., 0., 0.], [0., 0., 1.]])
    v = np.zeros((3, 3))
    box = (5., 5., 1.)                                      # particles.py:46-48 defaults
    iap = O.brute_build(3)
    drij, rij, rsq, dv = O.separations(iap, r, v, box)
    k = int(np.nonzero((iap[:, 0] == 0) & (iap[:, 1] == 1))[0][0])
    assert rij[k] == 1.0                                    # :24
    b = O.verlet_build(r, v, box, cutoff=10, tolerance=2)
    iap = O.compress(b["iap"], r, v, box, 10, 2)
    drij, rij, rsq, dv = O.separations(iap, r, v, box)
    assert rij[0] == 1.0                                    # :46
    assert O.ponder_rebuild(r, r, 2) is False               # :49
    r2 = r.copy()
    r2[0] = (100., 100., 100.)
    assert O.ponder_rebuild(r, r2, 2) is True               # :53
    assert C.ponder_rebuild(r, r2, 2) is True and C.ponder_rebuild(r, r, 2) is False


def test_two_particle_known_answer():
    """2 particles at distance 1, h = m = T = 1."""
    r = np.array([[0., 0., 0.], [1., 0., 0.]])
    out = O.sph_step(r, np.zeros((2, 3)), np.ones(2), np.ones(2), np.ones(2), (5., 5., 1.), 10, 2)
    assert out["rho"][0] == pytest.approx(2.088908628081126, rel=1e-15)
    assert out["p"][0] == pytest.approx(-46.990009252534314, rel=1e-14)
    assert out["pco"][0] == pytest.approx(-8.727078512943546, rel=1e-14)
    assert out["u"][0] == pytest.approx(-3.1778172561622524, rel=1e-14)
    assert np.all(out["vdot"] == 0.0)


def test_list_maintenance(golden_dir):
    g = np.load(os.path.join(golden_dir, "maintain_343.npz"))
    box = tuple(g["box"])
    b = O.verlet_build(g["r0"], g["v"], box, float(g["cutoff"]), float(g["tolerance"]))
    assert np.array_equal(b["iap"], g["iap_build"])
    c = O.compress(b["iap"], g["r1"], g["v"], box, float(g["cutoff"]), float(g["tolerance"]))
    assert np.array_equal(c, g["iap_compress"])
    assert O.ponder_rebuild(g["r0"], g["r1"], float(g["tolerance"])) == bool(g["rebuild"])


@pytest.mark.parametrize("shape,cutoff,tol", [((12, 12, 12), 2.0, 0.0), ((40, 40, 1), 2.0, 1.0),
                                               ((7, 5, 3), 2.0, 1.0)])
def test_c_cell_list_equals_brute_force(shape, cutoff, tol):
    """The C oracle's cell pruning must not lose or invent a pair (bit-exact set and order)."""
    r, v, box = O.lattice_workload(*shape, seed=7, jitter=0.3)
    if shape[2] == 1:
        box = (box[0], box[1], 20.0)
    a = O.verlet_build(r, v, box, cutoff, tol)["iap"]
    b = C.build_pairs(r, np.array(box), cutoff, tol)
    assert np.array_equal(a, b)


def test_empty_and_single():
    for n in (0, 1):
        r = np.zeros((n, 3))
        assert O.verlet_build(r, r, (5., 5., 5.), 2.0, 1.0)["iap"].shape == (0, 2)
        assert C.build_pairs(r, np.array([5., 5., 5.]), 2.0, 1.0).shape == (0, 2)


def test_builder_defined_viscous_statement_invariants():
    """oracle.gradv_two_pass / newtonian_stress / viscous_force are NOT restatements of the reference (its
    viscous arithmetic lives in the absent Fortran sphforce3d): they state the formulas the CUDA path defines,
    and the GPU suite compares the kernels with them.  Here: the invariants those formulas must have."""
    L = 10
    r, v, box = O.lattice_workload(L, L, L, seed=3, jitter=0.0)
    n = r.shape[0]
    m, h, t = np.ones(n), np.full(n, 2.0), np.ones(n)
    iap = C.build_pairs(r, np.array(box), 2.0, 0.0).astype(np.int64)
    A = np.array([[0.0, 0.02, -0.01], [-0.02, 0.0, 0.03], [0.01, -0.03, 0.0]])       # rigid rotation
    vrot = r @ A.T
    drij, rij, rsq, dv = O.separations(iap, r, vrot, box)
    pr = O.spam_properties(n, m, h, t, iap, rij, drij)
    g = O.gradv_two_pass(n, m, pr["rho"], iap, dv, pr["dwij"])
    bulk = np.all((r > 3.0) & (r < np.array(box) - 3.0), axis=1)
    c = g[bulk][0, 0, 1] / A[0, 1]
    assert -1.1 < c < -0.9                                           # gradv ~ -grad v (dW taken w.r.t. r_j - r_i)
    assert np.allclose(g[bulk], c * A, rtol=0, atol=1e-13)
    assert np.abs(O.newtonian_stress(g[bulk], 1.0, 0.1)).max() < 1e-13          # no stress under rigid rotation
    # uniform translation: no gradient at all
    drij, rij, rsq, dv0 = O.separations(iap, r, np.tile([0.3, -0.2, 0.1], (n, 1)), box)
    assert np.abs(O.gradv_two_pass(n, m, pr["rho"], iap, dv0, pr["dwij"])).max() == 0.0
    # a shear wave is decelerated and heats the fluid; the pair force is antisymmetric
    vs = np.stack([0.1 * np.sin(2 * np.pi * r[:, 1] / L), np.zeros(n), np.zeros(n)], axis=1)
    drij, rij, rsq, dv = O.separations(iap, r, vs, box)
    g = O.gradv_two_pass(n, m, pr["rho"], iap, dv, pr["dwij"])
    vd, ud = O.viscous_force(n, m, O.newtonian_stress(g, 1.0, 0.0), pr["rho"], iap, rij, pr["dwij"], dv)
    assert np.abs(vd.sum(axis=0)).max() < 1e-13 * np.abs(vd).sum()
    assert ud.sum() > 0.0 and (vd[:, 0] * vs[:, 0]).sum() < 0.0
    assert abs(ud.sum() + (vd * vs).sum()) < 1e-12 * abs(ud.sum())               # kinetic energy lost = heat gained


@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_force_variants_match_reference(golden_dir, impl):
    """forces.CohesiveSpamForce, SpamForce2d, CohesiveSpamForce2d and two forces stacked, as the reference
    computes them (tests/golden/make_golden.py:force_variants); the long-range inputs come from the fixture."""
    g = np.load(os.path.join(golden_dir, "force_variants_343.npz"))
    n = g["r"].shape[0]
    iap = g["iap"].astype(np.int64)
    m, rij, dv = g["m"], g["rij"], g["dv"]
    F = (lambda *a, **k: O.spam_force(*a, **k)) if impl == "numpy" else \
        (lambda n_, m_, pr, rho, ia, rij_, dw, dv_, cutoff=5.0, dim=3, vdot=None, udot=None:
         C.force(n_, m_, pr, rho, ia, rij_, dw, dv_, cutoff, dim, vdot, udot))
    # long-range density and kernel gradient restated from the pair list (the oracle's Lucy kernel with h = hl)
    hl = float(g["hl"])
    w_lr, dw_lr = O.lucy_kernel_pairs(rij, g["drij"] if "drij" in g.files else _drij(g), np.full(iap.shape[0], hl))
    rho_lr = np.full(n, O.lucy_kernel(0.0, (0., 0., 0.), hl)[0])
    O._scatter_pairs(rho_lr, iap, w_lr * m[iap[:, 1]], w_lr * m[iap[:, 0]])
    assert _close(dw_lr, g["dwij_lr"]) and _close(rho_lr, g["rho_lr"])
    cases = [("cohesive", g["pco"], g["rho_lr"], g["dwij_lr"], 3), ("spam2d", g["p"], g["rho"], g["dwij"], 2),
             ("cohesive2d", g["p"], g["rho"], g["dwij"], 2), ("cohesive_short", g["pco"], g["rho_lr"], g["dwij_lr"], 3)]
    for name, press, rho, dw, dim in cases:
        vd, ud = F(n, m, press, rho, iap, rij, dw, dv, cutoff=float(g["fcut_" + name]), dim=dim)
        assert _close(vd, g["vdot_" + name]) and _close(ud, g["udot_" + name]), name
    assert np.any(g["vdot_cohesive_short"] != g["vdot_cohesive"])       # the force's own cutoff does filter pairs
    vd, ud = F(n, m, g["p"], g["rho"], iap, rij, g["dwij"], dv, cutoff=5.0)
    vd, ud = F(n, m, g["pco"], g["rho_lr"], iap, rij, g["dwij_lr"], dv, cutoff=10.0, vdot=vd, udot=ud)
    assert _close(vd, g["vdot_stacked"]) and _close(ud, g["udot_stacked"])


def _drij(g):
    """Pair separations of the fixture's list (the fixture keeps rij only)."""
    box = tuple(float(x) for x in g["box"])
    return O.separations(g["iap"].astype(np.int64), g["r"], g["v"], box)[0]

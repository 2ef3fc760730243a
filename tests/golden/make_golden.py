#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE ITSELF (oracle/_ref, built from
/root/reference by oracle/make_ref.py) on seeded inputs.

Run here (the container that has /root/reference):
    python oracle/make_ref.py && python tests/golden/make_golden.py
The .npz fixtures are committed; this script is committed so they can be regenerated.
The reference modules used are its pure-Python fp64 path:
    neighbour_list.VerletList.build            (neighbour_list.py:160-189)
    neighbour_list.NeighbourList.separations   (neighbour_list.py:63-83, fp64, wrap-then-norm)
    properties.spam_properties                 (properties.py:63-120)
    forces.SpamForce / forces.Force.apply      (forces.py:38-42,321-368)
    spkernel.lucy_kernel, properties.vdw*      (spkernel.py:86-118, properties.py:38-49)
plus VerletList.compress / ponder_rebuild for the list-maintenance fixture.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))

import forces            # noqa: E402
import neighbour_list    # noqa: E402
import particles         # noqa: E402
import properties        # noqa: E402
import spkernel          # noqa: E402


def run_case(name, r, v, m, h, t, box, cutoff, tolerance, fcutoff=5.0):
    n = r.shape[0]
    p = particles.SmoothParticleSystem(n, d=3, maxn=n, xmax=box[0], ymax=box[1], zmax=box[2])
    p.r[:, :] = r
    p.v[:, :] = v
    p.m[:] = m
    p.h[:] = h
    p.t[:] = t
    nl = neighbour_list.VerletList(p, cutoff=cutoff, tolerance=tolerance)
    nl.build()
    k = nl.nip
    rsq_build = nl.rsq[:k].copy()
    neighbour_list.NeighbourList.separations(nl)        # fp64 consistent separations
    properties.spam_properties(p, nl)
    p.vdot[:, :] = 0.0
    p.udot[:] = 0.0
    f = forces.SpamForce(p, nl, cutoff=fcutoff)
    f.apply()
    out = dict(r=r, v=v, m=m, h=h, t_in=t, box=np.array(box, dtype=np.float64),
               cutoff=cutoff, tolerance=tolerance, fcutoff=fcutoff,
               iap=nl.iap[:k].astype(np.int32), rsq_build=rsq_build,
               drij=nl.drij[:k].copy(), rij=nl.rij[:k].copy(), dv=nl.dv[:k].copy(),
               wij=nl.wij[:k].copy(), dwij=nl.dwij[:k].copy(),
               rho=p.rho.copy(), p=p.p.copy(), pco=p.pco.copy(), u=p.u.copy(), t_out=p.t.copy(),
               vdot=p.vdot.copy(), udot=p.udot.copy())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("%-22s n=%5d pairs=%6d  mean rho=%.6f  max|vdot|=%.4e" %
          (name, n, k, p.rho.mean(), np.abs(p.vdot).max()))


def lattice(nx, ny, nz, seed, jitter=0.1, vmax=0.1):
    rng = np.random.default_rng(seed)
    n = nx * ny * nz
    idx = np.arange(n)
    r = np.empty((n, 3))
    r[:, 0] = idx % nx + 0.5
    r[:, 1] = (idx // nx) % ny + 0.5
    r[:, 2] = idx // (nx * ny) + 0.5
    r += rng.uniform(-jitter, jitter, size=(n, 3))
    v = vmax * (rng.random((n, 3)) - 0.5)
    return r, v


def main():
    # -- known-answer values straight from the reference functions
    kat = {}
    for i, (r, dx, h) in enumerate([(0.0, (0., 0., 0.), 2.0), (0.0, (0., 0., 0.), 1.0),
                                    (1.0, (1., 0., 0.), 2.0), (1.5, (0.9, 1.2, 0.), 2.0),
                                    (1.0, (1., 0.), 2.0), (0.5, (0.5,), 2.0),
                                    (2.0, (2., 0., 0.), 2.0), (2.5, (1.5, 2., 0.), 2.0),
                                    (0.3, (0.1, -0.2, 0.2), 1.3)]):
        w, dw = spkernel.lucy_kernel(r, dx, h)
        dwa = np.zeros(3)
        dwa[:len(dx)] = np.atleast_1d(np.asarray(dw, dtype=np.float64))
        dxa = np.zeros(3)
        dxa[:len(dx)] = dx
        kat["lucy_%d" % i] = np.array([r, h, len(dx), dxa[0], dxa[1], dxa[2], w, dwa[0], dwa[1], dwa[2]])
    kat["vdw_1_1"] = np.array(properties.vdw(1.0, 1.0))
    kat["vdw_05_15"] = np.array(properties.vdw(0.5, 1.5))
    kat["vdw_energy_1_5"] = np.array(properties.vdw_energy(1.0, 5.0))
    kat["vdw_temp_1_3"] = np.array(properties.vdw_temp(1.0, 3.0))
    np.savez_compressed(os.path.join(HERE, "kat.npz"), **kat)

    # -- C1-style sheet: 20x20x1, box 20^3, default Verlet tolerance (SURVEY 8d "C1")
    r, v = lattice(20, 20, 1, 20261)
    n = r.shape[0]
    run_case("sheet_400", r, v, np.ones(n), np.full(n, 2.0), np.ones(n), (20., 20., 20.), 2.0, 1.0)

    # -- small periodic cube: only 2 cells per side -> exercises the degenerate-grid path
    r, v = lattice(6, 6, 6, 20262)
    n = r.shape[0]
    run_case("cube_216", r, v, np.ones(n), np.full(n, 2.0), np.ones(n), (6., 6., 6.), 2.0, 1.0)

    # -- periodic cube 9^3 with cutoff=h=2, tolerance 0 (the bench's physical setting)
    r, v = lattice(9, 9, 9, 20263)
    n = r.shape[0]
    run_case("cube_729", r, v, np.ones(n), np.full(n, 2.0), np.ones(n), (9., 9., 9.), 2.0, 0.0)

    # -- random gas: non-uniform m, h, t; anisotropic box; some particles outside [0,L]
    rng = np.random.default_rng(20264)
    n = 500
    box = (11.0, 9.5, 7.25)
    r = rng.random((n, 3)) * np.array(box)
    r[:25] += rng.uniform(-1.5, 1.5, size=(25, 3))          # strays beyond the faces
    r[25] = (0.0, 0.0, 0.0)
    r[26] = (11.0, 9.5, 7.25)                               # exactly on the far corner
    r[27] = r[28]                                           # coincident pair (r == 0)
    v = rng.normal(size=(n, 3))
    m = rng.uniform(0.5, 1.5, n)
    h = rng.uniform(1.4, 2.2, n)
    t = rng.uniform(0.5, 1.5, n)
    run_case("gas_500", r, v, m, h, t, box, 2.0, 1.0, fcutoff=1.9)

    # -- list maintenance: build -> move -> compress -> ponder_rebuild
    r, v = lattice(7, 7, 7, 20265)
    n = r.shape[0]
    p = particles.SmoothParticleSystem(n, d=3, maxn=n, xmax=7., ymax=7., zmax=7.)
    p.r[:, :] = r
    p.v[:, :] = v
    nl = neighbour_list.VerletList(p, cutoff=2.0, tolerance=0.5)
    nl.build()
    iap0 = nl.iap[:nl.nip].copy()
    rng = np.random.default_rng(20266)
    r1 = r + rng.uniform(-0.2, 0.2, size=r.shape)
    p.r[:, :] = r1
    nl.compress()
    np.savez_compressed(os.path.join(HERE, "maintain_343.npz"), r0=r, r1=r1, v=v,
                        box=np.array([7., 7., 7.]), cutoff=2.0, tolerance=0.5,
                        iap_build=iap0.astype(np.int32), iap_compress=nl.iap[:nl.nip].astype(np.int32),
                        rebuild=np.array(nl.rebuild_list))
    print("maintain_343           build=%d compress=%d rebuild=%s" % (iap0.shape[0], nl.nip, nl.rebuild_list))


def c1_trajectory():
    """BASELINE config 1 restated with in-repo arithmetic: 20x20x1 sheet,
    VerletList(cutoff=2, tolerance=1) built once, pure-Python spam_properties + forces.SpamForce,
    imp_euler, dt=0.05.  Separations go through the fp64 base class (NeighbourList.separations)
    so the trajectory is the self-consistent fp64 one."""
    r, v = lattice(20, 20, 1, 20261)
    n = r.shape[0]
    particles.SPROPS = True
    p = particles.SmoothParticleSystem(n, d=3, maxn=n, xmax=20., ymax=20., zmax=20., hshort=2.0, hlong=4.0,
                                       integrator='ieuler')
    p.r[:, :] = r
    p.v[:, :] = v
    nl = neighbour_list.VerletList(p, cutoff=2.0, tolerance=1.0)
    nl.separations = lambda: neighbour_list.NeighbourList.separations(nl)
    p.nlists.append(nl)
    p.nl_default = nl
    p.forces.append(forces.SpamForce(p, nl))
    nl.build()
    nl.separations()
    properties.spam_properties(p, nl)
    out = dict(r0=r, v0=v, dt=0.05)
    for step in range(1, 21):
        p.update(0.05)
        if step in (1, 5, 20):
            out["r%d" % step] = p.r.copy()
            out["v%d" % step] = p.v.copy()
            out["rho%d" % step] = p.rho.copy()
            out["u%d" % step] = p.u.copy()
    particles.SPROPS = False
    np.savez_compressed(os.path.join(HERE, "c1_trajectory.npz"), **out)
    print("c1_trajectory          20 steps, max|v|=%.4f" % np.abs(p.v).max())


def conduction():
    """c_forces.SpamConduction.apply (c_forces.pyx:196-239), the compiled Cython original, on the
    cube_729 fixture with a synthetic heat-flux field."""
    import c_forces
    g = np.load(os.path.join(HERE, "cube_729.npz"))
    n = g["r"].shape[0]
    box = g["box"]
    p = particles.SmoothParticleSystem(n, d=3, maxn=n, xmax=box[0], ymax=box[1], zmax=box[2])
    p.r[:, :] = g["r"]
    p.v[:, :] = g["v"]
    p.m[:] = g["m"]
    p.h[:] = g["h"]
    p.t[:] = g["t_in"]
    nl = neighbour_list.VerletList(p, cutoff=float(g["cutoff"]), tolerance=float(g["tolerance"]))
    nl.build()
    neighbour_list.NeighbourList.separations(nl)
    properties.spam_properties(p, nl)
    rng = np.random.default_rng(20267)
    p.jq[:, :] = rng.normal(size=(n, 3))
    p.udot[:] = 0.0
    c_forces.SpamConduction(p, nl).apply()
    np.savez_compressed(os.path.join(HERE, "conduction_729.npz"), jq=p.jq.copy(), udot=p.udot.copy())
    print("conduction_729         max|udot|=%.4e" % np.abs(p.udot).max())


def force_variants():
    """The reference's other pressure-force classes on one seeded system: forces.CohesiveSpamForce (forces.py:371-405),
    forces.SpamForce2d (:246-274), forces.CohesiveSpamForce2d (:277-318, which as shipped applies the repulsive
    pressure) and the accumulation of two forces into the same vdot / udot.  The long-range inputs the cohesive
    force reads (p.rho_lr, nl.dwij_lr) are only ever filled by the reference's Fortran path; here they are set from
    the reference's own lucy_kernel with the long smoothing length, per pair, in list order, and saved with the
    outputs, so the fixture pins the force arithmetic for given inputs."""
    r, v = lattice(7, 7, 7, seed=20270, jitter=0.3)
    n = r.shape[0]
    rng = np.random.default_rng(20271)
    m, t = rng.uniform(0.8, 1.2, n), rng.uniform(0.8, 1.2, n)
    box = (7.0, 7.0, 7.0)
    hs, hl = 2.0, 3.0
    p = particles.SmoothParticleSystem(n, d=3, maxn=n, xmax=box[0], ymax=box[1], zmax=box[2])
    p.r[:, :] = r
    p.v[:, :] = v
    p.m[:] = m
    p.h[:] = hs
    p.hlr[:] = hl
    p.t[:] = t
    nl = neighbour_list.VerletList(p, cutoff=3.0, tolerance=0.0)
    nl.build()
    k = nl.nip
    neighbour_list.NeighbourList.separations(nl)
    properties.spam_properties(p, nl)
    p.rho_lr[:] = spkernel.lucy_kernel(0.0, np.zeros(3), hl)[0]
    for q in range(k):
        i, j = nl.iap[q]
        w, dw = spkernel.lucy_kernel(nl.rij[q], nl.drij[q], hl)
        nl.wij_lr[q] = w
        nl.dwij_lr[q, :] = dw
        p.rho_lr[i] += w * p.m[j]
        p.rho_lr[j] += w * p.m[i]
    out = dict(r=r, v=v, m=m, t_in=t, box=np.array(box), hs=hs, hl=hl, cutoff=3.0, tolerance=0.0,
               iap=nl.iap[:k].astype(np.int32), rij=nl.rij[:k].copy(), dv=nl.dv[:k].copy(), dwij=nl.dwij[:k].copy(),
               dwij_lr=nl.dwij_lr[:k].copy(), rho=p.rho.copy(), rho_lr=p.rho_lr.copy(), p=p.p.copy(), pco=p.pco.copy())
    for name, cls, cut in (("cohesive", forces.CohesiveSpamForce, 10.0), ("spam2d", forces.SpamForce2d, 5.0),
                           ("cohesive2d", forces.CohesiveSpamForce2d, 10.0), ("cohesive_short", forces.CohesiveSpamForce, 2.5)):
        p.vdot[:, :] = 0.0
        p.udot[:] = 0.0
        cls(p, nl, cutoff=cut).apply()
        out["vdot_" + name], out["udot_" + name], out["fcut_" + name] = p.vdot.copy(), p.udot.copy(), cut
    p.vdot[:, :] = 0.0
    p.udot[:] = 0.0
    forces.SpamForce(p, nl, cutoff=5.0).apply()
    forces.CohesiveSpamForce(p, nl, cutoff=10.0).apply()
    out["vdot_stacked"], out["udot_stacked"] = p.vdot.copy(), p.udot.copy()
    np.savez_compressed(os.path.join(HERE, "force_variants_343.npz"), **out)
    print("force_variants_343     pairs=%d  max|vdot| cohesive %.4e  2d %.4e" %
          (k, np.abs(out["vdot_cohesive"]).max(), np.abs(out["vdot_spam2d"]).max()))


def integrator_state0():
    return np.array([[0.3, -1.2, 0.8, 2.0], [0.0, 0.5, -0.7, 0.1], [1.0, 1.5, 0.25, -0.4]])


def integrator_rhs(x):
    """A small nonlinear system (three coupled rows) for the stepper fixture."""
    return np.stack([x[1], -np.sin(x[0]) - 0.1 * x[1] * x[2], 0.5 * x[0] * x[1] - 0.2 * x[2]])


def integrators():
    """integrator.euler / imp_euler / rk4 (integrator.py:14-103) driven through their callback protocol."""
    import integrator
    out = {"x0": integrator_state0(), "dt": 0.05, "steps": 6}
    for name in ("euler", "imp_euler", "rk4"):
        box = {"x": integrator_state0(), "xdot": None}

        def calc():
            box["xdot"] = integrator_rhs(box["x"])

        def setx(x):
            box["x"] = x.copy()

        for _ in range(out["steps"]):
            getattr(integrator, name)(lambda: box["x"], calc, lambda: box["xdot"], setx, out["dt"])
        out[name] = box["x"].copy()
    np.savez_compressed(os.path.join(HERE, "integrators.npz"), **out)
    print("integrators            euler %.6f  imp_euler %.6f  rk4 %.6f" % tuple(out[k][0, 0] for k in ("euler", "imp_euler", "rk4")))


PS_ARGS = dict(n=27, d=3, maxn=40, xmax=6.0, ymax=5.0, zmax=4.0, vmax=0.0, mass=0.1386, temperature=0.8,
               thermostat_temp=0.9, thermostat=True, hshort=1.5, hlong=3.0, integrator='rk4')


def particle_system():
    """SmoothParticleSystem storage (particles.py:69-169,256-359): the defaults its constructor leaves in the
    per-particle arrays, and the [11, maxn] state / derivative mappings (particles.py:496-542)."""
    kw = dict(PS_ARGS)
    n = kw.pop("n")
    p = particles.SmoothParticleSystem(n, **kw)
    out = {"defaults_" + k: getattr(p, k).copy() for k in ("m", "v", "t", "u", "h", "hlr", "rho", "p", "pco", "udot", "vdot")}
    out["step_name"] = p.step.__name__
    out["box_type"] = type(p.box).__name__
    out["timing_keys"] = np.array(sorted(p.timing))
    rng = np.random.default_rng(20272)
    fields = {}
    for k in ("r", "v", "rdot", "vdot"):
        fields[k] = rng.normal(size=(p.maxn, 3))
    for k in ("m", "mdot", "rho", "rhodot", "p", "pco", "u", "udot"):
        fields[k] = rng.normal(size=p.maxn)
    for k, a in fields.items():
        getattr(p, k)[...] = a
        out["in_" + k] = a
    out["x"] = p.gather_state().copy()
    out["xdot"] = p.gather_derivatives().copy()
    x2 = rng.normal(size=p.x.shape)
    out["x2"] = x2
    p.scatter_state(x2)
    for k in ("m", "r", "v", "rho", "p", "pco", "u"):
        out["scattered_" + k] = getattr(p, k).copy()
    np.savez_compressed(os.path.join(HERE, "particle_system.npz"), **out)
    print("particle_system        x %s  step %s  box %s" % (out["x"].shape, out["step_name"], out["box_type"]))


if __name__ == "__main__":
    if "--only-particle-system" in sys.argv:
        particle_system()
        sys.exit(0)
    if "--only-integrators" in sys.argv:
        integrators()
        sys.exit(0)
    if "--only-force-variants" not in sys.argv:
        main()
        c1_trajectory()
        conduction()
        integrators()
        particle_system()
    force_variants()

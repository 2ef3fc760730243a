"""Pair forces -- pyticles `forces` / `c_forces` surface for the SPH path.

    Force(particles, nl, cutoff=100.0)             forces.py:22-52
    SpamForce(p, nl, cutoff=5.0)                   forces.py:321-368, c_forces.pyx:54-115
    CohesiveSpamForce(p, nl, cutoff=10.0)          forces.py:371-405, c_forces.pyx:125-182
    SpamForce2d / CohesiveSpamForce2d              forces.py:246-318
    SpamConduction(p, nl)                          c_forces.pyx:185-239

`apply()` runs the CUDA force pass (sph_force) over the neighbour structure and ACCUMULATES
into p.vdot / p.udot, so several forces stack exactly as in the reference
(particles.py:549-550 zeroes them once per evaluation).  The acceleration carries no mass
factor while udot does (forces.py:353-368) -- kept.  Hooke and gravity forces are toy forces outside
the SPH hot path and are not provided; FortranCollisionForce is here only because
run_scripts/bspana.py:46 appends it next to SpamComplete.
"""
import torch

from . import properties as _properties


class Force(object):
    """A generic pairwise particle force (forces.py:22-52)."""

    def __init__(self, particles, nl, cutoff=100.0):
        self.p = particles
        self.nl = nl
        self.cutoff = cutoff
        self.cutoffsq = cutoff * cutoff

    def apply(self):
        """Pairs with rij^2 <= cutoff^2 get apply_force (forces.py:38-42)."""
        nl = self.nl
        use = (nl.rij ** 2 <= self.cutoffsq).nonzero().flatten().tolist()
        for k in use:
            self.apply_force(k)

    apply_sorted = apply

    def apply_force(self, k):
        pass


class _SpamBase(Force):
    _dim = 3
    _cohesive = False

    def apply(self):
        p, nl = self.p, self.nl
        nl._refresh_sorted()
        be = nl.backend
        if self._cohesive:
            press, rho, h = p.pco, p.rho_lr, p.hlr
            reuse = False
        else:
            press, rho, h = p.p, p.rho, p.h
            key = (p.p.data_ptr(), p.p._version, p.rho.data_ptr(), p.rho._version)
            reuse = be.press_ready and getattr(be, "press_key", None) == key
        be.force(press, rho, h, _properties._h_uniform(p, h), self.cutoff, self._dim, p.vdot, p.udot,
                 reuse_press=reuse)

    def apply_force(self, k):
        """One pair, host driven (forces.py:338-368) -- for scripts that call it directly."""
        p, nl = self.p, self.nl
        i = int(nl.iap[k, 0])
        j = int(nl.iap[k, 1])
        if self._cohesive:
            press, rho, dwdx = p.pco, p.rho_lr, nl.dwij_lr[k, :]
        else:
            press, rho, dwdx = p.p, p.rho, nl.dwij[k, :]
        dv = nl.dv[k, :]
        ps = press[i] / rho[i] ** 2 + press[j] / rho[j] ** 2
        a = ps * dwdx
        if self._dim == 2:
            a = a.clone()
            a[2] = 0.0
        p.vdot[i, :] += a
        p.vdot[j, :] -= a
        du = 0.5 * (a * dv).sum()
        p.udot[i] += du * p.m[j]
        p.udot[j] += du * p.m[i]


class SpamForce(_SpamBase):
    def __init__(self, particles, neighbour_list, cutoff=5.0):
        Force.__init__(self, particles, neighbour_list, cutoff=cutoff)


class CohesiveSpamForce(_SpamBase):
    _cohesive = True

    def __init__(self, particles, neighbour_list, cutoff=10.0):
        Force.__init__(self, particles, neighbour_list, cutoff=cutoff)


class SpamForce2d(_SpamBase):
    _dim = 2

    def __init__(self, particles, neighbour_list, cutoff=5.0):
        Force.__init__(self, particles, neighbour_list, cutoff=cutoff)


class CohesiveSpamForce2d(_SpamBase):
    """forces.py:277-318 -- as shipped this class applies the REPULSIVE pressure (p, rho,
    dwij) in two dimensions (the cohesive body is commented out at :287-299); kept."""
    _dim = 2

    def __init__(self, particles, neighbour_list, cutoff=10.0):
        Force.__init__(self, particles, neighbour_list, cutoff=cutoff)


class SpamConduction(Force):
    """Heat conduction using the full heat flux vector p.jq (c_forces.pyx:185-239): accumulates
    -(q_i/rho_i^2 + q_j/rho_j^2) . dW m_j into udot_i (and the mirror term into udot_j).  As in the
    reference there is no separate cutoff: the kernel gradient vanishes beyond h."""

    def __init__(self, particles, nl):
        Force.__init__(self, particles, nl)

    def apply(self):
        p, nl = self.p, self.nl
        nl._refresh_sorted()
        nl.backend.conduction(p.jq, p.rho, p.h, _properties._h_uniform(p, p.h), p.udot)


class FortranCollisionForce(Force):
    """forces.py:454-473: pairs with rij^2 <= cutoff^2 get `collision.collide3d(v_i, v_j, m_i, m_j, drij, rsq)`, a
    routine of the external Fortran module the reference does not ship.  BUILDER-DEFINED stand-in, so that the
    reference's front end runs unchanged (run_scripts/bspana.py:46): an instantaneous elastic collision of two hard
    spheres along their line of centres -- approaching pairs exchange the impulse 2 m_i m_j / (m_i + m_j) (dv . n) n,
    receding pairs are left alone -- applied pair by pair in list order like the reference's Force.apply loop
    (forces.py:38-42).  Not on the SPH hot path: it works on the exported pair list with plain tensor operations and
    touches the few pairs inside the (small) collision cutoff only."""

    def __init__(self, particles, neighbour_list, cutoff=1.0):
        Force.__init__(self, particles, neighbour_list, cutoff=cutoff)

    def apply(self):
        nl = self.nl
        if nl.nip == 0:
            return
        use = (nl.rsq[:nl.nip].as_subclass(torch.Tensor) <= self.cutoffsq).nonzero().flatten().tolist()
        for k in use:
            self.apply_force(k)

    apply_sorted = apply

    def apply_force(self, k):
        p, nl = self.p, self.nl
        i, j = int(nl.iap[k, 0]), int(nl.iap[k, 1])
        d = nl.drij[k].as_subclass(torch.Tensor)
        rr = torch.sqrt(nl.rsq[k].as_subclass(torch.Tensor))
        if float(rr) == 0.0:
            return
        nhat = d / rr
        vi, vj = p.v[i].as_subclass(torch.Tensor), p.v[j].as_subclass(torch.Tensor)
        closing = torch.dot(vj - vi, nhat)
        if float(closing) >= 0.0:
            return
        mi, mj = float(p.m[i]), float(p.m[j])
        imp = (2.0 * mi * mj / (mi + mj)) * closing * nhat
        p.v[i] = vi + imp / mi
        p.v[j] = vj - imp / mj

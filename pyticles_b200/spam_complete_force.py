"""SpamComplete -- pyticles `spam_complete_force` surface (spam_complete_force.py:19-183).

The reference class marshals everything to `sphforce3d.calc_sphforce3d`, a Fortran routine
that is NOT in the reference repository, so only its
argument / return contract can be matched.  What this class computes:

  pinned by the in-repo Python twins
    rho, rho_lr      short and long smoothing-length summation densities   (:128-131)
    p, pco, t, u     van der Waals EOS                                      (:134,151,168)
    vdot, udot       reversible pressure force, short-range repulsive part with (p, rho, h)
                     plus long-range cohesive part with (pco, rho_lr, hlr)  (:158-165)
    i.e. exactly what the reference's showcase run asks of it (eta = zeta = cgrad = 0:
    nanobox_quench.py:92-93);

  builder-defined (parity unpinned: the arithmetic lives only in the absent Fortran)
    gradv, pi_irr    eta / zeta: order-independent velocity gradient (reference sign convention:
                     gradv ~ -grad v) and the Newtonian stress
                     pi = 2 eta symmetric_traceless(gradv) + zeta tr(gradv) I  (tensor.py:5-16),
                     applied as a tensor pair force with the conventions of forces.py:353-368
                     (sph_gradv / sph_viscous_force in include/pyticles_b200.h);
    P                p I + pco I + pi_irr                                   (:177)

Non-zero cgrad / sigma / rcoef -- the capillary (density-gradient) and repulsive-core terms --
raise NotImplementedError instead of silently doing something invented: nothing in the reference
tree says what they multiply.
"""
import torch

from . import properties
from .forces import Force


class SpamComplete(Force):
    def __init__(self, particles, neighbour_list, adash=2.0, bdash=0.5, kbdash=1.0, sigma=0.0, rcoef=0.0,
                 cgrad=1.0, eta=1.0, zeta=0.1, kernel_type=2, cutoff=5.0):
        Force.__init__(self, particles, neighbour_list, cutoff=cutoff)
        self.adash, self.bdash, self.kbdash = adash, bdash, kbdash
        self.sigma, self.rcoef, self.cgrad = sigma, rcoef, cgrad
        self.eta, self.zeta = eta, zeta
        self.kernel_type = 2                                  # spam_complete_force.py:59

    def apply(self):
        if self.cgrad or self.sigma or self.rcoef:
            raise NotImplementedError(
                "SpamComplete: the cgrad/sigma/rcoef terms are defined only by the external Fortran "
                "sphforce3d, which the reference does not ship; set them to 0 (as nanobox_quench.py does)")
        p, nl = self.p, self.nl
        be = nl.backend
        n = p.n
        eos = (self.adash, self.bdash, self.kbdash)
        properties.spam_properties(p, nl, eos=eos, long_range=True)
        hu = properties._h_uniform
        # the reference overwrites vdot / udot with the routine's output (:171-181)
        p.vdot[:, :] = 0.0
        p.udot[:] = 0.0
        be.force(p.p, p.rho, p.h, hu(p, p.h), self.cutoff, 3, p.vdot, p.udot, reuse_press=True)
        be.force(p.pco, p.rho_lr, p.hlr, hu(p, p.hlr), self.cutoff, 3, p.vdot, p.udot)
        eye = torch.eye(3, dtype=p.P.dtype, device=p.P.device)
        p.P[:n] = (p.p[:n] + p.pco[:n])[:, None, None] * eye
        if self.eta or self.zeta:
            properties.spam_gradv(p, nl)
            be.viscous_force(p.gradv, p.rho, self.eta, self.zeta, p.h, hu(p, p.h), self.cutoff, p.vdot, p.udot)
            p.P[:n] += stress_tensor(p.gradv[:n], self.eta, self.zeta)


def stress_tensor(gradv, eta, zeta):
    """pi = 2 eta symmetric_traceless(gradv) + zeta tr(gradv) I  (tensor.py:5-16), batched; gradv is
    minus the velocity gradient (reference convention), so this is -2 eta S - zeta (div v) I."""
    g = gradv.as_subclass(torch.Tensor)
    tr = g.diagonal(dim1=-2, dim2=-1).sum(-1)
    eye = torch.eye(3, dtype=g.dtype, device=g.device)
    sym = 0.5 * (g + g.transpose(-1, -2)) - (tr / 3.0)[:, None, None] * eye
    return 2.0 * eta * sym + zeta * tr[:, None, None] * eye

"""SpamComplete -- pyticles `spam_complete_force` surface (spam_complete_force.py:19-183).

The reference class marshals everything to `sphforce3d.calc_sphforce3d`, a Fortran routine
that is NOT in the reference repository, so only its
argument / return contract can be matched.  What this class computes is the part of that
contract the in-repo Python twins pin:
    rho, rho_lr      short and long smoothing-length summation densities   (:128-131)
    p, pco, t, u     van der Waals EOS                                      (:134,151,168)
    vdot, udot       reversible pressure force, short-range repulsive part with (p, rho, h)
                     plus long-range cohesive part with (pco, rho_lr, hlr)  (:158-165)
i.e. exactly what the reference's showcase run asks of it (eta = zeta = cgrad = 0:
nanobox_quench.py:92-93).  Non-zero eta / zeta / cgrad / sigma / rcoef -- the viscous,
capillary and core terms whose arithmetic lives only in the absent Fortran -- raise
NotImplementedError instead of silently doing something unpinned.
"""
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
        if self.eta or self.zeta or self.cgrad or self.sigma or self.rcoef:
            raise NotImplementedError(
                "SpamComplete: eta/zeta/cgrad/sigma/rcoef terms are defined only by the external Fortran "
                "sphforce3d, which the reference does not ship; set them to 0 (as nanobox_quench.py does)")
        p, nl = self.p, self.nl
        be = nl.backend
        eos = (self.adash, self.bdash, self.kbdash)
        properties.spam_properties(p, nl, eos=eos, long_range=True)
        hu = properties._h_uniform
        # the reference overwrites vdot / udot with the routine's output (:171-181)
        p.vdot[:, :] = 0.0
        p.udot[:] = 0.0
        be.force(p.p, p.rho, p.h, hu(p, p.h), self.cutoff, 3, p.vdot, p.udot, reuse_press=True)
        be.force(p.pco, p.rho_lr, p.hlr, hu(p, p.hlr), self.cutoff, 3, p.vdot, p.udot)

"""SpamComplete -- pyticles `spam_complete_force` surface (spam_complete_force.py:19-183).

The reference class marshals everything to `sphforce3d.calc_sphforce3d`, a Fortran routine
that is NOT in the reference repository, so only its
argument / return contract can be matched.  `SpamComplete(p, nl)` with the reference's default
arguments (cgrad = 1, eta = 1, zeta = 0.1) runs, as run_scripts/bspana.py:45 writes it.  What it computes:

  pinned by the in-repo Python twins
    rho, rho_lr      short and long smoothing-length summation densities   (:128-131)
    t                T = max((u + a rho) / kb, 0) from the INTEGRATED internal energy u, which is left
                     alone (:134,151-152,168: calc_vdw_temp, then T[T < 0] = 0; properties.py:49)
    p, pco           van der Waals pressures at that T                      (properties.py:38-41)
    vdot, udot       reversible pressure force, short-range repulsive part with (p, rho, h)
                     plus long-range cohesive part with (pco, rho_lr, hlr)  (:158-165)
    i.e. exactly what the reference's showcase run asks of it (eta = zeta = cgrad = 0:
    nanobox_quench.py:92-93);

  BUILDER-DEFINED (parity unpinned: the arithmetic lives only in the absent Fortran), each checked against
  its own numpy statement in oracle/oracle.py and by invariants (tests/test_gpu_viscous.py, test_gpu_complete.py)
    eta, zeta        order-independent velocity gradient (reference sign convention: gradv ~ -grad v) and
                     the Newtonian stress pi = 2 eta symmetric_traceless(gradv) + zeta tr(gradv) I
                     (tensor.py:5-16), applied as a tensor pair force with the conventions of
                     forces.py:353-368 (sph_gradv / sph_viscous_force in include/pyticles_b200.h);
    cgrad            density-gradient (capillary) term: the gradient part of the Korteweg stress of the
                     LONG smoothing length density, P_c = cgrad (g (x) g - |g|^2 I / 2) with
                     g_i = grad rho_lr = sum_j m_j grad_i W_lr(r_ij) -- the `grad_rho_lr` the reference hands
                     to the routine next to `cgrad` (:113-115,54,158-165) -- added to the long-range reversible
                     pressure tensor and applied with dW_lr and rho_lr (sph_gradient / sph_stress_force);
    sigma, rcoef     repulsive core (:28-29,52-53): phi(r) = rcoef (1 - r^2/sigma^2)^4 per unit mass for
                     r < sigma (Hoover's SPAM core), off when either is zero -- the reference default
                     (sph_core_force);
    jq               heat flux vector (:112,181): jq = -thermalk grad T with
                     grad T_i = sum_j (m_j / rho_j) (T_j - T_i) grad_i W_ij, and its divergence enters udot
                     through the reference's own SpamConduction form (c_forces.pyx:196-239).  `thermalk`
                     is an attribute of this class only (the reference's constructor has no conductivity
                     argument); the default 0 gives jq = 0 and no conduction.
    P                p_rev + p_rev_lr + pi_irr = p I + (pco I + P_c) + pi     (:177)
"""
import torch

from . import properties
from .forces import Force


class SpamComplete(Force):
    def __init__(self, particles, neighbour_list, adash=2.0, bdash=0.5, kbdash=1.0, sigma=0.0, rcoef=0.0,
                 cgrad=1.0, eta=1.0, zeta=0.1, kernel_type=2, cutoff=5.0):
        Force.__init__(self, particles, neighbour_list, cutoff=cutoff)
        self.adash, self.bdash, self.kbdash = adash, bdash, kbdash
        # the reference sets the equation-of-state constants of the (module-global) Fortran eos here
        # (spam_complete_force.py:49-51: feos.eos.adash = adash ...), which is what spam_properties and the
        # thermostat's get_vdw_u then use; the `properties` module constants play that part
        properties.ADASH, properties.BDASH, properties.KBDASH = adash, bdash, kbdash
        self.sigma, self.rcoef, self.cgrad = sigma, rcoef, cgrad
        self.eta, self.zeta = eta, zeta
        self.kernel_type = 2                                  # spam_complete_force.py:59
        self.thermalk = 0.0                                   # builder-defined, see the module docstring

    def apply(self):
        p, nl = self.p, self.nl
        be = nl.backend
        n = p.n
        nl._refresh_sorted()
        eos = (self.adash, self.bdash, self.kbdash)
        hu = properties._h_uniform
        # densities, T from u, pressures (:128-152); u is the integrated state and is not touched
        be.density_eos(eos, p.h, hu(p, p.h), p.rho, p.p, p.pco, p.u, p.t, from_energy=True)
        be.density_eos(eos, p.hlr, hu(p, p.hlr), p.rho_lr, None, None, None, None, long_range=True)
        be.press_ready = True
        for name in ("wij", "dwij", "wij_lr", "dwij_lr"):
            nl._pairs.pop(name, None)
        # the reference overwrites vdot / udot with the routine's output (:171-181)
        be.force(p.p, p.rho, p.h, hu(p, p.h), self.cutoff, 3, p.vdot, p.udot, reuse_press=True, first_force=True)
        be.force(p.pco, p.rho_lr, p.hlr, hu(p, p.hlr), self.cutoff, 3, p.vdot, p.udot)
        T = lambda x: x.as_subclass(torch.Tensor)
        eye = torch.eye(3, dtype=p.P.dtype, device=p.P.device)
        P = (T(p.p)[:n] + T(p.pco)[:n])[:, None, None] * eye
        if self.eta or self.zeta:
            properties.spam_gradv(p, nl)
            be.viscous_force(p.gradv, p.rho, self.eta, self.zeta, p.h, hu(p, p.h), self.cutoff, p.vdot, p.udot)
            P = P + stress_tensor(T(p.gradv)[:n], self.eta, self.zeta)
        if self.cgrad:
            grho = torch.zeros((p.maxn, 3), dtype=torch.float64, device=p.P.device)
            be.gradient(None, T(p.m), False, p.hlr, hu(p, p.hlr), grho)
            pc = torch.zeros((p.maxn, 3, 3), dtype=torch.float64, device=p.P.device)
            pc[:n] = capillary_stress(grho[:n], self.cgrad)
            be.stress_force(pc, p.rho_lr, p.hlr, hu(p, p.hlr), self.cutoff, p.vdot, p.udot)
            self.grad_rho_lr = grho
            P = P + pc[:n]
        if self.sigma and self.rcoef:
            be.core_force(self.sigma, self.rcoef, p.vdot, p.udot)
        if self.thermalk:
            vol = torch.zeros(p.maxn, dtype=torch.float64, device=p.P.device)
            vol[:n] = T(p.m)[:n] / T(p.rho)[:n]
            gt = torch.zeros((p.maxn, 3), dtype=torch.float64, device=p.P.device)
            be.gradient(T(p.t), vol, True, p.h, hu(p, p.h), gt)
            p.jq[:n] = -self.thermalk * gt[:n]
            be.conduction(p.jq, p.rho, p.h, hu(p, p.h), p.udot)
        else:
            p.jq[:n] = 0.0
        p.P[:n] = P


def stress_tensor(gradv, eta, zeta):
    """pi = 2 eta symmetric_traceless(gradv) + zeta tr(gradv) I  (tensor.py:5-16), batched; gradv is
    minus the velocity gradient (reference convention), so this is -2 eta S - zeta (div v) I."""
    g = gradv.as_subclass(torch.Tensor)
    tr = g.diagonal(dim1=-2, dim2=-1).sum(-1)
    eye = torch.eye(3, dtype=g.dtype, device=g.device)
    sym = 0.5 * (g + g.transpose(-1, -2)) - (tr / 3.0)[:, None, None] * eye
    return 2.0 * eta * sym + zeta * tr[:, None, None] * eye


def capillary_stress(grad_rho, cgrad):
    """P_c = cgrad (g (x) g - |g|^2 I / 2): the gradient part of the Korteweg (square-gradient) stress, batched.
    Builder-defined, see the module docstring."""
    g = grad_rho.as_subclass(torch.Tensor)
    eye = torch.eye(3, dtype=g.dtype, device=g.device)
    return cgrad * (g[:, :, None] * g[:, None, :] - 0.5 * (g * g).sum(-1)[:, None, None] * eye)

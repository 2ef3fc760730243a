"""Smooth-particle properties -- pyticles `properties` / `f_properties` surface.

    spam_properties(p, nl[, hs, hl])   properties.py:63-120, f_properties.py:16-147

Density summation and the van der Waals equation of state run as ONE CUDA pass over the
neighbour structure (sph_density_eos): rho, p, pco, u, t are overwritten in place, as the
reference does.  Reference quirks that are kept because they are the semantics
(SURVEY.md fact 8): the self term is W(0; h[0]) for every particle, the kernel normalisation
is always 3-D, a pair uses h of its first member.  `gradv` is order dependent in the
reference (running rho, properties.py:95-98): spam_properties leaves it alone, and `spam_gradv`
computes the order-independent two-pass version (builder-defined, see sph_gradv in the header).
"""
import torch

ADASH = 2.0          # properties.py:18
BDASH = 0.5          # properties.py:19
KBDASH = 1.0         # properties.py:20
RHONAUGHT = 1.0      # properties.py:21
ADKE = False


def ideal_isothermal(rho, t):
    return rho * KBDASH


def art_water(rho, t):
    return (rho - RHONAUGHT) * KBDASH


def vdw(rho, t):
    """properties.py:38-41 -> (repulsive pressure, cohesive pressure)."""
    return (rho * KBDASH * t) / (1 - rho * BDASH), - ADASH * rho * rho


def vdw_energy(rho, t):
    """properties.py:44-46."""
    return t * KBDASH - ADASH * rho


def vdw_temp(rho, u):
    """properties.py:48-49."""
    return (u + ADASH * rho) / KBDASH


calc_pressure = vdw


def hamiltonian(p):
    """properties.py:55-60."""
    n = p.n
    H = (0.5 * p.m[:n] * (p.v[:n] ** 2).sum(dim=1) + p.u[:n]).sum()
    print(float(H))


def _h_uniform(p, h):
    key = (h.data_ptr(), h._version, p.n)
    cache = p.__dict__.setdefault("_h_uniform_cache", {})
    if cache.get("key") != key:
        n = p.n
        cache["key"] = key
        cache["val"] = bool((h[:n] == h[0]).all().item()) if n > 0 else True
    return cache["val"]


def spam_properties(p, nl, hs=None, hl=None, eos=None, long_range=False):
    """Kernel sums, densities, pressures, internal energy for every particle.

    `hs`, `hl` are accepted for f_properties compatibility (particles.py:559-561) and ignored
    like the pure-Python version ignores them; smoothing lengths come from p.h / p.hlr.
    `long_range=True` additionally fills p.rho_lr with the hlr density (f_properties.py:106-107).
    """
    be = nl.backend
    nl._refresh_sorted()
    n = p.n
    eos = (ADASH, BDASH, KBDASH) if eos is None else eos
    be.density_eos(eos, p.h, _h_uniform(p, p.h), p.rho, p.p, p.pco, p.u, p.t)
    be.press_key = (p.p.data_ptr(), p.p._version, p.rho.data_ptr(), p.rho._version)
    if long_range:
        be.density_eos(eos, p.hlr, _h_uniform(p, p.hlr), p.rho_lr, None, None, None, None, long_range=True)
        be.press_ready = True
    if p.maxn > n:
        # the reference applies these two lines to the whole arrays (properties.py:119-120)
        p.u[n:] = vdw_energy(p.rho[n:], p.t[n:])
        p.t[n:] = vdw_temp(p.rho[n:], p.u[n:])
    for name in ("wij", "dwij", "wij_lr", "dwij_lr"):
        nl._pairs.pop(name, None)


def spam_gradv(p, nl):
    """p.gradv[i, a, b] = sum_j (m_j / rho_j) (v_j - v_i)_a dW_ij/dx_b with the FINAL density p.rho
    (call after spam_properties).  Builder-defined: properties.py:95-98 uses the running density, so
    its result depends on the pair order; the weight is the one of c_properties.pyx:166-188.  As in the
    reference dW_ij is the gradient with respect to r_j - r_i, so this is MINUS the velocity gradient."""
    nl._refresh_sorted()
    nl.backend.gradv(p.rho, p.h, _h_uniform(p, p.h), p.gradv)


def spam_properties_ls(p, nl):
    """Short + long smoothing length variant (properties.py:123-187)."""
    spam_properties(p, nl, long_range=True)

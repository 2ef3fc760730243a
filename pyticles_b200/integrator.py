"""Explicit steppers -- pyticles `integrator` surface (integrator.py:14-103).

Same callback protocol: (get_state, calc_derivs, get_derivs, set_state, dt); the state is
whatever get_state returns (here a [11, maxn] CUDA tensor), so all arithmetic below runs
on the device.  State updates use the sph_axpy kernel where the expression is x0 + s*xdot.
"""
import torch

from . import backend as _backend


def _axpy(a, b, s):
    if isinstance(a, torch.Tensor) and a.is_cuda and a.dtype == torch.float64 and a.is_contiguous() \
            and isinstance(b, torch.Tensor) and b.is_contiguous() and b.shape == a.shape:
        x = torch.empty_like(a)
        _backend.axpy(x, a, b, s)
        return x
    return a + b * s


def euler(get_state, calc_derivs, get_derivs, set_state, dt):
    """integrator.py:14-41."""
    calc_derivs()
    x = get_state()
    xdot = get_derivs()
    set_state(_axpy(x, xdot, dt))


def imp_euler(get_state, calc_derivs, get_derivs, set_state, dt):
    """Improved Euler, two-stage predictor-corrector (integrator.py:44-59)."""
    calc_derivs()
    x_start = get_state().clone()
    c1 = get_derivs() * dt
    set_state(x_start + c1)
    calc_derivs()
    c2 = get_derivs() * dt
    set_state(x_start + (c1 + c2) / 2)


def rk4(get_state, calc_derivs, get_derivs, set_state, dt):
    """Fourth-order Runge-Kutta (integrator.py:62-95)."""
    calc_derivs()
    x_start = get_state().clone()
    c1 = get_derivs() * dt
    set_state(x_start + c1 / 2.0)
    calc_derivs()
    c2 = get_derivs() * dt
    set_state(x_start + c2 / 2.0)
    calc_derivs()
    c3 = get_derivs() * dt
    set_state(x_start + c3)
    calc_derivs()
    c4 = get_derivs() * dt
    set_state(x_start + (1.0 / 6.0) * (c1 + 2. * c2 + 2. * c3 + c4))


def _touch(*tensors):
    """Kernels launched through the C ABI write through raw pointers; bump the tensors' version
    counters so that the neighbour lists notice the particle state changed."""
    for t in tensors:
        t[:0].zero_()


def fused_imp_euler(p, dt):
    """Improved Euler for a SmoothParticleSystem without the [11, maxn] state matrices: the same
    arithmetic as imp_euler above driven through gather_state / scatter_state (integrator.py:44-59
    with particles.py:496-542), written directly on p.r, p.v, p.u with the sph_axpy kernel.
    Kept identical to the generic path in what it leaves behind: rho, p, pco end at their
    first-stage values (their state derivatives are zero, particles.py:538-540) and m is untouched."""
    n = p.n
    p.derivatives()                                      # stage 1: rdot = v0, vdot0, udot0
    r0, v0, u0 = p.r[:n].clone(), p.v[:n].clone(), p.u[:n].clone()
    vd0, ud0 = p.vdot[:n].clone(), p.udot[:n].clone()
    keep = [getattr(p, k)[:n].clone() for k in ("rho", "p", "pco")]
    # predictor: x = x_start + xdot * dt
    _backend.axpy(p.r[:n], r0, v0, dt)
    _backend.axpy(p.v[:n], v0, vd0, dt)
    _backend.axpy(p.u[:n], u0, ud0, dt)
    _touch(p.r, p.v, p.u)
    p.derivatives()                                      # stage 2 at the predicted state
    v1 = p.v[:n].clone()
    # corrector: x = x_start + (c1 + c2) / 2
    _backend.axpy(p.r[:n], r0, v0 + v1, 0.5 * dt)
    _backend.axpy(p.v[:n], v0, vd0 + p.vdot[:n], 0.5 * dt)
    _backend.axpy(p.u[:n], u0, ud0 + p.udot[:n], 0.5 * dt)
    _touch(p.r, p.v, p.u)
    for k, val in zip(("rho", "p", "pco"), keep):
        getattr(p, k)[:n] = val


def fused_rk4(p, dt):
    """Fourth-order Runge-Kutta for a SmoothParticleSystem on p.r, p.v, p.u directly (integrator.py:62-95 driven
    through gather_state / scatter_state, particles.py:496-542, without the [11, maxn] state matrices): the four
    derivative evaluations are the CUDA hot path, the stage updates are sph_axpy launches, and nothing leaves the
    device.  Like the generic path it leaves rho, p, pco at their first-stage values (zero state derivatives,
    particles.py:538-540)."""
    n = p.n
    p.derivatives()                                                  # k1
    r0, v0, u0 = p.r[:n].clone(), p.v[:n].clone(), p.u[:n].clone()
    keep = [getattr(p, k)[:n].clone() for k in ("rho", "p", "pco")]
    k1 = (v0.clone(), p.vdot[:n].clone(), p.udot[:n].clone())

    def stage(kprev, scale):
        """state = start + scale * dt * k_prev; returns the derivatives there."""
        _backend.axpy(p.r[:n], r0, kprev[0], scale * dt)
        _backend.axpy(p.v[:n], v0, kprev[1], scale * dt)
        _backend.axpy(p.u[:n], u0, kprev[2], scale * dt)
        _touch(p.r, p.v, p.u)
        p.derivatives()
        return (p.v[:n].clone(), p.vdot[:n].clone(), p.udot[:n].clone())

    k2 = stage(k1, 0.5)
    k3 = stage(k2, 0.5)
    k4 = stage(k3, 1.0)
    # x = x_start + (1/6) (c1 + 2 c2 + 2 c3 + c4) with c_i = k_i dt   (integrator.py:95)
    for x, x0, idx in ((p.r, r0, 0), (p.v, v0, 1), (p.u, u0, 2)):
        total = k1[idx] + 2. * k2[idx] + 2. * k3[idx] + k4[idx]
        _backend.axpy(x[:n], x0, total, dt * (1.0 / 6.0))
    _touch(p.r, p.v, p.u)
    for k, val in zip(("rho", "p", "pco"), keep):
        getattr(p, k)[:n] = val


def fused_euler(p, dt):
    """Forward Euler on p.r, p.v, p.u directly (integrator.py:14-41 through gather_state / scatter_state)."""
    n = p.n
    p.derivatives()
    keep = [getattr(p, k)[:n].clone() for k in ("rho", "p", "pco")]
    v0 = p.v[:n].clone()
    _backend.axpy(p.r[:n], p.r[:n].clone(), v0, dt)
    _backend.axpy(p.v[:n], v0, p.vdot[:n], dt)
    _backend.axpy(p.u[:n], p.u[:n].clone(), p.udot[:n], dt)
    _touch(p.r, p.v, p.u)
    for k, val in zip(("rho", "p", "pco"), keep):
        getattr(p, k)[:n] = val

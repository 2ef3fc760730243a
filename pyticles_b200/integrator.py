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

"""Particle systems -- pyticles `particles` surface on torch CUDA storage.

    ParticleSystem(n, d, maxn, controllers, xmax, ymax, zmax, vmax, simbox, mass, rinit,
                   source, side, integrator, spacing)                 particles.py:69-253
    SmoothParticleSystem(n, d, maxn, ..., temperature, thermostat_temp, thermostat, hshort,
                   hlong, source, integrator, simbox)                 particles.py:256-574

Attributes keep the reference's names and shapes (r, v, rdot, vdot [maxn,3]; m, rho, rho_lr,
t, u, udot, h, hlr, p, pco [maxn]; x, xdot [11,maxn]; n, maxn, dim, box, nlists, forces,
controllers, nl_default, timing, steps); they are float64 CUDA tensors instead of numpy
arrays.  `update(dt)` follows particles.py:459-494: rebuild_lists -> derivatives -> step ->
box.apply -> thermostat.  Refinement/amalgamation (check_refine, split, amalgamate) are
experimental 2-D code paths switched off in every shipped script (particles.py:55-57) and
are not provided.
"""
from time import time as _wall

import numpy as np
import torch

from . import box as _box
from . import configuration
from . import properties
from .array import from_numpy, ones, zeros
from .integrator import _touch, euler, fused_euler, fused_imp_euler, fused_rk4, imp_euler, rk4

XMAX = 5
YMAX = 5
ZMAX = 1
VMAX = 0.1
ADVECTIVE = False
SPROPS = False
FUSED = False            # SmoothParticleSystem.update: the steppers on p.r / p.v / p.u without the [11, maxn] state matrices
TIMING_SYNC = False      # True: synchronise the device before every clock read, so that p.timing holds device times
                         # (a bspana_profile.py-style run); False: the reference's plain wall-clock deltas, which on an
                         # asynchronous device are enqueue times


def time():
    """Clock of the p.timing entries (particles.py:163-169,474-493,552-567)."""
    if TIMING_SYNC and torch.cuda.is_available():
        torch.cuda.synchronize()
    return _wall()
DEVICE = "cuda"


def get_vdw_u(t, rho):
    """eos.get_vdw_u as used by the thermostat (particles.py:456)."""
    return properties.vdw_energy(rho, t)


class ParticleSystem(object):
    """A group of similar particles with basic mechanical properties."""

    def __init__(self, n, d=3, maxn=125, controllers=[], xmax=XMAX, ymax=YMAX, zmax=ZMAX, vmax=VMAX,
                 simbox=None, mass=1.0, rinit=None, source=None, side=(5, 5, 5), integrator='rk4',
                 spacing=0.1, device=None):
        self.n = n
        self.dim = d
        self.maxn = maxn
        self.dt = 0.0
        self.steps = 0
        self.device = torch.device(device or DEVICE)
        if not simbox:
            self.box = _box.MirrorBox(p=self, xmax=xmax, ymax=ymax, zmax=zmax)
        else:
            self.box = simbox
        self.step = {'euler': euler, 'ieuler': imp_euler, 'rk4': rk4}[integrator]

        dev = self.device
        # the hot path always works on three columns (neighbour_list.py:14)
        cols = 3
        r0 = np.zeros((maxn, cols))
        v0 = np.zeros((maxn, cols))
        r0[:, :d] = self.box.xmax * np.random.random([maxn, d])          # particles.py:121
        v0[:, :d] = vmax * (np.random.random([maxn, d]) - 0.5)          # particles.py:123
        if rinit == 'grid':
            r0[0:n, :] = configuration.grid3d(n, side, (xmax / 2., ymax / 2., zmax / 2.), spacing=spacing)
        elif rinit == 'fcc':
            raise NotImplementedError("rinit='fcc': the reference's fcc3d generator is not in its repository")
        self.r = from_numpy(r0, dev)
        self.v = from_numpy(v0, dev)
        self.m = zeros(maxn, dev)
        self.rdot = zeros((maxn, cols), dev)
        self.vdot = zeros((maxn, cols), dev)
        self.mdot = zeros(maxn, dev)
        self.m[:] = mass
        if rinit == 'load':
            # particles.py:139-143: restart from the last frame of $SPDATA/<source>
            import os
            from .spam_nc import read_step
            read_step(os.path.join(os.environ.get('SPDATA', '.'), source), self, step='last')
        self.colour = 1.0, 0.0, 0.0

        n_variables = 7
        self.x = zeros((n_variables, maxn), dev)
        self.xdot = zeros((n_variables, maxn), dev)
        self.nlists = []
        self.forces = []
        self.controllers = controllers
        for controller in self.controllers:
            controller.bind_particles(self)
        self.timing = {'force time': -1, 'deriv time': -1, 'pairsep time': -1, 'update time': -1,
                       'integrate time': -1}

    def create_particle(self, r, v=(0.0, 0.0, 0.0)):
        """particles.py:171-178."""
        self.r[self.n] = r
        self.m[self.n] = self.m[self.n - 1]
        self.v[self.n] = v
        self.n = self.n + 1
        for nl in self.nlists:
            nl.rebuild_list = True
        self.rebuild_lists()

    def rebuild_lists(self):
        """particles.py:180-185."""
        for nl in self.nlists:
            if nl.rebuild_list:
                nl.build()

    def update(self, dt):
        """particles.py:187-195."""
        self.rebuild_lists()
        self.step(self.gather_state, self.derivatives, self.gather_derivatives, self.scatter_state, dt)
        self.box.apply(self)
        _touch(self.r, self.v)
        self.steps += 1

    def gather_state(self):
        n = self.n
        self.x[0, 0:n] = self.m[0:n]
        self.x[1:4, 0:n] = self.r[0:n, :].t()
        self.x[4:7, 0:n] = self.v[0:n, :].t()
        return self.x

    def scatter_state(self, x):
        n = self.n
        self.m[0:n] = x[0, 0:n]
        self.r[0:n, :] = x[1:4, 0:n].t()
        self.v[0:n, :] = x[4:7, 0:n].t()

    def gather_derivatives(self):
        n = self.n
        self.xdot[0, 0:n] = self.mdot[0:n]
        self.xdot[1:4, 0:n] = self.rdot[0:n, :].t()
        self.xdot[4:7, 0:n] = self.vdot[0:n, :].t()
        return self.xdot

    def derivatives(self):
        """particles.py:235-253."""
        self.rdot = self.v
        self.vdot[:, :] = 0.0
        for nl in self.nlists:
            nl.separations()
        for force in self.forces:
            force.apply()
        for controller in self.controllers:
            controller.apply()


class SmoothParticleSystem(ParticleSystem):
    """A particle system with the extra fields of the smooth particle equations of motion."""

    def __init__(self, n, d=3, maxn=100, controllers=[], xmax=XMAX, ymax=YMAX, zmax=ZMAX, vmax=VMAX,
                 rinit=None, side=None, mass=1.0, spacing=None, temperature=1.0, thermostat_temp=1.0,
                 thermostat=False, hshort=1.0, hlong=3.0, source=None, integrator='ieuler', simbox=None,
                 device=None):
        # the reference drops `simbox` here (particles.py:291-302); it is honoured instead,
        # since a PeriodicBox is what the periodic scripts ask for (nanobox_quench.py:82-88)
        ParticleSystem.__init__(self, n=n, d=d, xmax=xmax, ymax=ymax, zmax=zmax, rinit=rinit, side=side,
                                mass=mass, source=source, spacing=spacing, integrator=integrator, vmax=vmax,
                                maxn=maxn, controllers=controllers, simbox=simbox, device=device)
        dev = self.device
        maxn = self.maxn
        self.rho = zeros(maxn, dev)
        self.rho_lr = zeros(maxn, dev)
        self.rhodot = zeros(maxn, dev)
        self.gradv = zeros((maxn, self.dim, self.dim), dev)
        self.jq = zeros((maxn, self.dim), dev)
        self.t = ones(maxn, dev)
        self.t[:] = temperature
        self.thermostat = thermostat
        self.u = ones(maxn, dev)
        self.udot = zeros(maxn, dev)
        self.h = zeros(maxn, dev)
        self.hlr = zeros(maxn, dev)
        self.h[:] = hshort
        self.hlr[:] = hlong
        self.p = zeros(maxn, dev)
        self.pco = zeros(maxn, dev)
        self.P = zeros((maxn, self.dim, self.dim), dev)
        self.thermostat_temp = thermostat_temp
        n_variables = 11
        self.x = zeros((n_variables, maxn), dev)
        self.xdot = zeros((n_variables, maxn), dev)
        self.timing['SPAM time'] = -1
        self.timing['nlist rebuild time'] = -1
        self.nl_default = None

    def apply_thermostat(self, target_temp):
        """Scaling thermostat (particles.py:450-457)."""
        tav = self.t.mean()
        self.t *= target_temp / tav
        self.u[:] = get_vdw_u(self.t, self.rho)

    def update(self, dt):
        """particles.py:459-494."""
        t1 = time()
        t = time()
        self.rebuild_lists()
        self.timing['nlist rebuild time'] = time() - t
        t = time()
        self.derivatives()
        self.timing['deriv time'] = time() - t
        t = time()
        if FUSED and self.step is imp_euler:
            fused_imp_euler(self, dt)
        elif FUSED and self.step is rk4:
            fused_rk4(self, dt)
        elif FUSED and self.step is euler:
            fused_euler(self, dt)
        else:
            self.step(self.gather_state, self.derivatives, self.gather_derivatives, self.scatter_state, dt)
        self.timing['integrate time'] = time() - t
        self.box.apply(self)
        _touch(self.r, self.v)
        if self.thermostat:
            self.apply_thermostat(self.thermostat_temp)
        self.timing['update time'] = time() - t1
        self.steps += 1

    def gather_state(self):
        n = self.n
        ParticleSystem.gather_state(self)
        self.x[7, 0:n] = self.rho[0:n]
        self.x[8, 0:n] = self.p[0:n]
        self.x[9, 0:n] = self.pco[0:n]
        self.x[10, 0:n] = self.u[0:n]
        return self.x

    def scatter_state(self, x):
        n = self.n
        ParticleSystem.scatter_state(self, x)
        self.rho[0:n] = x[7, 0:n]
        self.p[0:n] = x[8, 0:n]
        self.pco[0:n] = x[9, 0:n]
        self.u[0:n] = x[10, 0:n]

    def gather_derivatives(self):
        n = self.n
        ParticleSystem.gather_derivatives(self)
        self.xdot[7, 0:n] = self.rhodot[0:n]
        self.xdot[8, 0:n] = 0
        self.xdot[9, 0:n] = 0
        self.xdot[10, 0:n] = self.udot[0:n]
        return self.xdot

    def derivatives(self):
        """particles.py:544-570."""
        self.rdot = self.v
        self.vdot[:, :] = 0.0
        self.udot[:] = 0.0
        t = time()
        for nl in self.nlists:
            nl.separations()
        self.timing['pairsep time'] = time() - t
        t = time()
        if SPROPS:
            properties.spam_properties(self, self.nl_default, self.h[0:self.n], self.hlr[0:self.n])
        self.timing['SPAM time'] = time() - t
        t = time()
        for force in self.forces:
            force.apply()
        self.timing['force time'] = time() - t
        if ADVECTIVE:
            self.rdot = torch.zeros_like(self.v)

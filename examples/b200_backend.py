"""The ctypes stub of INTEGRATION.md section 2 as a real module: pyticles' OWN classes (numpy storage, the
reference's `SmoothParticleSystem`, `VerletList`) bound to the B200 C ABI.  A pyticles maintainer would drop it next
to `pairsep.pyx` / `c_forces.pyx` and call it from `VerletList.build` (neighbour_list.py:160-189),
`properties.spam_properties` (properties.py:63-120) and `SpamForce.apply` (forces.py:327-335); the numpy arrays stay
the caller-visible storage and are copied to device buffers per call, as `c_forces.pyx:81-88` already copies with
`.astype`.  tests/test_gpu_integration_stub.py runs it against the reference built in oracle/_ref.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))      # run from a checkout

from pyticles_b200.backend import NeighbourBackend      # noqa: E402  (ctypes shim over include/pyticles_b200.h)


class B200(object):
    def __init__(self, p, cutoff, tolerance, eos=(2.0, 0.5, 1.0)):
        self.p, self.be = p, NeighbourBackend("cuda")
        self.cutoff, self.tol, self.eos = cutoff, tolerance, eos
        self.dev = {}

    def _up(self, name, a):                       # numpy -> device tensor (H2D)
        a = np.ascontiguousarray(a, dtype=np.float64)
        t = self.dev.get(name)
        if t is None or t.shape != a.shape:
            t = self.dev[name] = torch.empty(a.shape, dtype=torch.float64, device="cuda")
        t.copy_(torch.from_numpy(a))
        return t

    def build(self, nl):                          # VerletList.build
        p = self.p
        n = p.n
        r, v, m = self._up("r", p.r[:n]), self._up("v", p.v[:n]), self._up("m", p.m[:n])
        self.be.plan((p.box.xmax, p.box.ymax, p.box.zmax), self.cutoff, self.tol, n, r)
        self.be.ensure(n)
        self.be.cells_and_list(r, v, m)           # sph_status_reset, sph_cells_build, sph_gather, sph_nlist_build
        iap = self.be.export_pairs().cpu().numpy()          # sph_pairs_count / sph_pairs_fill
        nl.nip = iap.shape[0]
        nl.iap[:nl.nip] = iap
        nl.r_old[:n] = p.r[:n]                    # neighbour_list.py:187

    def spam_properties(self):                    # properties.spam_properties
        p, d, n = self.p, self.dev, self.p.n
        for k in ("h", "t", "rho", "p", "pco", "u"):
            self._up(k, getattr(p, k)[:n])
        uniform = bool(np.all(p.h[:n] == p.h[0]))
        self.be.density_eos(self.eos, d["h"], uniform, d["rho"], d["p"], d["pco"], d["u"], d["t"])   # sph_density_eos
        for k in ("rho", "p", "pco", "u", "t"):
            getattr(p, k)[:n] = d[k].cpu().numpy()

    def spam_force(self, cutoff=5.0):             # forces.SpamForce.apply
        p, d, n = self.p, self.dev, self.p.n
        vdot, udot = self._up("vdot", p.vdot[:n]), self._up("udot", p.udot[:n])
        uniform = bool(np.all(p.h[:n] == p.h[0]))
        self.be.force(d["p"], d["rho"], d["h"], uniform, cutoff, 3, vdot, udot, reuse_press=True)     # sph_force
        p.vdot[:n] = vdot.cpu().numpy()
        p.udot[:n] = udot.cpu().numpy()

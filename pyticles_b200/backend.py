"""Device-side neighbour structure: owns the torch buffers the C ABI works on.

One NeighbourBackend = one cell grid + one Morton-sorted working set + one ELL neighbour
structure (include/pyticles_b200.h: sph_grid, sph_buffers).  All memory is torch-owned; the
CUDA library never allocates.  Nothing here computes on the CPU.
"""
import ctypes

import torch

from . import _lib
from ._lib import SphBuffers, SphEos, SphGrid, SphStatus, box3, check


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f64(t, what):
    if t.dtype != torch.float64 or not t.is_cuda or not t.is_contiguous():
        raise _lib.SphError("%s must be a contiguous CUDA float64 tensor" % what)
    return t


class NeighbourBackend(object):
    def __init__(self, device, max_nbrs=None):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.SphError("pyticles_b200 needs a CUDA device (got %s); there is no CPU path" % device)
        self.grid = SphGrid()
        self.grid_key = None
        self.buf = SphBuffers()
        self.n = 0
        self.K = 0
        self.user_max_nbrs = max_nbrs
        self.t = {}
        self.status_t = torch.zeros(16, dtype=torch.int32, device=self.device)
        self.built = False
        self._tab_key = None
        self.fresh = False          # sorted positions are the ones the list was built from
        self.press_ready = False    # vel4[.,3] holds press/rho^2 of the last density pass

    # ------------------------------------------------------------------ planning / memory
    def plan(self, box, cutoff, tolerance, n, r=None, slab=None, occ=None, n_hint=None):
        """Choose the cell grid.  Re-planned only when (box, cutoff, tolerance, n, slab) change;
        the occupancy extents (one reduction + sync, or `occ` = (lo, hi) given by the caller) only
        steer cell coarsening, never results.  `n_hint` overrides n as the table-size budget (the
        slab decomposition passes the global particle count so every rank plans the same grid)."""
        key = (tuple(float(b) for b in box), float(cutoff), float(tolerance),
               int(n) if n_hint is None else int(n_hint), slab)
        if key == self.grid_key:
            return
        lo = hi = None
        if occ is not None:
            lo, hi = box3(occ[0]), box3(occ[1])
        elif r is not None and n > 0:
            ext = torch.stack([r[:n].amin(dim=0), r[:n].amax(dim=0)]).cpu()
            if bool(torch.isfinite(ext).all()):
                lo = box3(ext[0].tolist())
                hi = box3(ext[1].tolist())
        check(self.lib.sph_grid_plan(box3(box), float(cutoff), float(tolerance),
                                     int(n) if n_hint is None else int(n_hint), lo, hi,
                                     ctypes.byref(self.grid)), "sph_grid_plan")
        if slab is not None:
            check(self.lib.sph_grid_restrict_x(ctypes.byref(self.grid), int(slab[0]), int(slab[1])),
                  "sph_grid_restrict_x")
        self.grid_key = key
        self.built = False

    def expected_nbrs(self, n):
        g = self.grid
        vol = g.box[0] * g.box[1] * g.box[2]
        rl = g.thr ** 0.5
        return 4.18879 * rl ** 3 * (n / vol if vol > 0 else 0.0)

    def _alloc(self, name, numel, dtype):
        t = self.t.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype:
            t = torch.empty(max(int(numel), 1), dtype=dtype, device=self.device)
            self.t[name] = t
        return t

    def ensure(self, n, K=None):
        g = self.grid
        if K is None:
            K = self.user_max_nbrs
        if K is None:
            K = self.K if (self.K and n == self.n) else int(1.35 * self.expected_nbrs(n)) + 16
        K = max(8, min(int(K), max(n - 1, 8)))
        K = (K + 3) // 4 * 4
        self.n, self.K = int(n), K
        ncode = int(g.ncode)
        b = self.buf
        b.n, b.max_nbrs = self.n, self.K
        b.cell_count = _ptr(self._alloc("cell_count", ncode + 1, torch.int32))
        b.cell_start = _ptr(self._alloc("cell_start", ncode + 2, torch.int32))
        b.scan_tmp = _ptr(self._alloc("scan_tmp", self.lib.sph_scan_tmp_elems(ncode + 1), torch.int32))
        b.code = _ptr(self._alloc("code", n, torch.int32))
        b.rank = _ptr(self._alloc("rank", n, torch.int32))
        b.perm = _ptr(self._alloc("perm", n, torch.int32))
        b.pos4 = _ptr(self._alloc("pos4", 4 * n, torch.float64))
        b.vel4 = _ptr(self._alloc("vel4", 4 * n, torch.float64))
        b.rel4 = _ptr(self._alloc("rel4", 4 * n, torch.float32))
        b.nbr = _ptr(self._alloc("nbr", self.lib.sph_nbr_elems(n, self.K), torch.int32))
        b.cnt = _ptr(self._alloc("cnt", n, torch.int32))
        b.status = _ptr(self.status_t)
        # per-grid table of the cell-group neighbour kernel (geometry only: refilled when the grid changes)
        elems = int(self.lib.sph_group_tab_elems(ctypes.byref(g)))
        if elems:
            old = self.t.get("group_tab")
            tab = self._alloc("group_tab", elems, torch.int32)
            b.group_tab = _ptr(tab)
            if tab is not old or self._tab_key != bytes(g):
                check(self.lib.sph_group_table(ctypes.byref(g), ctypes.byref(b), _stream()), "sph_group_table")
                self._tab_key = bytes(g)
        else:
            b.group_tab = None

    # ------------------------------------------------------------------ the hot path
    def cells_and_gather(self, r, v, m):
        """sph_status_reset + sph_cells_build + sph_gather on the current stream."""
        L, g, b, s = self.lib, ctypes.byref(self.grid), ctypes.byref(self.buf), _stream()
        _f64(r, "r"), _f64(v, "v"), _f64(m, "m")
        check(L.sph_status_reset(_ptr(self.status_t), s), "sph_status_reset")
        check(L.sph_cells_build(g, b, _ptr(r), s), "sph_cells_build")
        check(L.sph_gather(g, b, _ptr(r), _ptr(v), _ptr(m), s), "sph_gather")
        self.built = False

    def nlist(self):
        """sph_nlist_build on the current stream (after cells_and_gather)."""
        check(self.lib.sph_nlist_build(ctypes.byref(self.grid), ctypes.byref(self.buf), _stream()), "sph_nlist_build")
        self.built = True
        self.fresh = True
        self.press_ready = False

    def cells_and_list(self, r, v, m, check_overflow=True):
        """The whole neighbour build: cell list, Morton reorder, neighbour pass."""
        self.cells_and_gather(r, v, m)
        self.nlist()
        if check_overflow:
            self.resolve_overflow()

    def resolve_overflow(self):
        """Read the status block (one small D2H copy); grow the ELL capacity and redo the
        neighbour pass if some particle had more neighbours than max_nbrs."""
        st = self.status()
        while st.flags & _lib.SPH_F_NBR_OVERFLOW:
            need = int(st.max_count)
            if need <= self.K or need > max(self.n - 1, 8):
                raise _lib.SphError("neighbour pass reports %d neighbours for one of %d particles "
                                    "(capacity %d): inconsistent status" % (need, self.n, self.K))
            self.user_max_nbrs = None
            self.ensure(self.n, K=need + max(4, need // 8))
            L, s = self.lib, _stream()
            self.status_t[0] &= ~_lib.SPH_F_NBR_OVERFLOW
            check(L.sph_nlist_build(ctypes.byref(self.grid), ctypes.byref(self.buf), s), "sph_nlist_build")
            st = self.status()
        return st

    def regather(self, r, v, m, moved):
        check(self.lib.sph_gather(ctypes.byref(self.grid), ctypes.byref(self.buf), _ptr(_f64(r, "r")),
                                  _ptr(_f64(v, "v")), _ptr(_f64(m, "m")), _stream()), "sph_gather")
        if moved:
            self.fresh = False
        self.press_ready = False

    def density_eos(self, eos, h, h_uniform, rho, p, pco, u, t, long_range=False, from_energy=False):
        """`long_range`: rho only, with this h (the hlr density).  `from_energy`: SpamComplete's direction --
        u is read, T = max((u + a rho) / kb, 0) is written to t and the pressures follow from it."""
        e = SphEos(float(eos[0]), float(eos[1]), float(eos[2]))
        mode = 1 if long_range else (2 if from_energy else 0)
        check(self.lib.sph_density_eos(ctypes.byref(self.grid), ctypes.byref(self.buf), ctypes.byref(e),
                                       _ptr(_f64(h, "h")), int(bool(h_uniform)), int(self.fresh),
                                       mode, _ptr(rho), _ptr(p), _ptr(pco), _ptr(u), _ptr(t),
                                       _stream()), "sph_density_eos")
        self.press_ready = not long_range

    def force(self, press, rho, h, h_uniform, fcutoff, dim, vdot, udot, reuse_press=False, first_force=False, part=0):
        """`first_force`: vdot / udot would be all zero here (the evaluation's first force): results are stored,
        the caller need not zero them (particles.py:549-550).  `part`: 0 all particles, 1 all but the slab's
        boundary layers, 2 only those (see sph_force)."""
        if reuse_press and self.press_ready:
            pp = rp = ctypes.c_void_p(0)
        else:
            pp, rp = _ptr(_f64(press, "press")), _ptr(_f64(rho, "rho"))
            self.press_ready = False
        check(self.lib.sph_force(ctypes.byref(self.grid), ctypes.byref(self.buf), pp, rp, _ptr(_f64(h, "h")),
                                 int(bool(h_uniform)), int(self.fresh), float(fcutoff), int(dim),
                                 int(bool(first_force)), int(part), _ptr(_f64(vdot, "vdot")),
                                 _ptr(_f64(udot, "udot")), _stream()), "sph_force")

    def pressure_term(self, press, rho, first_orig):
        """vel4[., 3] = press/rho^2 for the particles with original index >= first_orig (ghosts)."""
        check(self.lib.sph_pressure_term(ctypes.byref(self.buf), _ptr(_f64(press, "press")), _ptr(_f64(rho, "rho")),
                                         int(first_orig), _stream()), "sph_pressure_term")

    def conduction(self, jq, rho, h, h_uniform, udot):
        aux4 = self._alloc("aux4", 4 * self.n, torch.float64)
        check(self.lib.sph_conduction(ctypes.byref(self.grid), ctypes.byref(self.buf), _ptr(_f64(jq, "jq")),
                                      _ptr(_f64(rho, "rho")), _ptr(_f64(h, "h")), int(bool(h_uniform)),
                                      int(self.fresh), _ptr(aux4), _ptr(_f64(udot, "udot")), _stream()),
              "sph_conduction")

    def gradv(self, rho, h, h_uniform, gradv):
        """sph_gradv: velocity gradient tensor [n,3,3] with the final density (builder-defined, see the header)."""
        aux4 = self._alloc("aux4", 4 * self.n, torch.float64)
        check(self.lib.sph_gradv(ctypes.byref(self.grid), ctypes.byref(self.buf), _ptr(_f64(rho, "rho")),
                                 _ptr(_f64(h, "h")), int(bool(h_uniform)), int(self.fresh), _ptr(aux4),
                                 _ptr(_f64(gradv, "gradv")), _stream()), "sph_gradv")

    def viscous_force(self, gradv, rho, eta, zeta, h, h_uniform, fcutoff, vdot, udot):
        """sph_viscous_force: Newtonian stress pair force, accumulated into vdot / udot."""
        aux8 = self._alloc("aux8", 8 * self.n, torch.float64)
        check(self.lib.sph_viscous_force(ctypes.byref(self.grid), ctypes.byref(self.buf), _ptr(_f64(gradv, "gradv")),
                                         _ptr(_f64(rho, "rho")), float(eta), float(zeta), _ptr(_f64(h, "h")),
                                         int(bool(h_uniform)), int(self.fresh), float(fcutoff), _ptr(aux8),
                                         _ptr(_f64(vdot, "vdot")), _ptr(_f64(udot, "udot")), _stream()),
              "sph_viscous_force")

    def gradient(self, f, wgt, subtract_self, h, h_uniform, out):
        """sph_gradient: out_i = sum_j wgt_j (f_j - [subtract_self] f_i) grad_i W_ij (f None: f = 1); builder-defined."""
        aux4 = self._alloc("aux4", 4 * self.n, torch.float64)
        check(self.lib.sph_gradient(ctypes.byref(self.grid), ctypes.byref(self.buf),
                                    _ptr(_f64(f, "f")) if f is not None else ctypes.c_void_p(0), _ptr(_f64(wgt, "wgt")),
                                    int(bool(subtract_self)), _ptr(_f64(h, "h")), int(bool(h_uniform)), int(self.fresh),
                                    _ptr(aux4), _ptr(_f64(out, "out")), _stream()), "sph_gradient")

    def stress_force(self, stress, rho, h, h_uniform, fcutoff, vdot, udot):
        """sph_stress_force: tensor pair force of a symmetric stress [n,3,3], accumulated into vdot / udot."""
        aux8 = self._alloc("aux8", 8 * self.n, torch.float64)
        check(self.lib.sph_stress_force(ctypes.byref(self.grid), ctypes.byref(self.buf), _ptr(_f64(stress, "stress")),
                                        _ptr(_f64(rho, "rho")), _ptr(_f64(h, "h")), int(bool(h_uniform)),
                                        int(self.fresh), float(fcutoff), _ptr(aux8), _ptr(_f64(vdot, "vdot")),
                                        _ptr(_f64(udot, "udot")), _stream()), "sph_stress_force")

    def core_force(self, sigma, rcoef, vdot, udot):
        """sph_core_force: the repulsive core of SpamComplete (sigma, rcoef), accumulated into vdot / udot."""
        check(self.lib.sph_core_force(ctypes.byref(self.grid), ctypes.byref(self.buf), float(sigma), float(rcoef),
                                      int(self.fresh), _ptr(_f64(vdot, "vdot")), _ptr(_f64(udot, "udot")), _stream()),
              "sph_core_force")

    def compress(self):
        check(self.lib.sph_compress(ctypes.byref(self.grid), ctypes.byref(self.buf), _stream()), "sph_compress")

    # ------------------------------------------------------------------ host-visible results
    def status(self):
        raw = self.status_t.cpu().numpy().tobytes()
        return SphStatus.from_buffer_copy(raw)

    def ponder_rebuild(self, r_old, r, n, tol_sq):
        check(self.lib.sph_ponder_rebuild(_ptr(r_old), _ptr(r), int(n), float(tol_sq), _ptr(self.status_t),
                                          _stream()), "sph_ponder_rebuild")
        return bool(self.status().rebuild)

    def export_pairs(self):
        """Lexicographic i<j pair list in original indices: int32 tensor [nip, 2]."""
        n = self.n
        L, b, s = self.lib, ctypes.byref(self.buf), _stream()
        if n == 0:
            return torch.zeros((0, 2), dtype=torch.int32, device=self.device)
        row_count = torch.empty(n, dtype=torch.int32, device=self.device)
        check(L.sph_pairs_count(b, _ptr(row_count), s), "sph_pairs_count")
        # 64-bit offsets: more than 2^32 pairs must not wrap (the scan is a torch op: this is the export, not the hot path)
        row_start = torch.zeros(n + 1, dtype=torch.int64, device=self.device)
        torch.cumsum(row_count.to(torch.int64), dim=0, out=row_start[1:])
        nip = int(row_start[n].item())
        iap = torch.empty((nip, 2), dtype=torch.int32, device=self.device)
        check(L.sph_pairs_fill(b, _ptr(row_start), _ptr(iap), nip, s), "sph_pairs_fill")
        return iap

    def neighbour_rows(self, idx):
        """Neighbour rows of the particles with ORIGINAL indices `idx` (int64 tensor [m]) as original indices,
        [m, K] int64 padded with -1.  A diagnostic read-out of the ELL structure (bench.py's parity gate, tests);
        not on the hot path."""
        n, K = self.n, self.K
        perm = self.t["perm"][:n].to(torch.int64)
        inv = torch.empty(n, dtype=torch.int64, device=self.device)
        inv[perm] = torch.arange(n, dtype=torch.int64, device=self.device)
        a = inv[idx]
        cnt = self.t["cnt"][:n].to(torch.int64)[a].clamp(max=K)
        k = torch.arange(K, dtype=torch.int64, device=self.device)
        off = ((a >> 5)[:, None] * K + k[None, :]) * 32 + (a & 31)[:, None]
        rows = perm[self.t["nbr"][off.clamp(max=self.t["nbr"].numel() - 1)].to(torch.int64).clamp(0, n - 1)]
        return torch.where(k[None, :] < cnt[:, None], rows, torch.full_like(rows, -1))

    def count_links(self):
        return int(self.t["cnt"][:self.n].clamp(max=self.K).sum(dtype=torch.int64).item())

    def separations(self, box, iap, r, v):
        nip = iap.shape[0]
        dev = self.device
        drij = torch.empty((nip, 3), dtype=torch.float64, device=dev)
        dv = torch.empty((nip, 3), dtype=torch.float64, device=dev)
        rij = torch.empty(nip, dtype=torch.float64, device=dev)
        rsq = torch.empty(nip, dtype=torch.float64, device=dev)
        check(self.lib.sph_separations(box3(box), _ptr(iap), nip, _ptr(r), _ptr(v), _ptr(drij), _ptr(rij),
                                       _ptr(rsq), _ptr(dv), _stream()), "sph_separations")
        return drij, rij, rsq, dv

    def pair_kernels(self, iap, rij, drij, h):
        nip = iap.shape[0]
        wij = torch.empty(nip, dtype=torch.float64, device=self.device)
        dwij = torch.empty((nip, 3), dtype=torch.float64, device=self.device)
        check(self.lib.sph_pair_kernels(_ptr(iap), nip, _ptr(rij), _ptr(drij), _ptr(_f64(h, "h")), _ptr(wij),
                                        _ptr(dwij), _stream()), "sph_pair_kernels")
        return wij, dwij


def axpy(x, a, b, s):
    """x <- a + s*b (integrator.py:40,53,57) on the current stream."""
    check(_lib.load().sph_axpy(_ptr(x), _ptr(a), _ptr(b), float(s), x.numel(), _stream()), "sph_axpy")


def box_apply(box, kind, r, v, n):
    check(_lib.load().sph_box_apply(box3(box), int(kind), _ptr(r), _ptr(v), int(n), _stream()), "sph_box_apply")

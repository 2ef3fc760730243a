"""Neighbour lists -- the pyticles `neighbour_list` module surface on the B200 backend.

Same classes, attributes and call sequence as the reference (neighbour_list.py):
    NeighbourList(particle)                         :16-124
    VerletList(particle, cutoff=2.0, tolerance=1.0) :127-252
    SortedVerletList                                :262-301
with `build / compress / separations / ponder_rebuild / find_pair / minimum_image /
apply_minimum_image / sort_by_r` and the per-pair arrays `iap, rij, rsq, drij, dv, wij,
wij_lr, dwij, dwij_lr` plus `nip, rebuild_list, cutoff_radius, cutoff_radius_sq,
tolerance_sq, r_old, max_interactions, particle`.

What is different underneath: the reference scans all n^2/2 pairs in Python and
pre-allocates n^2/2 slots per array.  Here `build()` runs the CUDA cell-list + neighbour
pass (sph_cells_build / sph_gather / sph_nlist_build) and keeps the pair set as a
device-resident neighbour structure that the density and force passes consume directly.
The per-pair arrays are materialised on the device only when somebody reads them; they are
torch CUDA tensors of exactly `nip` rows, in the reference's lexicographic i<j order.
Separations follow the fp64 base-class semantics (wrap, then norm).
"""
import numpy as np
import torch

from .array import PArray
from .backend import NeighbourBackend

DIM = 3

_PAIR_ARRAYS = ("iap", "rij", "rsq", "drij", "dv", "wij", "wij_lr", "dwij", "dwij_lr")


def _ver(t):
    return (t.data_ptr(), t._version)


class NeighbourList(object):
    """All-pairs list (neighbour_list.py:16-124).  `build()` lists every i<j pair."""

    def __init__(self, particle, max_nbrs=None):
        self.particle = particle
        self.max_interactions = (particle.maxn * particle.maxn) // 2 - 1
        self.rebuild_list = False
        self.nforce = 0
        self.forces = []
        self.backend = NeighbourBackend(particle.device, max_nbrs=max_nbrs)
        self._pairs = {}
        self._nip = None
        self._built_for = None
        self._gathered = None
        self._sorted_idx = None
        self.cutoff_radius = None
        self.cutoff_radius_sq = None
        self.tolerance_sq = 0.0
        self.defer_status = False

    # ---- pair threshold: every finite separation passes (brute-force list)
    def _threshold(self):
        return (1.0e150, 0.0)

    # ------------------------------------------------------------------ build
    def build(self):
        """Pair list from scratch (neighbour_list.py:46-56; VerletList: :160-189)."""
        p = self.particle
        n = p.n
        cutoff, tol = self._threshold()
        box = (p.box.xmax, p.box.ymax, p.box.zmax)
        be = self.backend
        be.plan(box, cutoff, tol, n, p.r, slab=getattr(p, "slab", None))
        be.ensure(n, K=(n - 1 if self.cutoff_radius is None else None))
        self.rebuild_list = False
        if hasattr(self, "r_old"):
            self.r_old[:, :] = p.r[:, :]                       # neighbour_list.py:167
        be.cells_and_list(p.r, p.v, p.m, check_overflow=not self.defer_status)
        self._built_for = _ver(p.r)
        self._gathered = (_ver(p.r), _ver(p.v), _ver(p.m))
        self._invalidate()

    def _invalidate(self, keep_iap=False):
        iap = self._pairs.get("iap") if keep_iap else None
        self._pairs = {}
        self._sorted_idx = None
        if iap is not None:
            self._pairs["iap"] = iap
        else:
            self._nip = None

    def _refresh_sorted(self):
        """Bring the Morton-sorted working copy up to date with p.r / p.v / p.m."""
        p = self.particle
        be = self.backend
        if not be.built:
            raise RuntimeError("neighbour list used before build()")
        now = (_ver(p.r), _ver(p.v), _ver(p.m))
        if now != self._gathered:
            be.regather(p.r, p.v, p.m, moved=(now[0] != self._built_for))
            self._gathered = now
            return True
        return False

    def compress(self):
        """There is no concept of compression for a brute force list (neighbour_list.py:58-61)."""
        print('Cannot compress a brute list')

    def separations(self):
        """neighbour_list.py:63-83 -- refresh the separations of the listed pairs.  The pair
        geometry is recomputed on the fly by every consumer from the sorted positions, so this
        only re-gathers them if p.r / p.v changed; drij, rij, rsq, dv are produced when read."""
        self._refresh_sorted()
        self._invalidate(keep_iap=True)

    # ------------------------------------------------------------------ lazily materialised pair arrays
    @property
    def nip(self):
        if self._nip is None:
            if not self.backend.built:
                return 0
            if "iap" in self._pairs:
                self._nip = int(self._pairs["iap"].shape[0])
            else:
                if self.defer_status:
                    self.backend.resolve_overflow()
                self._nip = self.backend.count_links() // 2
        return self._nip

    @nip.setter
    def nip(self, value):
        self._nip = int(value)

    def _materialise(self, name):
        be = self.backend
        p = self.particle
        if not be.built:
            return torch.zeros((0, 2) if name == "iap" else (0,), device=p.device,
                               dtype=torch.int32 if name == "iap" else torch.float64)
        if name == "iap":
            if self.defer_status:
                be.resolve_overflow()
            self._pairs["iap"] = be.export_pairs()
            self._nip = int(self._pairs["iap"].shape[0])
        elif name in ("drij", "rij", "rsq", "dv"):
            iap = self.iap
            box = (p.box.xmax, p.box.ymax, p.box.zmax)
            d = be.separations(box, iap, p.r, p.v)
            self._pairs.update(drij=d[0], rij=d[1], rsq=d[2], dv=d[3])
        elif name in ("wij", "dwij"):
            w = be.pair_kernels(self.iap, self.rij, self.drij, p.h)
            self._pairs.update(wij=w[0], dwij=w[1])
        elif name in ("wij_lr", "dwij_lr"):
            w = be.pair_kernels(self.iap, self.rij, self.drij, p.hlr)
            self._pairs.update(wij_lr=w[0], dwij_lr=w[1])
        if self._sorted_idx is not None and name != "iap":
            pass
        return self._pairs[name]

    def __getattr__(self, name):
        if name in _PAIR_ARRAYS:
            pairs = self.__dict__.get("_pairs")
            if pairs is None:
                raise AttributeError(name)
            if name in pairs:
                return pairs[name]
            return self._materialise(name)
        raise AttributeError(name)

    # ------------------------------------------------------------------ small helpers of the reference API
    def find_pair(self, i, j):
        """neighbour_list.py:85-96: index k of pair (i, j) or -1."""
        iap = self.iap
        hit = torch.nonzero((iap[:, 0] == i) & (iap[:, 1] == j))
        return int(hit[0, 0]) if hit.numel() else -1

    def apply_minimum_image(self):
        """neighbour_list.py:98-103."""
        p = self.particle
        self._pairs["drij"] = _minimum_image_rows(self.drij, p.box.xmax, p.box.ymax, p.box.zmax)

    def minimum_image(self, dr, xmax, ymax, zmax):
        """neighbour_list.py:105-123, in place on a 3-vector (tensor, ndarray or list)."""
        drx, dry, drz = (float(dr[0]), float(dr[1]), float(dr[2]))
        if (drx > xmax / 2.):
            drx = drx - xmax
        if (dry > ymax / 2.):
            dry = dry - ymax
        if (drz > zmax / 2.):
            drz = drz - zmax
        if (drx < -xmax / 2.):
            drx = drx + xmax
        if (dry < -ymax / 2.):
            dry = dry + ymax
        if (drz < -zmax / 2.):
            drz = drz + zmax
        dr[0], dr[1], dr[2] = drx, dry, drz


def _minimum_image_rows(d, xmax, ymax, zmax):
    d = d.clone()
    for c, L in enumerate((xmax, ymax, zmax)):
        col = d[:, c]
        col = torch.where(col > L / 2., col - L, col)
        col = torch.where(col < -L / 2., col + L, col)
        d[:, c] = col
    return d


class VerletList(NeighbourList):
    """Pairs inside cutoff^2 + tolerance^2 (neighbour_list.py:127-252)."""

    def __init__(self, particle, cutoff=2.0, tolerance=1.0, max_nbrs=None):
        NeighbourList.__init__(self, particle, max_nbrs=max_nbrs)
        self.cutoff_radius = cutoff
        self.cutoff_radius_sq = cutoff ** 2
        self.tolerance = tolerance
        self.tolerance_sq = tolerance * tolerance
        self.r_old = torch.zeros_like(self.particle.r).as_subclass(PArray)

    def _threshold(self):
        return (self.cutoff_radius, self.tolerance)

    def compress(self):
        """neighbour_list.py:191-223: drop pairs now outside the list radius, then ponder."""
        self._refresh_sorted()
        self.backend.compress()
        self._invalidate()
        self.ponder_rebuild()

    def ponder_rebuild(self):
        """neighbour_list.py:225-234: rebuild_list = max |r_old - r|^2 > tolerance^2."""
        p = self.particle
        if self.backend.ponder_rebuild(self.r_old, p.r, p.r.shape[0], self.tolerance_sq):
            self.rebuild_list = True


class SortedVerletList(VerletList):
    """Pair arrays ordered by descending separation (neighbour_list.py:262-301)."""

    def separations(self):
        VerletList.separations(self)
        self.sort_by_r()

    def sort_by_r(self):
        """neighbour_list.py:271-285."""
        for name in _PAIR_ARRAYS:
            getattr(self, name)
        idx = torch.argsort(self._pairs["rsq"], descending=True, stable=True)
        for name in _PAIR_ARRAYS:
            self._pairs[name] = self._pairs[name][idx]
        self._sorted_idx = idx

    def build(self):
        VerletList.build(self)
        self.sort_by_r()

    def compress(self):
        VerletList.compress(self)
        self.sort_by_r()


def as_numpy(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)

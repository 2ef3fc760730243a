"""The cell-group (tile) neighbour kernel against the general kernel and the CPU oracle.

The tile kernel (pyticles_b200/csrc/sph_tiles.cu) is the fast path of the neighbour pass; the
general warp-per-cell kernel behind it redoes the pass when a fixed capacity is exceeded
(SPH_F_TILE_FALLBACK).  Both must give the reference's pair set bit-exactly, and the fields computed
from either structure must agree with the oracle (<= 1e-10) and with each other (rows hold the same
sets in a different order, so sums differ by rounding only).
"""
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle as C
from oracle import oracle as O

pytestmark = pytest.mark.gpu

from test_gpu_parity import _np, check_against, make_system, rel_err, run_step  # noqa: E402


class tiles(object):
    """with tiles(False): the general kernel only (SPH_TILES is read at every call)."""

    def __init__(self, on):
        self.on = on

    def __enter__(self):
        self.old = os.environ.get("SPH_TILES")
        os.environ["SPH_TILES"] = "1" if self.on else "0"

    def __exit__(self, *a):
        if self.old is None:
            os.environ.pop("SPH_TILES", None)
        else:
            os.environ["SPH_TILES"] = self.old


def fallback_flag(nl):
    from pyticles_b200 import _lib
    return bool(nl.backend.status().flags & _lib.SPH_F_TILE_FALLBACK)


def fields(p, n):
    return {k: _np(getattr(p, k))[:n].copy() for k in ("rho", "p", "pco", "u", "t", "vdot", "udot")}


@pytest.mark.parametrize("shape,cutoff,tol,jitter", [
    ((32, 32, 32), 2.0, 0.0, 0.1),       # C3-like: 15 layers per dimension, cells of 8 .. 27 particles
    ((26, 22, 18), 2.0, 0.0, 0.1),       # odd layer counts (13, 11, 9): the periodic seam falls inside a group
    ((14, 10, 6), 2.0, 0.0, 0.2),        # 7, 5 and 3 layers: every cell is a boundary cell in z
    ((24, 24, 24), 2.0, 1.0, 0.45),      # heavy jitter, default Verlet tolerance (~47 neighbours)
    ((128, 128, 1), 2.0, 0.0, 0.1),      # C2-like sheet in a deep box
    ((20, 20, 20), 1.5, 0.0, 0.3),       # short cutoff: 13 layers of ~3.6 particles, Q = 8 streams
])
def test_tile_kernel_matches_oracle_and_general_kernel(shape, cutoff, tol, jitter):
    r, v, box = O.lattice_workload(*shape, seed=31, jitter=jitter)
    if shape[2] == 1:
        box = (box[0], box[1], float(shape[0]))
    n = r.shape[0]
    rng = np.random.default_rng(5)
    m, h, t = rng.uniform(0.8, 1.2, n), np.full(n, 2.0), rng.uniform(0.8, 1.2, n)
    ref = C.sph_step(r, v, m, h, t, np.array(box), cutoff, tol, 5.0)
    out = {}
    for on in (True, False):
        with tiles(on):
            p = make_system(r, v, m, h, t, box)
            nl = run_step(p, cutoff, tol, 5.0)
            check_against(p, nl, ref, n)
            assert fallback_flag(nl) is False, "the tile kernel gave up on a case it is meant to handle" if on \
                else "SPH_TILES=0 must not touch the tile kernel"
            out[on] = fields(p, n)
    for k in out[True]:
        assert rel_err(out[True][k], out[False][k]) < 1e-11, k


def test_tile_kernel_rows_hold_the_same_sets():
    """Row by row: the ELL structure of the tile kernel is a permutation of the general kernel's."""
    from pyticles_b200 import neighbour_list
    r, v, box = O.lattice_workload(20, 20, 20, seed=33, jitter=0.3)
    n = r.shape[0]
    rows = {}
    for on in (True, False):
        with tiles(on):
            p = make_system(r, v, np.ones(n), np.full(n, 2.0), np.ones(n), box)
            nl = neighbour_list.VerletList(p, cutoff=2.0, tolerance=0.0)
            nl.build()
            be = nl.backend
            K = be.K
            nbr = _np(be.t["nbr"])[: ((n + 31) // 32) * 32 * K].reshape(-1, K, 32)
            cnt = _np(be.t["cnt"])[:n]
            perm = _np(be.t["perm"])[:n]
            rows[on] = (nbr.copy(), cnt.copy(), perm.copy())
    (na, ca, pa), (nb, cb, pb) = rows[True], rows[False]
    assert np.array_equal(pa, pb) and np.array_equal(ca, cb)
    for a in range(0, n, 37):
        ra = np.sort(na[a >> 5, : ca[a], a & 31])
        rb = np.sort(nb[a >> 5, : cb[a], a & 31])
        assert np.array_equal(ra, rb), a


def test_tile_kernel_is_reproducible():
    r, v, box = O.lattice_workload(20, 20, 20, seed=33, jitter=0.3)
    n = r.shape[0]
    m, h, t = np.ones(n), np.full(n, 2.0), np.ones(n)
    runs = []
    for _ in range(2):
        p = make_system(r, v, m, h, t, box)
        nl = run_step(p, 2.0, 0.0, 5.0)
        assert not fallback_flag(nl)
        runs.append(fields(p, n))
    for k in runs[0]:
        assert np.array_equal(runs[0][k], runs[1][k]), k


@pytest.mark.parametrize("kind", ["crowded_cell", "crowded_window", "long_lists"])
def test_capacity_limits_fall_back_to_the_general_kernel(kind):
    rng = np.random.default_rng(7)
    r, v, box = O.lattice_workload(12, 12, 12, seed=35, jitter=0.1)
    if kind == "crowded_cell":          # 70 more particles in one cell (> 64 per cell)
        extra = np.array([5.0, 5.0, 5.0]) + rng.uniform(0.0, 0.5, size=(70, 3))
    elif kind == "crowded_window":      # 4.5 particles per unit volume: > 1280 candidates around a group
        extra = rng.uniform(0.0, 12.0, size=(6000, 3))
    else:                               # a tight cluster of 45: streams of its cell (and of the cells around) list > 32 hits even at Q = 4
        extra = np.array([3.05, 3.05, 3.05]) + rng.uniform(0.0, 0.9, size=(45, 3))
    r = np.vstack([r, extra])
    v = np.vstack([v, rng.uniform(-0.05, 0.05, size=extra.shape)])
    n = r.shape[0]
    m, h, t = np.ones(n), np.full(n, 2.0), np.ones(n)
    ref = C.sph_step(r, v, m, h, t, np.array(box), 2.0, 0.0, 5.0)
    p = make_system(r, v, m, h, t, box)
    nl = run_step(p, 2.0, 0.0, 5.0)
    if kind != "long_lists":
        assert fallback_flag(nl)
    check_against(p, nl, ref, n, pair_arrays=False)


def test_long_rows_grow_the_ell_capacity_without_fallback():
    """cutoff 3: ~110 neighbours per particle.  Rows longer than max_nbrs are an ELL overflow (the host
    grows the capacity and repeats the pass), not a tile-kernel limit."""
    from pyticles_b200 import neighbour_list
    r, v, box = O.lattice_workload(15, 15, 15, seed=41, jitter=0.1)
    n = r.shape[0]
    p = make_system(r, v, np.ones(n), np.full(n, 2.0), np.ones(n), box)
    nl = neighbour_list.VerletList(p, cutoff=3.0, tolerance=0.0, max_nbrs=16)
    nl.build()
    assert nl.backend.K > 100
    assert np.array_equal(_np(nl.iap), C.build_pairs(r, np.array(box), 3.0, 0.0))


def test_evaluator_against_oracle():
    """stepper.SphEvaluator (what bench.py times) on the tile kernel, against the oracle."""
    import torch
    from pyticles_b200 import _lib, forces, neighbour_list, particles, stepper
    from pyticles_b200.array import parray
    r, v, box = O.lattice_workload(32, 32, 32, seed=37, jitter=0.1)
    n = r.shape[0]
    ref = C.sph_step(r, v, np.ones(n), np.full(n, 2.0), np.ones(n), np.array(box), 2.0, 0.0, 5.0)
    p = particles.SmoothParticleSystem(n, d=3, maxn=n, xmax=box[0], ymax=box[1], zmax=box[2], hshort=2.0)
    p.r = parray(torch.as_tensor(r, device="cuda"))
    p.v = parray(torch.as_tensor(v, device="cuda"))
    nl = neighbour_list.VerletList(p, cutoff=2.0, tolerance=0.0)
    nl.defer_status = True
    p.nlists.append(nl)
    p.nl_default = nl
    f = forces.SpamForce(p, nl, cutoff=5.0)
    p.forces.append(f)
    ev = stepper.SphEvaluator(p, nl, f)
    for _ in range(2):
        ev.evaluate()
    st = ev.check()
    assert not st.flags & _lib.SPH_F_TILE_FALLBACK
    for name in ("rho", "p", "pco", "u", "vdot", "udot"):
        assert rel_err(_np(getattr(p, name))[:n], ref[name]) < 1e-10, name
    assert np.array_equal(_np(nl.backend.export_pairs()).astype(np.int64), ref["iap"].astype(np.int64))


def test_positions_slightly_outside_the_box_stay_on_the_tile_kernel():
    """Particles up to 0.3 outside [0, L) (an integrator step before the box is applied): binned by their
    periodic image, paired under the reference's single-shift minimum image -- no fallback needed."""
    from pyticles_b200 import _lib
    r, v, box = O.lattice_workload(16, 14, 12, seed=43, jitter=0.1)
    rng = np.random.default_rng(13)
    n = r.shape[0]
    edge = np.any((r < 0.8) | (r > np.array(box) - 0.8), axis=1)
    r[edge] += rng.uniform(-0.7, 0.7, size=(int(edge.sum()), 3))
    assert (r < 0).any() and (r > np.array(box)).any()
    m, h, t = np.ones(n), np.full(n, 2.0), np.ones(n)
    ref = C.sph_step(r, v, m, h, t, np.array(box), 2.0, 0.0, 5.0)
    p = make_system(r, v, m, h, t, box)
    nl = run_step(p, 2.0, 0.0, 5.0)
    flags = nl.backend.status().flags
    assert flags & _lib.SPH_F_OUT_OF_BOX and not flags & (_lib.SPH_F_OUT_OF_RANGE | _lib.SPH_F_TILE_FALLBACK)
    check_against(p, nl, ref, n)


def test_default_verlet_tolerance_stays_on_the_tile_kernel():
    """VerletList's default tolerance (1.0, neighbour_list.py:153): ~47 neighbours per particle, rows of
    up to ~70 -- the launcher starts at 8 particles per pass and nothing falls back."""
    r, v, box = O.lattice_workload(27, 27, 27, seed=45, jitter=0.3)
    n = r.shape[0]
    m, h, t = np.ones(n), np.full(n, 2.0), np.ones(n)
    ref = C.sph_step(r, v, m, h, t, np.array(box), 2.0, 1.0, 5.0)
    p = make_system(r, v, m, h, t, box)
    nl = run_step(p, 2.0, 1.0, 5.0)
    assert not fallback_flag(nl)
    check_against(p, nl, ref, n, pair_arrays=False)


def _pairs(r, box, cutoff, tol, on, slab=None):
    from pyticles_b200 import neighbour_list
    n = r.shape[0]
    with tiles(on):
        p = make_system(r, np.zeros_like(r), np.ones(n), np.full(n, 2.0), np.ones(n), box)
        nl = neighbour_list.VerletList(p, cutoff=cutoff, tolerance=tol)
        nl.build()
        return _np(nl.iap).astype(np.int64), fallback_flag(nl)


@pytest.mark.parametrize("jitter", [0.0, 1e-9, 1e-7])
def test_pairs_at_the_cutoff_are_decided_by_the_fp64_predicate(jitter):
    """Unit lattice, cutoff 2: six neighbours of every particle sit at distance 2 (+- jitter), inside the fp32
    error band.  `rsq < cutoff^2` is strict (neighbour_list.py:178): at jitter 0 they are all out."""
    r, v, box = O.lattice_workload(12, 12, 12, seed=61, jitter=0.0)
    if jitter:
        r = r + np.random.default_rng(5).uniform(-jitter, jitter, size=r.shape)
    ref = C.build_pairs(r, np.array(box), 2.0, 0.0).astype(np.int64)
    if jitter == 0.0:
        assert ref.shape[0] == r.shape[0] * 13            # 26 neighbours within r < 2, none at r == 2
    got, fb = _pairs(r, box, 2.0, 0.0, True)
    assert not fb
    assert np.array_equal(got, ref)


def test_random_geometries_tile_and_general_kernels_agree():
    """Forty random boxes -- anisotropic, odd layer counts, lattices, gases and clustered gases, several
    cutoffs and tolerances: the two neighbour kernels must produce the same lexicographic pair list, and
    the first of each kind is also checked against the C oracle."""
    rng = np.random.default_rng(2026)
    kinds_checked = set()
    n_tile = 0
    for case in range(40):
        kind = ("lattice", "gas", "clustered")[case % 3]
        cutoff = float(rng.choice([1.5, 2.0, 2.0, 2.5]))
        tol = float(rng.choice([0.0, 0.0, 0.5, 1.0]))
        dims = rng.integers(7, 19, size=3)
        if kind == "lattice":
            r, v, box = O.lattice_workload(int(dims[0]), int(dims[1]), int(dims[2]), seed=100 + case,
                                           jitter=float(rng.choice([0.0, 0.1, 0.3, 0.45])))
        else:
            box = tuple(float(d) for d in dims)
            n = int(np.prod(dims) * rng.uniform(0.3, 1.5))
            r = rng.random((n, 3)) * np.array(box)
            if kind == "clustered":                        # a third of the particles in a few tight blobs
                k = n // 3
                centres = rng.random((5, 3)) * np.array(box)
                r[:k] = centres[rng.integers(0, 5, k)] + rng.normal(scale=0.6, size=(k, 3))
                r = np.mod(r, np.array(box))
        a, fb_a = _pairs(r, box, cutoff, tol, True)
        b, fb_b = _pairs(r, box, cutoff, tol, False)
        assert not fb_b
        assert np.array_equal(a, b), (case, kind, cutoff, tol, dims)
        n_tile += not fb_a
        if kind not in kinds_checked:
            kinds_checked.add(kind)
            assert np.array_equal(a, C.build_pairs(r, np.array(box), cutoff, tol).astype(np.int64)), (case, kind)
    assert n_tile >= 25            # the tile kernel is the one that ran in most cases


def test_blocks_work_out_their_window_without_the_group_table():
    """sph_buffers.group_tab is optional: with NULL every block of the cell-group kernel derives the cell codes of its
    window itself (the path a caller that never calls sph_group_table gets).  Same rows, bit for bit."""
    from pyticles_b200 import neighbour_list
    r, v, box = O.lattice_workload(22, 14, 18, seed=77, jitter=0.3)       # odd layer counts: half-empty last groups
    n = r.shape[0]
    with tiles(True):
        p = make_system(r, np.zeros_like(r), np.ones(n), np.full(n, 2.0), np.ones(n), box)
        nl = neighbour_list.VerletList(p, cutoff=2.0, tolerance=0.0)
        nl.build()
        be = nl.backend
        assert be.buf.group_tab, "the backend allocates the table for a grid with cell groups"
        assert not fallback_flag(nl)
        with_tab = (be.t["nbr"].clone(), be.t["cnt"].clone())
        be.buf.group_tab = None
        be.t["nbr"].zero_()
        be.t["cnt"].zero_()
        be.nlist()
        torch.cuda.synchronize()
        assert not fallback_flag(nl)
        assert torch.equal(be.t["cnt"][:n], with_tab[1][:n])
        # entry k of sorted particle a sits at ((a >> 5) K + k) 32 + (a & 31); slots at k >= cnt[a] are never written
        K, nw = be.K, (n + 31) // 32
        cnt = torch.zeros(nw * 32, dtype=torch.int32, device=be.t["cnt"].device)
        cnt[:n] = be.t["cnt"][:n]
        live = torch.arange(K, device=cnt.device)[None, :, None] < cnt.view(nw, 1, 32)
        rows_a = be.t["nbr"][:nw * K * 32].view(nw, K, 32)
        rows_b = with_tab[0][:nw * K * 32].view(nw, K, 32)
        assert int(live.sum()) == 2 * _np(nl.iap).shape[0]
        assert torch.equal(rows_a[live], rows_b[live])
    ref = C.build_pairs(r, np.array(box), 2.0, 0.0).astype(np.int64)
    assert np.array_equal(_np(nl.iap).astype(np.int64), ref)

"""CPU suite (gloo, world_size 2, 3 and 4): the host-side logic of the slab decomposition --
ownership by cell layer, migration, the fixed-capacity ring exchange A/B (the protocol the CUDA path uses) and the
pair-ownership rule -- with the CPU oracle as the local "kernel": the global pair set must come out bit-exact and
rho / p / vdot / udot of every particle within 1e-10 of the oracle on the whole box.  With three or more ranks the
left and right neighbours differ, and rank 0 and rank W-1 meet across the periodic face.  The CUDA passes
themselves are covered by the gpu suite."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import c_oracle as C
from oracle import oracle as O

BOX = (24.0, 8.0, 8.0)
CUTOFF, TOL = 2.0, 0.5


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _inputs():
    r, v, box = O.lattice_workload(24, 8, 8, seed=31, jitter=0.3)
    n = r.shape[0]
    rng = np.random.default_rng(9)
    m = 1.0 + 0.1 * rng.random(n)
    t = 1.0 + 0.2 * rng.random(n)
    return r, v, m, np.full(n, 2.0), t


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pyticles_b200 import distributed as D
        r, v, m, h, t = _inputs()
        n = r.shape[0]
        gid = np.arange(n)
        mine = gid % world == rank                      # deliberately NOT spatial: migrate() must fix it
        T = torch.from_numpy
        rows = D.make_rows(T(r[mine]), T(v[mine]), T(m[mine]), T(h[mine]), T(t[mine]), T(gid[mine]))
        dec = D.SlabDecomposition(BOX, CUTOFF, TOL, n)
        res = {"rank": rank, "nc": dec.nc, "bounds": dec.bounds, "nbrs": (dec.left, dec.right)}
        own = dec.migrate(rows)
        lay = dec.layer_of(own[:, D.C_R])
        res["own_ok"] = bool(((lay >= dec.lay0) & (lay < dec.lay1)).all())
        ghosts = dec.halo_exchange(own)
        gl = dec.layer_of(ghosts[:, D.C_R])
        li, ri, cl, cr = dec._halo
        # the left neighbour's last layer first, then the right neighbour's first layer
        res["ghost_ok"] = bool((gl[:cl] == (dec.lay0 - 1) % dec.nc).all() and (gl[cl:] == dec.lay1 % dec.nc).all()
                               and cl + cr == ghosts.shape[0] and cl > 0 and cr > 0)
        # exchange B must deliver columns of the same particles in the same order
        tag = torch.stack([own[:, D.C_GID] * 2 + 1, own[:, D.C_GID] * 3], dim=1)
        got = dec.halo_exchange_again(tag)
        res["b_ok"] = bool((got[:, 0] == ghosts[:, D.C_GID] * 2 + 1).all() and (got[:, 1] == ghosts[:, D.C_GID] * 3).all())
        # one derivative evaluation as SlabSphEvaluator.evaluate runs it, the oracle standing in for the kernels
        loc = torch.cat([own, ghosts]).numpy()
        no, nl = own.shape[0], own.shape[0] + ghosts.shape[0]
        bx = np.array(BOX)
        lr, lv = np.ascontiguousarray(loc[:, 0:3]), np.ascontiguousarray(loc[:, 3:6])
        iap = C.build_pairs(lr, bx, CUTOFF, TOL)
        drij, rij, rsq, dv = C.separations(iap, lr, lv, bx)
        pr = C.density_eos(nl, loc[:, D.C_M], loc[:, D.C_H], loc[:, D.C_T], iap, rij, drij)
        pb = dec.halo_exchange_again(torch.from_numpy(np.stack([pr["p"][:no], pr["rho"][:no]], axis=1)))    # B
        press, rho = pr["p"].copy(), pr["rho"].copy()
        press[no:], rho[no:] = pb[:, 0].numpy(), pb[:, 1].numpy()
        vdot, udot = C.force(nl, loc[:, D.C_M], press, rho, iap, rij, pr["dwij"], dv, 5.0)
        res.update(rho=pr["rho"][:no], p=pr["p"][:no], vdot=vdot[:no], udot=udot[:no])
        # the ownership rule on the local pair list
        iap = torch.from_numpy(iap.astype(np.int64))
        g = torch.from_numpy(loc[:, D.C_GID]).to(torch.int64)
        gi, gj = g[iap[:, 0]], g[iap[:, 1]]
        keep = dec.owns_pair(gi, gj, iap[:, 0] < no, iap[:, 1] < no)
        pairs = torch.stack([torch.minimum(gi, gj)[keep], torch.maximum(gi, gj)[keep]], dim=1).numpy()
        res["pairs"] = pairs
        res["gids"] = own[:, D.C_GID].to(torch.int64).numpy()
        # move everything by 1.7 cells along x (periodic wrap) and migrate again
        moved = own.clone()
        moved[:, D.C_R] = torch.remainder(moved[:, D.C_R] + 1.7 / dec.inv_w, BOX[0])
        own2 = dec.migrate(moved)
        lay2 = dec.layer_of(own2[:, D.C_R])
        res["own2_ok"] = bool(((lay2 >= dec.lay0) & (lay2 < dec.lay1)).all())
        res["gids2"] = own2[:, D.C_GID].to(torch.int64).numpy()
        gathered = [None] * world
        dist.all_gather_object(gathered, res)
        if rank == 0:
            out.put(gathered)
    finally:
        dist.destroy_process_group()


def _spawn(target, world, timeout=240):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=target, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    import queue
    import time
    res, t0 = None, time.time()
    try:
        while res is None:
            try:
                res = q.get(timeout=2.0)
            except queue.Empty:
                if time.time() - t0 > timeout or any(p.exitcode not in (None, 0) for p in procs):
                    for p in procs:
                        if p.is_alive():
                            p.kill()
                    raise AssertionError("a rank failed or timed out")
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.kill()
    for p in procs:
        assert p.exitcode == 0
    return res


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 3, 4])
def test_slab_decomposition(world):
    res = _spawn(_worker, world)
    r, v, m, h, t = _inputs()
    n = r.shape[0]
    ref = C.sph_step(r, v, m, h, t, np.array(BOX), CUTOFF, TOL, 5.0)
    assert all(x["nc"] == res[0]["nc"] and x["bounds"] == res[0]["bounds"] for x in res)
    assert min(b - a for a, b in zip(res[0]["bounds"], res[0]["bounds"][1:])) >= 2
    for k, x in enumerate(res):
        assert x["own_ok"] and x["ghost_ok"] and x["b_ok"] and x["own2_ok"]
        assert x["nbrs"] == ((k - 1) % world, (k + 1) % world)
    if world > 2:
        assert all(x["nbrs"][0] != x["nbrs"][1] for x in res)          # a ring with distinct neighbours
    # every particle owned exactly once, before and after the move
    for key in ("gids", "gids2"):
        allg = np.sort(np.concatenate([x[key] for x in res]))
        assert np.array_equal(allg, np.arange(n))
    # union of the per-rank pair lists == global pair set, each pair exactly once
    allp = np.concatenate([x["pairs"] for x in res])
    allp = allp[np.lexsort((allp[:, 1], allp[:, 0]))]
    assert allp.shape == ref["iap"].shape
    assert np.array_equal(allp, ref["iap"].astype(np.int64))
    # pairs across the periodic face between rank 0 and rank W-1 are part of it
    nc, lay = res[0]["nc"], np.floor(r[:, 0] * (res[0]["nc"] / BOX[0])).astype(int)
    assert np.any((lay[ref["iap"][:, 0]] == 0) & (lay[ref["iap"][:, 1]] == nc - 1))
    gid = np.concatenate([x["gids"] for x in res])
    for k in ("rho", "p", "vdot", "udot"):
        full = np.empty_like(ref[k])
        full[gid] = np.concatenate([x[k] for x in res])
        scale = np.maximum(np.abs(ref[k]), 1e-3 * np.max(np.abs(ref[k])))
        assert np.max(np.abs(full - ref[k]) / scale) < 1e-10, k


def test_single_rank_decomposition_is_identity():
    from pyticles_b200 import distributed as D
    dec = D.SlabDecomposition(BOX, CUTOFF, TOL, 100)
    assert dec.world == 1 and dec.slab is None
    rows = torch.zeros((5, D.NCOL), dtype=torch.float64)
    assert dec.migrate(rows) is rows
    assert dec.halo_exchange(rows).shape[0] == 0


# ---------------------------------------------------------------------------------------------------------------
# SlabStepper: improved Euler + periodic box + thermostat over two ranks against the same arithmetic on one process.
# The derivative evaluation is a stand-in (every rank gathers the whole box and asks the C oracle), so that the
# test pins the stepper's own logic: predictor / corrector bookkeeping across a migration, box, global thermostat.
STEP_DIMS, STEP_DT, STEP_N, STEP_T = (16, 6, 6), 0.05, 3, 1.3


class _OracleSim(object):
    """The storage protocol of SlabSphEvaluator on CPU tensors."""

    def __init__(self, rows, box, cutoff, tol, fcut, n_total):
        from pyticles_b200 import distributed as D
        self.D = D
        self.dec = D.SlabDecomposition(box, cutoff, tol, n_total)
        self.cutoff, self.tol, self.fcut = cutoff, tol, fcut
        self._load_rows(self.dec.migrate(rows))

    def _load_rows(self, rows):
        D, n = self.D, rows.shape[0]
        self.n_owned = n
        z = lambda *shape: torch.zeros(shape, dtype=torch.float64)
        self.S = dict(r=rows[:, D.C_R:D.C_R + 3].clone(), v=rows[:, D.C_V:D.C_V + 3].clone(), m=rows[:, D.C_M].clone(),
                      h=rows[:, D.C_H].clone(), t=rows[:, D.C_T].clone(), gid=rows[:, D.C_GID].to(torch.int64),
                      rho=z(n), p=z(n), pco=z(n), u=z(n), vdot=z(n, 3), udot=z(n))

    def rows(self):
        S = self.S
        return self.D.make_rows(S["r"], S["v"], S["m"], S["h"], S["t"], S["gid"])

    def evaluate(self):
        D, own = self.D, self.rows().numpy()
        if self.dec.world > 1:
            parts = [None] * self.dec.world
            dist.all_gather_object(parts, own)
            own_all = np.concatenate(parts)
        else:
            own_all = own
        g = own_all[np.argsort(own_all[:, D.C_GID])]
        ref = C.sph_step(np.ascontiguousarray(g[:, D.C_R:D.C_R + 3]), np.ascontiguousarray(g[:, D.C_V:D.C_V + 3]),
                         np.ascontiguousarray(g[:, D.C_M]), np.ascontiguousarray(g[:, D.C_H]),
                         np.ascontiguousarray(g[:, D.C_T]), np.array(self.dec.box), self.cutoff, self.tol, self.fcut)
        mine = self.S["gid"].numpy()
        for k in ("rho", "p", "pco", "u", "vdot", "udot"):
            self.S[k] = torch.from_numpy(np.ascontiguousarray(ref[k][mine]))


def _step_inputs():
    r, v, box = O.lattice_workload(*STEP_DIMS, seed=77, jitter=0.3, vmax=4.0)    # fast: some cross a slab face
    n = r.shape[0]
    rng = np.random.default_rng(5)
    t = 1.0 + 0.2 * rng.random(n)
    return r, v, box, t


def _step_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pyticles_b200 import distributed as D
        r, v, box, t = _step_inputs()
        n = r.shape[0]
        gid = np.arange(n)
        mine = gid % world == rank
        k = int(mine.sum())
        rows = D.make_rows(torch.from_numpy(r[mine]), torch.from_numpy(v[mine]), torch.ones(k, dtype=torch.float64),
                           torch.full((k,), 2.0, dtype=torch.float64), torch.from_numpy(t[mine]), torch.from_numpy(gid[mine]))
        sim = _OracleSim(rows, box, CUTOFF, 0.0, 5.0, n)
        st = D.SlabStepper(sim, box_kind="periodic", thermostat_temp=STEP_T)
        moved = 0
        for _ in range(STEP_N):
            before = set(sim.S["gid"].tolist())
            st.step(STEP_DT)
            moved += len(set(sim.S["gid"].tolist()) - before)
        lay = sim.dec.layer_of(sim.S["r"][:, 0])
        res = {"gid": sim.S["gid"].numpy(), "moved": moved,
               "own_ok": bool(((lay >= sim.dec.lay0) & (lay < sim.dec.lay1)).all())}
        for kx in ("r", "v", "t", "u", "rho", "p", "pco"):
            res[kx] = sim.S[kx].numpy()
        parts = [None] * world
        dist.all_gather_object(parts, res)
        if rank == 0:
            out.put(parts)
    finally:
        dist.destroy_process_group()


def _one_process_steps():
    """SmoothParticleSystem.update with imp_euler, PeriodicBox and the thermostat, written out on global arrays
    (particles.py:450-494, integrator.py:44-59, box.py:35-47)."""
    r, v, box, t = _step_inputs()
    n = r.shape[0]
    m, h, bx = np.ones(n), np.full(n, 2.0), np.array(box)
    for _ in range(STEP_N):
        d1 = C.sph_step(r, v, m, h, t, bx, CUTOFF, 0.0, 5.0)
        r0, v0, u0 = r.copy(), v.copy(), d1["u"].copy()
        r, v = r0 + v0 * STEP_DT, v0 + d1["vdot"] * STEP_DT
        d2 = C.sph_step(r, v, m, h, t, bx, CUTOFF, 0.0, 5.0)
        v1 = v.copy()
        r = r0 + (v0 + v1) * (0.5 * STEP_DT)
        v = v0 + (d1["vdot"] + d2["vdot"]) * (0.5 * STEP_DT)
        u = u0 + (d1["udot"] + d2["udot"]) * (0.5 * STEP_DT)
        for d in range(3):
            hi, lo = r[:, d] > box[d], r[:, d] < 0
            r[hi, d] = 0.0
            r[lo, d] = box[d]
        t = t * (STEP_T / t.mean())
        u = t * 1.0 - 2.0 * d1["rho"]
    return dict(r=r, v=v, t=t, u=u, rho=d1["rho"], p=d1["p"], pco=d1["pco"])


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 3])
def test_slab_stepper_matches_one_process(world):
    res = _spawn(_step_worker, world)
    ref = _one_process_steps()
    n = ref["r"].shape[0]
    gid = np.concatenate([x["gid"] for x in res])
    assert np.array_equal(np.sort(gid), np.arange(n))                 # every particle owned exactly once
    assert all(x["own_ok"] for x in res)
    assert sum(x["moved"] for x in res) > 0                           # the run did carry particles across a face
    for k in ("r", "v", "t", "u", "rho", "p", "pco"):
        got = np.concatenate([x[k] for x in res])
        full = np.empty_like(ref[k])
        full[gid] = got
        assert np.allclose(full, ref[k], rtol=1e-12, atol=1e-12), k


def test_slab_stepper_single_rank():
    """world 1: no process group, migrate is the identity; same numbers as the written-out update."""
    from pyticles_b200 import distributed as D
    r, v, box, t = _step_inputs()
    n = r.shape[0]
    rows = D.make_rows(torch.from_numpy(r), torch.from_numpy(v), torch.ones(n, dtype=torch.float64),
                       torch.full((n,), 2.0, dtype=torch.float64), torch.from_numpy(t), torch.arange(n))
    sim = _OracleSim(rows, box, CUTOFF, 0.0, 5.0, n)
    st = D.SlabStepper(sim, box_kind="periodic", thermostat_temp=STEP_T)
    for _ in range(STEP_N):
        st.step(STEP_DT)
    ref = _one_process_steps()
    for k in ("r", "v", "t", "u", "rho", "p", "pco"):
        assert np.allclose(sim.S[k].numpy(), ref[k], rtol=1e-12, atol=1e-12), k
    with pytest.raises(ValueError):
        D.SlabStepper(sim, box_kind="torus")

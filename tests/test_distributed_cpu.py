"""CPU suite (gloo, world_size 2): the host-side logic of the slab decomposition --
ownership by cell layer, migration, ghost exchange A/B and the pair-ownership rule --
checked against the CPU oracle.  The CUDA passes themselves are covered by the gpu suite."""
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


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pyticles_b200 import distributed as D
        r, v, box = O.lattice_workload(24, 8, 8, seed=31, jitter=0.3)
        n = r.shape[0]
        gid = np.arange(n)
        mine = gid % world == rank                      # deliberately NOT spatial: migrate() must fix it
        rows = D.make_rows(torch.from_numpy(r[mine]), torch.from_numpy(v[mine]), torch.ones(mine.sum(), dtype=torch.float64),
                           torch.full((int(mine.sum()),), 2.0, dtype=torch.float64), torch.ones(mine.sum(), dtype=torch.float64),
                           torch.from_numpy(gid[mine]))
        dec = D.SlabDecomposition(BOX, CUTOFF, TOL, n)
        res = {"rank": rank, "nc": dec.nc, "bounds": dec.bounds}
        own = dec.migrate(rows)
        lay = dec.layer_of(own[:, D.C_R])
        res["own_ok"] = bool(((lay >= dec.lay0) & (lay < dec.lay1)).all())
        ghosts = dec.halo_exchange(own)
        gl = dec.layer_of(ghosts[:, D.C_R])
        res["ghost_ok"] = bool(((gl == (dec.lay0 - 1) % dec.nc) | (gl == dec.lay1 % dec.nc)).all())
        # exchange B must deliver columns of the same particles in the same order
        tag = torch.stack([own[:, D.C_GID] * 2 + 1, own[:, D.C_GID] * 3], dim=1)
        got = dec.halo_exchange_again(tag)
        res["b_ok"] = bool((got[:, 0] == ghosts[:, D.C_GID] * 2 + 1).all() and (got[:, 1] == ghosts[:, D.C_GID] * 3).all())
        # local pairs by the oracle on owned + ghost particles, then the ownership rule
        loc = torch.cat([own, ghosts])
        no = own.shape[0]
        iap = torch.from_numpy(C.build_pairs(loc[:, 0:3].numpy(), np.array(BOX), CUTOFF, TOL).astype(np.int64))
        g = loc[:, D.C_GID].to(torch.int64)
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


@pytest.mark.timeout(300)
def test_slab_decomposition_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, _free_port_once(), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    r, v, box = O.lattice_workload(24, 8, 8, seed=31, jitter=0.3)
    ref = C.build_pairs(r, np.array(BOX), CUTOFF, TOL).astype(np.int64)
    assert res[0]["nc"] == res[1]["nc"] and res[0]["bounds"] == res[1]["bounds"]
    for x in res:
        assert x["own_ok"] and x["ghost_ok"] and x["b_ok"] and x["own2_ok"]
    # every particle owned exactly once, before and after the move
    for key in ("gids", "gids2"):
        allg = np.sort(np.concatenate([x[key] for x in res]))
        assert np.array_equal(allg, np.arange(r.shape[0]))
    # union of the per-rank pair lists == global pair set, each pair exactly once
    allp = np.concatenate([x["pairs"] for x in res])
    allp = allp[np.lexsort((allp[:, 1], allp[:, 0]))]
    assert allp.shape == ref.shape
    assert np.array_equal(allp, ref)


_PORT = []


def _free_port_once():
    if not _PORT:
        _PORT.append(_free_port())
    return _PORT[0]


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
STEP_DIMS, STEP_DT, STEP_N, STEP_T = (12, 6, 6), 0.05, 3, 1.3


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
def test_slab_stepper_world2_matches_one_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_step_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
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

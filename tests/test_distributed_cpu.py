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

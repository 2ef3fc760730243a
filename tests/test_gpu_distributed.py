"""GPU suite, needs >= 2 GPUs (skipped otherwise): the slab-decomposed evaluation over NCCL
against the CPU oracle on the whole box -- global pair set bit-exact (each pair reported once),
rho / p / vdot / udot of every particle within 1e-10."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DIMS = (48, 16, 16)
CUTOFF, TOL, FCUT, H = 2.0, 0.0, 5.0, 2.0


def _port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from oracle import oracle as O
        from pyticles_b200 import distributed as D
        r, v, box = O.lattice_workload(*DIMS, seed=41, jitter=0.25)
        n = r.shape[0]
        gid = np.arange(n)
        mine = (gid // (n // world)) == rank
        k = int(mine.sum())
        one = torch.ones(k, dtype=torch.float64, device=dev)
        rows = D.make_rows(torch.from_numpy(r[mine]).to(dev), torch.from_numpy(v[mine]).to(dev), one, one * H, one,
                           torch.from_numpy(gid[mine]).to(dev))
        ev = D.SlabSphEvaluator(rows, box, CUTOFF, TOL, FCUT, (2.0, 0.5, 1.0), n, dev)
        ev.evaluate()
        ev.check()
        res = {"gid": ev.own_gid.cpu().numpy(),
               "pairs": ev.local_pairs_global_ids().cpu().numpy()}
        for kx in ("rho", "p", "vdot", "udot"):
            res[kx] = ev.result[kx].cpu().numpy()
        out = [None] * world
        dist.all_gather_object(out, res)
        if rank == 0:
            q.put(out)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.timeout(600)
def test_two_gpu_slab_evaluation_matches_oracle():
    import torch.multiprocessing as mp
    from oracle import c_oracle as C
    from oracle import oracle as O
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=500)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    r, v, box = O.lattice_workload(*DIMS, seed=41, jitter=0.25)
    n = r.shape[0]
    ref = C.sph_step(r, v, np.ones(n), np.full(n, H), np.ones(n), np.array(box), CUTOFF, TOL, FCUT)
    pairs = np.concatenate([x["pairs"] for x in res])
    pairs = pairs[np.lexsort((pairs[:, 1], pairs[:, 0]))]
    assert np.array_equal(pairs, ref["iap"].astype(np.int64))
    gid = np.concatenate([x["gid"] for x in res])
    assert np.array_equal(np.sort(gid), np.arange(n))
    for k in ("rho", "p", "vdot", "udot"):
        got = np.concatenate([x[k] for x in res])
        full = np.empty_like(ref[k])
        full[gid] = got
        scale = np.maximum(np.abs(ref[k]), 1e-3 * np.max(np.abs(ref[k])))
        assert np.max(np.abs(full - ref[k]) / scale) < 1e-10, k

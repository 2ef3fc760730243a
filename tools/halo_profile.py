#!/usr/bin/env python
"""Where the time of one slab-decomposed evaluation goes, sub-step by sub-step (CUDA events between every launch
group of distributed.SlabSphEvaluator.evaluate): the binning of the owned particles, exchange A (pack, ring
send/recv, unpack), the rest of the cell list, the three pair passes and exchange B.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \\
        tools/halo_profile.py [--workload c3|c4] [--steps 10]

Prints one table per rank 0 (mean over the steps, max over the ranks).  The events serialise nothing (everything is
on one stream anyway), so the sum is the evaluation time of bench.py.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist

import bench
from pyticles_b200 import stepper


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sim = stepper.make_bench_system(bench.WORKLOADS[a.workload], world, rank, dev, bench.SEED, bench.H, bench.CUTOFF,
                                    bench.TOL, bench.FCUT, bench.EOS)
    for _ in range(3):
        sim.evaluate()
    sim.check()
    if not hasattr(sim, "fine_times"):
        raise SystemExit("one GPU: there is no halo; run under torchrun with N > 1")
    sim.reset_pass_timers()
    if world > 1:
        dist.barrier()
    for _ in range(a.steps):
        sim.evaluate(timed="fine")
    torch.cuda.synchronize()
    rows = sim.fine_times()
    t = torch.tensor([ms for _, ms in rows], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("# %s on %d GPUs, %d particles here, halo capacity %d rows per side, ghosts %s; mean ms over %d evaluations, "
              "max over ranks" % (a.workload, world, sim.n_owned, sim.halo_cap, getattr(sim, "ghosts", None), a.steps))
        for (label, _), ms in zip(rows, t.tolist()):
            print("%-44s %8.4f" % (label, ms))
        print("%-44s %8.4f" % ("sum", float(t.sum())))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Quench of a periodic SPH box spread over several GPUs: the liquid-vapour set-up of pyticles' nanobox_quench.py
(:57-101 -- scaled van der Waals constants, lattice start, thermostat) with the box cut into x slabs, one process per
GPU, ghost exchange and migration over NCCL (pyticles_b200.distributed), time stepping by distributed.SlabStepper.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \\
        examples/slab_quench.py [--side 64] [--steps 20] [--dt 0.01]

Prints, per step, the global particle count (must not change), the mean temperature (held by the thermostat), the
density range and how many particles changed rank.  Only the short-range van der Waals pressure force is applied
(forces.SpamForce over SlabSphEvaluator); the long-range cohesive pass of the single-GPU quench example is not
part of the slab evaluator yet.
"""
import argparse
import os
import sys
from time import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))      # run from a checkout

import torch
import torch.distributed as dist

from pyticles_b200 import distributed as D


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=64, help="lattice planes per dimension on EACH rank (x) / in total (y, z)")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--dt", type=float, default=0.01)
    ap.add_argument("--spacing", type=float, default=1.0)
    ap.add_argument("--temperature", type=float, default=0.8)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # this rank's block of the lattice (x fastest), small random velocities
    nx, ny, nz = a.side, a.side, a.side
    n = nx * ny * nz
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    idx = torch.arange(n, device=dev, dtype=torch.int64)
    r = torch.empty((n, 3), dtype=torch.float64, device=dev)
    r[:, 0] = (idx % nx + rank * nx).to(torch.float64) + 0.5
    r[:, 1] = ((idx // nx) % ny).to(torch.float64) + 0.5
    r[:, 2] = (idx // (nx * ny)).to(torch.float64) + 0.5
    r *= a.spacing
    v = (torch.rand((n, 3), dtype=torch.float64, device=dev, generator=g) - 0.5) * 0.1
    one = torch.ones(n, dtype=torch.float64, device=dev)
    h = 2.0 * a.spacing
    rows = D.make_rows(r, v, one, one * h, one * a.temperature, idx + rank * n)
    box = (nx * world * a.spacing, ny * a.spacing, nz * a.spacing)
    eos = (2.0, 0.5, 1.0)                                  # properties.py:18-20
    sim = D.SlabSphEvaluator(rows, box, cutoff=h, tol=0.0, fcut=5.0 * a.spacing, eos=eos, n_total=n * world, device=dev)
    st = D.SlabStepper(sim, box_kind="periodic", thermostat_temp=a.temperature, eos=eos)
    if rank == 0:
        print("ranks %d  particles %d  box %s" % (world, n * world, box))
        print("STEP  seconds  particles  mean T  min rho  max rho  changed rank")
    t0 = time()
    for k in range(a.steps):
        before = sim.own_gid.clone()
        st.step(a.dt)
        no = sim.n_owned
        S = sim.S
        moved = no - int(torch.isin(sim.own_gid, before).sum())
        bad = float(torch.isnan(S["r"][:no]).any())
        acc = torch.tensor([float(no), float(S["t"][:no].sum()), float(moved), bad], dtype=torch.float64, device=dev)
        lo = S["rho"][:no].min().reshape(1)
        hi = S["rho"][:no].max().reshape(1)
        if world > 1:
            dist.all_reduce(acc)
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if float(acc[3]) > 0.0:                      # decided by ALL ranks together: nobody is left waiting in a collective
            if bad:
                print("rank %d: stopping due to nan" % rank)
            break
        if rank == 0:
            print("%4d  %7.3f  %9d  %.6f  %.4f  %.4f  %d" % (k, time() - t0, int(acc[0]), float(acc[1] / acc[0]),
                                                           float(lo), float(hi), int(acc[2])))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

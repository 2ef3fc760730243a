"""Where the halo exchange of the slab decomposition spends its time (torchrun, >= 2 GPUs):
wall clock of each sub-step with a device sync on both sides, mean over the timed evaluations.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/halo_profile.py
"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import bench
    from pyticles_b200 import stepper
    sim = stepper.make_bench_system(bench.WORKLOADS["c3"], world, rank, dev, bench.SEED, bench.H, bench.CUTOFF,
                                    bench.TOL, bench.FCUT, bench.EOS)
    for _ in range(3):
        sim.evaluate()
    sim.check()
    dec = sim.dec
    acc = {}

    def timed(name, fn):
        def wrap(*a, **k):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = fn(*a, **k)
            torch.cuda.synchronize()
            acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
            return out
        return wrap

    dec.halo_select_x = timed("halo_select_x", dec.halo_select_x)
    dec._exchange = timed("_exchange", dec._exchange)
    sim._pack = timed("_pack", sim._pack)
    sim._halo_a = timed("halo_a (all)", sim._halo_a)
    dec.halo_exchange_again = timed("halo_exchange_again (all)", dec.halo_exchange_again)
    sim.be.pressure_term = timed("pressure_term", sim.be.pressure_term)
    steps = 10
    for _ in range(steps):
        sim.evaluate()
    torch.cuda.synchronize()
    if rank == 0:
        for k, v in acc.items():
            print("%-28s %.3f ms" % (k, 1e3 * v / steps))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

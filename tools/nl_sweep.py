#!/usr/bin/env python
"""Neighbour-list build-only sweep (BASELINE.json configs[4]; SURVEY.md section 8d "C5").

    python tools/nl_sweep.py [--max-n 67108864] [--reps 5] > profiles/nl_sweep.txt

For lattice-plus-jitter boxes of N particles at number density rho and list cutoff rc (tolerance 0)
it times one neighbour build -- cell list + Morton reorder + neighbour pass -- with CUDA events and
prints pairs, pairs/s, particles/s and the fraction of the HBM roofline on the algorithmic bytes
24 N + 8 P.  The reference side of this config (test/time_nlist.py protocol) is `--reference`:
VerletList(cutoff=10, tolerance=2) build / compress, then timed separations(), n = 3..99.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def reference_sweep():
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    import neighbour_list
    import particles
    rng = np.random.default_rng(0)
    print("# reference protocol of test/time_nlist.py:16-26 (oracle/_ref, 1 core): n, pairs, seconds per separations()")
    for n in (3, 10, 30, 60, 99):
        p = particles.ParticleSystem(n, d=3, maxn=n)
        p.r[:, :] = rng.random((n, 3)) * 10
        nl = neighbour_list.VerletList(p, cutoff=10, tolerance=2)
        nl.build()
        nl.compress()
        t = time.perf_counter()
        nl.separations()
        dt = time.perf_counter() - t
        print("%4d %6d %.6f  (%.3e pairs/s)" % (n, nl.nip, dt, nl.nip / dt if dt > 0 else 0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-n", type=int, default=1 << 26)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--reference", action="store_true")
    ap.add_argument("--logn", type=int, nargs="*", default=[20, 22, 24, 26, 28], help="log2 of the particle counts")
    a = ap.parse_args()
    if a.reference:
        return reference_sweep()
    import torch
    import bench
    from pyticles_b200 import neighbour_list, particles
    from pyticles_b200.array import parray
    peak, _ = bench.peaks()
    dev = torch.device("cuda", 0)
    print("# N, density, cutoff, pairs/particle, ms per build, particles/s, pairs/s, alg GB/s, frac of %.0f GB/s, "
          "neighbour kernel (tile = cell-group kernel, general = warp-per-cell kernel after SPH_F_TILE_FALLBACK)" % peak)
    for logn in a.logn:
        n_target = 1 << logn
        if n_target > a.max_n:
            break
        for rho, rc in ((1.0, 1.5), (1.0, 2.0), (1.0, 2.5), (1.0, 3.0), (0.5, 2.0), (2.0, 2.0)):
            if logn >= 26 and (rc > 2.0 or rho > 1.0):
                continue                                   # keep the big boxes to the headline setting
            side = round(n_target ** (1.0 / 3.0))
            dims = (side, side, side)
            n = side ** 3
            spacing = rho ** (-1.0 / 3.0)
            r, v = bench.lattice_on_device(dims, 0, dev, 1)
            r *= spacing
            box = tuple(float(s * spacing) for s in dims)
            p = particles.SmoothParticleSystem(n, d=3, maxn=n, xmax=box[0], ymax=box[1], zmax=box[2], device=dev)
            p.r, p.v = parray(r), parray(v)
            nl = neighbour_list.VerletList(p, cutoff=rc, tolerance=0.0)
            nl.build()                                     # sizes the neighbour capacity
            torch.cuda.synchronize()
            nl.defer_status = True
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.reps):
                nl.build()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.reps
            pairs = nl.nip
            gbs = (24.0 * n + 8.0 * pairs) / (ms * 1e-3) / 1e9
            which = "general" if nl.backend.status().flags & 32 else "tile"
            print("%10d %4.1f %4.1f %7.2f %9.3f %.3e %.3e %8.1f %.4f %s" %
                  (n, rho, rc, pairs / n, ms, n / (ms * 1e-3), pairs / (ms * 1e-3), gbs, gbs / peak, which), flush=True)
            del p, nl, r, v
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

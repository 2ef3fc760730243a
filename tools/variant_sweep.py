#!/usr/bin/env python
"""A/B timing of compile-time variants of the density / force passes (csrc/sph_kernels.cu tunables).

    python tools/variant_sweep.py build            # here (no GPU): one library per variant under variants/
    python tools/variant_sweep.py run [names...]   # on the GPU box: bench.py per variant, pass-time table

Every variant computes each particle's sums in the same order with the same operations (only launch
geometry, register caps, pipeline depth and cache hints differ), so the results are bit-identical and
the parity tests need not be repeated per variant; `run` still checks rho/vdot checksums against the
default build.  The table goes to stdout and gpurun_out/variant_sweep.txt.
"""
import json
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "variants")

R1E = dict(SPH_PP_BLOCK=256, SPH_DENS_MINB=1, SPH_FORCE_MINB=1, SPH_ROW_U=4, SPH_ROW_UF=2, SPH_IDX_AHEAD=1)   # defaults up to r1e
W = dict(SPH_PP_BLOCK=128, SPH_DENS_MINB=7, SPH_FORCE_MINB=7, SPH_ROW_U=4, SPH_ROW_UF=1, SPH_IDX_AHEAD=2)     # defaults since r1f


def _v(base, **kw):
    d = dict(base)
    d.update(kw)
    return ["-D%s=%s" % kv for kv in sorted(d.items())]


VARIANTS = {
    "default": [],
    # first sweep (profiles/r1f_variant_sweep_gen1.txt): around the r1e defaults
    "base": _v(R1E),
    "ahead2": _v(R1E, SPH_IDX_AHEAD=2),
    "b128": _v(R1E, SPH_PP_BLOCK=128),
    "b128_r72_80": _v(R1E, SPH_PP_BLOCK=128, SPH_DENS_MINB=7, SPH_FORCE_MINB=6),
    "b128_r64_64": _v(R1E, SPH_PP_BLOCK=128, SPH_DENS_MINB=8, SPH_FORCE_MINB=8),
    "u2_uf1_b128_r48_72": _v(R1E, SPH_ROW_U=2, SPH_ROW_UF=1, SPH_PP_BLOCK=128, SPH_DENS_MINB=10, SPH_FORCE_MINB=7),
    "u4_uf1_b128_r72_72_ahead2": _v(W),
    "u6_uf3": _v(R1E, SPH_ROW_U=6, SPH_ROW_UF=3),
    "b512": _v(R1E, SPH_PP_BLOCK=512),
    "b64": _v(R1E, SPH_PP_BLOCK=64),
    # second sweep (profiles/r1f_variant_sweep_gen2.txt): on top of the first sweep's winner
    "w": _v(W),
    "w_f_ahead3": _v(W, SPH_IDX_AHEAD_F=3),
    "w_f_ahead4": _v(W, SPH_IDX_AHEAD_F=4),
    "w_f_r64": _v(W, SPH_FORCE_MINB=8),
    # round 2: window staged by TMA bulk copies (cp.async.bulk + mbarrier) in the cell-group neighbour kernel
    "tile_tma": ["-DSPH_TILE_TMA=1"],
    # round 2: the tensor-core variant of the neighbour kernel (csrc/sph_tiles_mma.cu; run with SPH_TILES=2)
    "mma_b5": ["-DSPH_TILE_BLOCKS=5"],    # 5 blocks per SM (48 registers, spills), window capacity 1024
}
# The other rows of the r1f sweep tables (noalloc, keep, *_smq, *_maxl1, w_f_pipe_r80, w_stride8, w_pairload, w_intra)
# were variants whose code was removed after they lost; they can be rebuilt from commit dac84c7.


def lib_path(name):
    return os.path.join(VDIR, "libpyticles_b200_%s.so" % name)


def build_one(name):
    from pyticles_b200 import build as B
    cmd = [B.nvcc()] + B.NVCC_FLAGS + VARIANTS[name] + B.SRC + ["-o", lib_path(name)]
    subprocess.check_call(cmd)
    return name


def cmd_build(names):
    os.makedirs(VDIR, exist_ok=True)
    with ThreadPoolExecutor(max_workers=6) as ex:
        for nm in ex.map(build_one, names):
            print("built", lib_path(nm))


DIGEST = r"""
import hashlib, sys
sys.path.insert(0, %r)
import numpy as np, torch
from oracle import oracle as O
from pyticles_b200 import forces, neighbour_list, particles, properties
r, v, box = O.lattice_workload(24, 20, 16, seed=5)
n = r.shape[0]
p = particles.SmoothParticleSystem(n, d=3, maxn=n, xmax=box[0], ymax=box[1], zmax=box[2], hshort=2.0, device="cuda:0")
p.r[0:n, :] = r
p.v[0:n, :] = v
nl = neighbour_list.VerletList(p, cutoff=2.0, tolerance=0.0)
nl.build(); nl.separations()
properties.spam_properties(p, nl)
forces.SpamForce(p, nl).apply()
torch.cuda.synchronize()
h = hashlib.sha256()
for k in ("rho", "p", "u", "vdot", "udot"):
    h.update(getattr(p, k).cpu().numpy()[:n].tobytes())
print("DIGEST", h.hexdigest()[:16])
""" % ROOT


def digest(env):
    p = subprocess.run([sys.executable, "-c", DIGEST], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=300)
    for l in p.stdout.splitlines():
        if l.startswith("DIGEST"):
            return l.split()[1]
    return "failed: " + p.stderr[-200:].replace("\n", " | ")


def cmd_run(names, steps, budget_s):
    """Times the variants in the given order until `budget_s` is spent; the table is rewritten after every
    variant (a cut-off call still leaves what was measured) and the fastest build (whole evaluation) is named in
    gpurun_out/variant_winner.txt."""
    import time
    t_start = time.time()
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    head = "%-32s %9s %9s %9s %9s %9s  %s" % ("variant", "step ms", "cells", "neighbour", "density", "force", "sm MHz")
    lines, best = [head], (None, 1e30)
    ref_digest = digest(dict(os.environ))
    for nm in names:
        if time.time() - t_start > budget_s:
            lines.append("%-32s skipped (time budget)" % nm)
            continue
        if not os.path.exists(lib_path(nm)):
            lines.append("%-32s not built" % nm)
            continue
        env = dict(os.environ, PYTICLES_B200_LIB=lib_path(nm))
        p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", str(steps), "--warmup", "3",
                            "--no-e2e", "--no-cpu-baseline"], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                           text=True, timeout=300)
        line = [l for l in p.stdout.splitlines() if l.startswith("{")]
        if p.returncode != 0 or not line:
            lines.append("%-32s failed rc=%d %s" % (nm, p.returncode, p.stderr[-300:].replace("\n", " | ")))
        else:
            js = json.loads(line[-1])
            d = digest(env)
            same = d == ref_digest
            ps = js["roofline"]["passes"]
            lines.append("%-32s %9.3f %9.3f %9.3f %9.3f %9.3f  %s  %s" % (
                nm, js["ms_per_step"], ps["cells+reorder"]["ms"], ps["neighbour"]["ms"], ps["density"]["ms"],
                ps["force"]["ms"], (js.get("clocks") or {}).get("sm_mhz"),
                "same bits as the default build" if same else "DIGEST %s != %s" % (d, ref_digest)))
            t = js["ms_per_step"]
            if same and t < best[1]:
                best = (nm, t)
        with open(os.path.join(out_dir, "variant_sweep.txt"), "w") as fh:
            fh.write("\n".join(lines) + "\n")
        with open(os.path.join(out_dir, "variant_winner.txt"), "w") as fh:
            fh.write((best[0] or "base") + "\n")
    print("\n".join(lines))
    print("winner:", best[0])


if __name__ == "__main__":
    args = sys.argv[1:]
    mode = args[0] if args else "run"
    names = [a for a in args[1:] if not a.startswith("-")] or list(VARIANTS)
    budget = [float(a.split("=")[1]) for a in args if a.startswith("--budget=")]
    if mode == "build":
        cmd_build(names)
    else:
        cmd_run(names, 6, budget[0] if budget else 600.0)

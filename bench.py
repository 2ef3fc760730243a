#!/usr/bin/env python
"""Benchmark of the SPH step hot path (BASELINE.json: "SPH particle-updates/s").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c4|small] [--impl reference]

One "step" = one derivative evaluation of the hot path over one synthetic lattice-plus-jitter
box: cell list + Morton reorder + neighbour pass + density/EOS pass +
force pass.  `value` = particles * steps / time with inputs resident in HBM; `e2e` = the same
through the public API with HOST buffers (pinned H2D of r, v, m, h, t and D2H of rho, p, vdot,
udot every step).  Prints ONE JSON line (rank 0).

Workloads (number density 1, h = cutoff = 2.0, tolerance 0, force cutoff 5, EOS a=2 b=.5 kb=1):
    c3     3-D 256^3 = 16 777 216 particles on ONE GPU (BASELINE configs[2]); N>1 keeps 16 Mi per GPU
           (weak scaling: the box grows along x, slab-decomposed; 4 GPUs = the 64 Mi box of configs[3])
    c2     2-D sheet 1024x1024x1 in a 1024-deep box (configs[1])
    c4     3-D 512x512x256 = 64 Mi particles in total, strong-scaled over the GPUs (configs[3])
    c5     neighbour-list build only (test/time_nlist.py's protocol: the list, not the forces), 2^26 particles per GPU
           (configs[4])
    small  3-D 64^3 (quick functional run)
--all-configs appends a `configs` block to the JSON line (outside the timed headline): c2 with its own roofline,
c4 strong-scaled when N > 1, and the c5 build-only point.

--impl reference times the reference's own CPU implementation (oracle/_ref, built from the
unmodified reference sources by oracle/make_ref.py) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "SPH particle-updates/s"
UNIT = "particle-updates/s"
SEED = 20263

WORKLOADS = {
    #        per-GPU lattice (nx, ny, nz), box z override, scaling
    "c3": dict(dims=(256, 256, 256), zbox=None, scaling="weak",
               name="3D periodic SPH box 256^3 (16 Mi particles) per GPU, fixed h"),
    "c2": dict(dims=(1024, 1024, 1), zbox=1024.0, scaling="weak",
               name="2D periodic SPH sheet 1024x1024 (1 Mi particles) per GPU in a 1024-deep box"),
    "c4": dict(dims=(512, 512, 256), zbox=None, scaling="strong",
               name="3D periodic SPH box 512x512x256 (64 Mi particles) in total"),
    "c5": dict(dims=(512, 512, 256), zbox=None, scaling="weak", build_only=True,
               name="neighbour-list build only (cell list + Morton reorder + neighbour pass), 2^26 particles per GPU"),
    "small": dict(dims=(64, 64, 64), zbox=None, scaling="weak", name="3D periodic SPH box 64^3 per GPU"),
}
H, CUTOFF, TOL, FCUT = 2.0, 2.0, 0.0, 5.0
EOS = (2.0, 0.5, 1.0)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------ clocks
class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples = []
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.samples:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except Exception:
                continue
            for k, nm in enumerate(names):
                if len(f) > 3 + k and f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ synthetic input
def lattice_on_device(dims, x0, device, seed):
    """(i+.5, j+.5, k+.5) + U(-.1,.1), v ~ U(-.05,.05); x fastest; lattice column offset x0."""
    import torch
    nx, ny, nz = dims
    n = nx * ny * nz
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    idx = torch.arange(n, device=device, dtype=torch.int64)
    r = torch.empty((n, 3), dtype=torch.float64, device=device)
    r[:, 0] = (idx % nx + x0).to(torch.float64) + 0.5
    r[:, 1] = ((idx // nx) % ny).to(torch.float64) + 0.5
    r[:, 2] = (idx // (nx * ny)).to(torch.float64) + 0.5
    del idx
    r += (torch.rand((n, 3), dtype=torch.float64, device=device, generator=g) - 0.5) * 0.2
    v = (torch.rand((n, 3), dtype=torch.float64, device=device, generator=g) - 0.5) * 0.1
    return r, v


# ------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's own CPU path (oracle/_ref): ospana.py wiring -- VerletList + c_forces.SpamForce
    (which computes spam_properties itself) driven by SmoothParticleSystem.derivatives()."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    wl = WORKLOADS[args.workload]
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    kind = "reference"
    side = (12, 12, 12) if wl["dims"][2] > 1 else (40, 40, 1)
    from oracle import oracle as O
    r, v, box = O.lattice_workload(side[0], side[1], side[2], seed=SEED)
    if wl["zbox"]:
        box = (box[0], box[1], float(side[0]))
    n = r.shape[0]
    sample = "%dx%dx%d lattice-plus-jitter box (n=%d) of the same workload, list built once, %s" % (
        side[0], side[1], side[2], n, "then timed derivative evaluations (separations + properties + force)")
    step = None
    if os.path.isdir(ref_dir) and any(f.startswith("particles.") for f in os.listdir(ref_dir)):
        sys.path.insert(0, ref_dir)
        import c_forces
        import neighbour_list
        import particles
        p = particles.SmoothParticleSystem(n, d=3, maxn=n, xmax=box[0], ymax=box[1], zmax=box[2],
                                           hshort=H, hlong=2 * H)
        p.r[:, :] = r
        p.v[:, :] = v
        nl = neighbour_list.VerletList(p, cutoff=CUTOFF, tolerance=TOL)
        p.nlists.append(nl)
        p.nl_default = nl
        p.forces.append(c_forces.SpamForce(p, nl))
        nl.build()                     # O(n^2) pure Python, once, untimed (the reference never rebuilds)
        step = p.derivatives
    else:
        kind = "port"
        from oracle import c_oracle as C
        m, h, t = np.ones(n), np.full(n, H), np.ones(n)
        bx = np.array(box)

        def step():
            C.sph_step(r, v, m, h, t, bx, CUTOFF, TOL, FCUT)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = n * args.steps / dt
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
           "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "config": {"workload": wl["name"], "sample": sample},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)


def cpu_baseline():
    """The oracle port (scalar C, 1 core) on a bounded sample of the workload, ~10-20 s."""
    import numpy as np
    from oracle import c_oracle as C
    from oracle import oracle as O
    side = (64, 64, 32)
    r, v, box = O.lattice_workload(*side, seed=SEED)
    n = r.shape[0]
    m, h, t, bx = np.ones(n), np.full(n, H), np.ones(n), np.array(box)
    C.sph_step(r, v, m, h, t, bx, CUTOFF, TOL, FCUT)
    reps, t0 = 0, time.perf_counter()
    while reps < 3 or time.perf_counter() - t0 < 8.0:
        C.sph_step(r, v, m, h, t, bx, CUTOFF, TOL, FCUT)
        reps += 1
    dt = time.perf_counter() - t0
    return {"value": n * reps / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "%dx%dx%d box (n=%d), %d full evaluations (cell-list build + separations + density/EOS "
                      "+ force) of oracle/sph_oracle.c" % (side[0], side[1], side[2], n, reps)}


# ------------------------------------------------------------------ parity gate (the oracle as the checker)
def parity_gate(sim, wl, world, rank, device):
    """Before any timing is reported: one evaluation of the benched system is compared with the CPU oracle
    (oracle/sph_oracle.c) on a sub-box -- a core of up to 48^3 lattice cells plus a margin of two cutoffs, so that
    every core particle and every neighbour of it has its complete neighbourhood inside the sample; with more than
    one GPU the core straddles the slab face between rank 0 and rank 1.  Core particles: neighbour SETS bit-exact,
    rho / p / vdot / udot within 1e-10 (normalised as in the tests).  Returns the `parity` block of the JSON line."""
    import numpy as np
    import torch
    import torch.distributed as dist
    nx, ny, nz = wl["dims"]
    nxl = nx // world if wl["scaling"] == "strong" else nx
    dims = (nxl * world, ny, nz)
    pad = 2.0 * CUTOFF + 0.5
    centre = [float(nxl) if world > 1 else dims[0] / 2.0, dims[1] / 2.0, dims[2] / 2.0]
    half = []
    for d in range(3):
        room = (nxl if (d == 0 and world > 1) else dims[d] / 2.0) - pad - 0.5
        half.append(min(24.0, room) if dims[d] > 1 else None)        # None: a sheet, every z belongs to the core
    if any(h is not None and h < 2.0 for h in half):
        return {"checked": 0, "skipped": "box too small for a sub-box with a margin of two cutoffs"}
    st = sim.owned_state()
    t_in = st["t"].clone()
    sim.evaluate()
    sim.check()
    res = sim.owned_results()
    r = st["r"]
    in_core = torch.ones(r.shape[0], dtype=torch.bool, device=device)
    in_samp = torch.ones(r.shape[0], dtype=torch.bool, device=device)
    for d in range(3):
        if half[d] is None:
            continue
        in_core &= (r[:, d] >= centre[d] - half[d]) & (r[:, d] < centre[d] + half[d])
        in_samp &= (r[:, d] >= centre[d] - half[d] - pad) & (r[:, d] < centre[d] + half[d] + pad)
    idx = torch.nonzero(in_samp).flatten()
    core_idx = torch.nonzero(in_core).flatten()
    mine = {k: st[k][idx].cpu().numpy() for k in ("r", "v", "gid")}
    mine["t"] = t_in[idx].cpu().numpy()
    mine["core"] = in_core[idx].cpu().numpy()
    for k in ("rho", "p", "vdot", "udot"):
        mine[k] = res[k][idx].cpu().numpy()
    mine["rows"] = sim.neighbour_gids(core_idx).cpu().numpy()
    mine["rows_gid"] = st["gid"][core_idx].cpu().numpy()
    parts = [mine]
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, mine)
    if rank != 0:
        return None
    from oracle import c_oracle as C
    cat = {k: np.concatenate([x[k] for x in parts]) for k in mine if k != "rows"}
    width = max(x["rows"].shape[1] for x in parts)
    rows = np.concatenate([np.pad(x["rows"], ((0, 0), (0, width - x["rows"].shape[1])), constant_values=-1)
                           for x in parts])
    order = np.argsort(cat["gid"], kind="stable")
    gid = cat["gid"][order]
    n = gid.shape[0]
    box = np.array([float(dims[0]), float(dims[1]), float(wl["zbox"] or dims[2])])
    ref = C.sph_step(np.ascontiguousarray(cat["r"][order]), np.ascontiguousarray(cat["v"][order]), np.ones(n),
                     np.full(n, H), np.ascontiguousarray(cat["t"][order]), box, CUTOFF, TOL, FCUT, *EOS)
    core = cat["core"][order]
    # neighbour sets of the core particles, as sorted global ids
    iap = ref["iap"].astype(np.int64)
    both = np.concatenate([iap, iap[:, ::-1]])
    both = both[core[both[:, 0]]]
    both = both[np.lexsort((gid[both[:, 1]], both[:, 0]))]
    want_cnt = np.bincount(both[:, 0], minlength=n)[core]
    ro = np.argsort(cat["rows_gid"], kind="stable")
    got = np.sort(np.where(rows[ro] < 0, np.iinfo(np.int64).max, rows[ro]), axis=1)
    got_cnt = (rows[ro] >= 0).sum(axis=1)
    sets_ok = bool(np.array_equal(cat["rows_gid"][ro], gid[core]) and np.array_equal(got_cnt, want_cnt) and
                   np.array_equal(got[got < np.iinfo(np.int64).max], gid[both[:, 1]]))
    errs = {}
    for k in ("rho", "p", "vdot", "udot"):
        a, b = cat[k][order][core], ref[k][core]
        scale = np.maximum(np.abs(b), 1e-3 * np.max(np.abs(b)))
        errs[k] = float(np.max(np.abs(a - b) / scale))
    out = {"checked": int(core.sum()), "sample": int(n), "pairs_checked": int(both.shape[0] // 2),
           "neighbour_sets_equal": sets_ok, "max_rel": max(errs.values()), "rel": errs, "tol": 1e-10,
           "oracle": "oracle/sph_oracle.c on a %s sub-box%s" % (
               "x".join("%d" % (2 * h) if h is not None else "all" for h in half),
               " straddling the rank 0 / rank 1 slab face" if world > 1 else ""),
           "ok": bool(sets_ok and max(errs.values()) < 1e-10)}
    return out


def ncu_traffic(workload, world):
    """dram__bytes_read + dram__bytes_write per launch of each pass from the committed `ncu --set full` capture
    (profiles/r2_traffic.json, regenerated by profiles/summarise.py from the .ncu-rep of the same command); None
    when there is no capture for this workload / GPU count."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as fh:
            doc = json.load(fh)
        return doc.get("%s_g%d" % (workload, world)), doc.get("source")
    except Exception:
        return None, None


def side_run(name, world, rank, device, steps=5, warmup=3):
    """One more BASELINE config, measured like the headline (CUDA events, max over ranks) but shorter: for the
    `configs` block of --all-configs."""
    import torch
    import torch.distributed as dist
    from pyticles_b200 import stepper
    wl = WORKLOADS[name]
    build_only = bool(wl.get("build_only"))
    sim = stepper.make_bench_system(wl, world, rank, device, SEED, H, CUTOFF, TOL, FCUT, EOS)
    for _ in range(warmup):
        sim.evaluate(build_only=build_only)
    sim.check()
    sim.reset_pass_timers()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        sim.evaluate(timed=True, build_only=build_only)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        tm = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm.item())
    st = sim.check()
    passes, pairs = sim.pass_times(), sim.pairs_per_particle()
    n_local, n_total = sim.n_owned, sim.n_total
    peak, _ = peaks()
    alg = {"cells+reorder": 24.0 + 2 * 72.0, "neighbour": 24.0 + 8.0 * pairs, "density": 80.0, "force": 112.0}
    pass_out = {}
    for k, t_ms in passes.items():
        if build_only and k in ("density", "force"):
            continue
        pass_out[k] = {"ms": round(t_ms, 4)}
        if k in alg and t_ms > 0:
            gbs = alg[k] * n_local / (t_ms * 1e-3) / 1e9
            pass_out[k].update(achieved_gbs=round(gbs, 1), frac=round(gbs / peak, 4))
    dom = max((k for k in pass_out if k in alg), key=lambda k: pass_out[k]["ms"])
    out = {"workload": wl["name"], "scaling": wl["scaling"], "particles_total": n_total, "steps": steps,
           "ms_per_step": ms / steps, "value": n_total * steps / (ms * 1e-3),
           "unit": "particles/s (list builds)" if build_only else UNIT,
           "pairs_per_particle": round(pairs, 3), "pairs_per_s": round(pairs * n_total * steps / (ms * 1e-3), 1),
           "tile_fallback": bool(st.flags & 32),
           "roofline": {"pass": dom, "frac": pass_out[dom].get("frac"), "achieved": pass_out[dom].get("achieved_gbs"),
                        "unit": "GB/s", "passes": pass_out}}
    del sim
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------ our arm
_REAL_STDOUT = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version
    line at the first collective when NCCL_DEBUG is set), so everything but our line goes to stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(out):
    f = _REAL_STDOUT or sys.stdout
    f.write(json.dumps(out) + "\n")
    f.flush()


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="skip the oracle parity gate (on by default)")
    ap.add_argument("--all-configs", action="store_true",
                    help="also measure c2, c4 (strong, N > 1) and the c5 build-only point; reported under `configs`")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU path")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    if args.warmup < 3:
        args.warmup = 3

    from pyticles_b200 import stepper
    wl = WORKLOADS[args.workload]
    sim = stepper.make_bench_system(wl, world, rank, device, SEED, H, CUTOFF, TOL, FCUT, EOS)
    n_local, n_total = sim.n_owned, sim.n_total

    for _ in range(args.warmup):
        sim.evaluate()
    sim.check()
    torch.cuda.synchronize()
    parity = None
    if not args.no_check:
        parity = parity_gate(sim, wl, world, rank, device)
        if rank == 0 and not parity.get("ok", True):
            sys.stderr.write("PARITY GATE FAILED: %s\n" % json.dumps(parity))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    sim.reset_pass_timers()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()                 # all ranks enter the timed region together
        torch.cuda.synchronize()
    t0 = time.time()
    e0.record()
    for _ in range(args.steps):
        sim.evaluate(timed=True)
    e1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        tm = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm.item())
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    st = sim.check()
    tile_fallback = bool(st.flags & 32)     # SPH_F_TILE_FALLBACK: the general neighbour kernel did the pass
    passes = sim.pass_times()           # per-pass mean ms over the timed steps (CUDA events)
    pairs = sim.pairs_per_particle()

    e2e = None
    if not args.no_e2e:
        e2e = sim.run_e2e(args.steps, max(1, min(args.warmup, 3)))
        if world > 1:
            tm = torch.tensor([e2e["ms"]], dtype=torch.float64, device=device)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            e2e["ms"] = float(tm.item())

    configs = None
    if args.all_configs:
        configs = {}
        names = ["c2"] + (["c4"] if (world > 1 and 512 % world == 0) else []) + ["c5"]
        for nm in names:
            if nm == args.workload:
                continue
            try:
                configs[nm] = side_run(nm, world, rank, device)
            except Exception as exc:                                # e.g. out of memory next to the headline system
                configs[nm] = {"error": "%s: %s" % (type(exc).__name__, exc)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    traffic, traffic_src = ncu_traffic(args.workload, world)
    value = n_total * args.steps / (ms * 1e-3)
    # algorithmic bytes per particle per launch (SURVEY.md section 8d / DESIGN.md)
    alg = {"cells+reorder": 24.0 + 2 * 72.0, "neighbour": 24.0 + 8.0 * pairs, "density": 80.0, "force": 112.0}
    pass_out = {}
    for k, t_ms in passes.items():
        if k in alg and t_ms > 0:
            gbs = alg[k] * n_local / (t_ms * 1e-3) / 1e9
            pass_out[k] = {"ms": round(t_ms, 4), "alg_bytes_per_particle": round(alg[k], 1),
                           "achieved_gbs": round(gbs, 1), "frac": round(gbs / peak, 4)}
        else:
            pass_out[k] = {"ms": round(t_ms, 4)}
    dom = max((k for k in pass_out if k in alg), key=lambda k: pass_out[k]["ms"])
    step_bytes = (216.0 + 8.0 * pairs) * n_local
    roof = {"bound": "hbm", "kernel": sim.kernel_names[dom], "pass": dom,
            "achieved": pass_out[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
            "frac": pass_out[dom]["frac"],
            "traffic": (traffic or {}).get(dom), "traffic_source": traffic_src if traffic else None,
            "algorithmic_bytes": round(alg[dom] * n_local, 1),
            "peak_source": peak_src,
            "whole_step": {"alg_bytes_per_particle": round(216.0 + 8.0 * pairs, 1),
                           "achieved_gbs": round(step_bytes / (ms / args.steps * 1e-3) / 1e9, 1),
                           "frac": round(step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak, 4)},
            "passes": pass_out, "pairs_per_particle": round(pairs, 3),
            "pairs_per_s": round(pairs * value, 1)}
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
           "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": wl["name"], "particles_total": n_total, "particles_per_gpu": n_local,
                      "h": H, "cutoff": CUTOFF, "tolerance": TOL, "force_cutoff": FCUT,
                      "parallelism": "single GPU" if world == 1 else "x-slab decomposition over %d GPUs, NCCL halo exchange" % world,
                      "l2": "inputs larger than L2 (no flush)" if n_local * 350 > 3 * 126e6 else "working set near L2 size; no flush",
                      "max_nbrs": sim.max_nbrs},
           "roofline": roof, "gpu_launches": sim.launches_per_eval * args.steps}
    out["config"]["tile_fallback"] = tile_fallback
    if parity is not None:
        out["parity"] = parity
    if clocks:
        out["clocks"] = clocks
    if e2e:
        out["e2e"] = {"value": n_total * e2e["steps"] / (e2e["ms"] * 1e-3), "unit": UNIT,
                      "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                      "static_bytes": e2e["static"], "ms_per_step": e2e["ms"] / e2e["steps"],
                      "note": "r, v, t in and rho, p, vdot, udot out every step; m and h (constant between the "
                              "evaluations of a run) uploaded once = static_bytes"}
    if configs:
        out["configs"] = configs
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline()
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

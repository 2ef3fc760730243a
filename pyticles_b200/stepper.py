"""Fused driver of one derivative evaluation of the hot path, as bench.py and long runs use it.

`SphEvaluator.evaluate()` enqueues, on the current stream and without any host sync,
    status reset -> cell list -> Morton gather -> neighbour pass -> density/EOS -> force
for a SmoothParticleSystem, i.e. what SmoothParticleSystem.derivatives() does
(particles.py:544-570) with the list rebuilt at every evaluation.  Neighbour-capacity overflow
is detected from the device status block at the next `check()`.
"""
import torch

from . import _lib, forces, neighbour_list, particles, properties
from .array import parray

# launches of OUR kernels per evaluation: status_reset, bin, scan x3, scatter, cell_sort, gather (8);
# flags_clear, tile_list, nlist (the general kernel behind the tile kernel: launched, returns at once
# unless that one gave up) (3); density, force (2).  The cudaMemset of the cell counters is not counted.
LAUNCHES_PER_EVAL = 13


class SphEvaluator(object):
    kernel_names = {"cells+reorder": "bin_kernel+scan+scatter_kernel+cell_sort_kernel+gather_kernel",
                    "neighbour": "tile_list_kernel", "density": "density_kernel<true>", "force": "force_kernel<true>"}

    def __init__(self, p, nl, force, eos=(2.0, 0.5, 1.0)):
        self.p, self.nl, self.force, self.eos = p, nl, force, eos
        self.n_owned = p.n
        self.n_total = p.n
        self.launches_per_eval = LAUNCHES_PER_EVAL
        self._events = []
        self._planned = False

    @property
    def max_nbrs(self):
        return self.nl.backend.K

    def _plan(self):
        p, nl = self.p, self.nl
        be = nl.backend
        box = (p.box.xmax, p.box.ymax, p.box.zmax)
        be.plan(box, nl.cutoff_radius, nl.tolerance, p.n, p.r)
        be.ensure(p.n)
        self.h_uniform = properties._h_uniform(p, p.h)
        self._planned = True

    def evaluate(self, timed=False, io=None, build_only=False):
        """One derivative evaluation.  `io` (dict of tensors r v m h t rho p pco u vdot udot) selects
        other input/output buffers than the particle system's own (used by the streamed e2e path).
        `build_only`: stop after the neighbour pass (BASELINE configs[4], the list-build sweep)."""
        if not self._planned:
            self._plan()
        p, be = self.p, self.nl.backend
        x = io if io is not None else {k: getattr(p, k) for k in ("r", "v", "m", "h", "t", "rho", "p", "pco", "u",
                                                                    "vdot", "udot")}
        ev = None
        if timed:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            ev[0].record()
        be.cells_and_gather(x["r"], x["v"], x["m"])
        if timed:
            ev[1].record()
        be.nlist()
        if timed:
            ev[2].record()
        if build_only:
            if timed:
                for k in (3, 4):
                    ev[k].record()
                self._events.append(ev)
            return
        be.density_eos(self.eos, x["h"], self.h_uniform, x["rho"], x["p"], x["pco"], x["u"], x["t"])
        if timed:
            ev[3].record()
        be.force(None, None, x["h"], self.h_uniform, self.force.cutoff, 3, x["vdot"], x["udot"], reuse_press=True,
                 first_force=True)      # stores: the zeroing of particles.py:549-550 is folded into the kernel
        if timed:
            ev[4].record()
            self._events.append(ev)

    def check(self):
        """Sync and look at the status block; grow the neighbour capacity if it overflowed."""
        torch.cuda.synchronize()
        be = self.nl.backend
        st = be.status()
        if st.flags & _lib.SPH_F_NBR_OVERFLOW:
            if int(st.max_count) <= be.K:
                raise _lib.SphError("inconsistent neighbour overflow status")
            be.user_max_nbrs = None
            be.ensure(be.n, K=int(st.max_count) + max(4, int(st.max_count) // 8))
            self.evaluate()
            return self.check()
        if st.flags & _lib.SPH_F_OUT_OF_SLAB:
            raise _lib.SphError("a particle left the local cell-layer range")
        self.nl._built_for = (p_ver(self.p.r))
        return st

    def reset_pass_timers(self):
        self._events = []

    def pass_times(self):
        names = ["cells+reorder", "neighbour", "density", "force"]
        tot = dict.fromkeys(names, 0.0)
        for ev in self._events:
            for k, nm in enumerate(names):
                tot[nm] += ev[k].elapsed_time(ev[k + 1])
        k = max(1, len(self._events))
        return {nm: tot[nm] / k for nm in names}

    def pairs_per_particle(self):
        return self.nl.backend.count_links() / 2.0 / max(1, self.n_owned)

    # ------------------------------------------------------------------ read-outs for bench.py's parity gate
    def owned_state(self):
        """r, v, t and the global id (= index) of the owned particles."""
        p, n = self.p, self.p.n
        T = lambda x: x.as_subclass(torch.Tensor)[:n]
        return {"r": T(p.r), "v": T(p.v), "t": T(p.t), "gid": torch.arange(n, dtype=torch.int64, device=p.r.device)}

    def owned_results(self):
        p, n = self.p, self.p.n
        return {k: getattr(p, k).as_subclass(torch.Tensor)[:n] for k in ("rho", "p", "vdot", "udot")}

    def neighbour_gids(self, idx):
        return self.nl.backend.neighbour_rows(idx)

    # ------------------------------------------------------------------ end to end with host buffers
    def run_e2e(self, steps, warmup):
        """Host-buffer path: every step copies that step's r, v, t from pinned host memory to
        the device, evaluates, and copies rho, p, vdot, udot back to pinned host memory.  m and h do not
        change between the derivative evaluations of a run (the reference never touches them after set-up:
        particles.py:122-127,323-343), so they are uploaded once and reported as `static`.  Frames are
        independent, so the three stages run as a pipeline over two device slots (copy-in of frame
        k+1 and copy-out of frame k-1 overlap the kernels of frame k on separate streams)."""
        p = self.p
        ins, static, outs, alls = ("r", "v", "t"), ("m", "h"), ("rho", "p", "vdot", "udot"), \
            ("r", "v", "m", "h", "t", "rho", "p", "pco", "u", "vdot", "udot")
        base = {k: getattr(p, k).as_subclass(torch.Tensor) for k in alls}
        slots = [base, {k: (v if k in static else torch.empty_like(v)) for k, v in base.items()}]
        h_in = {k: torch.empty(base[k].shape, dtype=base[k].dtype, pin_memory=True).copy_(base[k]) for k in ins}
        h_out = {k: torch.empty(base[k].shape, dtype=base[k].dtype, pin_memory=True) for k in outs}
        h2d = sum(h_in[k].numel() * h_in[k].element_size() for k in ins)
        d2h = sum(h_out[k].numel() * h_out[k].element_size() for k in outs)
        s_comp = torch.cuda.current_stream()
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        comp_done = [None, None]
        out_done = [None, None]

        def frame(k):
            sl = slots[k % 2]
            with torch.cuda.stream(s_in):
                if comp_done[k % 2] is not None:
                    s_in.wait_event(comp_done[k % 2])          # slot inputs no longer being read
                for name in ins:
                    sl[name].copy_(h_in[name], non_blocking=True)
                in_done = torch.cuda.Event()
                in_done.record(s_in)
            s_comp.wait_event(in_done)
            if out_done[k % 2] is not None:
                s_comp.wait_event(out_done[k % 2])             # slot outputs already copied out
            self.evaluate(io=sl)
            comp_done[k % 2] = torch.cuda.Event()
            comp_done[k % 2].record(s_comp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(comp_done[k % 2])
                for name in outs:
                    h_out[name].copy_(sl[name], non_blocking=True)
                out_done[k % 2] = torch.cuda.Event()
                out_done[k % 2].record(s_out)

        for k in range(warmup):
            frame(k)
        torch.cuda.synchronize()
        comp_done[:] = [None, None]
        out_done[:] = [None, None]
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record(s_comp)
        s_in.wait_event(t0)
        for k in range(steps):
            frame(k)
        s_comp.wait_stream(s_out)
        s_comp.wait_stream(s_in)
        t1.record(s_comp)
        torch.cuda.synchronize()
        stat = sum(base[k].numel() * base[k].element_size() for k in static)
        return {"ms": t0.elapsed_time(t1), "steps": steps, "h2d": h2d, "d2h": d2h, "static": stat,
                "out": {k: h_out[k] for k in outs}}


def p_ver(t):
    return (t.data_ptr(), t._version)


def make_bench_system(wl, world, rank, device, seed, h, cutoff, tol, fcut, eos):
    """Synthetic lattice-plus-jitter system of bench.py for this rank."""
    if world > 1:
        from . import distributed
        return distributed.make_bench_system(wl, world, rank, device, seed, h, cutoff, tol, fcut, eos)
    import bench
    nx, ny, nz = wl["dims"]
    n = nx * ny * nz
    box = (float(nx), float(ny), float(wl["zbox"] or nz))
    p = particles.SmoothParticleSystem(n, d=3, maxn=n, xmax=box[0], ymax=box[1], zmax=box[2], hshort=h,
                                       hlong=2 * h, device=device)
    r, v = bench.lattice_on_device((nx, ny, nz), 0, device, seed)
    p.r = parray(r)
    p.v = parray(v)
    nl = neighbour_list.VerletList(p, cutoff=cutoff, tolerance=tol)
    nl.defer_status = True
    p.nlists.append(nl)
    p.nl_default = nl
    f = forces.SpamForce(p, nl, cutoff=fcut)
    p.forces.append(f)
    return SphEvaluator(p, nl, f, eos)

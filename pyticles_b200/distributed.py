"""Multi-GPU slab decomposition of the SPH step.

The reference is single process; this is the B200-native scaling of its hot path.  The periodic
box is cut along x into slabs of whole cell layers, one process per GPU (torch.distributed:
nccl on GPUs, gloo in the CPU tests).  Positions keep their GLOBAL coordinates everywhere, so the
reference's minimum image and pair predicate apply unchanged on every rank.

Per derivative evaluation
    A  ghost exchange: the two boundary cell layers of each slab (r, v, m, h, t, global id) go to the
       x-neighbours (ring, periodic)                                     -> all_to_all_single
    1  cell list + neighbour pass + density/EOS over owned + ghost particles (local grid =
       owned layers + one ghost layer each side, sph_grid_restrict_x)
    B  ghost exchange of (p, rho) for the same particles, same order     -> all_to_all_single
    2  force pass; results of owned particles are kept
After integration `migrate()` re-homes particles whose cell layer changed owner.  There is no
other collective on the data path.  A pair is reported by the rank that owns its lower-global-id
member, so the union of the per-rank pair lists is the global i<j set exactly once.

`SlabDecomposition` is pure torch + torch.distributed (device agnostic: it is what the gloo tests
exercise on CPU); `SlabSphEvaluator` adds the CUDA passes.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib

NCOL = 10          # r(3) v(3) m h t gid
C_R, C_V, C_M, C_H, C_T, C_GID = 0, 3, 6, 7, 8, 9


class SlabDecomposition(object):
    def __init__(self, box, cutoff, tolerance, n_total, occ=None, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.box = tuple(float(b) for b in box)
        self.cutoff, self.tolerance, self.n_total = float(cutoff), float(tolerance), int(n_total)
        self.occ = occ
        g = _lib.SphGrid()
        lo = _lib.box3(occ[0]) if occ is not None else None
        hi = _lib.box3(occ[1]) if occ is not None else None
        _lib.check(_lib.load().sph_grid_plan(_lib.box3(self.box), self.cutoff, self.tolerance, self.n_total,
                                             lo, hi, ctypes.byref(g)), "sph_grid_plan")
        self.nc = int(g.nc[0])
        self.inv_w = float(g.inv_w[0])
        W = self.world
        if W > 1 and self.nc < 2 * W + 1:
            raise _lib.SphError("box has %d cell layers along x: too few for %d slabs" % (self.nc, W))
        self.bounds = [(k * self.nc) // W for k in range(W + 1)]
        self.lay0, self.lay1 = self.bounds[self.rank], self.bounds[self.rank + 1]
        # local grid: one ghost layer on each side of the owned layers
        self.slab = ((self.lay0 - 1) % self.nc, (self.lay1 - self.lay0) + 2) if W > 1 else None
        self.left, self.right = (self.rank - 1) % W, (self.rank + 1) % W
        self._halo = None

    # ------------------------------------------------------------------ geometry
    def layer_of(self, x):
        """Global x cell layer, with the kernel's own formula (floor(x * inv_w) mod nc)."""
        return torch.floor(x * self.inv_w).to(torch.int64) % self.nc

    def owner_of_layer(self, layer):
        b = torch.tensor(self.bounds[1:], dtype=torch.int64, device=layer.device)
        return torch.searchsorted(b, layer, right=True)

    # ------------------------------------------------------------------ exchange primitive
    def _exchange(self, rows, dest, splits=None, send_counts=None):
        """Send row k to rank dest[k]; returns (received rows, (send order, send splits, recv splits)).
        `send_counts` (host list, one count per rank) says the rows are already grouped by destination
        rank in rank order: no sort, no histogram, one host round trip (the receive counts) instead of two."""
        W = self.world
        if splits is not None:
            order, send_splits, recv_splits = splits
        elif send_counts is not None:
            order, send_splits = None, [int(c) for c in send_counts]
            counts = torch.tensor(send_splits, dtype=torch.int64, device=rows.device)
            rc = torch.empty_like(counts)
            dist.all_to_all_single(rc, counts, group=self.group)
            recv_splits = rc.tolist()
        else:
            order = torch.argsort(dest, stable=True)
            counts = torch.bincount(dest, minlength=W)
            rc = torch.empty_like(counts)
            dist.all_to_all_single(rc, counts, group=self.group)
            send_splits, recv_splits = counts.tolist(), rc.tolist()
        send = rows.contiguous() if order is None else rows[order].contiguous()
        recv = torch.empty((sum(recv_splits), rows.shape[1]), dtype=rows.dtype, device=rows.device)
        dist.all_to_all_single(recv, send, output_split_sizes=recv_splits, input_split_sizes=send_splits,
                               group=self.group)
        return recv, (order, send_splits, recv_splits)

    def migrate(self, own):
        """Re-home rows (n, NCOL) whose cell layer belongs to another rank."""
        if self.world == 1:
            return own
        dest = self.owner_of_layer(self.layer_of(own[:, C_R]))
        recv, _ = self._exchange(own, dest)
        return recv

    def halo_select(self, own):
        """Indices (into own) and destinations of the boundary-layer particles."""
        return self.halo_select_x(own[:, C_R])

    def halo_select_x(self, x):
        """Same, from the (possibly strided) x-coordinate column.  On CUDA the selection is one
        pass of sph_slab_select; on CPU (gloo tests) it is plain torch."""
        if x.is_cuda:
            n = x.shape[0]
            while True:
                buf = getattr(self, "_sel_buf", None)
                if buf is None or buf.device != x.device:
                    cap = max(4096, int(4.0 * n / max(1, self.lay1 - self.lay0)))
                    buf = self._sel_buf = torch.empty((2, cap), dtype=torch.int32, device=x.device)
                    self._sel_cnt = torch.zeros(2, dtype=torch.int32, device=x.device)
                cap = buf.shape[1]
                _lib.check(_lib.load().sph_slab_select(
                    ctypes.c_void_p(x.data_ptr()), int(x.stride(0)), int(n), self.inv_w, self.nc, int(self.lay0),
                    int(self.lay1 - 1), ctypes.c_void_p(buf[0].data_ptr()), ctypes.c_void_p(buf[1].data_ptr()),
                    int(cap), ctypes.c_void_p(self._sel_cnt.data_ptr()),
                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "sph_slab_select")
                nl, nr = (int(c) for c in self._sel_cnt.tolist())
                if max(nl, nr) <= cap:
                    break
                self._sel_buf = torch.empty((2, 2 * max(nl, nr)), dtype=torch.int32, device=x.device)
            li = torch.sort(buf[0, :nl]).values.to(torch.int64)
            ri = torch.sort(buf[1, :nr]).values.to(torch.int64)
        else:
            layer = self.layer_of(x)
            li = torch.nonzero(layer == self.lay0).flatten()
            ri = torch.nonzero(layer == self.lay1 - 1).flatten()
        # rows grouped by destination rank, in rank order (what all_to_all_single sends)
        parts = [(self.left, li), (self.right, ri)]
        if self.right < self.left:
            parts.reverse()
        idx = torch.cat([parts[0][1], parts[1][1]])
        # the CUDA path sends by counts (self._send_counts); the per-row destinations are for the generic exchange
        dest = None if x.is_cuda else torch.cat([torch.full_like(parts[0][1], parts[0][0]),
                                                 torch.full_like(parts[1][1], parts[1][0])])
        self._send_counts = [0] * self.world
        for rk, sel in parts:
            self._send_counts[rk] += int(sel.shape[0])
        return idx, dest

    def halo_exchange(self, own):
        """Exchange A: returns the ghost rows; remembers the pattern for halo_exchange_again."""
        if self.world == 1:
            self._halo = None
            return own[:0]
        idx, dest = self.halo_select(own)
        ghosts, pattern = self._exchange(own[idx], dest)
        self._halo = (idx, pattern)
        return ghosts

    def halo_exchange_again(self, cols):
        """Exchange B: per-particle columns of the same ghosts, in the same order.  `cols` is an
        (n_own, C) tensor or a list of C per-particle vectors (gathered before stacking, so only
        the boundary particles are touched)."""
        idx, pattern = self._halo if self._halo is not None else (None, None)
        if isinstance(cols, (list, tuple)):
            if self.world == 1:
                return cols[0].new_zeros((0, len(cols)))
            send = torch.stack([c[idx] for c in cols], dim=1)
        else:
            if self.world == 1:
                return cols[:0]
            send = cols[idx]
        out, _ = self._exchange(send, None, splits=pattern)
        return out

    def owns_pair(self, gid_i, gid_j, owned_i, owned_j):
        """A pair belongs to the rank that owns its lower-global-id member."""
        return torch.where(gid_i < gid_j, owned_i, owned_j)


class SlabSphEvaluator(object):
    """bench.py / long-run driver: one rank's share of the distributed derivative evaluation.
    Owned particles live at the front of persistent structure-of-arrays tensors; the ghosts of
    the current evaluation are appended behind them."""
    kernel_names = {"cells+reorder": "bin_kernel+scan+scatter_kernel+cell_sort_kernel+gather_kernel",
                    "neighbour": "tile_list_kernel", "density": "density_kernel<true>", "force": "force_kernel<true>",
                    "halo": "slab_select_kernel + all_to_all_single (NCCL)"}
    ncu_traffic = {}
    IN = ("r", "v", "m", "h", "t")
    OUT = ("rho", "p", "pco", "u", "vdot", "udot")

    def __init__(self, own, box, cutoff, tol, fcut, eos, n_total, device, occ=None):
        from .backend import NeighbourBackend
        self.dec = SlabDecomposition(box, cutoff, tol, n_total, occ=occ)
        self.device = torch.device(device)
        self.box, self.cutoff, self.tol, self.fcut, self.eos = box, cutoff, tol, fcut, eos
        self.S = {}
        self.cap = 0
        self._load_rows(self.dec.migrate(own))
        cnt = torch.tensor([self.n_owned], dtype=torch.int64, device=self.device)
        if self.dec.world > 1:
            dist.all_reduce(cnt)
        self.n_total = int(cnt.item())
        self.be = NeighbourBackend(self.device)
        vol = box[0] * box[1] * box[2]
        rl = (cutoff * cutoff + tol * tol) ** 0.5
        self.be.user_max_nbrs = int(1.35 * 4.18879 * rl ** 3 * self.n_total / vol) + 16
        self.launches_per_eval = 19     # the 13 of one GPU + slab_select, halo pack / unpack (x2), pressure_term (ghosts)
        self._events = []
        self.n_local = self.n_owned

    # ------------------------------------------------------------------ storage
    def _reserve(self, cap, S=None):
        S = self.S if S is None else S
        have = S["m"].shape[0] if "m" in S else 0
        if cap <= have:
            return
        cap = int(cap * 1.06) + 4096
        shapes = dict(r=(cap, 3), v=(cap, 3), m=(cap,), h=(cap,), t=(cap,), gid=(cap,), rho=(cap,), p=(cap,),
                      pco=(cap,), u=(cap,), vdot=(cap, 3), udot=(cap,))
        for k, shp in shapes.items():
            new = torch.zeros(shp, dtype=torch.int64 if k == "gid" else torch.float64, device=self.device)
            old = S.get(k)
            if old is not None:
                new[:old.shape[0]] = old
            S[k] = new
        if S is self.S:
            self.cap = cap

    def _load_rows(self, rows):
        n = rows.shape[0]
        self.cap = 0
        self.S = {}
        self._reserve(n)
        S = self.S
        S["r"][:n], S["v"][:n] = rows[:, C_R:C_R + 3], rows[:, C_V:C_V + 3]
        S["m"][:n], S["h"][:n], S["t"][:n] = rows[:, C_M], rows[:, C_H], rows[:, C_T]
        S["gid"][:n] = rows[:, C_GID].to(torch.int64)
        self.n_owned = int(n)

    def _pack(self, idx, S=None):
        S = self.S if S is None else S
        return make_rows(S["r"][idx], S["v"][idx], S["m"][idx], S["h"][idx], S["t"][idx], S["gid"][idx])

    def rows(self):
        """Owned particles as (n, NCOL) rows (for migrate / checkpoints)."""
        return self._pack(torch.arange(self.n_owned, device=self.device))

    def migrate(self):
        """Re-home owned particles whose cell layer changed owner (call after integration)."""
        self._load_rows(self.dec.migrate(self.rows()))

    @property
    def own_gid(self):
        return self.S["gid"][:self.n_owned]

    @property
    def max_nbrs(self):
        return self.be.K

    # ------------------------------------------------------------------ one evaluation
    def _halo_a(self, S):
        dec, no = self.dec, self.n_owned
        if dec.world == 1:
            dec._halo = None
            return 0
        idx, dest = dec.halo_select_x(S["r"][:no, 0])
        L, st = _lib.load(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        rows = torch.empty((idx.shape[0], NCOL), dtype=torch.float64, device=self.device)
        _lib.check(L.sph_halo_pack(P(idx), int(idx.shape[0]), P(S["r"]), P(S["v"]), P(S["m"]), P(S["h"]), P(S["t"]),
                                   P(S["gid"]), P(rows), st), "sph_halo_pack")
        ghosts, pattern = dec._exchange(rows, dest, send_counts=dec._send_counts)
        dec._halo = (idx, pattern)
        ng = int(ghosts.shape[0])
        self._reserve(no + ng, S)
        _lib.check(L.sph_halo_unpack(P(ghosts), ng, no, P(S["r"]), P(S["v"]), P(S["m"]), P(S["h"]), P(S["t"]),
                                     P(S["gid"]), st), "sph_halo_unpack")
        return ng

    def evaluate(self, timed=False, S=None):
        """One distributed derivative evaluation on the storage slot S (default: the primary one)."""
        dec, be = self.dec, self.be
        S = self.S if S is None else S
        ev = None
        if timed:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
            ev[0].record()
        ng = self._halo_a(S)                                  # A
        no = self.n_owned
        n = no + ng
        r, v, m, h, t = (S[k][:n] for k in self.IN)
        rho, p, pco, u, vdot, udot = (S[k][:n] for k in self.OUT)
        if timed:
            ev[1].record()
        be.plan(self.box, self.cutoff, self.tol, n, slab=dec.slab, occ=dec.occ, n_hint=dec.n_total)
        if be.n != n or not be.K:
            be.ensure(n, K=be.K or None)
        be.cells_and_gather(r, v, m)
        if timed:
            ev[2].record()
        be.nlist()
        if timed:
            ev[3].record()
        vdot.zero_()
        udot.zero_()
        be.density_eos(self.eos, h, True, rho, p, pco, u, t)
        if timed:
            ev[4].record()
        if ng:
            idx, pattern = dec._halo                                                 # B
            L, st = _lib.load(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            P = lambda t: ctypes.c_void_p(t.data_ptr())
            send = torch.empty((idx.shape[0], 2), dtype=torch.float64, device=self.device)
            _lib.check(L.sph_halo_pack2(P(idx), int(idx.shape[0]), P(p), P(rho), P(send), st), "sph_halo_pack2")
            pr, _ = dec._exchange(send, None, splits=pattern)
            _lib.check(L.sph_halo_unpack2(P(pr), int(pr.shape[0]), no, P(p), P(rho), st), "sph_halo_unpack2")
            be.pressure_term(p, rho, no)                      # only the ghosts need their p/rho^2 refreshed
        if timed:
            ev[5].record()
        be.force(p, rho, h, True, self.fcut, 3, vdot, udot, reuse_press=True)
        if timed:
            ev[6].record()
            self._events.append(ev)
        self.n_local = n
        self.result = dict(rho=rho[:no], p=p[:no], pco=pco[:no], u=u[:no], vdot=vdot[:no], udot=udot[:no])

    def check(self):
        torch.cuda.synchronize()
        st = self.be.status()
        if st.flags & _lib.SPH_F_NBR_OVERFLOW:
            if int(st.max_count) <= self.be.K:
                raise _lib.SphError("inconsistent neighbour overflow status")
            self.be.user_max_nbrs = None
            self.be.ensure(self.be.n, K=int(st.max_count) + max(4, int(st.max_count) // 8))
            self.evaluate()
            return self.check()
        if st.flags & _lib.SPH_F_OUT_OF_SLAB:
            raise _lib.SphError("a particle lies outside this rank's cell layers (missing migrate()?)")
        return st

    def reset_pass_timers(self):
        self._events = []

    def pass_times(self):
        names = ["halo", "cells+reorder", "neighbour", "density", "halo_b", "force"]
        tot = dict.fromkeys(names, 0.0)
        for ev in self._events:
            for k, nm in enumerate(names):
                tot[nm] += ev[k].elapsed_time(ev[k + 1])
        k = max(1, len(self._events))
        out = {nm: tot[nm] / k for nm in names}
        out["halo"] += out.pop("halo_b")
        return out

    def pairs_per_particle(self):
        return self.be.count_links() / 2.0 / max(1, self.n_local)

    def local_pairs_global_ids(self):
        """(gid_i, gid_j) of the pairs this rank reports (lower-gid member owned here), sorted.
        Call after evaluate()."""
        no = self.n_owned
        gid = self.S["gid"][:self.n_local]
        iap = self.be.export_pairs().to(torch.int64)
        gi, gj = gid[iap[:, 0]], gid[iap[:, 1]]
        keep = self.dec.owns_pair(gi, gj, iap[:, 0] < no, iap[:, 1] < no)
        lo, hi = torch.minimum(gi, gj)[keep], torch.maximum(gi, gj)[keep]
        key = lo * (self.n_total + 1) + hi
        order = torch.argsort(key)
        return torch.stack([lo[order], hi[order]], dim=1)

    def run_e2e(self, steps, warmup):
        """Host-buffer path: every step copies that step's r, v, m, h, t of the owned particles from
        pinned host memory, evaluates (halo exchanges included), and copies rho, p, vdot, udot back.
        Frames are independent, so copy-in of frame k+1 and copy-out of frame k-1 overlap the kernels
        of frame k (two storage slots, three streams) -- as stepper.SphEvaluator.run_e2e does."""
        no = self.n_owned
        S2 = {}
        self._reserve(self.S["m"].shape[0], S2)
        S2["gid"][:no] = self.S["gid"][:no]
        slots = [self.S, S2]
        h_in = {k: torch.empty(self.S[k][:no].shape, dtype=torch.float64, pin_memory=True).copy_(self.S[k][:no])
                for k in self.IN}
        outs = ("rho", "p", "vdot", "udot")
        h_out = {k: torch.empty(self.S[k][:no].shape, dtype=torch.float64, pin_memory=True) for k in outs}
        h2d = sum(t.numel() * t.element_size() for t in h_in.values())
        d2h = sum(t.numel() * t.element_size() for t in h_out.values())
        s_comp = torch.cuda.current_stream()
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        comp_done, out_done = [None, None], [None, None]

        def frame(k):
            S = slots[k % 2]
            with torch.cuda.stream(s_in):
                if comp_done[k % 2] is not None:
                    s_in.wait_event(comp_done[k % 2])
                for name in self.IN:
                    S[name][:no].copy_(h_in[name], non_blocking=True)
                in_done = torch.cuda.Event()
                in_done.record(s_in)
            s_comp.wait_event(in_done)
            if out_done[k % 2] is not None:
                s_comp.wait_event(out_done[k % 2])
            self.evaluate(S=S)
            S = slots[k % 2]                                   # (evaluate may have regrown the slot)
            comp_done[k % 2] = torch.cuda.Event()
            comp_done[k % 2].record(s_comp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(comp_done[k % 2])
                for name in outs:
                    h_out[name].copy_(S[name][:no], non_blocking=True)
                out_done[k % 2] = torch.cuda.Event()
                out_done[k % 2].record(s_out)

        for k in range(warmup):
            frame(k)
        torch.cuda.synchronize()
        comp_done[:] = [None, None]
        out_done[:] = [None, None]
        if self.dec.world > 1:
            dist.barrier()
            torch.cuda.synchronize()
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
        tot = torch.tensor([h2d, d2h], dtype=torch.int64, device=self.device)
        if self.dec.world > 1:
            dist.all_reduce(tot)
        return {"ms": t0.elapsed_time(t1), "steps": steps, "h2d": int(tot[0]), "d2h": int(tot[1])}


class SlabStepper(object):
    """Time stepping of a slab-decomposed system: what SmoothParticleSystem.update does (particles.py:459-494:
    improved Euler of integrator.py:44-59 over r, v, u, then box.apply, then the scaling thermostat of
    particles.py:450-457) with the particles spread over the ranks.

    `sim` is a SlabSphEvaluator (or anything with its storage protocol: `dec`, `n_owned`, `S` with the owned
    particles in front, `rows()`, `_load_rows(rows)`, `evaluate()`).  The predictor can carry a particle across
    a slab face, and the second stage has to find it on the rank that owns its new cell layer: the stage-1 data
    the corrector needs (start state, first-stage derivatives) travel with the particle through `migrate`, as extra
    columns behind the NCOL state columns.  Like the reference's state vector (particles.py:538-540: zero state
    derivatives for rho, p, pco) the step leaves rho, p, pco at their first-stage values.  The evaluation the
    reference runs before its integrator (particles.py:478: "Is this derivative step required?") is not repeated:
    the integrator's first stage recomputes the same derivatives from the same state."""

    def __init__(self, sim, box_kind="periodic", thermostat_temp=None, eos=(2.0, 0.5, 1.0)):
        if box_kind not in ("periodic", "mirror", "none"):
            raise ValueError("box_kind is 'periodic', 'mirror' or 'none'")
        self.sim, self.box_kind, self.thermostat_temp, self.eos = sim, box_kind, thermostat_temp, eos
        self.steps = 0

    def _migrate_with(self, extra):
        """Re-home the owned particles together with per-particle columns `extra` (n_owned, C)."""
        sim = self.sim
        rows = torch.cat([sim.rows(), extra], dim=1)
        rows = sim.dec.migrate(rows)
        sim._load_rows(rows[:, :NCOL])
        return rows[:, NCOL:]

    def _evaluate(self):
        """One derivative evaluation, settled: a neighbour-capacity overflow is grown and re-evaluated (check())
        before its truncated sums can enter the step."""
        self.sim.evaluate()
        if hasattr(self.sim, "check"):
            self.sim.check()

    def step(self, dt):
        sim = self.sim
        self._evaluate()                                                   # stage 1
        n, S = sim.n_owned, sim.S
        # start state, first-stage derivatives and the fields the step leaves at their first-stage values
        cols = [S["r"][:n], S["v"][:n], S["u"][:n, None], S["vdot"][:n], S["udot"][:n, None],
                S["rho"][:n, None], S["p"][:n, None], S["pco"][:n, None]]
        carry = torch.cat([c.reshape(n, -1) for c in cols], dim=1).clone()
        R0, V0, U0, VD0, UD0, RHO, P, PCO = 0, 3, 6, 7, 10, 11, 12, 13
        # predictor: x = x_start + xdot * dt   (integrator.py:52-53)
        S["r"][:n] = carry[:, R0:R0 + 3] + carry[:, V0:V0 + 3] * dt
        S["v"][:n] = carry[:, V0:V0 + 3] + carry[:, VD0:VD0 + 3] * dt
        carry = self._migrate_with(carry)
        self._evaluate()                                                   # stage 2 at the predicted state
        n, S = sim.n_owned, sim.S
        v1 = S["v"][:n].clone()
        # corrector: x = x_start + (c1 + c2) / 2   (integrator.py:56-59)
        S["r"][:n] = carry[:, R0:R0 + 3] + (carry[:, V0:V0 + 3] + v1) * (0.5 * dt)
        S["v"][:n] = carry[:, V0:V0 + 3] + (carry[:, VD0:VD0 + 3] + S["vdot"][:n]) * (0.5 * dt)
        S["u"][:n] = carry[:, U0] + (carry[:, UD0] + S["udot"][:n]) * (0.5 * dt)
        S["rho"][:n], S["p"][:n], S["pco"][:n] = carry[:, RHO], carry[:, P], carry[:, PCO]
        self._box_apply(S["r"][:n], S["v"][:n])
        if self.thermostat_temp is not None:
            self._thermostat(S, n)
        # final positions decide the owner for the next step; u, rho, p, pco go along
        keep = torch.stack([S["u"][:n], S["rho"][:n], S["p"][:n], S["pco"][:n]], dim=1)
        keep = self._migrate_with(keep)
        n, S = sim.n_owned, sim.S
        S["u"][:n], S["rho"][:n], S["p"][:n], S["pco"][:n] = keep[:, 0], keep[:, 1], keep[:, 2], keep[:, 3]
        self.steps += 1

    def _box_apply(self, r, v):
        box = self.sim.dec.box
        for d in range(3):
            x = r[:, d]
            if self.box_kind == "periodic":                # box.py:35-47: reset to the opposite face, not wrapped
                hi, lo = x > box[d], x < 0
                x[hi] = 0.0
                x[lo] = box[d]
            elif self.box_kind == "mirror":                # box.py:53-73
                hi, lo = x > box[d], x < 0
                x[hi] = box[d]
                x[lo] = 0.0
                v[:, d][hi | lo] *= -1.0

    def _thermostat(self, S, n):
        """particles.py:450-457 with the mean taken over all ranks."""
        acc = torch.stack([S["t"][:n].sum(), torch.tensor(float(n), dtype=torch.float64, device=S["t"].device)])
        if self.sim.dec.world > 1:
            dist.all_reduce(acc, group=self.sim.dec.group)
        tav = acc[0] / acc[1]
        S["t"][:n] *= self.thermostat_temp / tav
        a, _, kb = self.eos
        S["u"][:n] = S["t"][:n] * kb - a * S["rho"][:n]                   # eos.get_vdw_u = vdw_energy (properties.py:46)


def make_rows(r, v, m, h, t, gid):
    n = r.shape[0]
    rows = torch.empty((n, NCOL), dtype=torch.float64, device=r.device)
    rows[:, C_R:C_R + 3] = r
    rows[:, C_V:C_V + 3] = v
    rows[:, C_M], rows[:, C_H], rows[:, C_T] = m, h, t
    rows[:, C_GID] = gid.to(torch.float64)
    return rows


def make_bench_system(wl, world, rank, device, seed, h, cutoff, tol, fcut, eos):
    """This rank's share of bench.py's lattice-plus-jitter box (weak: the box grows along x with
    the number of ranks; strong: the fixed box is cut along x)."""
    import bench
    nx, ny, nz = wl["dims"]
    if wl["scaling"] == "strong":
        if nx % world:
            raise _lib.SphError("strong-scaled workload needs nx divisible by the number of GPUs")
        nxl, nx_tot = nx // world, nx
    else:
        nxl, nx_tot = nx, nx * world
    box = (float(nx_tot), float(ny), float(wl["zbox"] or nz))
    n = nxl * ny * nz
    r, v = bench.lattice_on_device((nxl, ny, nz), rank * nxl, device, seed + rank)
    gid = torch.arange(n, device=device, dtype=torch.int64) + rank * n
    one = torch.ones(n, dtype=torch.float64, device=device)
    rows = make_rows(r, v, one, one * h, one, gid)
    del r, v, one, gid
    occ = ((0.0, 0.0, 0.0), (box[0], box[1], float(nz))) if wl["zbox"] else None
    return SlabSphEvaluator(rows, box, cutoff, tol, fcut, eos, n * world, device, occ=occ)

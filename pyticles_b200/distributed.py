"""Multi-GPU slab decomposition of the SPH step.

The reference is single process; this is the B200-native scaling of its hot path.  The periodic
box is cut along x into slabs of whole cell layers, one process per GPU (torch.distributed:
nccl on GPUs, gloo in the CPU tests).  Positions keep their GLOBAL coordinates everywhere, so the
reference's minimum image and pair predicate apply unchanged on every rank.

Per derivative evaluation (nothing in it waits for the host)
    0  the owned particles are binned into the local cell grid; the same pass lists the particles of the two
       boundary cell layers
    A  ghost exchange: those particles (r, v, m, h, t, global id) go to the two x-neighbours of the ring in
       fixed-capacity buffers with the count in a header row          -> ncclSend/ncclRecv, one group
    1  the ghosts are binned behind them (unused ghost slots fall into a spare cell), then cell list +
       neighbour pass + density/EOS over owned + ghost particles (local grid = owned layers + one ghost layer
       each side, sph_grid_restrict_x); no rows, densities or forces are computed FOR ghosts
    B  ghost exchange of (p, rho) for the same particles in the same order
    2  force pass; results of owned particles are kept
After integration `migrate()` re-homes particles whose cell layer changed owner.  There is no
other collective on the data path.  A pair is reported by the rank that owns its lower-global-id
member, so the union of the per-rank pair lists is the global i<j set exactly once.  Capacity overflows
(neighbour rows, halo buffers) raise flags in the device status block; `check()` settles them COLLECTIVELY:
every rank grows and re-evaluates when any rank overflowed.

`SlabDecomposition` is pure torch + torch.distributed (device agnostic: it is what the gloo tests
exercise on CPU, with the same buffers and ring protocol); `SlabSphEvaluator` adds the CUDA passes.
"""
import ctypes
import os

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
        if W > 1 and self.nc < 2 * W:
            raise _lib.SphError("box has %d cell layers along x: too few for %d slabs of at least two" % (self.nc, W))
        self.bounds = [(k * self.nc) // W for k in range(W + 1)]
        self.lay0, self.lay1 = self.bounds[self.rank], self.bounds[self.rank + 1]
        # local grid: one ghost layer on each side of the owned layers
        self.slab = ((self.lay0 - 1) % self.nc, (self.lay1 - self.lay0) + 2) if W > 1 else None
        self.left, self.right = (self.rank - 1) % W, (self.rank + 1) % W
        self.halo_cap = 0
        self._halo = None

    # ------------------------------------------------------------------ geometry
    def layer_of(self, x):
        """Global x cell layer, with the kernel's own formula (floor(x * inv_w) mod nc)."""
        return torch.floor(x * self.inv_w).to(torch.int64) % self.nc

    def owner_of_layer(self, layer):
        b = torch.tensor(self.bounds[1:], dtype=torch.int64, device=layer.device)
        return torch.searchsorted(b, layer, right=True)

    # ------------------------------------------------------------------ ring exchange (the data-path collective)
    def ring_exchange(self, send_left, send_right, recv_left, recv_right, group=None):
        """send_left goes to the left neighbour, send_right to the right one; recv_left receives what the left
        neighbour sent rightwards, recv_right what the right neighbour sent leftwards.  All four buffers have the
        same fixed size on every rank: one ncclGroup of two sends and two receives, no counts, no host sync.
        With two ranks both neighbours are the same peer; messages are matched in issue order (and by tag on gloo:
        0 travels leftwards, 1 rightwards)."""
        group = self.group if group is None else group
        ops = [dist.P2POp(dist.isend, send_left, self.left, group, 0),
               dist.P2POp(dist.isend, send_right, self.right, group, 1),
               dist.P2POp(dist.irecv, recv_right, self.right, group, 0),
               dist.P2POp(dist.irecv, recv_left, self.left, group, 1)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()

    def all_max(self, values, device):
        """Element-wise maximum of a few host integers over the ranks (capacity decisions are collective)."""
        t = torch.tensor([int(v) for v in values], dtype=torch.int64, device=device)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return [int(v) for v in t.tolist()]

    # ------------------------------------------------------------------ generic exchange (migration)
    def _exchange(self, rows, dest):
        """Send row k to rank dest[k] (all_to_all with counts; used once per time step by migrate, not by the
        derivative evaluation)."""
        W = self.world
        order = torch.argsort(dest, stable=True)
        counts = torch.bincount(dest, minlength=W)
        rc = torch.empty_like(counts)
        dist.all_to_all_single(rc, counts, group=self.group)
        send_splits, recv_splits = counts.tolist(), rc.tolist()
        send = rows[order].contiguous()
        recv = torch.empty((sum(recv_splits), rows.shape[1]), dtype=rows.dtype, device=rows.device)
        dist.all_to_all_single(recv, send, output_split_sizes=recv_splits, input_split_sizes=send_splits,
                               group=self.group)
        return recv

    def migrate(self, own):
        """Re-home rows (n, NCOL) whose cell layer belongs to another rank."""
        if self.world == 1:
            return own
        dest = self.owner_of_layer(self.layer_of(own[:, C_R]))
        return self._exchange(own, dest)

    # ------------------------------------------------------------------ halo exchange in plain torch (CPU tests)
    def halo_select(self, own):
        """Indices (into own) of the particles in the first and in the last owned cell layer."""
        layer = self.layer_of(own[:, C_R])
        return torch.nonzero(layer == self.lay0).flatten(), torch.nonzero(layer == self.lay1 - 1).flatten()

    def _fixed(self, rows, cols):
        """(halo_cap + 1, cols) buffer: header row {count, ...} + the rows."""
        buf = rows.new_zeros((self.halo_cap + 1, cols))
        buf[0, 0] = rows.shape[0]
        buf[1:rows.shape[0] + 1] = rows
        return buf

    def halo_exchange(self, own):
        """Exchange A with the protocol of the CUDA path (fixed-capacity buffers, header count, ring
        send/recv): returns the ghost rows, the left neighbour's first; remembers the lists for
        halo_exchange_again.  The capacity grows collectively when a layer does not fit."""
        if self.world == 1:
            self._halo = None
            return own[:0]
        li, ri = self.halo_select(own)
        need = self.all_max([max(li.shape[0], ri.shape[0])], own.device)[0]
        if need > self.halo_cap:
            self.halo_cap = need + need // 8 + 8
        sl, sr = self._fixed(own[li], own.shape[1]), self._fixed(own[ri], own.shape[1])
        rl, rr = torch.empty_like(sl), torch.empty_like(sr)
        self.ring_exchange(sl, sr, rl, rr)
        cl, cr = int(rl[0, 0]), int(rr[0, 0])
        self._halo = (li, ri, cl, cr)
        return torch.cat([rl[1:cl + 1], rr[1:cr + 1]])

    def halo_exchange_again(self, cols):
        """Exchange B: per-particle columns of the same ghosts, in the same order.  `cols` is an
        (n_own, C) tensor."""
        if self.world == 1:
            return cols[:0]
        li, ri, cl, cr = self._halo
        sl, sr = self._fixed(cols[li], cols.shape[1]), self._fixed(cols[ri], cols.shape[1])
        rl, rr = torch.empty_like(sl), torch.empty_like(sr)
        self.ring_exchange(sl, sr, rl, rr)
        return torch.cat([rl[1:cl + 1], rr[1:cr + 1]])

    def owns_pair(self, gid_i, gid_j, owned_i, owned_j):
        """A pair belongs to the rank that owns its lower-global-id member."""
        return torch.where(gid_i < gid_j, owned_i, owned_j)


def _P(t):
    return ctypes.c_void_p(t.data_ptr())


class SlabSphEvaluator(object):
    """bench.py / long-run driver: one rank's share of the distributed derivative evaluation.
    Owned particles live at the front of persistent structure-of-arrays tensors; behind them sits a ghost
    region of fixed capacity (2 * halo_cap slots), filled by exchange A of the current evaluation."""
    _warned = False
    # "nccl": ncclSend/Recv through torch (one group of two sends and two receives per exchange).  "peer": the packed
    # rows are copied into the neighbours' symmetric-memory buffers by a kernel with remote stores over NVLink and
    # flagged -- 0.11 instead of 0.15 ms per exchange on two B200s, most of either being the wait for the slower
    # neighbour, but the end-to-end path with host buffers LOSES 4-6 ms per evaluation with it (24.4 -> 28.4-30.8 ms,
    # profiles/r2_halo.txt), so NCCL is the default.
    halo_transport = "nccl"
    kernel_names = {"cells+reorder": "bin_kernel+scan+scatter_kernel+cell_sort_kernel+gather_kernel",
                    "neighbour": "tile_list_kernel", "density": "density_kernel<true>", "force": "force_kernel<true>",
                    "halo": "halo_pack/unpack kernels + peer-to-peer copies into the neighbours' symmetric-memory buffers"}
    IN = ("r", "v", "m", "h", "t")
    OUT = ("rho", "p", "pco", "u", "vdot", "udot")

    def __init__(self, own, box, cutoff, tol, fcut, eos, n_total, device, occ=None, halo_cap=None):
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
        # 13 of one GPU + second binning pass (ghosts) + halo pack / unpack (x2) + pressure_term
        self.launches_per_eval = 13 if self.dec.world == 1 else 19
        self._events = []
        self._nvalid = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._halo_buf = None
        self.overlap_b = os.environ.get("SPH_OVERLAP_B", "0") == "1"
        self._group_hp = None
        if self.dec.world > 1:
            if halo_cap is None:
                # a boundary layer holds n_owned / owned layers particles on average
                halo_cap = int(1.25 * self.n_total / self.dec.nc) + 1024
            self._set_halo_cap(self.dec.all_max([halo_cap], self.device)[0])
        self.check_uniform_h()

    # ------------------------------------------------------------------ storage
    @property
    def halo_cap(self):
        return self.dec.halo_cap

    @property
    def n_slots(self):
        """Particle slots the passes run over: owned + the fixed ghost region."""
        return self.n_owned + (2 * self.halo_cap if self.dec.world > 1 else 0)

    def _set_halo_cap(self, cap):
        """(Re)allocate the halo buffers for `cap` rows per side.  COLLECTIVE when there is more than one rank: the
        receive buffers are symmetric memory (torch.distributed._symmetric_memory: every rank maps every rank's
        buffer), so that a neighbour's pack kernel stores its rows straight into them over NVLink and no NCCL call is
        on the path; two generations of them alternate from evaluation to evaluation (a rank can only be one
        evaluation ahead of its neighbours, whose data it needs).  If the rendezvous is not possible the exchange
        falls back to ncclSend/Recv (SlabDecomposition.ring_exchange) and says so once."""
        cap = int(cap)
        self.dec.halo_cap = cap
        dev = self.device
        self._halo_buf = dict(idx=torch.zeros((2, cap), dtype=torch.int32, device=dev),
                              send_a=torch.zeros((2, cap + 1, NCOL), dtype=torch.float64, device=dev),
                              recv_a=torch.zeros((2, cap + 1, NCOL), dtype=torch.float64, device=dev),
                              send_b=torch.zeros((2, cap, 2), dtype=torch.float64, device=dev),
                              recv_b=torch.zeros((2, cap, 2), dtype=torch.float64, device=dev))
        self._symm = None
        if self.halo_transport == "peer" and dev.type == "cuda" and self.dec.world > 1:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                na, nb = (cap + 1) * NCOL, cap * 2                       # doubles per side: exchange A, exchange B
                per_gen = 2 * na + 2 * nb
                t = symm_mem.empty(2 * per_gen, dtype=torch.float64, device=dev)     # never zeroed here: an asynchronous
                # fill could land after a fast neighbour's first rows; every slot is written before it is signalled
                group = self.dec.group if self.dec.group is not None else dist.group.WORLD
                hdl = symm_mem.rendezvous(t, group.group_name)
                # element offsets inside a generation: [A from left | A from right | B from left | B from right]
                self._symm = dict(t=t, hdl=hdl, per_gen=per_gen, off=dict(a_l=0, a_r=na, b_l=2 * na, b_r=2 * na + nb),
                                  na=na, nb=nb, gen=0)
            except Exception as exc:                                     # pragma: no cover (depends on the machine)
                if not SlabSphEvaluator._warned:
                    SlabSphEvaluator._warned = True
                    import sys
                    sys.stderr.write("pyticles_b200: symmetric memory not available (%s: %s); the halo exchange uses "
                                     "ncclSend/Recv\n" % (type(exc).__name__, exc))
                self._symm = None

    def _peer_exchange(self, which, send_l, send_r):
        """One ring exchange over peer memory: the packed rows travelling leftwards are copied into the slot "from
        right" of the LEFT neighbour's receive buffer, the others into the slot "from left" of the RIGHT
        neighbour's -- two coalesced copy kernels with remote stores over NVLink (a pack kernel storing its 8-byte
        columns straight into the remote buffer took 330 us) -- then one flag per direction tells the neighbours the
        rows are there.  Everything is stream ordered, nothing waits on the host.  Returns (rows received from the
        left neighbour, from the right neighbour) as local tensors."""
        sy, dec = self._symm, self.dec
        hdl, o = sy["hdl"], sy["off"]
        base = sy["gen"] * sy["per_gen"]
        n = sy["na"] if which == "a" else sy["nb"]
        key_l, key_r = which + "_l", which + "_r"
        total = 2 * sy["per_gen"]
        remote = lambda rank, key: hdl.get_buffer(rank, (total,), torch.float64)[base + o[key]:base + o[key] + n]
        # an elementwise kernel with remote stores, not a memcpy: the copy engines belong to the host transfers of the
        # end-to-end path (a cudaMemcpyPeer here cost it 4 ms per evaluation)
        torch.mul(send_l.reshape(-1), 1.0, out=remote(dec.left, key_r))
        torch.mul(send_r.reshape(-1), 1.0, out=remote(dec.right, key_l))
        ch = 0 if which == "a" else 2
        hdl.put_signal(dec.left, channel=ch)                  # travelling leftwards
        hdl.put_signal(dec.right, channel=ch + 1)             # travelling rightwards
        hdl.wait_signal(dec.right, channel=ch)                # what my right neighbour sent leftwards
        hdl.wait_signal(dec.left, channel=ch + 1)
        t = sy["t"]
        return t[base + o[key_l]:base + o[key_l] + n], t[base + o[key_r]:base + o[key_r] + n]

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
        self._reserve(n + 2 * self.dec.halo_cap)
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
        self.check_uniform_h()                                # arrivals carry their own h

    def check_uniform_h(self):
        """The slab passes use one smoothing length for every pair (h_uniform): refuse anything else, on every
        rank together, instead of computing with the first local particle's h."""
        h = self.S["h"][:self.n_owned]
        big = torch.finfo(torch.float64).max
        lohi = torch.stack([h.min() if self.n_owned else h.new_tensor(big),
                            -h.max() if self.n_owned else h.new_tensor(big)])
        if self.dec.world > 1:
            dist.all_reduce(lohi, op=dist.ReduceOp.MIN)
        lo, hi = float(lohi[0]), -float(lohi[1])
        if lo != hi:
            raise _lib.SphError("the slab evaluator needs one global smoothing length (h ranges over [%r, %r])" % (lo, hi))
        self.h_global = torch.full((1,), lo, dtype=torch.float64, device=self.device)

    @property
    def own_gid(self):
        return self.S["gid"][:self.n_owned]

    @property
    def max_nbrs(self):
        return self.be.K

    # ------------------------------------------------------------------ one evaluation
    def _fields(self, S):
        f = _lib.SphFields()
        f.r, f.v, f.m, f.h, f.t, f.gid = (_P(S[k]) for k in ("r", "v", "m", "h", "t", "gid"))
        return f

    def evaluate(self, timed=False, S=None, build_only=False):
        """One distributed derivative evaluation on the storage slot S (default: the primary one).  Everything is
        enqueued on the current stream; nothing waits for the host."""
        dec, be = self.dec, self.be
        S = self.S if S is None else S
        multi = dec.world > 1
        no, cap = self.n_owned, self.halo_cap
        n = self.n_slots
        self._reserve(n, S)
        be.plan(self.box, self.cutoff, self.tol, n, slab=dec.slab, occ=dec.occ, n_hint=dec.n_total)
        if be.n != n or not be.K:
            be.ensure(n, K=be.K or None)
        b = be.buf
        b.n_valid = _P(self._nvalid) if multi else ctypes.c_void_p(0)
        b.sort_key = _P(S["gid"]) if multi else ctypes.c_void_p(0)
        b.n_owned = no if multi else 0
        L, g, bp = _lib.load(), ctypes.byref(be.grid), ctypes.byref(be.buf)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        stat = _P(be.status_t)
        r, v, m, t = (S[k][:n] for k in ("r", "v", "m", "t"))
        rho, p, pco, u, vdot, udot = (S[k][:n] for k in self.OUT)
        ev = None
        if timed:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
            ev[0].record()
        fine = [] if timed == "fine" else None           # (label, event) marks for tools/halo_profile.py

        def mark(label):
            if fine is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                fine.append((label, e))

        mark("start")
        _lib.check(L.sph_status_reset(stat, st), "sph_status_reset")
        if multi:
            hb = self._halo_buf
            idx, sa, ra, sb, rb = hb["idx"], hb["send_a"], hb["recv_a"], hb["send_b"], hb["recv_b"]
            f = self._fields(S)
            _lib.check(L.sph_cells_begin(g, bp, _P(r), 0, no, _P(idx[0]), _P(idx[1]), cap, st), "sph_cells_begin")
            if timed:
                ev[1].record()
            mark("bin owned (+ boundary lists)")
            _lib.check(L.sph_halo_pack(ctypes.byref(f), _P(idx[0]), _P(idx[1]), cap, _P(sa[0]), _P(sa[1]), stat, st),
                       "sph_halo_pack")                                                        # A
            mark("A: pack")
            if self._symm is not None:
                self._symm["gen"] ^= 1
                from_l, from_r = self._peer_exchange("a", sa[0], sa[1])
                mark("A: peer copies + signals")
            else:
                dec.ring_exchange(sa[0], sa[1], ra[0], ra[1])
                mark("A: ring send/recv")
                from_l, from_r = ra[0], ra[1]
            _lib.check(L.sph_halo_unpack(ctypes.byref(f), _P(from_l), _P(from_r), cap, no, _P(self._nvalid), stat, st),
                       "sph_halo_unpack")
            mark("A: unpack")
            if timed:
                ev[2].record()
            _lib.check(L.sph_cells_add(g, bp, _P(r), no, n - no, st), "sph_cells_add")
        else:
            _lib.check(L.sph_cells_begin(g, bp, _P(r), 0, n, None, None, 0, st), "sph_cells_begin")
            if timed:
                ev[1].record()
                ev[2].record()
        _lib.check(L.sph_cells_finish(g, bp, st), "sph_cells_finish")
        _lib.check(L.sph_gather(g, bp, _P(r), _P(v), _P(m), st), "sph_gather")
        be.built = False
        mark("bin ghosts, scan, scatter, sort, gather")
        if timed:
            ev[3].record()
        be.nlist()
        mark("neighbour pass")
        if timed:
            ev[4].record()
        if build_only:                                        # BASELINE configs[4]: the list, not the forces
            if timed:
                for k in (5, 6, 7):
                    ev[k].record()
                self._events.append(ev)
            return
        be.density_eos(self.eos, self.h_global, True, rho, p, pco, u, t)
        mark("density / EOS")
        if timed:
            ev[5].record()
        if multi and self.overlap_b:
            # B on a side stream while the force pass runs over every particle that has no ghost neighbour (all but
            # the two boundary cell layers); the boundary layers follow.  Measured on 2 B200s it LOSES: the two force
            # launches cost 0.5 ms more than the one, exchange B only 0.2 ms (profiles/r2_halo.txt) -- off by default.
            main = torch.cuda.current_stream()
            side = self._side_stream()
            done_density = torch.cuda.Event()
            done_density.record(main)
            with torch.cuda.stream(side):
                side.wait_event(done_density)
                self._exchange_b(L, idx, sb, rb, cap, no, p, rho, stat, ctypes.c_void_p(side.cuda_stream))
                done_b = torch.cuda.Event()
                done_b.record(side)
            be.force(p, rho, self.h_global, True, self.fcut, 3, vdot, udot, reuse_press=True, first_force=True, part=1)
            mark("force, all but the boundary layers (exchange B beside it)")
            if timed:
                ev[6].record()
            main.wait_event(done_b)
            mark("wait for exchange B")
            be.force(p, rho, self.h_global, True, self.fcut, 3, vdot, udot, reuse_press=True, first_force=True, part=2)
            mark("force, boundary layers")
        else:
            if multi:                                                                          # B
                self._exchange_b(L, idx, sb, rb, cap, no, p, rho, stat, st, mark)
            if timed:
                ev[6].record()
            be.force(p, rho, self.h_global, True, self.fcut, 3, vdot, udot, reuse_press=True, first_force=True)
            mark("force")
        if timed:
            ev[7].record()
            self._events.append(ev)
        if fine is not None:
            self._fine = getattr(self, "_fine", []) + [fine]
        self.result = dict(rho=rho[:no], p=p[:no], pco=pco[:no], u=u[:no], vdot=vdot[:no], udot=udot[:no])

    def _exchange_b(self, L, idx, sb, rb, cap, no, p, rho, stat, st, mark=lambda label: None):
        """(p, rho) of the boundary-layer particles to the neighbours' ghost slots, same particles and order as A."""
        _lib.check(L.sph_halo_pack2(_P(idx[0]), _P(idx[1]), cap, _P(p), _P(rho), _P(sb[0]), _P(sb[1]), stat, st),
                   "sph_halo_pack2")
        mark("B: pack")
        if self._symm is not None:
            from_l, from_r = self._peer_exchange("b", sb[0], sb[1])
            mark("B: peer copies + signals")
        else:
            self.dec.ring_exchange(sb[0], sb[1], rb[0], rb[1], group=self._group_hp if self.overlap_b else None)
            mark("B: ring send/recv")
            from_l, from_r = rb[0], rb[1]
        _lib.check(L.sph_halo_unpack2(_P(from_l), _P(from_r), cap, no, _P(p), _P(rho), stat, st), "sph_halo_unpack2")
        mark("B: unpack")
        self.be.pressure_term(p, rho, no)                     # only the ghosts need their p/rho^2 refreshed
        mark("B: ghost pressure term")

    def _side_stream(self):
        if self._group_hp is None and self.device.type == "cuda":
            # (COLLECTIVE, once: every rank runs the same evaluate().)  NCCL's own stream has to be a high-priority one
            # too, or the send/recv kernel queues behind the blocks of the force pass.
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            ranks = dist.get_process_group_ranks(self.dec.group) if self.dec.group is not None else None
            self._group_hp = dist.new_group(ranks=ranks, pg_options=opts)
        if getattr(self, "_side", None) is None:
            # high priority: its few small blocks (pack, the NCCL send/recv kernel, unpack) must get SMs while the force
            # pass still has hundreds of thousands of blocks queued -- at equal priority they ran AFTER it (measured)
            self._side = torch.cuda.Stream(device=self.device, priority=-1)
        return self._side

    def check(self, _depth=0):
        """Sync, read the status block and settle capacity overflows COLLECTIVELY: when any rank overflowed
        (neighbour rows or halo buffers) every rank grows to the largest need and re-evaluates, so the ranks
        stay in step with each other's collectives; errors are raised on all ranks together."""
        torch.cuda.synchronize()
        st = self.be.status()
        fl = int(st.flags)
        mine = [1 if fl & _lib.SPH_F_NBR_OVERFLOW else 0, int(st.max_count), int(self.be.K),
                1 if fl & _lib.SPH_F_HALO_OVERFLOW else 0, max(int(st.halo_count[0]), int(st.halo_count[1])),
                1 if fl & _lib.SPH_F_OUT_OF_SLAB else 0]
        nbr_over, need_k, have_k, halo_over, need_h, lost = self.dec.all_max(mine, self.device)
        if lost:
            raise _lib.SphError("a particle lies outside some rank's cell layers (missing migrate()?)")
        if nbr_over or halo_over:
            if _depth >= 4:
                raise _lib.SphError("capacity overflow not settled after %d re-evaluations" % _depth)
            if nbr_over:
                # every rank ends up with the same capacity: the largest need seen anywhere plus a margin
                self.be.user_max_nbrs = None
                self.be.ensure(self.be.n, K=max(have_k, need_k + max(4, need_k // 8)))
            if halo_over:
                if need_h <= self.halo_cap:
                    raise _lib.SphError("inconsistent halo overflow status")
                self._set_halo_cap(need_h + need_h // 8 + 64)
            self.evaluate()
            return self.check(_depth + 1)
        self.ghosts = (int(st.ghost_count[0]), int(st.ghost_count[1]))
        return st

    def reset_pass_timers(self):
        self._events = []
        self._fine = []

    def fine_times(self):
        """Mean ms of every marked sub-step of evaluate(timed="fine") (tools/halo_profile.py)."""
        tot, order = {}, []
        for marks in getattr(self, "_fine", []):
            for (_, e0), (label, e1) in zip(marks, marks[1:]):
                if label not in tot:
                    tot[label] = 0.0
                    order.append(label)
                tot[label] += e0.elapsed_time(e1)
        k = max(1, len(getattr(self, "_fine", [])))
        return [(label, tot[label] / k) for label in order]

    def pass_times(self):
        """Mean ms per pass.  `halo` is exchange A + exchange B (pack, ring send/recv, unpack); with `overlap_b`
        exchange B hides under the interior part of the force pass and shows up in `force` instead."""
        names = ["cells_own", "halo", "cells+reorder", "neighbour", "density", "halo_b", "force"]
        tot = dict.fromkeys(names, 0.0)
        for ev in self._events:
            for k, nm in enumerate(names):
                tot[nm] += ev[k].elapsed_time(ev[k + 1])
        k = max(1, len(self._events))
        out = {nm: tot[nm] / k for nm in names}
        if self.overlap_b:
            out["force"] += out.pop("halo_b")                 # interior part; the boundary part is in "force"
        else:
            out["halo"] += out.pop("halo_b")
        out["cells+reorder"] += out.pop("cells_own")
        return out

    def links(self):
        """Directed neighbour links in the rows of the OWNED particles (ghosts have no rows)."""
        return self.be.count_links()

    def pairs_per_particle(self):
        """Global pairs per particle: every pair is listed once from each member, by the member's owner."""
        tot = torch.tensor([self.links(), self.n_owned], dtype=torch.int64, device=self.device)
        if self.dec.world > 1:
            dist.all_reduce(tot)
        return float(tot[0]) / 2.0 / max(1, int(tot[1]))

    # ------------------------------------------------------------------ read-outs for bench.py's parity gate
    def owned_state(self):
        no = self.n_owned
        return {k: self.S[k][:no] for k in ("r", "v", "t", "gid")}

    def owned_results(self):
        return {k: v for k, v in self.result.items() if k in ("rho", "p", "vdot", "udot")}

    def neighbour_gids(self, idx):
        """Neighbour rows of the owned particles `idx` (local indices) as global ids, -1 padded."""
        rows = self.be.neighbour_rows(idx)
        gid = self.S["gid"][:self.n_slots]
        return torch.where(rows >= 0, gid[rows.clamp(min=0)], rows)

    def local_pairs_global_ids(self):
        """(gid_i, gid_j) of the pairs this rank reports (lower-gid member owned here), sorted.
        Call after evaluate() + check()."""
        no = self.n_owned
        gid = self.S["gid"][:self.n_slots]
        iap = self.be.export_pairs().to(torch.int64)
        gi, gj = gid[iap[:, 0]], gid[iap[:, 1]]
        keep = self.dec.owns_pair(gi, gj, iap[:, 0] < no, iap[:, 1] < no)
        lo, hi = torch.minimum(gi, gj)[keep], torch.maximum(gi, gj)[keep]
        key = lo * (self.n_total + 1) + hi
        order = torch.argsort(key)
        return torch.stack([lo[order], hi[order]], dim=1)

    def run_e2e(self, steps, warmup):
        """Host-buffer path: every step copies that step's r, v, t of the owned particles from pinned host
        memory, evaluates (halo exchanges included), and copies rho, p, vdot, udot back.  m and h do not
        change between the derivative evaluations of a run (the reference never touches them after set-up),
        so they are uploaded once and reported as `static`.  Frames are independent, so copy-in of frame k+1
        and copy-out of frame k-1 overlap the kernels of frame k (two storage slots, three streams) -- as
        stepper.SphEvaluator.run_e2e does."""
        no = self.n_owned
        S2 = {}
        self._reserve(self.S["m"].shape[0], S2)
        ins, static, outs = ("r", "v", "t"), ("m", "h"), ("rho", "p", "vdot", "udot")
        for k in static + ("gid",):
            S2[k][:no] = self.S[k][:no]
        slots = [self.S, S2]
        h_in = {k: torch.empty(self.S[k][:no].shape, dtype=torch.float64, pin_memory=True).copy_(self.S[k][:no])
                for k in ins}
        h_out = {k: torch.empty(self.S[k][:no].shape, dtype=torch.float64, pin_memory=True) for k in outs}
        h2d = sum(t.numel() * t.element_size() for t in h_in.values())
        d2h = sum(t.numel() * t.element_size() for t in h_out.values())
        stat = sum(self.S[k][:no].numel() * 8 for k in static)
        s_comp = torch.cuda.current_stream()
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        comp_done, out_done = [None, None], [None, None]

        def frame(k):
            S = slots[k % 2]
            with torch.cuda.stream(s_in):
                if comp_done[k % 2] is not None:
                    s_in.wait_event(comp_done[k % 2])
                for name in ins:
                    S[name][:no].copy_(h_in[name], non_blocking=True)
                in_done = torch.cuda.Event()
                in_done.record(s_in)
            s_comp.wait_event(in_done)
            if out_done[k % 2] is not None:
                s_comp.wait_event(out_done[k % 2])
            self.evaluate(S=S)
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
        tot = torch.tensor([h2d, d2h, stat], dtype=torch.int64, device=self.device)
        if self.dec.world > 1:
            dist.all_reduce(tot)
        return {"ms": t0.elapsed_time(t1), "steps": steps, "h2d": int(tot[0]), "d2h": int(tot[1]),
                "static": int(tot[2]), "out": {k: h_out[k] for k in outs}}


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

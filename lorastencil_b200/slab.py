"""Multi-GPU slab decomposition: one process per GPU, nearest-neighbour halo exchange per launch
(2-D / 3-D) or per temporal block (1-D).

New functionality (the reference is single-GPU: no cudaSetDevice / NCCL / MPI anywhere in src/).  The
grid is cut along its OUTERMOST axis into `world` contiguous slabs; every rank keeps the reference's
two ping-pong buffers for its slab, padded with the reference's storage halo (4 elements / 4 rows /
1 plane -- S1 of SURVEY.md section 8a), which is >= the stencil radius (4 / 3 / 1).

Per launch i (src = buf[i%2], dst = buf[(i+1)%2]) a rank
  1. computes its two edge bands (the first and last `halo` interior rows of dst) on the comm stream,
  2. posts send/recv of those bands with its neighbours (torch.distributed P2P = ncclSend/ncclRecv over
     NVLink) on the comm stream -- the bands land directly in the neighbours' halo rows of dst,
  3. computes the rest of the interior on the main stream, overlapping the exchange,
  4. joins the two streams.
Halo rows on the outer faces of the global grid are never written, so they keep the reference's
semantics (S2): caller's halo in buf[0], zeros in buf[1].  Results are bit-identical to a single-GPU
run because every cell sees the same operands in the same order.

1-D with temporal blocking (tb launches fused per sweep, lorastencil_b200/csrc/stencil1d_tb.cu): sides
that face a neighbour carry a GHOST zone of 4*tb_max cells instead of the 4-cell halo.  A fused launch
reads up to 4*tb cells beyond the slab, so one exchange of 4*tb_max cells per temporal block replaces tb
exchanges of 4 cells; sides that face the end of the global line keep the 4-cell halo, which the kernel
treats as virtual (caller's halo at even times, zero at odd times).

The compute step is pluggable (`step_fn(src, dst, lo, hi)`, `fused_fn(...)`): the product uses `Plan.step`
/ `Plan.step_fused` (CUDA); the CPU tests (gloo, world_size 2-3) inject the oracle to exercise the
partition / exchange / scheduling logic.
"""
from __future__ import annotations

import os

import numpy as np

from .plan import DEFAULT_TB_1D, HALO, MAX_TB_1D


def temporal_schedule(times: int, max_tb: int):
    """Temporal blocks for `times` launches (same rule as lora_plan_run): blocks of max_tb, the remainder,
    and one block split in two when the number of fused launches would not have the parity of `times`
    (the result has to land in buf[times % 2], S3)."""
    tbs, left = [], times
    while left > 0:
        t = min(left, max_tb)
        tbs.append(t)
        left -= t
    if len(tbs) % 2 != times % 2:
        for i in range(len(tbs) - 1, -1, -1):
            if tbs[i] >= 2:
                a = tbs[i] // 2
                tbs[i:i + 1] = [a, tbs[i] - a]
                break
    return tbs


class SlabGeometry:
    """Contiguous balanced split of the outermost interior axis (multiples of `align` except the tail).
    `ghost` = cells kept beyond the slab on a side that faces a neighbour (>= halo)."""

    def __init__(self, dims, world: int, rank: int, align: int = 1, ghost: int | None = None):
        self.dims = tuple(int(d) for d in dims)
        self.dim = len(self.dims)
        self.world, self.rank = world, rank
        self.halo = HALO[self.dim][0]
        n0 = self.dims[0]
        units = -(-n0 // align)  # balanced split in units of `align` cells (the last unit may be partial)
        self.bounds = [min(n0, align * (units * r // world)) for r in range(world + 1)]
        self.lo, self.hi = self.bounds[rank], self.bounds[rank + 1]
        self.prev = rank - 1 if rank > 0 else None
        self.next = rank + 1 if rank < world - 1 else None
        g = self.halo if ghost is None else ghost
        self.wl = g if self.prev is not None else self.halo   # cells stored left of the slab
        self.wr = g if self.next is not None else self.halo
        if self.hi - self.lo < max(self.wl, self.wr):
            raise ValueError(f"slab of {self.hi - self.lo} is thinner than its halo/ghost zone: use fewer ranks")
        self.slab = self.hi - self.lo
        # what the plan sees: an array whose 'interior' also covers the ghost cells beyond the halo width
        self.off = self.wl - self.halo                        # first slab cell in plan-interior coordinates
        self.local_dims = (self.slab + self.off + (self.wr - self.halo),) + self.dims[1:]
        self.local_padded = tuple(d + 2 * h for d, h in zip(self.local_dims, HALO[self.dim]))

    def global_rows(self):
        """Rows of the GLOBAL padded array this rank's padded buffer mirrors."""
        return slice(self.lo + self.halo - self.wl, self.hi + self.halo + self.wr)


class SlabRunner:
    def __init__(self, shape: str, global_dims, params=None, mode: int = 0, group=None, device=None, step_fn=None,
                 fused_fn=None, temporal_block: int | None = None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.shape = shape
        dim = len(global_dims)
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.cuda = self.device.type == "cuda"
        injected = step_fn is not None
        if temporal_block is None:
            temporal_block = int(os.environ.get("LORA_TB", str(DEFAULT_TB_1D))) if (dim == 1 and (not injected or fused_fn)) else 1
        self.max_tb = max(1, min(MAX_TB_1D, temporal_block)) if dim == 1 else 1
        if injected and fused_fn is None:
            self.max_tb = 1
        ghost = 4 * self.max_tb if (dim == 1 and self.max_tb > 1) else None
        self.geo = SlabGeometry(global_dims, self.world, self.rank, align=16 if dim == 1 else 1, ghost=ghost)
        if not injected:
            from .plan import Plan
            self.plan = Plan(shape, self.geo.local_dims, params=params, mode=mode)
            step_fn, fused_fn = self.plan.step, self.plan.step_fused
        else:
            self.plan = None
        self.step_fn, self.fused_fn = step_fn, fused_fn
        self.buf = [torch.zeros(self.geo.local_padded, dtype=torch.float64, device=self.device) for _ in range(2)]
        self.launch = 0   # kernel sweeps issued: the result sits in buf[launch % 2]
        self.time = 0     # time steps applied
        if self.cuda:
            self.comm_stream = torch.cuda.Stream(device=self.device)
            self.ev_main = torch.cuda.Event()
            self.ev_comm = torch.cuda.Event()

    # ---- data movement helpers (tests / parity; not on the timed path) ----
    def load_global(self, a_global: np.ndarray):
        """Every rank takes its slab (with halo / ghost rows) out of the same global padded array."""
        t = self.torch.from_numpy(np.ascontiguousarray(a_global[self.geo.global_rows()]))
        self.buf[0].copy_(t)
        self.buf[1].zero_()
        self.launch = self.time = 0

    def result(self):
        return self.buf[self.launch % 2]

    def gather_global(self, a_global_shape):
        """Rank 0 reassembles the global padded result (interior rows from their owners, outer halo
        rows from the end ranks)."""
        dist = self.dist
        h = self.geo.halo
        mine = (self.result().cpu().numpy(), self.geo.wl, self.geo.wr)
        pieces = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(pieces, mine, group=self.group)
        else:
            pieces = [mine]
        out = np.zeros(a_global_shape, dtype=np.float64)
        for r, (p, wl, wr) in enumerate(pieces):
            lo, hi = self.geo.bounds[r], self.geo.bounds[r + 1]
            out[lo + h:hi + h] = p[wl:wl + hi - lo]
            if r == 0:
                out[:h] = p[:h]
            if r == self.world - 1:
                out[hi + h:] = p[wl + hi - lo:]
        return out

    # ---- the timed path ----
    def _exchange(self, dst):
        """My first / last `w` slab rows go to the neighbours' ghost rows; theirs arrive in mine."""
        dist = self.dist
        g = self.geo
        ops = []
        if g.prev is not None:
            ops.append(dist.P2POp(dist.isend, dst[g.wl:2 * g.wl], g.prev, self.group))
            ops.append(dist.P2POp(dist.irecv, dst[0:g.wl], g.prev, self.group))
        if g.next is not None:
            e = g.wl + g.slab
            ops.append(dist.P2POp(dist.isend, dst[e - g.wr:e], g.next, self.group))
            ops.append(dist.P2POp(dist.irecv, dst[e:e + g.wr], g.next, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def _sweep(self, tb: int):
        """One kernel sweep of `tb` time steps over the slab + exchange of the edge bands."""
        g = self.geo
        src, dst = self.buf[self.launch % 2], self.buf[(self.launch + 1) % 2]
        lo, hi = g.off, g.off + g.slab  # the slab in plan-interior coordinates

        if self.max_tb > 1:
            def compute(a, b, stream=None):
                kw = {} if stream is None else {"stream": stream}
                self.fused_fn(src, dst, self.buf[0], a, b, tb, self.time, g.prev is None, g.next is None, **kw)
        else:
            def compute(a, b, stream=None):
                kw = {} if stream is None else {"stream": stream}
                self.step_fn(src, dst, a, b, **kw)

        if self.world == 1 or not self.cuda:
            compute(lo, hi)
            if self.world > 1:
                self._exchange(dst)
        else:
            torch = self.torch
            main = torch.cuda.current_stream(self.device)
            self.ev_main.record(main)
            self.comm_stream.wait_event(self.ev_main)  # src is complete (previous sweep + its ghost rows)
            top = min(lo + g.wl, hi) if g.prev is not None else lo
            bot = max(hi - g.wr, top) if g.next is not None else hi
            with torch.cuda.stream(self.comm_stream):
                if top > lo:
                    compute(lo, top, self.comm_stream)
                if bot < hi:
                    compute(bot, hi, self.comm_stream)
                self._exchange(dst)
                self.ev_comm.record(self.comm_stream)
            if bot > top:
                compute(top, bot, main)
            main.wait_event(self.ev_comm)
        self.launch += 1
        self.time += tb

    def step(self):
        self._sweep(1)

    def run(self, times: int):
        """`times` launches of the reference operator; the result is in buf[times % 2] like the reference's."""
        if self.world == 1 and self.plan is not None and self.launch % 2 == 0 and self.time % 2 == 0:
            res = self.plan.run(self.buf[0], self.buf[1], times)  # whole line on one device: the plan schedules it
            self.launch += times
            self.time += times
            return res
        if self.max_tb > 1:
            assert self.launch % 2 == self.time % 2, "fused runs must start from a parity-consistent state"
            for tb in temporal_schedule(times, self.max_tb):
                self._sweep(tb)
        else:
            for _ in range(times):
                self._sweep(1)
        return self.result()

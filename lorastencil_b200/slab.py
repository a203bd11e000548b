"""Multi-GPU slab decomposition: one process per GPU, nearest-neighbour halo exchange per launch.

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

The compute step is pluggable (`step_fn(src, dst, lo, hi)`): the product uses `Plan.step` (CUDA); the
CPU tests (gloo, world_size 2) inject the oracle to exercise the partition / exchange logic.
"""
from __future__ import annotations

import numpy as np

from .plan import HALO


class SlabGeometry:
    """Contiguous balanced split of the outermost interior axis (multiples of `align` except the tail)."""

    def __init__(self, dims, world: int, rank: int, align: int = 1):
        self.dims = tuple(int(d) for d in dims)
        self.dim = len(self.dims)
        self.world, self.rank = world, rank
        self.halo = HALO[self.dim][0]
        n0 = self.dims[0]
        per = -(-n0 // world)
        per = -(-per // align) * align
        self.bounds = [min(n0, r * per) for r in range(world + 1)]
        self.lo, self.hi = self.bounds[rank], self.bounds[rank + 1]
        if self.hi - self.lo < self.halo:
            raise ValueError(f"slab of {self.hi - self.lo} is thinner than the halo {self.halo}: use fewer ranks")
        self.local_dims = (self.hi - self.lo,) + self.dims[1:]
        self.local_padded = tuple(d + 2 * h for d, h in zip(self.local_dims, HALO[self.dim]))
        self.prev = rank - 1 if rank > 0 else None
        self.next = rank + 1 if rank < world - 1 else None

    def global_rows(self):
        """Rows of the GLOBAL padded array this rank's padded buffer mirrors."""
        return slice(self.lo, self.hi + 2 * self.halo)


class SlabRunner:
    def __init__(self, shape: str, global_dims, params=None, mode: int = 0, group=None, device=None, step_fn=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.shape = shape
        self.geo = SlabGeometry(global_dims, self.world, self.rank, align=4 if len(global_dims) == 1 else 1)
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.cuda = self.device.type == "cuda"
        if step_fn is None:
            from .plan import Plan
            self.plan = Plan(shape, self.geo.local_dims, params=params, mode=mode)
            step_fn = self.plan.step
        else:
            self.plan = None
        self.step_fn = step_fn
        self.buf = [torch.zeros(self.geo.local_padded, dtype=torch.float64, device=self.device) for _ in range(2)]
        self.launch = 0
        if self.cuda:
            self.comm_stream = torch.cuda.Stream(device=self.device)
            self.ev_main = torch.cuda.Event()
            self.ev_comm = torch.cuda.Event()

    # ---- data movement helpers (tests / parity; not on the timed path) ----
    def load_global(self, a_global: np.ndarray):
        """Every rank takes its slab (with halo rows) out of the same global padded array."""
        t = self.torch.from_numpy(np.ascontiguousarray(a_global[self.geo.global_rows()]))
        self.buf[0].copy_(t)
        self.buf[1].zero_()
        self.launch = 0

    def result(self):
        return self.buf[self.launch % 2]

    def gather_global(self, a_global_shape):
        """Rank 0 reassembles the global padded result (interior rows from their owners, outer halo
        rows from the end ranks)."""
        dist, torch = self.dist, self.torch
        h = self.geo.halo
        mine = self.result().cpu()
        pieces = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(pieces, mine.numpy(), group=self.group)
        else:
            pieces = [mine.numpy()]
        out = np.zeros(a_global_shape, dtype=np.float64)
        for r, p in enumerate(pieces):
            lo, hi = self.geo.bounds[r], self.geo.bounds[r + 1]
            out[lo + h:hi + h] = p[h:h + hi - lo]
            if r == 0:
                out[:h] = p[:h]
            if r == self.world - 1:
                out[hi + h:] = p[h + hi - lo:]
        return out

    # ---- the timed path ----
    def _exchange(self, dst):
        dist = self.dist
        g, h = self.geo, self.geo.halo
        L = g.local_dims[0]
        ops = []
        if g.prev is not None:
            ops.append(dist.P2POp(dist.isend, dst[h:2 * h], g.prev, self.group))
            ops.append(dist.P2POp(dist.irecv, dst[0:h], g.prev, self.group))
        if g.next is not None:
            ops.append(dist.P2POp(dist.isend, dst[L:L + h], g.next, self.group))
            ops.append(dist.P2POp(dist.irecv, dst[L + h:L + 2 * h], g.next, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def step(self):
        g, h = self.geo, self.geo.halo
        L = g.local_dims[0]
        src, dst = self.buf[self.launch % 2], self.buf[(self.launch + 1) % 2]
        if self.world == 1:
            self.step_fn(src, dst, 0, L)
        elif not self.cuda:
            self.step_fn(src, dst, 0, L)
            self._exchange(dst)
        else:
            torch = self.torch
            main = torch.cuda.current_stream(self.device)
            self.ev_main.record(main)
            self.comm_stream.wait_event(self.ev_main)  # src is complete (previous launch + its halos)
            top = min(h, L)
            bot = max(L - h, top)
            with torch.cuda.stream(self.comm_stream):
                self.step_fn(src, dst, 0, top, stream=self.comm_stream)
                if bot < L:
                    self.step_fn(src, dst, bot, L, stream=self.comm_stream)
                self._exchange(dst)
                self.ev_comm.record(self.comm_stream)
            if bot > top:
                self.step_fn(src, dst, top, bot, stream=main)
            main.wait_event(self.ev_comm)
        self.launch += 1

    def run(self, times: int):
        for _ in range(times):
            self.step()
        return self.result()

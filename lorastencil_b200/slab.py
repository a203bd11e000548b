"""Multi-GPU slab decomposition: one process per GPU, nearest-neighbour halo exchange once per sweep (= once per
launch, or once per temporal block where launches are fused: 1-D, 2-D star forms).

New functionality (the reference is single-GPU: no cudaSetDevice / NCCL / MPI anywhere in src/).  The
grid is cut along its OUTERMOST axis into `world` contiguous slabs; every rank keeps the reference's
two ping-pong buffers for its slab, padded with the reference's storage halo (4 elements / 4 rows /
1 plane -- S1 of SURVEY.md section 8a), which is >= the stencil radius (4 / 3 / 1).

Per sweep i (src = buf[i%2], dst = buf[(i+1)%2]) a rank
  1. computes its two edge bands (the first and last ghost-width interior rows of dst) on the comm stream,
  2. gets those bands into the neighbours' ghost rows of dst:
       "p2p"  (default on GPUs): the whole sweep is ONE kernel launch inside the library (NativeSlab, csrc/slab.cu):
              its band tasks run first and store every cell a second time into the neighbour's buffer mapped over
              NVLink (CUDA IPC), the last band task raises a 64-bit flag in the neighbour's memory, the interior
              overlaps all that, and the next sweep's launch waits for the flags in stream order.  No communication
              library on the data path, no Python per sweep;
       "nccl" (LORA_HALO=nccl, and the CPU tests with gloo): send/recv of the bands (torch.distributed P2P),
  3. computes the rest of the interior on the main stream -- it reads the rank's own rows only and never waits
     for a neighbour,
  4. joins the two streams.
Halo rows on the outer faces of the global grid are never written, so they keep the reference's
semantics (S2): caller's halo in buf[0], zeros in buf[1].  Results are bit-identical to a single-GPU
run because every cell sees the same operands in the same order.

Temporal blocking: sides that face a neighbour carry a GHOST zone of radius x tb_max (1-D: 4 x 15 = 60 cells,
lorastencil_b200/csrc/stencil1d_tb.cu; 2-D cross / diamond forms: 3 x 3 = 9 rows, stencil2d_tb.cu) instead of the
storage halo.  A fused launch reads up to radius x tb cells beyond the slab, so one exchange per temporal block
replaces tb exchanges; sides that face the end of the global grid keep the storage halo, which the kernels treat as
virtual (caller's halo at even times, zero at odd times).

The compute step is pluggable (`step_fn(src, dst, lo, hi)`, `fused_fn(...)`): the product uses `Plan.step`
/ `Plan.step_fused` (CUDA); the CPU tests (gloo, world_size 2-3) inject the oracle to exercise the
partition / exchange / scheduling logic.
"""
from __future__ import annotations

import os

import numpy as np

from . import _lib
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


def temporal_schedule_2d(times: int, max_tb: int):
    """2-D fuses 3 launches (cross form), 2 (diamond / pyramid) or none.  Sweeps of 3, then the remainder one by one:
    every sweep advances an odd number of steps, so sweep k reads buf[k % 2] at a time of parity k % 2 (the source
    buffer's own halo ring is the right one for level 0) and the result lands in buf[times % 2] (S3)."""
    if max_tb >= 3:
        return [3] * (times // 3) + [1] * (times % 3)
    if max_tb == 2 and times >= 4:
        # sweeps of two launches (diamond / pyramid forms, native slab driver only): an even number of them, so that the
        # data is back in buffer 0 when the remaining launches run one by one (csrc/slab.cu: schedule_for)
        a = times // 2
        a -= a % 2
        return [2] * a + [1] * (times - 2 * a)
    return [1] * times


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


class _DevMem:
    """Zero-copy torch view of device memory the C library owns (lora_peer_alloc)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class NativeSlab:
    """This rank's slab inside the library (csrc/slab.cu, lora_slab_*): buffers, flags, ghost-zone geometry and the
    sweep loop are native; Python only carries the three CUDA-IPC handles to the neighbours (torch.distributed
    all_gather_object) and calls lora_slab_run.  One kernel launch per sweep: band tasks first, their cells stored a
    second time into the neighbour's ghost zone over NVLink, a flag raised there by the last band task."""

    def __init__(self, runner, temporal_block: int):
        import ctypes
        import torch
        self.L = _lib.lib()
        self.torch = torch
        dist = runner.dist
        g = runner.geo
        self._h = ctypes.c_void_p()
        d = (ctypes.c_longlong * 3)(*g.dims, *([0] * (3 - g.dim)))
        p = runner.params
        pp = None if p is None else p.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        rc = self.L.lora_slab_create(ctypes.byref(self._h), _lib.SHAPE_IDS[runner.shape], int(runner.mode), pp, d,
                                     runner.world, runner.rank, int(temporal_block))
        handles = None
        if rc == 0:
            hb = ctypes.create_string_buffer(192)
            if self.L.lora_slab_export(self._h, hb) == 0:
                handles = hb.raw
        self.err = None if handles is not None else self.L.lora_last_error().decode()
        everyone = [None] * runner.world
        if runner.world > 1:
            dist.all_gather_object(everyone, handles, group=runner.group)
        else:
            everyone = [handles]
        if any(h is None for h in everyone):  # collective decision: all ranks give up together
            self.destroy()
            raise _lib.LoraError(f"lora_slab_create / export failed on some rank ({self.err})")
        ok = True
        for side, r in ((0, g.prev), (1, g.next)):
            if r is not None and self.L.lora_slab_connect_ipc(self._h, side, ctypes.create_string_buffer(everyone[r], 192)) != 0:
                self.err = self.L.lora_last_error().decode()
                ok = False
        self.ok = ok
        info = (ctypes.c_longlong * 10)()
        _lib.check(self.L.lora_slab_info(self._h, info), "lora_slab_info")
        assert (info[0], info[1], info[2], info[3], info[4]) == (g.lo, g.hi, g.wl, g.wr, g.off), "native and Python slab geometry differ"
        assert tuple(info[5:5 + g.dim]) == tuple(g.local_padded)
        self.max_tb = int(info[8])
        self.buf = [torch.as_tensor(_DevMem(self.L.lora_slab_buffer(self._h, i), g.local_padded, "<f8"), device=runner.device)
                    for i in range(2)]

    def run(self, times: int, stream):
        from ctypes import c_void_p
        _lib.check(self.L.lora_slab_run(self._h, int(times), c_void_p(stream.cuda_stream)), "lora_slab_run")

    def sweep(self, tb: int, stream):
        from ctypes import c_void_p
        _lib.check(self.L.lora_slab_sweep(self._h, int(tb), c_void_p(stream.cuda_stream)), "lora_slab_sweep")

    def reset(self):
        _lib.check(self.L.lora_slab_reset(self._h), "lora_slab_reset")

    @property
    def result_index(self) -> int:
        return int(self.L.lora_slab_result_index(self._h))

    @property
    def plan_handle(self):
        return self.L.lora_slab_plan(self._h)

    def destroy(self):
        if self._h:
            self.L.lora_slab_destroy(self._h)
            self._h = None


class SlabRunner:
    def __init__(self, shape: str, global_dims, params=None, mode: int = 0, group=None, device=None, step_fn=None,
                 fused_fn=None, temporal_block: int | None = None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.shape, self.mode = shape, mode
        self.params = None if params is None else np.ascontiguousarray(np.asarray(params, dtype=np.float64).reshape(-1))
        dim = len(global_dims)
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.cuda = self.device.type == "cuda"
        injected = step_fn is not None
        if temporal_block is None:
            if dim == 1 and (not injected or fused_fn):
                temporal_block = int(os.environ.get("LORA_TB", str(DEFAULT_TB_1D)))
            elif dim >= 2 and not injected:
                from .plan import Plan
                # 2-D: 3 for the cross form, 2 for diamond / pyramid; 3-D: 2 -- unless the column count is odd (no
                # tensor map: direct-tap kernel, no fusion)
                probe_dims = (16,) + tuple(int(x) for x in global_dims[1:])
                temporal_block = Plan(shape, probe_dims, params=params, mode=mode).temporal_block
            else:
                temporal_block = 1
        if dim == 1:
            self.max_tb = max(1, min(MAX_TB_1D, temporal_block))
        elif dim == 2:
            # 2-D fuses 3 launches, or 2 (which borrow buffer 1's halo ring: the native peer-memory driver only), or none
            pairs_ok = self.cuda and not injected and os.environ.get("LORA_HALO", "p2p") == "p2p"
            self.max_tb = 3 if temporal_block >= 3 else (2 if temporal_block == 2 and pairs_ok else 1)
        else:  # 3-D fuses 2 launches (native peer-memory driver only) or none
            pairs_ok = self.cuda and not injected and os.environ.get("LORA_HALO", "p2p") == "p2p"
            self.max_tb = 2 if temporal_block >= 2 and pairs_ok else 1
        if dim >= 2 and self.max_tb == 2 and self.world > 1 and int(global_dims[0]) // self.world < 4 * (3 if dim == 2 else 1):
            self.max_tb = 1  # slabs too thin for the bands of a two-launch sweep (same rule in csrc/slab.cu)
        if injected and fused_fn is None:
            self.max_tb = 1
        # ghost zone towards a neighbour: radius x deepest temporal block (1-D: 4 x tb cells, 2-D: 3 x 3 rows), never
        # less than the reference's storage halo
        ghost = None
        if dim == 1 and self.max_tb > 1:
            ghost = 4 * self.max_tb
        elif dim == 2 and self.max_tb > 1:
            ghost = 3 * self.max_tb
        elif dim == 3 and self.max_tb > 1:
            ghost = self.max_tb
        self.ghost, self.align = ghost, (16 if dim == 1 else 1)
        self.geo = SlabGeometry(global_dims, self.world, self.rank, align=self.align, ghost=ghost)
        g = self.geo
        # halo exchange: "p2p" (default on GPUs) = the native slab driver (NativeSlab / csrc/slab.cu): band cells stored
        # straight into the neighbours' ghost zones over NVLink peer memory by the sweep's own kernel launch;
        # "nccl" (LORA_HALO=nccl, cross-check) = edge-band launches + ncclSend/ncclRecv of the bands
        # (torch.distributed); the CPU tests (gloo, injected compute) always take the latter route
        self.halo_mode = "nccl"
        self.native = None
        self.plan = None
        if self.cuda and not injected and os.environ.get("LORA_HALO", "p2p") == "p2p":
            thin = [SlabGeometry(global_dims, self.world, r, align=self.align, ghost=ghost) for r in range(self.world)]
            if self.world == 1 or all(t.slab >= t.wl + t.wr for t in thin):  # bands of the two sides must not overlap
                self.halo_mode = "p2p"
        if self.halo_mode == "p2p":
            # all ranks switch together: if CUDA IPC / peer mapping fails anywhere (container without IPC, GPUs
            # without peer access), everybody falls back to NCCL send/recv
            err = None
            try:
                self.native = NativeSlab(self, self.max_tb)
                if not self.native.ok:
                    err = self.native.err
            except Exception as e:  # noqa: BLE001 -- whatever went wrong, the collective decision is what matters
                err = e
            ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=self.device)
            if self.world > 1:
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 1:
                from .plan import Plan
                self.buf = self.native.buf
                self.plan = Plan.borrowed(self.native.plan_handle, shape, g.local_dims)
                assert self.native.max_tb == self.max_tb
            else:
                if self.native is not None:
                    self.native.destroy()
                    self.native = None
                self.halo_mode = "nccl"
                if self.rank == 0:
                    print(f"lorastencil_b200.slab: peer-memory halo exchange unavailable ({err}); using NCCL", flush=True)
        if self.halo_mode != "p2p":
            if dim >= 2 and self.max_tb == 2:
                self.max_tb = 1  # pairs need the native driver; the ghost rows / planes stay as wide as they are
            if not injected:
                from .plan import Plan
                self.plan = Plan(shape, g.local_dims, params=params, mode=mode)
                if dim == 2:
                    self.plan.temporal_block = self.max_tb
                step_fn, fused_fn = self.plan.step, self.plan.step_fused
            self.buf = [torch.zeros(g.local_padded, dtype=torch.float64, device=self.device) for _ in range(2)]
        self.step_fn, self.fused_fn = step_fn, fused_fn
        self.launch = 0   # kernel sweeps issued: the result sits in buf[launch % 2]
        self.time = 0     # time steps applied
        if self.cuda:
            self.comm_stream = torch.cuda.Stream(device=self.device)
            self.ev_main = torch.cuda.Event()
            self.ev_comm = torch.cuda.Event()

    def close(self):
        """Unmap the neighbours' buffers and free the slab (p2p mode); the runner is unusable afterwards."""
        if self.native is not None:
            self.sync_ranks()   # nobody is still storing into anybody's ghost zone
            self.buf = None
            self.plan = None
            self.native.destroy()
            self.native = None
            self.sync_ranks()

    # ---- data movement helpers (tests / parity; not on the timed path) ----
    def load_global(self, a_global: np.ndarray):
        """Every rank takes its slab (with halo / ghost rows) out of the same global padded array."""
        self.sync_ranks()  # a neighbour's band tasks of an earlier run may still be storing into my ghost zone
        t = self.torch.from_numpy(np.ascontiguousarray(a_global[self.geo.global_rows()]))
        self.buf[0].copy_(t)
        self.buf[1].zero_()
        self.launch = self.time = 0
        if self.native is not None:
            self.native.reset()
        self.sync_ranks()

    def sync_ranks(self):
        """Host-side rendezvous around (re)filling the buffers from outside: no neighbour may store into my ghost
        rows while I am writing them myself (p2p mode)."""
        if self.cuda:
            self.torch.cuda.synchronize(self.device)
        if self.world > 1:
            self.dist.barrier(group=self.group)

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
        if self.native is not None:
            self.native.sweep(tb, self.torch.cuda.current_stream(self.device))
            self.launch += 1
            self.time += tb
            return
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
        if self.native is not None:
            self.native.run(times, self.torch.cuda.current_stream(self.device))  # the whole sweep loop is native
            if self.world == 1 and self.launch % 2 == 0 and self.time % 2 == 0:
                self.launch += times  # lora_slab_run hands a whole grid on one device to lora_plan_run
            elif self.geo.dim == 1:
                self.launch += len(temporal_schedule(times, self.max_tb))
            else:
                self.launch += len(temporal_schedule_2d(times, self.max_tb))
            self.time += times
            assert self.launch % 2 == self.native.result_index
            return self.result()
        if self.world == 1 and self.plan is not None and self.launch % 2 == 0 and self.time % 2 == 0:
            res = self.plan.run(self.buf[0], self.buf[1], times)  # whole line on one device: the plan schedules it
            self.launch += times
            self.time += times
            return res
        if self.max_tb > 1:
            assert self.launch % 2 == self.time % 2, "fused runs must start from a parity-consistent state"
            sched = temporal_schedule(times, self.max_tb) if self.geo.dim == 1 else temporal_schedule_2d(times, self.max_tb)
            for tb in sched:
                self._sweep(tb)
        else:
            for _ in range(times):
                self._sweep(1)
        return self.result()


# ---- host-resident 1-D lines on N GPUs without any exchange -------------------------------------------------
def host_segment(n_global: int, world: int, rank: int, times: int):
    """Split a global 1-D line of n_global cells for `world` independent drop-in operator calls.

    A cell after `times` launches depends on 4 x times cells either side only, so a rank that is handed its slab
    PLUS a margin of that width can run the whole job on its own, no halo exchange at all: its operator call treats
    the ends of its segment as ends of a line, and that error travels 4 cells per launch -- it never leaves the
    margin.  Returns (lo, hi, gl, gr): the slab [lo, hi) and the margins kept left / right of it (0 at the ends of
    the global line, where the segment's end IS the line's end)."""
    units = -(-n_global // 16)
    bounds = [min(n_global, 16 * (units * r // world)) for r in range(world + 1)]
    lo, hi = bounds[rank], bounds[rank + 1]
    margin = 4 * times + 8
    return lo, hi, min(margin, lo), min(margin, n_global - hi)


def run_host_segment(shape: str, seg_in, seg_out, params, times: int):
    """One rank's share of a host-resident 1-D job: `seg_in` = padded segment (4 halo + margin + slab + margin + 4
    halo doubles, cut out of the global padded line), `seg_out` likewise; the reference-facing operator runs on it
    (chunked, copies overlapped with launches).  The slab part of seg_out is exact; the margins are not."""
    from . import ops
    n_seg = int(seg_in.numel() if hasattr(seg_in, "numel") else seg_in.size) - 8
    return ops.BY_SHAPE[shape](seg_in, seg_out, params, times, n_seg)

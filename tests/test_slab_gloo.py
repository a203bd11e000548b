"""CPU tests (gloo, world_size 2 and 3) of the multi-GPU slab logic: partition, per-launch halo exchange
and the preserved outer-halo semantics (S2).  The compute step is injected (the CPU oracle), so what is
under test is lorastencil_b200/slab.py's host logic -- exactly the code the GPU path runs around
Plan.step."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from lorastencil_b200.slab import SlabGeometry, SlabRunner


def test_geometry_covers_the_axis_without_overlap():
    for dims, world, align in (((1 << 20,), 8, 4), ((40960, 40960), 8, 1), ((1024, 1024, 1024), 4, 1), ((100, 64), 3, 1),
                               ((1001,), 2, 4)):
        gs = [SlabGeometry(dims, world, r, align) for r in range(world)]
        assert gs[0].lo == 0 and gs[-1].hi == dims[0]
        for a, b in zip(gs[:-1], gs[1:]):
            assert a.hi == b.lo
        assert gs[0].prev is None and gs[-1].next is None and gs[0].next == 1
        for g in gs:
            assert g.local_padded[0] == g.hi - g.lo + 2 * g.halo
    with pytest.raises(ValueError):
        SlabGeometry((6, 64), 4, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, shape, dims, times, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d = oracle.dim_of(shape)
        halo = oracle.HALO[d]
        eff = oracle.effective_params(shape)

        def step_fn(src, dst, lo, hi, stream=None):
            full = oracle.step(shape, src.numpy(), eff)
            sl = (slice(halo[0] + lo, halo[0] + hi),) + tuple(slice(h, -h) for h in halo[1:])
            dst.numpy()[sl] = full[sl]

        a = oracle.fill_rand(shape, dims)  # every rank generates the same global input
        runner = SlabRunner(shape, dims, step_fn=step_fn)
        runner.load_global(a)
        runner.run(times)
        got = runner.gather_global(a.shape)
        if rank == 0:
            ref = oracle.run(shape, a, eff, times)
            ok = np.array_equal(got[:-1], ref[:-1]) if d == 1 else np.array_equal(got, ref)
            ret.put(bool(ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape,dims,world,times", [
    ("1d2r", (4096,), 2, 3), ("box2d3r", (48, 64), 2, 4), ("star2d1r", (40, 36), 3, 3),
    ("box3d1r", (12, 8, 64), 2, 3), ("star3d1r", (9, 5, 30), 3, 2), ("box2d1r", (32, 64), 2, 1)])
def test_slab_run_equals_single_domain(shape, dims, world, times):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, shape, dims, times, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ret.get(timeout=10) is True

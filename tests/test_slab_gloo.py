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
    # ghost zones (1-D temporal blocking, tb = 4 here): 16 cells towards neighbours, the 4-cell halo towards the ends
    gs = [SlabGeometry((4096,), 3, r, 16, ghost=16) for r in range(3)]
    assert [(g.wl, g.wr) for g in gs] == [(4, 16), (16, 16), (16, 4)]
    assert [g.off for g in gs] == [0, 12, 12]
    assert gs[1].local_padded[0] == gs[1].slab + 32 and gs[0].local_padded[0] == gs[0].slab + 20
    with pytest.raises(ValueError):
        SlabGeometry((6, 64), 4, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_fused_fn(eff):
    """CPU stand-in for Plan.step_fused (lora_plan_step_fused): tb oracle steps with the virtual halo."""
    def fused(src, dst, halo_src, lo, hi, tb, t0, virt_lo, virt_hi, stream=None):
        cur = src.numpy().copy()
        H = halo_src.numpy()
        for s in range(tb):
            use_h = (t0 + s) % 2 == 0
            if virt_lo:
                cur[:4] = H[:4] if use_h else 0.0
            if virt_hi:
                cur[-4:] = H[-4:] if use_h else 0.0
            nxt = cur.copy()
            nxt[4:-4] = oracle.step(1, cur, eff)[4:-4]
            cur = nxt
        dst.numpy()[4 + lo:4 + hi] = cur[4 + lo:4 + hi]
    return fused


def _oracle_fused_fn_2d(eff):
    """CPU stand-in for the 2-D lora_plan_step_fused: tb oracle steps; the halo ring is virtual (caller's at even
    times, zero at odd) on the left / right always and on the top / bottom where the slab ends the global grid; rows
    towards a neighbour are ordinary data (valid only inside the dependency cone, like in the kernel)."""
    def fused(src, dst, halo_src, lo, hi, tb, t0, virt_lo, virt_hi, stream=None):
        cur = src.numpy().copy()
        H = halo_src.numpy()
        for s in range(tb):
            use_h = (t0 + s) % 2 == 0
            if s > 0:  # level 0 reads the source buffer's own ring
                cur[:, :4] = H[:, :4] if use_h else 0.0
                cur[:, -4:] = H[:, -4:] if use_h else 0.0
                if virt_lo:
                    cur[:4] = H[:4] if use_h else 0.0
                if virt_hi:
                    cur[-4:] = H[-4:] if use_h else 0.0
            # like the kernel, intermediate levels are computed on EVERY stored row (rows of the ghost zone that sit
            # in the storage halo included), with zeros beyond the array
            ext = np.zeros((cur.shape[0] + 8, cur.shape[1]))
            ext[4:-4] = cur
            nxt = cur.copy()
            nxt[:, 4:-4] = oracle.step(2, ext, eff)[4:-4, 4:-4]
            cur = nxt
        dst.numpy()[4 + lo:4 + hi, 4:-4] = cur[4 + lo:4 + hi, 4:-4]
    return fused


def _worker(rank, world, port, shape, dims, times, ret, fused=False):
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
        if d == 2 and fused:
            runner = SlabRunner(shape, dims, step_fn=step_fn, fused_fn=_oracle_fused_fn_2d(eff), temporal_block=3)
            assert runner.max_tb == 3 and (runner.geo.wl == 9 or runner.geo.prev is None)
        else:
            runner = SlabRunner(shape, dims, step_fn=step_fn, fused_fn=_oracle_fused_fn(eff) if fused else None,
                                temporal_block=4 if fused else None)
            assert runner.max_tb == (4 if fused else 1)
        runner.load_global(a)
        runner.run(times)
        got = runner.gather_global(a.shape)
        if rank == 0:
            ref = oracle.run(shape, a, eff, times)
            ok = np.array_equal(got[:-1], ref[:-1]) if d == 1 else np.array_equal(got, ref)
            ret.put(bool(ok))
    finally:
        dist.destroy_process_group()


def test_temporal_schedule_lands_in_the_reference_buffer():
    from lorastencil_b200.slab import temporal_schedule
    for max_tb in (1, 2, 3, 4, 5, 8):
        for times in range(0, 60):
            tbs = temporal_schedule(times, max_tb)
            assert sum(tbs) == times and all(1 <= t <= max_tb for t in tbs)
            assert len(tbs) % 2 == times % 2  # result in buf[times % 2] (S3)
    assert temporal_schedule(1000, 4) == [4] * 250
    assert temporal_schedule(1000, 8) == [8] * 124 + [4, 4]


@pytest.mark.parametrize("shape,dims,world,times,fused", [
    ("1d2r", (4096,), 2, 3, False), ("box2d3r", (48, 64), 2, 4, False), ("star2d1r", (40, 36), 3, 3, False),
    ("box3d1r", (12, 8, 64), 2, 3, False), ("star3d1r", (9, 5, 30), 3, 2, False), ("box2d1r", (32, 64), 2, 1, False),
    ("1d2r", (4096,), 2, 7, True), ("1d1r", (1000,), 3, 10, True), ("star2d3r", (60, 40), 2, 7, True),
    ("star2d1r", (66, 36), 3, 5, True)])
def test_slab_run_equals_single_domain(shape, dims, world, times, fused):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, shape, dims, times, ret, fused)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ret.get(timeout=10) is True


def test_peer_mirror_shifts_map_edge_bands_onto_neighbour_ghost_rows():
    """The element shift the p2p halo exchange adds to a store address (slab.PeerHalo.shift) must send a rank's
    first / last ghost-width slab rows exactly onto the neighbour's trailing / leading ghost rows: checked against
    the global-row bookkeeping, for 1-D ghost zones and 2-D / 3-D storage halos, ragged splits included."""
    from lorastencil_b200.slab import SlabGeometry
    for dims, world, align, ghost in (((1 << 20,), 4, 16, 60), ((100003,), 3, 16, 16), ((300, 258), 4, 1, None),
                                      ((40960, 64), 8, 1, None), ((33, 40, 136), 3, 1, None)):
        gs = [SlabGeometry(dims, world, r, align=align, ghost=ghost) for r in range(world)]
        rest = int(np.prod(gs[0].local_padded[1:])) if len(dims) > 1 else 1
        for r, g in enumerate(gs):
            first_global = g.global_rows().start  # global padded outer index of my local outer index 0
            if g.prev is not None:
                gp = gs[g.prev]
                shift = ((gp.wl + gp.slab) - g.wl) * rest          # PeerHalo.shift["prev"]
                assert shift % rest == 0
                for u in range(g.wl, g.wl + gp.wr):                # my top band, local outer index u
                    v = u + shift // rest                          # where it lands in prev's buffer
                    assert gp.wl + gp.slab <= v < gp.local_padded[0]
                    assert gp.global_rows().start + v == first_global + u  # same global row
            if g.next is not None:
                gn = gs[g.next]
                shift = (gn.wl - g.wl - g.slab) * rest             # PeerHalo.shift["next"]
                for u in range(g.wl + g.slab - gn.wl, g.wl + g.slab):  # my bottom band
                    v = u + shift // rest
                    assert 0 <= v < gn.wl
                    assert gn.global_rows().start + v == first_global + u


def test_2d_schedule_uses_odd_blocks_only():
    """2-D sweeps advance 3 or 1 launches: every prefix has time parity == sweep parity (the source buffer's own halo
    ring is then the right one for level 0), and the result lands in buf[times % 2] (S3)."""
    from lorastencil_b200.slab import temporal_schedule_2d
    for times in range(0, 50):
        for max_tb in (1, 3):
            tbs = temporal_schedule_2d(times, max_tb)
            assert sum(tbs) == times and all(t in (1, 3) for t in tbs) and (max_tb == 3 or all(t == 1 for t in tbs))
            done = 0
            for k, t in enumerate(tbs):
                assert done % 2 == k % 2
                done += t
            assert len(tbs) % 2 == times % 2


def test_host_segments_cover_the_line_and_carry_the_dependency_cone():
    """slab.host_segment: the slabs of all ranks tile [0, n) exactly; a margin is the dependency cone of `times`
    launches (4 cells per launch, + the 4-cell halo and alignment slack) or reaches the end of the line."""
    from lorastencil_b200.slab import host_segment
    for n, world, times in ((1 << 28, 8, 1000), (1 << 20, 3, 37), (300000, 2, 8), (1000, 4, 500), (17, 2, 1)):
        segs = [host_segment(n, world, r, times) for r in range(world)]
        assert segs[0][0] == 0 and segs[-1][1] == n
        for (lo, hi, gl, gr), nxt in zip(segs, segs[1:] + [None]):
            assert lo <= hi
            if nxt is not None:
                assert hi == nxt[0]
            assert gl == min(4 * times + 8, lo) and gr == min(4 * times + 8, n - hi)
            assert lo - gl >= 0 and hi + gr <= n


def test_pair_schedule_for_two_launch_sweeps():
    """Sweeps of two launches (2-D diamond / pyramid, 3-D): an EVEN number of them first -- every pair sweep starts at an
    even time with the data parity of the buffers restored at the end of the pairs -- then 0..3 single launches, so the
    result lands in buf[times % 2] (S3).  Fewer than 4 launches: no pairs at all (csrc/plan.cu: lora_plan_run,
    csrc/slab.cu: schedule_for, and this mirror must agree)."""
    from lorastencil_b200.slab import temporal_schedule_2d
    for times in range(0, 60):
        tbs = temporal_schedule_2d(times, 2)
        assert sum(tbs) == times and all(t in (1, 2) for t in tbs)
        pairs = [t for t in tbs if t == 2]
        assert len(pairs) % 2 == 0 and tbs[:len(pairs)] == pairs      # pairs first, an even number of them
        assert len(tbs) - len(pairs) <= 3                            # at most three single launches behind them
        assert (len(pairs) == 0) == (times < 4)
        assert len(tbs) % 2 == times % 2                             # one buffer flip per sweep
        done = 0
        for t in tbs:
            if t == 2:
                assert done % 2 == 0                                 # a pair sweep starts at an even time
            done += t

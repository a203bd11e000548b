"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle.

Bar: bit-exact (==) while every intermediate stays an exact integer below 2^53 (integer inputs and
weights, the first few launches); afterwards max relative error <= 1e-12 (BASELINE.json north_star).
The oracle is fed `effective_weights` = the weights the reference GPU operator applies."""
import hashlib
import json
import os

import numpy as np
import pytest

import lorastencil_b200 as ls
import oracle
from lorastencil_b200 import ops

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "single_step.json")))
RTOL = 1e-12  # north_star: max relative error <= 1e-12 for FP64


@pytest.fixture(scope="module", autouse=True)
def _quiet():
    prev = ops.set_verbose(False)
    yield
    ops.set_verbose(prev)


def interior(shape, dims):
    halo = oracle.HALO[oracle.dim_of(shape)]
    return tuple(slice(h, h + x) for h, x in zip(halo, dims))


def max_rel_err(got, ref):
    scale = np.abs(ref).max()
    return 0.0 if scale == 0 else float(np.abs(got - ref).max() / scale)


def run_dropin(shape, a, params, times, dims, fill=-7.0):
    out = np.full_like(a, fill)
    ops.BY_SHAPE[shape](a, out, params, times, *dims)
    return out


@pytest.mark.parametrize("g", GOLDEN, ids=lambda g: f"{g['shape']}-{'x'.join(map(str, g['dims']))}")
def test_golden_single_step(g):
    """One launch on the reference's own input reproduces the reference test_cpu golden values."""
    shape, dims = g["shape"], tuple(g["dims"])
    a = oracle.fill_rand(shape, dims)
    out = run_dropin(shape, a, oracle.reference_params(shape), 1, dims)
    inner = np.ascontiguousarray(out[interior(shape, dims)])
    assert float(inner.sum()) == g["interior_sum"]
    assert hashlib.sha256(inner.tobytes()).hexdigest() == g["interior_sha256"]


CASES = [
    ("1d1r", (1024,)), ("1d2r", (1024,)), ("1d2r", (130,)), ("1d1r", (7,)), ("1d2r", (100003,)), ("1d2r", (1 << 20,)),
    ("box2d1r", (64, 64)), ("box2d3r", (32, 64)), ("box2d3r", (50, 70)), ("box2d1r", (1, 2)), ("box2d3r", (300, 258)),
    ("star2d3r", (64, 64)), ("star2d3r", (33, 130)), ("star2d1r", (64, 64)), ("star2d1r", (40, 36)),
    ("box2d1r", (1024, 1024)),
    ("box3d1r", (16, 16, 64)), ("box3d1r", (5, 9, 30)), ("box3d1r", (70, 40, 136)), ("box3d1r", (1, 1, 2)),
    ("star3d1r", (16, 16, 64)), ("star3d1r", (7, 33, 132)), ("star3d1r", (64, 64, 64)),
    # odd column counts: no 16-byte row pitch, so no tensor map -- the direct-tap kernels (stencil_direct.cu)
    ("box2d3r", (50, 71)), ("star2d1r", (33, 129)), ("star2d3r", (300, 1)), ("box2d1r", (7, 1001)),
    ("box3d1r", (5, 9, 31)), ("star3d1r", (7, 6, 5)), ("box3d1r", (33, 20, 257)),
]


@pytest.mark.parametrize("shape,dims", CASES, ids=lambda v: v if isinstance(v, str) else "x".join(map(str, v)))
def test_multi_step_bit_exact_then_1e12(shape, dims):
    """S1-S4 over several launches, full padded output (halo alternation included)."""
    a = oracle.fill_rand(shape, dims)
    p = oracle.reference_params(shape)
    eff = ls.effective_weights(shape, ls.WEIGHTS_REFERENCE, p)
    assert np.array_equal(eff, oracle.effective_params(shape, p))
    exact_upto = {"1d1r": 8, "1d2r": 8, "box2d1r": 5, "box2d3r": 5, "star2d1r": 6, "star2d3r": 9,
                  "box3d1r": 8, "star3d1r": 15}[shape]
    for times in (0, 1, 2, 3, 4, 10):
        out = run_dropin(shape, a, p, times, dims)
        ref = np.full_like(a, -7.0)
        oracle.run(shape, a, eff, times, out=ref)
        if times <= exact_upto:
            assert np.array_equal(out, ref), (shape, dims, times)
        else:
            assert max_rel_err(out, ref) <= RTOL, (shape, dims, times)
        if oracle.dim_of(shape) == 1:
            assert out[-1] == -7.0  # S3: 1-D copies back cols-1 doubles


@pytest.mark.parametrize("shape,dims", [("1d2r", (4096,)), ("box2d3r", (96, 128)), ("star2d3r", (64, 192)),
                                        ("star2d1r", (64, 64)), ("box3d1r", (12, 32, 128)), ("star3d1r", (9, 40, 64))])
def test_float_data_long_run_relative_error(shape, dims):
    """Non-integer data, 40 launches (values grow by the weight sum each launch): <= 1e-12 relative."""
    rng = np.random.default_rng(11)
    a = rng.uniform(-1, 1, oracle.padded_shape(shape, dims))
    p = oracle.reference_params(shape)
    eff = oracle.effective_params(shape, p)
    for times in (1, 7, 40):
        out = run_dropin(shape, a, p, times, dims, fill=0.0)
        ref = oracle.run(shape, a, eff, times)
        assert max_rel_err(out, ref) <= RTOL, (shape, times, max_rel_err(out, ref))


def test_general_mode_honours_every_weight():
    """LORA_WEIGHTS_GENERAL == the reference's test_cpu for arbitrary tables: pyramid with a centre
    remainder, asymmetric pyramid, full-rank (direct taps), general cross / diamond, 3-D forms."""
    rng = np.random.default_rng(13)
    dims2, dims3 = (70, 200), (9, 40, 136)
    a2 = rng.uniform(-1, 1, oracle.padded_shape("box2d3r", dims2))
    a3 = rng.uniform(-1, 1, oracle.padded_shape("box3d1r", dims3))
    tables2 = []
    W = np.zeros((7, 7))
    for t in range(3):
        u, v = np.zeros(7), np.zeros(7)
        u[t:7 - t] = rng.uniform(0.5, 1.5, 7 - 2 * t)
        v[t:7 - t] = rng.uniform(0.5, 1.5, 7 - 2 * t)
        W += np.outer(u, v)
    W[3, 3] += 0.37
    tables2.append(("pyramid", W.ravel()))
    tables2.append(("direct49", rng.standard_normal(49)))
    cross = np.zeros((7, 7))
    cross[:, 3] = rng.standard_normal(7)
    cross[3, :] = rng.standard_normal(7)
    tables2.append(("cross", cross.ravel()))
    dia = oracle.reference_params("star2d1r").reshape(7, 7).copy()
    dia[3, 0] = 0.25
    dia[1, 1] = -3.0
    tables2.append(("diamond", dia.ravel()))
    # rank 2 / rank 3 with full support and no pyramidal structure: the LU (cross approximation) fallback, 28 / 42 taps
    for r in (2, 3):
        tables2.append((f"rank{r}", sum(np.outer(rng.uniform(0.5, 1.5, 7), rng.standard_normal(7)) for _ in range(r)).ravel()))
    for form, w in tables2:
        assert ls.decompose_2d("box2d3r", w)["form"] == form
        out = np.zeros_like(a2)
        ops.run_host("box2d3r", a2, out, w, 3, dims2, mode=ls.WEIGHTS_GENERAL)
        assert max_rel_err(out, oracle.run(2, a2, w, 3)) <= RTOL, form
    tables3 = [np.einsum("i,j,k->ijk", *[rng.uniform(0.5, 1.5, 3) for _ in range(3)]).ravel(), rng.standard_normal(27)]
    st = np.zeros(27)
    st[[13, 12, 14, 10, 16, 4, 22]] = rng.standard_normal(7)
    tables3.append(st)
    for w in tables3:
        out = np.zeros_like(a3)
        ops.run_host("box3d1r", a3, out, w, 3, dims3, mode=ls.WEIGHTS_GENERAL)
        assert max_rel_err(out, oracle.run(3, a3, w, 3)) <= RTOL
    w1 = rng.standard_normal(9)
    a1 = rng.uniform(-1, 1, (5000 + 8,))
    out = np.zeros_like(a1)
    ops.run_host("1d2r", a1, out, w1, 5, (5000,), mode=ls.WEIGHTS_GENERAL)
    assert max_rel_err(out[:-1], oracle.run(1, a1, w1, 5)[:-1]) <= RTOL


def test_plan_api_partial_ranges_and_device_buffers():
    """Layer 2: launches over sub-ranges of the outermost axis compose to the full step, halo cells of
    the destination are left alone, and the result of `run` sits in buf[times % 2]."""
    import torch
    for shape, dims, cuts in (("1d2r", (8192,), (0, 1024, 5000, 8192)), ("box2d3r", (100, 256), (0, 3, 64, 100)),
                              ("star2d1r", (64, 128), (0, 32, 64)), ("box3d1r", (20, 32, 128), (0, 1, 9, 20)),
                              ("star3d1r", (10, 33, 64), (0, 5, 10))):
        a = oracle.fill_rand(shape, dims)
        plan = ls.Plan(shape, dims)
        src = torch.from_numpy(a).cuda()
        dst = torch.full(plan.padded_shape, -3.0, dtype=torch.float64, device="cuda")
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            plan.step(src, dst, lo, hi)
        torch.cuda.synchronize()
        ref = np.full_like(a, -3.0)
        inner = interior(shape, dims)
        ref[inner] = oracle.step(shape, a, oracle.effective_params(shape))[inner]
        assert np.array_equal(dst.cpu().numpy(), ref), shape
        b0, b1 = torch.from_numpy(a).cuda(), plan.new_buffer()
        res = plan.run(b0, b1, 3)
        torch.cuda.synchronize()
        assert res is b1
        full = oracle.run(shape, a, oracle.effective_params(shape), 3)
        got = res.cpu().numpy()
        if oracle.dim_of(shape) == 1:
            assert np.array_equal(got[:-1], full[:-1])
        else:
            assert np.array_equal(got, full)
        # 1-D and the 2-D cross form fuse the 3 launches of `run` into one temporally blocked sweep; the 2-D diamond form
        # and 3-D fuse pairs (from 4 launches on), the FP64-bound 2-D box form launches once per step
        fused = oracle.dim_of(shape) == 1 or shape == "star2d3r"
        assert plan.launches == len(cuts) - 1 + (1 if fused else 3)


def test_linearity_and_shift_invariance_at_scale():
    """Size-independent properties at a BASELINE-sized grid (box2d3r 10240 x 10240 would need minutes of
    oracle time): stencil(a + 2b) == stencil(a) + 2 stencil(b) exactly for integer data, and a grid of
    ones maps to the weight sum everywhere away from the zero halo."""
    import torch
    shape, dims = "box2d3r", (2048, 10240)
    plan = ls.Plan(shape, dims)
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randint(0, 100, plan.padded_shape, generator=g, device="cuda").double()
    b = torch.randint(0, 100, plan.padded_shape, generator=g, device="cuda").double()
    outs = []
    for x in (a, b, a + 2 * b):
        o = plan.new_buffer()
        plan.step(x, o)
        outs.append(o)
    torch.cuda.synchronize()
    assert torch.equal(outs[2], outs[0] + 2 * outs[1])
    ones = torch.ones(plan.padded_shape, dtype=torch.float64, device="cuda")
    o = plan.new_buffer()
    plan.step(ones, o)
    torch.cuda.synchronize()
    assert torch.all(o[4:-4, 4:-4] == float(oracle.reference_params(shape).sum()))
    # spot-check rows of the big grid against the oracle on a slab cut out of it
    r0 = 1000
    slab = a[r0:r0 + 8 + 16].cpu().numpy()
    ref = oracle.step(2, np.ascontiguousarray(slab), oracle.effective_params(shape))
    assert np.array_equal(outs[0][r0 + 4:r0 + 4 + 16].cpu().numpy()[:, 4:-4], ref[4:-4, 4:-4])


@pytest.mark.parametrize("n", [7, 130, 1000, 4096, 100003, 1 << 20])
def test_temporal_blocking_1d_equals_unfused_launches(n):
    """Fused sweeps of 2..15 launches (intermediate levels in registers, virtual alternating halo) give the
    same bits as one launch per step and as the oracle, for every launch count parity and ragged sizes."""
    import torch
    shape = "1d2r"
    a = oracle.fill_rand(shape, (n,))
    rng = np.random.default_rng(n)
    af = rng.uniform(-1, 1, a.shape)
    eff = oracle.effective_params(shape)
    plan = ls.Plan(shape, (n,))
    assert plan.temporal_block == 15
    for data, exact_upto in ((a, 8), (af, 0)):
        for times in (1, 2, 3, 4, 5, 6, 7, 8, 9, 13, 40):
            ref = oracle.run(shape, data, eff, times)[:-1]
            results = []
            for tb in (1, 2, 3, 4, 5, 7, 8, 11, 12, 15):
                plan.temporal_block = tb
                b0, b1 = torch.from_numpy(data).cuda(), plan.new_buffer()
                res = plan.run(b0, b1, times)
                torch.cuda.synchronize()
                assert res is (b0 if times % 2 == 0 else b1)
                results.append(res.cpu().numpy()[:-1])
            for r in results[1:]:
                assert np.array_equal(r, results[0]), (n, times)  # same operation order => same bits
            if times <= exact_upto:
                assert np.array_equal(results[0], ref), (n, times)
            else:
                assert max_rel_err(results[0], ref) <= RTOL, (n, times)


@pytest.mark.parametrize("n,tb,lo,hi", [(20000, 3, 4096, 12000), (20000, 8, 4097, 11999), (3000, 5, 41, 2950),
                                         (70000, 8, 32, 69968), (600, 2, 100, 101), (5000, 15, 60, 4940)])
def test_fused_step_sub_ranges_with_real_halo_data(n, tb, lo, hi):
    """lora_plan_step_fused on interior sub-ranges with virt flags off (the inter-slab case): the fused launch
    equals tb single launches wherever the dependency cone stays inside the data that was provided."""
    import torch
    rng = np.random.default_rng(9)
    a = rng.integers(0, 10, (n + 8,)).astype(np.float64)  # small integers: tb <= 8 steps stay exact on both sides
    eff = oracle.effective_params("1d1r")
    plan = ls.Plan("1d1r", (n,))
    src = torch.from_numpy(a).cuda()
    dst = torch.full((n + 8,), -5.0, dtype=torch.float64, device="cuda")
    plan.step_fused(src, dst, None, lo, hi, tb, 0, False, False)
    torch.cuda.synchronize()
    # reference: tb plain steps of the whole line with its physical halo kept fixed (never re-zeroed):
    ref = a.copy()
    for _ in range(tb):
        nxt = ref.copy()
        nxt[4:-4] = oracle.step(1, ref, eff)[4:-4]
        ref = nxt
    got = dst.cpu().numpy()
    if tb <= 8:   # every intermediate is an integer below 2^53: bit-exact
        assert np.array_equal(got[4 + lo:4 + hi], ref[4 + lo:4 + hi])
    else:
        assert max_rel_err(got[4 + lo:4 + hi], ref[4 + lo:4 + hi]) <= RTOL
    assert np.all(got[:4 + lo] == -5.0) and np.all(got[4 + hi:] == -5.0)


def test_two_gpu_slabs_identical_to_one_gpu():
    """N > 1 on real GPUs (skipped on a 1-GPU box; tests/test_slab_gloo.py covers the logic on CPU)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517",
                        os.path.join(os.path.dirname(__file__), "multigpu_check.py")], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("shape,n,times,chunks", [("1d2r", 1 << 20, 40, 4), ("1d1r", 300001, 7, 3), ("1d2r", 70000, 8, 8),
                                                  ("1d2r", 5000, 100, 5), ("1d1r", 1 << 18, 0, 4), ("1d2r", 1 << 18, 1, 2)])
def test_chunked_copy_overlapped_operator_equals_plain(shape, n, times, chunks, monkeypatch):
    """The 1-D drop-in operator cut into ghost-margined chunks (H2D / launches / D2H overlapped) returns the
    same bits as the plain H2D -> launches -> D2H path, halo cells and the untouched last double included,
    and matches the oracle.  Includes chunks narrower than the dependency cone (margins clipped at the ends)."""
    a = oracle.fill_rand(shape, (n,))
    rng = np.random.default_rng(n + times)
    af = rng.uniform(-1, 1, a.shape)
    p = oracle.reference_params(shape)
    for data in (a, af):
        monkeypatch.setenv("LORA_CHUNKS", "0")
        plain = run_dropin(shape, data, p, times, (n,))
        assert ops.last_chunks() == 1
        monkeypatch.setenv("LORA_CHUNKS", str(chunks))
        chunked = run_dropin(shape, data, p, times, (n,))
        if times >= 1:
            assert ops.last_chunks() >= 2
        assert np.array_equal(chunked, plain), (shape, n, times, chunks)
        assert chunked[-1] == -7.0  # the reference copies back n + 7 doubles (src/1d/gpu_1r.cu:134)
        ref = oracle.run(shape, data, oracle.effective_params(shape, p), times)
        assert max_rel_err(chunked[:-1], ref[:-1]) <= RTOL


@pytest.mark.parametrize("shape", ["star2d3r", "star2d1r", "box2d1r"])
@pytest.mark.parametrize("dims", [(40, 130), (64, 64), (300, 258), (257, 1000), (1000, 130), (9, 8), (2, 2), (400, 230)])
def test_temporal_blocking_2d_equals_unfused_launches(shape, dims):
    """2-D sweeps of 3 fused launches (intermediate grids in registers, virtual alternating halo ring on all four
    sides, overlapped strips) give the same bits as one launch per step, and match the oracle, for every launch
    count residue and ragged sizes."""
    import torch
    a = oracle.fill_rand(shape, dims)
    rng = np.random.default_rng(dims[0] * 7 + dims[1])
    af = rng.uniform(-1, 1, a.shape)
    eff = oracle.effective_params(shape)
    plan = ls.Plan(shape, dims)
    # defaults: the pyramid and diamond forms fuse two launches, the cross form three
    assert plan.temporal_block == {"box2d1r": 2, "star2d1r": 2, "star2d3r": 3}[shape]
    exact_upto = {"box2d1r": 5, "star2d1r": 6, "star2d3r": 9}[shape]
    for data in (a, af):
        for times in (3, 4, 5, 6, 7, 9, 10):
            results = []
            for tb in (1, 3):
                plan.temporal_block = tb
                assert plan.temporal_block == tb
                b0, b1 = torch.from_numpy(data).cuda(), plan.new_buffer()
                n0 = plan.launches
                res = plan.run(b0, b1, times)
                torch.cuda.synchronize()
                assert res is (b0 if times % 2 == 0 else b1)
                assert plan.launches - n0 == (times if tb == 1 else times // 3 + times % 3)
                results.append(res.cpu().numpy())
            assert np.array_equal(results[0], results[1]), (shape, dims, times)  # same operation order => same bits
            ref = oracle.run(shape, data, eff, times)
            if data is a and times <= exact_upto:
                assert np.array_equal(results[1], ref), (shape, dims, times)
            else:
                assert max_rel_err(results[1], ref) <= RTOL, (shape, dims, times)


@pytest.mark.parametrize("shape", ["star2d1r", "box2d1r"])
@pytest.mark.parametrize("dims", [(40, 130), (64, 64), (300, 258), (257, 1000), (1000, 130), (9, 8), (2, 2), (400, 230)])
def test_temporal_blocking_2d_pairs_equal_unfused_launches(shape, dims):
    """2-D sweeps of TWO fused launches (the diamond and pyramid forms: no register spills where three launches spill):
    an even number of sweeps, the caller's ring copied into buffer 1 for their duration and cleared afterwards, the
    remaining launches one by one -- same bits as one launch per step, the oracle's values, and both halo rings left
    as the reference's ping-pong leaves them."""
    import torch
    a = oracle.fill_rand(shape, dims)
    rng = np.random.default_rng(dims[0] * 5 + dims[1])
    af = rng.uniform(-1, 1, a.shape)
    eff = oracle.effective_params(shape)
    plan = ls.Plan(shape, dims)
    inner = interior(shape, dims)
    exact_upto = {"box2d1r": 5, "star2d1r": 6}[shape]
    for data in (a, af):
        for times in (4, 5, 6, 7, 8, 9, 11, 12):
            results = []
            for tb in (1, 2):
                plan.temporal_block = tb
                assert plan.temporal_block == tb
                b0, b1 = torch.from_numpy(data).cuda(), plan.new_buffer()
                n0 = plan.launches
                res = plan.run(b0, b1, times)
                torch.cuda.synchronize()
                assert res is (b0 if times % 2 == 0 else b1)
                sweeps = (times // 2) - (times // 2) % 2
                assert plan.launches - n0 == (times if tb == 1 else sweeps + times - 2 * sweeps)
                results.append(res.cpu().numpy())
                h0, h1 = b0.cpu().numpy().copy(), b1.cpu().numpy().copy()
                d_h = data.copy()
                h0[inner] = 0.0
                d_h[inner] = 0.0
                h1[inner] = 0.0
                assert np.array_equal(h0, d_h) and not h1.any()
            assert np.array_equal(results[0], results[1]), (shape, dims, times)
            ref = oracle.run(shape, data, eff, times)
            if data is a and times <= exact_upto:
                assert np.array_equal(results[1], ref), (shape, dims, times)
            else:
                assert max_rel_err(results[1], ref) <= RTOL, (shape, dims, times)


@pytest.mark.parametrize("shape,dims,times", [("1d1r", (5000,), 31), ("1d2r", (100003,), 16), ("star2d1r", (70, 250), 7),
                                              ("box2d1r", (64, 130), 4), ("star2d3r", (300, 370), 10), ("box3d1r", (9, 34, 130), 4),
                                              ("star3d1r", (12, 8, 64), 5), ("star2d3r", (9, 8), 4)])
def test_no_write_outside_the_destination_interior(shape, dims, times):
    """compute-sanitizer is closed on this pool, so out-of-bounds stores are hunted with canaries: both ping-pong
    buffers are carved out of one allocation with guard zones around them, every guard double and every halo cell of
    the buffers must come back untouched (S2: a launch writes the interior of its destination only), for the plain
    and the fused (TMA-store, overlapped-strip, edge-task) paths."""
    import torch
    plan = ls.Plan(shape, dims)
    n = int(np.prod(plan.padded_shape))
    guard = 4096
    pool = torch.full((3 * guard + 2 * n + 64,), -123.25, dtype=torch.float64, device="cuda")
    off0 = guard
    off1 = (2 * guard + n + 3) // 4 * 4  # keep 32-byte alignment
    b0 = pool[off0:off0 + n].view(plan.padded_shape)
    b1 = pool[off1:off1 + n].view(plan.padded_shape)
    a = np.random.default_rng(11).uniform(-1, 1, plan.padded_shape)
    b0.copy_(torch.from_numpy(a))
    b1.zero_()
    res = plan.run(b0, b1, times)
    torch.cuda.synchronize()
    ref = oracle.run(shape, a, oracle.effective_params(shape), times)
    got = res.cpu().numpy()
    if oracle.dim_of(shape) == 1:
        got, ref = got[:-1], ref[:-1]
    assert max_rel_err(got, ref) <= RTOL
    host = pool.cpu().numpy()
    assert np.all(host[:off0] == -123.25) and np.all(host[off0 + n:off1] == -123.25) and np.all(host[off1 + n:] == -123.25)
    # the halo of buffer 0 still holds the caller's halo, the halo of buffer 1 is still zero
    inner = interior(shape, dims)
    h0, h1 = b0.cpu().numpy().copy(), b1.cpu().numpy().copy()
    a_h = a.copy()
    h0[inner] = 0.0
    a_h[inner] = 0.0
    h1[inner] = 0.0
    assert np.array_equal(h0, a_h) and not h1.any()


def test_chunked_operator_automatic_boundaries_at_scale(monkeypatch):
    """The chunk boundaries the operator picks on its own for a long line (quarter-length first / last chunk, not
    multiples of the kernel row): same bits as the plain path on a 2^26-point line, halo cells included."""
    shape, n, times = "1d2r", 1 << 26, 10
    rng = np.random.default_rng(3)
    a = rng.integers(0, 10, size=(n + 8,)).astype(np.float64)
    p = oracle.reference_params(shape)
    monkeypatch.setenv("LORA_CHUNKS", "0")
    plain = run_dropin(shape, a, p, times, (n,))
    assert ops.last_chunks() == 1
    monkeypatch.delenv("LORA_CHUNKS")
    auto = run_dropin(shape, a, p, times, (n,))
    assert ops.last_chunks() >= 4
    assert np.array_equal(auto, plain)


@pytest.mark.parametrize("shape", ["star2d3r", "star2d1r"])
def test_fused_2d_inner_strip_crossing_the_right_edge(shape):
    """n = 91 x 112 + 6: the last strip is narrower than 8 columns, so the 128-column window of the second-to-last
    (an INNER strip, which runs as a long task) crosses column n as well -- its halo cells cannot be staged in shared
    memory (task longer than the staging area) and come from global memory.  Fused == unfused, bit for bit, and a
    band of rows is checked against the oracle."""
    import torch
    dims = (3000, 91 * 112 + 6)
    plan = ls.Plan(shape, dims)
    rng = np.random.default_rng(17)
    a = rng.integers(0, 10, size=plan.padded_shape).astype(np.float64)
    results = []
    for tb in (1, 3):
        plan.temporal_block = tb
        b0, b1 = torch.from_numpy(a).cuda(), plan.new_buffer()
        res = plan.run(b0, b1, 6)
        torch.cuda.synchronize()
        results.append(res.cpu().numpy())
    assert np.array_equal(results[0], results[1])
    # right-hand columns of the first rows against the oracle on a cut-out (6 launches need 18 more rows / columns)
    sub = np.ascontiguousarray(a[:60 + 8, -(300 + 8):])
    ref = oracle.run(shape, sub, oracle.effective_params(shape), 6)
    assert np.array_equal(results[1][4:4 + 30, -(4 + 200):-4], ref[4:4 + 30, -(4 + 200):-4])


@pytest.mark.parametrize("shape,dims,times", [("box2d3r", (3000, 3072), 4), ("star2d3r", (2500, 4100), 7), ("star2d1r", (4096, 2048), 1),
                                              ("box3d1r", (200, 256, 256), 5), ("star3d1r", (130, 200, 264), 2), ("1d2r", (3000000,), 9),
                                              ("box2d1r", (4000, 2050), 0), ("star3d1r", (130, 200, 264), 9), ("star3d1r", (96, 130, 136), 7),
                                              ("star2d1r", (4096, 2048), 6), ("star2d1r", (3000, 1030), 8), ("box3d1r", (150, 64, 256), 4),
                                              ("star2d1r", (2000, 1500), 5)])
def test_copy_overlapped_operator_equals_plain_sequence(shape, dims, times, monkeypatch):
    """The 2-D / 3-D drop-in operators upload band by band under the first sweep and download band by band behind
    the last one (run_host_pipelined); pageable buffers go through pinned staging slots filled by worker threads
    (hostmove.cu).  LORA_BANDS=1 is the plain copy -> launch loop -> copy sequence: same bits, for pageable (numpy) and
    pinned (torch) caller buffers, halo rows included; and the oracle agrees -- also where the bands run sweeps of two
    launches (3-D, 2-D diamond), which borrow buffer 1's ring for the caller's halo and hand it back zeroed."""
    import torch
    rng = np.random.default_rng(5)
    a = rng.integers(0, 100, size=oracle.padded_shape(shape, dims)).astype(np.float64)
    p = oracle.reference_params(shape)
    monkeypatch.setenv("LORA_BANDS", "1")
    monkeypatch.setenv("LORA_CHUNKS", "0")
    plain = run_dropin(shape, a, p, times, dims)
    outs = []
    for bands in ("7", None):
        if bands:
            monkeypatch.setenv("LORA_BANDS", bands)
        else:
            monkeypatch.delenv("LORA_BANDS")
        outs.append(run_dropin(shape, a, p, times, dims))
        assert ops.last_bands() == (7 if bands else ops.last_bands())
        hin = torch.from_numpy(a).pin_memory()
        hout = torch.full(a.shape, -7.0, dtype=torch.float64).pin_memory()
        ops.BY_SHAPE[shape](hin, hout, p, times, *dims)
        outs.append(hout.numpy().copy())
    for o in outs:
        assert np.array_equal(o, plain), (shape, dims, times)
    ref = np.full_like(a, -7.0)
    oracle.run(shape, a, oracle.effective_params(shape, p), times, out=ref)
    if times in (1, 2):
        assert np.array_equal(plain, ref)
    else:  # sweeps of two or three launches inside the bands (ring of buffer 1 borrowed and given back): the oracle's values
        assert max_rel_err(plain, ref) <= RTOL
        halo = np.ones(a.shape, dtype=bool)
        halo[interior(shape, dims)] = False
        if oracle.dim_of(shape) > 1:
            assert np.array_equal(plain[halo], ref[halo])  # the ring: the caller's or zeros, exactly (S3)


@pytest.mark.parametrize("dims", [(16, 16, 64), (40, 50, 130), (7, 33, 132), (64, 64, 64), (30, 22, 120), (33, 23, 122), (5, 100, 400),
                                  (3, 3, 4), (70, 40, 250), (6, 30, 256), (4, 47, 384), (9, 45, 128), (5, 24, 258), (6, 31, 128), (5, 61, 130)])
@pytest.mark.parametrize("shape", ["star3d1r", "box3d1r"])
def test_temporal_blocking_3d_equals_unfused_launches(shape, dims):
    """3-D sweeps of 2 fused launches (level 1 handed on in registers, by warp shuffles and through two
    shared-memory rows per warp; tiles of 30 x 128 outputs, the level-1 columns beside a tile as extra cells; zero halo at the intermediate level, the caller's ring
    copied into buffer 1 for the odd sweeps and cleared afterwards) give the same bits as one launch per step, match
    the oracle, and leave both halo rings as the reference's ping-pong would."""
    import torch
    a = oracle.fill_rand(shape, dims)
    rng = np.random.default_rng(sum(dims))
    af = rng.uniform(-1, 1, a.shape)
    eff = oracle.effective_params(shape)
    plan = ls.Plan(shape, dims)
    assert plan.temporal_block == 2  # both forms are fused by default
    inner = interior(shape, dims)
    for data in (a, af):
        for times in (4, 5, 6, 7, 8, 9, 12, 15):
            results = []
            for tb in (1, 2):
                plan.temporal_block = tb
                assert plan.temporal_block == tb
                b0, b1 = torch.from_numpy(data).cuda(), plan.new_buffer()
                n0 = plan.launches
                res = plan.run(b0, b1, times)
                torch.cuda.synchronize()
                assert res is (b0 if times % 2 == 0 else b1)
                fused_sweeps = (times // 2) - (times // 2) % 2
                assert plan.launches - n0 == (times if tb == 1 else fused_sweeps + times - 2 * fused_sweeps)
                results.append(res.cpu().numpy())
                # the rings: buffer 0 still holds the caller's halo, buffer 1 zeros
                h0, h1 = b0.cpu().numpy().copy(), b1.cpu().numpy().copy()
                d_h = data.copy()
                h0[inner] = 0.0
                d_h[inner] = 0.0
                h1[inner] = 0.0
                assert np.array_equal(h0, d_h) and not h1.any()
            assert np.array_equal(results[0], results[1]), (dims, times)  # same operation order => same bits
            ref = oracle.run(shape, data, eff, times)
            if data is a and times <= (15 if shape == "star3d1r" else 8):
                assert np.array_equal(results[1], ref), (shape, dims, times)
            else:
                assert max_rel_err(results[1], ref) <= RTOL, (shape, dims, times)


@pytest.mark.parametrize("shape,dims,times", [("1d2r", (3_000_000,), 31), ("1d1r", (1 << 21,), 16), ("box2d1r", (1500, 2048), 9),
                                              ("star2d1r", (1400, 1930), 9), ("star2d3r", (1600, 2050), 10),
                                              ("box3d1r", (70, 40, 250), 9), ("star3d1r", (70, 40, 250), 9),
                                              ("box3d1r", (40, 150, 260), 5)])
def test_run_to_run_determinism(shape, dims, times):
    """Races between a TMA refill and reads that were issued but not yet performed (found twice: the fused 2-D ring in
    round 2, the 3-D stage release after the kernels' control flow became uniform) never showed as wrong values in a
    single comparison with the oracle at small sizes -- they show as RUN-TO-RUN differences on grids large enough to
    back the load / store unit up.  Every kernel, fused and unfused: 12 repeats must give the same bits."""
    import torch
    plan = ls.Plan(shape, dims)
    a = np.random.default_rng(3).integers(0, 100, plan.padded_shape).astype(np.float64)
    default_tb = plan.temporal_block
    for tb in sorted({1, default_tb}):
        plan.temporal_block = tb
        first = None
        for rep in range(12):
            b0, b1 = torch.from_numpy(a).cuda(), plan.new_buffer()
            res = plan.run(b0, b1, times).clone()
            torch.cuda.synchronize()
            if first is None:
                first = res
            else:
                assert torch.equal(first, res), (shape, dims, tb, rep)


def test_general_weight_tables_under_fusion():
    """Fused sweeps honour arbitrary tables too (LORA_WEIGHTS_GENERAL): a pyramid WITH a centre remainder (the unpruned
    form, sweeps of two), an asymmetric cross (three), a general diamond (two and three), a general 7-point star and a
    general separable 27-point table in 3-D (two) -- same bits as single launches, the oracle's values."""
    import torch
    rng = np.random.default_rng(17)
    W = np.zeros((7, 7))
    for t in range(3):
        u, v = np.zeros(7), np.zeros(7)
        u[t:7 - t] = rng.uniform(0.5, 1.5, 7 - 2 * t)
        v[t:7 - t] = rng.uniform(0.5, 1.5, 7 - 2 * t)
        W += np.outer(u, v)
    W[3, 3] += 0.37
    cross = np.zeros((7, 7))
    cross[:, 3] = rng.standard_normal(7)
    cross[3, :] = rng.standard_normal(7)
    dia = oracle.reference_params("star2d1r").reshape(7, 7).copy()
    dia[3, 0] = 0.25
    dia[1, 1] = -3.0
    st = np.zeros(27)
    st[[13, 12, 14, 10, 16, 4, 22]] = rng.standard_normal(7)
    sep = np.einsum("i,j,k->ijk", *[rng.uniform(0.5, 1.5, 3) for _ in range(3)]).ravel()
    cases = [("box2d3r", (300, 258), W.ravel(), "pyramid", (2,)), ("box2d3r", (257, 400), cross.ravel(), "cross", (3,)),
             ("box2d3r", (129, 386), dia.ravel(), "diamond", (2, 3)), ("box3d1r", (40, 50, 130), st, None, (2,)),
             ("box3d1r", (33, 61, 256), sep, None, (2,))]
    for shape, dims, w, form, tbs in cases:
        if form:
            assert ls.decompose_2d(shape, w)["form"] == form
        a = rng.uniform(-1, 1, oracle.padded_shape(shape, dims))
        plan = ls.Plan(shape, dims, params=w, mode=ls.WEIGHTS_GENERAL)
        for times in (4, 7, 9):
            results = []
            for tb in (1,) + tbs:
                plan.temporal_block = tb
                assert plan.temporal_block == tb, (shape, form, tb)
                b0, b1 = torch.from_numpy(a).cuda(), plan.new_buffer()
                n0 = plan.launches
                res = plan.run(b0, b1, times)
                torch.cuda.synchronize()
                if tb > 1:
                    assert plan.launches - n0 < times  # really fused
                results.append(res.cpu().numpy())
            for other in results[1:]:
                assert np.array_equal(results[0], other), (shape, form, times)
            assert max_rel_err(results[0], oracle.run(oracle.dim_of(shape), a, w, times)) <= RTOL, (shape, form, times)

"""Radius-2 3-D shapes box3d2r / star3d2r (the reference's shape list ends at radius 1,
src/3d/3d_utils.h:39-42, so there is no reference vector for them).

CPU: the checker (oracle.step_r2 / run_r2: the test_cpu protocol of src/3d/main.cu:33-68 on a 5x5x5 window, with the
S2 / S3 buffer semantics of src/3d/gpu_box.cu:190-223) is pinned against an independent implementation,
scipy.ndimage.correlate; the host decomposition picks its form by structure and reports the taps it applies.
GPU: the drop-in operators, the plan API and torch.ops against the checker."""
import numpy as np
import pytest

import lorastencil_b200 as ls
import oracle
from lorastencil_b200 import _lib, ops
from lorastencil_b200.plan import effective_weights, reference_table

RTOL = 1e-12
INNER = (slice(2, -2), slice(2, -2), slice(4, -4))


def tables(rng):
    """(name, shape, table, expected form)"""
    star = np.zeros((5, 5, 5))
    for ax in range(3):
        for d in (-2, -1, 1, 2):
            idx = [2, 2, 2]
            idx[ax] += d
            star[tuple(idx)] = rng.uniform(-1, 1)
    star[2, 2, 2] = rng.uniform(-1, 1)
    hsep = np.einsum("i,jk->ijk", rng.uniform(-1, 1, 5), rng.uniform(-1, 1, (5, 5)))
    dense = rng.uniform(-1, 1, (5, 5, 5))
    return [("default box", "box3d2r", oracle.reference_params_r2("box3d2r"), "hsep5"),
            ("default star", "star3d2r", oracle.reference_params_r2("star3d2r"), "star13"),
            ("general star", "star3d2r", star.reshape(-1), "star13"),
            ("rank 1 along h", "box3d2r", hsep.reshape(-1), "hsep5"),
            ("dense", "box3d2r", dense.reshape(-1), "direct125"),
            ("dense through the star entry point", "star3d2r", dense.reshape(-1), "direct125")]


@pytest.mark.parametrize("dims", [(6, 7, 9), (1, 1, 1), (3, 2, 64)])
def test_checker_equals_scipy_correlation(dims):
    from scipy import ndimage
    rng = np.random.default_rng(1)
    for _, shape, w, _ in tables(rng):
        wi = np.round(8 * w)  # integers: every sum is exact, so the comparison can be ==
        a = rng.integers(0, 100, oracle.padded_shape_r2(dims)).astype(np.float64)
        out = oracle.step_r2(a, wi)
        ref = ndimage.correlate(a, wi.reshape(5, 5, 5), mode="constant")
        assert np.array_equal(out[INNER], ref[INNER])
        ring = out.copy()
        ring[INNER] = 0
        assert not ring.any()  # interior only
    # S2 / S3: launch i reads buf[i % 2]; buffer 1 starts as zeros; the result is the whole padded buf[times % 2]
    w = oracle.reference_params_r2("star3d2r")
    a = rng.integers(0, 100, oracle.padded_shape_r2(dims)).astype(np.float64)
    b = [a.copy(), np.zeros_like(a)]
    for i in range(4):
        assert np.array_equal(oracle.run_r2(a, w, i), b[i % 2])
        b[(i + 1) % 2][INNER] = oracle.step_r2(b[i % 2], w)[INNER]


def test_default_tables_match_the_checker_restatement():
    for shape in oracle.R2_SHAPES:
        assert np.array_equal(reference_table(shape), oracle.reference_params_r2(shape))
        assert _lib.nparams(shape) == 125 and _lib.halo_of(shape) == oracle.HALO_R2


def test_every_weight_is_honoured_in_both_modes():
    rng = np.random.default_rng(2)
    for name, shape, w, form in tables(rng):
        for mode in (ls.WEIGHTS_REFERENCE, ls.WEIGHTS_GENERAL):
            eff = effective_weights(shape, mode, w)
            assert eff.shape == (125,)
            assert np.abs(eff - w).max() <= 64 * 2.3e-16 * np.abs(w).max(), name
            if form != "hsep5":
                assert np.array_equal(eff, w), name  # only the rank-1 form re-multiplies its factors


def test_form_follows_the_structure_of_the_table(monkeypatch):
    """decompose_3d_r2 (csrc/decompose.cpp): 13-point star / fully separable / rank 1 along the plane axis / 125 taps."""
    from lorastencil_b200.plan import decompose_3d_r2
    monkeypatch.delenv("LORA_R2_SEP5", raising=False)
    rng = np.random.default_rng(6)
    for name, shape, w, form in tables(rng):
        d = decompose_3d_r2(shape, w)
        want = {"default box": "sep5"}.get(name, form)  # fully separable tables take the shuffle kernel by default
        assert d["form"] == want, name
        assert d["macs_per_cell"] == {"star13": 13, "sep5": 15, "hsep5": 30, "direct125": 125}[want]
    # the default box factors into the integers it was made of: exact taps, exact first launches
    d = decompose_3d_r2("box3d2r", oracle.reference_params_r2("box3d2r"))
    for k in "abc":
        assert np.array_equal(d[k], [1, 2, 3, 2, 1]), k
    assert d["recon_err"] == 0.0 and np.array_equal(d["q"], np.outer(d["b"], d["c"]))
    # one off-axis weight ends the star; one perturbed entry ends the separability
    w = oracle.reference_params_r2("star3d2r").copy()
    w[0] = 0.5
    assert decompose_3d_r2("star3d2r", w)["form"] == "direct125"
    w = oracle.reference_params_r2("box3d2r").copy()
    w[7] += 1.0
    assert decompose_3d_r2("box3d2r", w)["form"] == "direct125"
    # a (x) Q with Q of rank 2: rank 1 along the plane axis only
    q = np.outer([1, 2, 3, 2, 1], [1, 2, 3, 2, 1]) + np.outer([1, 0, 0, 0, 1], [0, 1, 0, 1, 0])
    w = np.einsum("i,jk->ijk", [1.0, 2, 4, 2, 1], q).reshape(-1)
    d = decompose_3d_r2("box3d2r", w)
    assert d["form"] == "hsep5" and d["recon_err"] == 0.0 and np.array_equal(np.einsum("i,jk->ijk", d["a"], d["q"]).reshape(-1), w)
    # LORA_R2_SEP5=0 keeps separable tables on the two-cells-per-thread kernel
    monkeypatch.setenv("LORA_R2_SEP5", "0")
    assert decompose_3d_r2("box3d2r", oracle.reference_params_r2("box3d2r"))["form"] == "hsep5"
    with pytest.raises(_lib.LoraError):
        decompose_3d_r2("box3d1r", np.zeros(125))


def test_fully_separable_tables_factor_exactly_when_the_sep5_form_is_on(monkeypatch):
    """LORA_R2_SEP5=1: a (x) b (x) c tables take the 5 + 5 + 5 form; integer tables factor into integers (exact taps)."""
    monkeypatch.setenv("LORA_R2_SEP5", "1")
    w = oracle.reference_params_r2("box3d2r")
    assert np.array_equal(effective_weights("box3d2r", ls.WEIGHTS_GENERAL, w), w)
    rng = np.random.default_rng(4)
    w = np.einsum("i,j,k->ijk", rng.uniform(-1, 1, 5), rng.uniform(-1, 1, 5), rng.uniform(-1, 1, 5)).reshape(-1)
    assert np.abs(effective_weights("box3d2r", ls.WEIGHTS_GENERAL, w) - w).max() <= 64 * 2.3e-16 * np.abs(w).max()
    # rank 1 along the plane axis only: stays what it was
    w = np.einsum("i,jk->ijk", rng.uniform(-1, 1, 5), rng.uniform(-1, 1, (5, 5))).reshape(-1)
    assert np.abs(effective_weights("box3d2r", ls.WEIGHTS_GENERAL, w) - w).max() <= 64 * 2.3e-16 * np.abs(w).max()


def test_launch_geometry_covers_every_plane_in_deep_grids():
    """Host planning of the radius-2 launches (csrc/stencil3d_r2.cu: r2_planes_per_chunk): chunks cover all planes, none
    is empty, never more than ceil(h / 16) of them, about 32 CTAs per SM when the grid allows, z <= 65535."""
    from ctypes import c_longlong
    L = _lib.lib()
    out = (c_longlong * 4)()
    cols = {(11, 0): 128, (11, 2): 256, (12, 1): 128, (13, 2): 256, (14, 0): 112, (14, 2): 112}
    for (form, variant), c in cols.items():
        for h, m, n in ((512, 512, 512), (1, 1, 1), (17, 3, 64), (1024, 1024, 1024), (100000, 2, 2), (3000000, 1, 1), (40, 33, 130)):
            assert L.lora_debug_r2_grid(form, variant, h, m, n, 148, out) == 0
            gx, gy, chunks, per = list(out)
            assert gx == -(-n // c) and gy == -(-m // 2)
            assert 1 <= chunks <= 65535 and (chunks - 1) * per < h <= chunks * per  # every plane, no empty chunk
            assert 2 * per >= min(h, 16)  # never more than ceil(h / 16) chunks: a chunk re-reads 4 planes of warm-up
            if h >= 64 and gx * gy * (h // 16) >= 32 * 148:
                assert gx * gy * chunks >= 0.5 * 32 * 148, (form, variant, h, m, n)
    assert L.lora_debug_r2_grid(5, 0, 8, 8, 8, 148, out) != 0  # not a radius-2 form


def test_reference_only_entry_points_refuse_the_new_shapes():
    """Slabs (and everything else that asks shape_dim) do not know these shapes: an error, not a wrong layout."""
    from ctypes import POINTER, byref, c_double, c_longlong, c_void_p
    L = _lib.lib()
    out = c_void_p()
    d = (c_longlong * 3)(8, 8, 8)
    rc = L.lora_slabset_create(byref(out), _lib.SHAPE_IDS["box3d2r"], 0, None, d, 2, None)
    assert rc == 1 and not out.value and b"unknown shape" in L.lora_last_error()


VARIANTS = ["0", "1", "2"]  # csrc/stencil3d_r2.cu: one cell per thread / + read-only loads, 4 planes per trip / two cells


@pytest.mark.gpu
@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("dims", [(12, 10, 64), (1, 1, 1), (5, 9, 31), (40, 33, 130), (70, 6, 300), (100, 4, 2)])
def test_drop_in_operators_match_the_checker(dims, variant, monkeypatch):
    monkeypatch.setenv("LORA_R2_VARIANT", variant)  # read at every launch
    rng = np.random.default_rng(7)
    ops.set_verbose(False)
    for name, shape, w, _ in tables(rng):
        a = rng.uniform(-1, 1, oracle.padded_shape_r2(dims))
        for times in (0, 1, 2, 5):
            out = np.full_like(a, -7.0)
            ops.BY_SHAPE[shape](a, out, w, times, *dims)
            ref = oracle.run_r2(a, w, times)
            assert np.abs(out - ref).max() <= RTOL * max(np.abs(ref).max(), 1e-300), (name, dims, times)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", VARIANTS)
def test_integer_data_is_exact_for_the_first_launches(variant, monkeypatch):
    """Small-integer data and the default tables: every sum is exact in FP64 whatever its order (as for the reference's
    own shapes) -- bit-identical to the checker."""
    monkeypatch.setenv("LORA_R2_VARIANT", variant)
    dims = (20, 16, 128)
    ops.set_verbose(False)
    for shape, upto in (("box3d2r", 4), ("star3d2r", 8)):
        a = np.random.default_rng(3).integers(0, 100, oracle.padded_shape_r2(dims)).astype(np.float64)
        w = oracle.reference_params_r2(shape)
        for times in range(1, upto + 1):
            out = np.zeros_like(a)
            ops.BY_SHAPE[shape](a, out, w, times, *dims)
            assert np.array_equal(out, oracle.run_r2(a, w, times)), (shape, times)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", VARIANTS + [None])
def test_plan_api_ranges_boundaries_and_torch_op(variant, monkeypatch):
    import torch
    if variant is None:
        monkeypatch.delenv("LORA_R2_VARIANT", raising=False)  # the default kernel
    else:
        monkeypatch.setenv("LORA_R2_VARIANT", variant)
    import lorastencil_b200.torch_ops  # noqa: F401
    rng = np.random.default_rng(11)
    dims = (30, 20, 66)
    a = rng.uniform(-1, 1, oracle.padded_shape_r2(dims))
    for shape in oracle.R2_SHAPES:
        w = oracle.reference_params_r2(shape)
        plan = ls.Plan(shape, dims)
        assert plan.padded_shape == a.shape and plan.temporal_block == 1
        assert plan.describe.startswith("3d radius-2")
        # one launch cut into plane ranges == one launch over everything
        src, whole, parts = torch.from_numpy(a).cuda(), plan.new_buffer(), plan.new_buffer()
        plan.step(src, whole)
        for lo, hi in ((0, 7), (7, 8), (8, 30)):
            plan.step(src, parts, lo, hi)
        torch.cuda.synchronize()
        assert torch.equal(whole, parts)
        assert np.abs(whole.cpu().numpy() - oracle.step_r2(a, w)).max() <= RTOL * np.abs(a).max() * np.abs(w).sum()
        # fused sweeps do not exist for these shapes
        plan.temporal_block = 2
        assert plan.temporal_block == 1
        # periodic boundary on the radius-2 layout
        plan.boundary = "periodic"
        got = plan.run(torch.from_numpy(a).cuda(), plan.new_buffer(), 3).cpu().numpy()
        cur = np.pad(a[INNER], [(2, 2), (2, 2), (4, 4)], mode="wrap")
        for _ in range(3):
            cur = np.pad(oracle.step_r2(cur, w)[INNER], [(2, 2), (2, 2), (4, 4)], mode="wrap")
        assert np.abs(got - cur).max() <= RTOL * np.abs(cur).max()
        # torch op, reference halo semantics
        y = torch.ops.lorastencil.stencil3d(torch.from_numpy(a).cuda(), shape, 4)
        ref = oracle.run_r2(a, w, 4)
        assert np.abs(y.cpu().numpy() - ref).max() <= RTOL * np.abs(ref).max()


@pytest.mark.gpu
@pytest.mark.parametrize("dims", [(12, 10, 64), (1, 1, 1), (5, 9, 31), (40, 33, 130), (70, 6, 300), (9, 5, 27), (9, 5, 28),
                                  (9, 5, 29), (6, 3, 112), (6, 3, 113), (20, 16, 128)])
def test_fully_separable_form(dims, monkeypatch):
    """LORA_R2_SEP5=1 (csrc/stencil3d_r2.cu: k_stencil3d_r2_sep -- row sums exchanged between lanes by shuffle, warps of 28
    stored columns): the default box table (exact on integer data) and a random a (x) b (x) c table, column counts
    around the warp and CTA widths (28, 112)."""
    import torch
    monkeypatch.setenv("LORA_R2_SEP5", "1")
    ops.set_verbose(False)
    rng = np.random.default_rng(5)
    plan = ls.Plan("box3d2r", dims)
    assert "separable a(h) x b(m) x c(n)" in plan.describe
    w = oracle.reference_params_r2("box3d2r")
    a = rng.integers(0, 100, oracle.padded_shape_r2(dims)).astype(np.float64)
    for times in (1, 2, 3, 4):
        out = np.full_like(a, -7.0)
        ops.gpu_box_3d2r(a, out, w, times, *dims)
        assert np.array_equal(out, oracle.run_r2(a, w, times)), (dims, times)
    ws = np.einsum("i,j,k->ijk", rng.uniform(-1, 1, 5), rng.uniform(-1, 1, 5), rng.uniform(-1, 1, 5)).reshape(-1)
    af = rng.uniform(-1, 1, a.shape)
    plan = ls.Plan("box3d2r", dims, params=ws, mode=ls.WEIGHTS_GENERAL)
    assert "separable a(h) x b(m) x c(n)" in plan.describe
    for times in (1, 5):
        got = plan.run(torch.from_numpy(af).cuda(), plan.new_buffer(), times).cpu().numpy()
        ref = oracle.run_r2(af, ws, times)
        assert np.abs(got - ref).max() <= RTOL * np.abs(ref).max(), (dims, times)
    # plane ranges of one launch
    src, whole, parts = torch.from_numpy(af).cuda(), plan.new_buffer(), plan.new_buffer()
    plan.step(src, whole)
    cut = max(1, dims[0] // 3)
    for lo, hi in ((0, cut), (cut, dims[0])):
        plan.step(src, parts, lo, hi)
    torch.cuda.synchronize()
    assert torch.equal(whole, parts)

"""Periodic boundary mode (LORA_BOUNDARY_PERIODIC; SURVEY.md section 8(f)-4: the reference itself never updates a halo,
S2 at src/2d/gpu.cu:396-400).

CPU: (i) the checker -- oracle.run_periodic = the reference's test_cpu step on a torus -- is pinned against an
independent implementation, scipy.ndimage.correlate(mode='wrap') on the interior; (ii) the product's halo refresh
(the work items and axis order of csrc/boundary.cu, run on host memory through lora_debug_wrap_ring_host) equals numpy's
wrap padding.  GPU: lora_plan_run in periodic mode against the checker."""
from ctypes import POINTER, c_double, c_longlong

import numpy as np
import pytest

import lorastencil_b200 as ls
import oracle
from lorastencil_b200 import _lib

RTOL = 1e-12
CASES = [("1d2r", (5000,)), ("1d1r", (4,)), ("1d2r", (1027,)), ("star2d3r", (40, 66)), ("box2d3r", (4, 4)),
         ("star2d1r", (33, 71)), ("box2d1r", (64, 128)), ("box3d1r", (6, 10, 64)), ("star3d1r", (1, 2, 4)),
         ("box3d1r", (5, 9, 31)), ("star3d1r", (12, 32, 128))]


def weights_nd(shape, w):
    d = oracle.dim_of(shape)
    return np.asarray(w, dtype=np.float64).reshape({1: (9,), 2: (7, 7), 3: (3, 3, 3)}[d])


@pytest.mark.parametrize("shape,dims", CASES)
def test_periodic_oracle_equals_scipy_wrap_correlation(shape, dims):
    from scipy import ndimage
    rng = np.random.default_rng(11)
    a = rng.integers(0, 100, oracle.padded_shape(shape, dims)).astype(np.float64)  # integers: sums are exact
    w = np.round(rng.uniform(-4, 4, oracle.reference_params(shape).size))
    inner = tuple(slice(h, -h) for h in oracle.HALO[len(dims)])
    cur = a[inner].copy()
    for times in (1, 2, 3):
        cur = ndimage.correlate(cur, weights_nd(shape, w), mode="wrap")
        got = oracle.run_periodic(shape, a, w, times)
        assert np.array_equal(got[inner], cur), (shape, dims, times)
        assert np.array_equal(got, oracle.wrap_ring(got))  # the ring of the result is the image of its interior


@pytest.mark.parametrize("shape,dims", CASES)
def test_host_restatement_of_the_halo_refresh_equals_numpy_wrap_padding(shape, dims):
    rng = np.random.default_rng(5)
    a = rng.uniform(-1, 1, oracle.padded_shape(shape, dims))
    got = a.copy()
    d = (c_longlong * 3)(*dims, *([0] * (3 - len(dims))))
    _lib.check(_lib.lib().lora_debug_wrap_ring_host(len(dims), d, got.ctypes.data_as(POINTER(c_double))), "wrap")
    assert np.array_equal(got, oracle.wrap_ring(a))
    inner = tuple(slice(h, -h) for h in oracle.HALO[len(dims)])
    assert np.array_equal(got[inner], a[inner])  # the interior is never written


def test_grids_thinner_than_their_storage_halo_are_refused():
    d = (c_longlong * 3)(8, 3, 0)
    buf = np.zeros((16, 11))
    rc = _lib.lib().lora_debug_wrap_ring_host(2, d, buf.ctypes.data_as(POINTER(c_double)))
    assert rc == 3 and b"thinner" in _lib.lib().lora_last_error()


@pytest.mark.gpu
@pytest.mark.parametrize("shape,dims", CASES + [("1d2r", (70001,)), ("star2d3r", (300, 258)), ("box2d1r", (257, 400)),
                                                ("star3d1r", (40, 50, 130))])
def test_periodic_run_matches_the_checker(shape, dims):
    import torch
    rng = np.random.default_rng(3)
    a = rng.uniform(-1, 1, oracle.padded_shape(shape, dims))
    w = oracle.effective_params(shape)
    plan = ls.Plan(shape, dims)
    plan.boundary = "periodic"
    assert plan.boundary == "periodic"
    for times in (0, 1, 2, 5, 16):
        b0, b1 = torch.from_numpy(a).cuda(), plan.new_buffer()
        res = plan.run(b0, b1, times)
        torch.cuda.synchronize()
        got = res.cpu().numpy()
        ref = oracle.run_periodic(shape, a, w, times)
        assert np.abs(got - ref).max() <= RTOL * np.abs(ref).max(), (shape, dims, times)
    # fused sweeps are refused in this mode rather than computed with the wrong halo
    if len(dims) < 3 and dims[-1] % 2 == 0:
        with pytest.raises(_lib.LoraError, match="periodic"):
            plan.step_fused(b0, b1, b0, 0, dims[0], 3, 0, 1, 1)


@pytest.mark.gpu
def test_periodic_general_weights_and_wrap_ring_on_its_own():
    """An asymmetric dense 7x7 table (direct taps) on a torus; Plan.wrap_ring + Plan.step driven by hand equals run."""
    import torch
    rng = np.random.default_rng(9)
    shape, dims = "box2d3r", (96, 130)
    w = rng.uniform(-1, 1, 49)
    a = rng.uniform(-1, 1, oracle.padded_shape(shape, dims))
    plan = ls.Plan(shape, dims, params=w, mode=_lib.WEIGHTS_GENERAL)
    plan.boundary = "periodic"
    b0, b1 = torch.from_numpy(a).cuda(), plan.new_buffer()
    got = plan.run(b0, b1, 3).cpu().numpy()
    ref = oracle.run_periodic(shape, a, w, 3)
    assert np.abs(got - ref).max() <= RTOL * np.abs(ref).max()
    plan.boundary = "reference"
    c = [torch.from_numpy(a).cuda(), plan.new_buffer()]
    for i in range(3):
        plan.wrap_ring(c[i % 2])
        plan.step(c[i % 2], c[(i + 1) % 2])
    plan.wrap_ring(c[1])
    torch.cuda.synchronize()
    assert np.array_equal(c[1].cpu().numpy(), got)


@pytest.mark.parametrize("shape,dims", [("1d2r", (257,)), ("star2d1r", (12, 20)), ("box2d3r", (9, 16)), ("box3d1r", (5, 6, 12))])
def test_periodic_checker_is_shift_invariant_and_linear(shape, dims):
    """Size-independent properties of a stencil on a torus: rolling the interior rolls the result; the operator is linear.
    (Integer data and weights: both hold exactly.)"""
    rng = np.random.default_rng(21)
    d = len(dims)
    inner = tuple(slice(h, -h) for h in oracle.HALO[d])
    w = np.round(rng.uniform(-3, 3, oracle.reference_params(shape).size))

    def padded(interior):
        a = np.zeros(oracle.padded_shape(shape, dims))
        a[inner] = interior
        return a

    x = rng.integers(0, 50, dims).astype(np.float64)
    y = rng.integers(0, 50, dims).astype(np.float64)
    fx = oracle.run_periodic(shape, padded(x), w, 3)[inner]
    fy = oracle.run_periodic(shape, padded(y), w, 3)[inner]
    shift = tuple(int(s) for s in rng.integers(1, 5, d))
    rolled = oracle.run_periodic(shape, padded(np.roll(x, shift, axis=tuple(range(d)))), w, 3)[inner]
    assert np.array_equal(rolled, np.roll(fx, shift, axis=tuple(range(d))))
    assert np.array_equal(oracle.run_periodic(shape, padded(2 * x - 3 * y), w, 3)[inner], 2 * fx - 3 * fy)
    # the caller's halo values are never read
    junk = padded(x)
    ring = np.ones_like(junk, dtype=bool)
    ring[inner] = False
    junk[ring] = 1e30
    assert np.array_equal(oracle.run_periodic(shape, junk, w, 3)[inner], fx)

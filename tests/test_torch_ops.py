"""torch.ops.lorastencil.{stencil1d, stencil2d, stencil3d} (lorastencil_b200/torch_ops.py): registration and shape
propagation on CPU; numerics on the GPU against the oracle."""
import numpy as np
import pytest
import torch

import oracle
import lorastencil_b200 as ls
import lorastencil_b200.torch_ops  # noqa: F401  (registers the ops)


def test_ops_are_registered_with_meta_kernels_and_no_cpu_fallback():
    for d, shape, padded in ((1, "1d2r", (1032,)), (2, "box2d3r", (72, 136)), (3, "star3d1r", (10, 20, 72))):
        op = getattr(torch.ops.lorastencil, f"stencil{d}d")
        y = op(torch.empty(padded, dtype=torch.float64, device="meta"), shape, 3)
        assert y.shape == padded and y.dtype == torch.float64 and y.device.type == "meta"
        with pytest.raises(ls.LoraError):
            op(torch.zeros(padded, dtype=torch.float64), shape, 3)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,dims,times", [("1d2r", (5000,), 17), ("1d1r", (1024,), 4), ("box2d1r", (64, 130), 4),
                                              ("star2d3r", (70, 250), 7), ("star2d1r", (64, 64), 3),
                                              ("box3d1r", (9, 34, 130), 4), ("star3d1r", (12, 8, 64), 5)])
def test_ops_match_the_oracle(shape, dims, times):
    d = len(dims)
    a = oracle.fill_rand(shape, dims)
    x = torch.from_numpy(a).cuda()
    op = getattr(torch.ops.lorastencil, f"stencil{d}d")
    y = op(x, shape, times)
    torch.cuda.synchronize()
    assert torch.equal(x.cpu(), torch.from_numpy(a))  # the input is not written
    ref = oracle.run(shape, a, oracle.effective_params(shape), times)
    got = y.cpu().numpy()
    if d == 1:
        got, ref = got[:-1], ref[:-1]
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
    # general weights through `params` + mode 1: every weight honoured (== test_cpu)
    rng = np.random.default_rng(5)
    w = rng.standard_normal({1: 9, 2: 49, 3: 27}[d])
    af = rng.uniform(-1, 1, a.shape)
    y = op(torch.from_numpy(af).cuda(), shape, 2, torch.from_numpy(w), 1)
    ref = oracle.run(d, af, w, 2)
    got = y.cpu().numpy()
    if d == 1:
        got, ref = got[:-1], ref[:-1]
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()


def test_boundary_argument_is_validated_before_any_plan_is_made():
    op = torch.ops.lorastencil.stencil2d
    y = op(torch.empty((72, 136), dtype=torch.float64, device="meta"), "box2d3r", 3, None, 0, 3)
    assert y.shape == (72, 136)
    assert "int boundary=0" in str(op.default._schema)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,dims", [("1d2r", (5000,)), ("star2d3r", (70, 250)), ("box3d1r", (9, 34, 130))])
def test_ops_periodic_boundary(shape, dims):
    d = len(dims)
    rng = np.random.default_rng(17)
    a = rng.uniform(-1, 1, oracle.padded_shape(shape, dims))
    op = getattr(torch.ops.lorastencil, f"stencil{d}d")
    y = op(torch.from_numpy(a).cuda(), shape, 5, None, 0, 3)
    ref = oracle.run_periodic(shape, a, oracle.effective_params(shape), 5)
    assert np.abs(y.cpu().numpy() - ref).max() <= 1e-12 * np.abs(ref).max()
    # the cached plan goes back to the reference's halo semantics when the next call asks for them
    y = op(torch.from_numpy(a).cuda(), shape, 5)
    ref = oracle.run(shape, a, oracle.effective_params(shape), 5)
    got = y.cpu().numpy()
    if d == 1:
        got, ref = got[:-1], ref[:-1]
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
    with pytest.raises(ValueError, match="boundary"):
        op(torch.from_numpy(a).cuda(), shape, 1, None, 0, 7)

"""Boundary modes of the plan API (lora_plan_set_boundary): 'reference' = the reference's
alternating caller's / zero halo (S2), 'dirichlet' = the caller's halo values are the boundary condition of every
launch, 'zero' = zero halo for every launch.  Checked against a plain oracle loop (one test_cpu step per launch with the
halo ring rewritten by hand), through the unfused kernels and the fused sweeps (1-D: 15 launches, 2-D: 3 / 2, 3-D: 2)."""
import numpy as np
import pytest

import lorastencil_b200 as ls
import oracle

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def oracle_boundary(shape, a, w, times, mode):
    d = oracle.dim_of(shape)
    inner = tuple(slice(h, -h) for h in oracle.HALO[d])
    cur = a.copy()
    if mode == "zero":
        keep = cur[inner].copy()
        cur[...] = 0.0
        cur[inner] = keep
    for _ in range(times):
        nxt = cur.copy()  # the ring is carried over unchanged: fixed boundary values
        nxt[inner] = oracle.step(d, cur, w)[inner]
        cur = nxt
    return cur


@pytest.mark.parametrize("shape,dims", [("1d2r", (5000,)), ("1d1r", (70000,)), ("star2d3r", (300, 258)), ("star2d1r", (64, 130)),
                                        ("box2d3r", (96, 128)), ("box2d1r", (50, 71)), ("box3d1r", (12, 32, 128)),
                                        ("star3d1r", (9, 40, 64)), ("box2d1r", (300, 258)), ("star2d1r", (257, 400)),
                                        ("star3d1r", (40, 50, 130))])
@pytest.mark.parametrize("mode", ["dirichlet", "zero"])
def test_fixed_boundary_modes_match_an_oracle_loop(shape, dims, mode):
    import torch
    rng = np.random.default_rng(3)
    a = rng.uniform(-1, 1, oracle.padded_shape(shape, dims))
    w = oracle.effective_params(shape)
    plan = ls.Plan(shape, dims)
    assert plan.boundary == "reference"
    plan.boundary = mode
    assert plan.boundary == mode
    for times in (1, 2, 3, 7, 16, 31):
        results = []
        # fused sweeps of every kind: 1-D 15 launches, 2-D cross 3, 2-D diamond / pyramid 2 (3 for the diamond too), 3-D 2
        # (a Dirichlet boundary makes 3-D fall back to single launches); odd column counts have no fused kernels
        if len(dims) == 1:
            tbs = (1, 15)
        elif dims[-1] % 2:
            tbs = (1,)
        elif shape == "star2d3r":
            tbs = (1, 3)
        elif shape == "star2d1r":
            tbs = (1, 2, 3)
        else:
            tbs = (1, 2)
        for tb in tbs:
            plan.temporal_block = tb
            b0, b1 = torch.from_numpy(a).cuda(), plan.new_buffer()
            res = plan.run(b0, b1, times)
            torch.cuda.synchronize()
            results.append(res.cpu().numpy())
        ref = oracle_boundary(shape, a, w, times, mode)
        for got in results:
            g, r = (got[:-1], ref[:-1]) if len(dims) == 1 else (got, ref)
            assert np.abs(g - r).max() <= RTOL * np.abs(r).max(), (shape, mode, times)
        for other in results[1:]:
            assert np.array_equal(results[0], other), (shape, mode, times)  # fused == unfused, bit for bit


def test_reference_mode_is_unchanged_by_the_option():
    import torch
    shape, dims = "star2d3r", (70, 200)
    a = oracle.fill_rand(shape, dims)
    plan = ls.Plan(shape, dims)
    plan.boundary = "dirichlet"
    plan.boundary = "reference"
    b0, b1 = torch.from_numpy(a).cuda(), plan.new_buffer()
    res = plan.run(b0, b1, 7)
    torch.cuda.synchronize()
    assert np.array_equal(res.cpu().numpy(), oracle.run(shape, a, oracle.effective_params(shape), 7))

"""Small end-to-end pass over every kernel family, written for `compute-sanitizer --tool memcheck`:

    compute-sanitizer --tool memcheck --error-exitcode 66 python tests/sanitize_small.py

(compute-sanitizer is closed on the GPU pool this was developed on -- the plain run passes; out-of-bounds stores are
hunted with canaries instead, tests/test_parity_gpu.py::test_no_write_outside_the_destination_interior.)

Covers: the drop-in operators of all 8 shapes at ragged sizes, fused 1-D sweeps (every temporal block depth, virtual
halo, sub-ranges), the chunked copy-overlapped 1-D operator, fused 2-D sweeps (edge strips, narrow grids), general
weights (direct49 / direct27 forms), the radius-2 3-D shapes (every form and kernel variant), the periodic halo
refresh.  Results are checked against the oracle as well."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import lorastencil_b200 as ls  # noqa: E402
import oracle  # noqa: E402
from lorastencil_b200 import ops  # noqa: E402

ops.set_verbose(False)
RTOL = 1e-12


def check(shape, got, ref, what):
    if oracle.dim_of(shape) == 1:
        got, ref = got[:-1], ref[:-1]
    err = np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300)
    assert err <= RTOL, (what, err)
    print(f"ok {what}: max rel err {err:.2g}", flush=True)


for shape, dims, times in [("1d1r", (3000,), 17), ("1d2r", (70001,), 31), ("star2d1r", (70, 250), 7), ("box2d1r", (64, 130), 4),
                           ("star2d3r", (130, 40), 10), ("box2d3r", (33, 66), 3), ("box3d1r", (9, 34, 130), 4),
                           ("star3d1r", (12, 8, 64), 5), ("star2d3r", (400, 370), 6), ("star2d1r", (9, 8), 4)]:
    a = np.random.default_rng(1).uniform(-1, 1, oracle.padded_shape(shape, dims))
    p = oracle.reference_params(shape)
    out = np.zeros_like(a)
    ops.BY_SHAPE[shape](a, out, p, times, *dims)
    check(shape, out, oracle.run(shape, a, oracle.effective_params(shape, p), times), f"{shape} {dims} x{times}")

# chunked 1-D operator
os.environ["LORA_CHUNKS"] = "3"
a = np.random.default_rng(2).uniform(-1, 1, (50000 + 8,))
out = np.zeros_like(a)
ops.gpu_1d2r(a, out, oracle.reference_params("1d2r"), 20, 50000)
assert ops.last_chunks() == 3
check("1d2r", out, oracle.run("1d2r", a, oracle.effective_params("1d2r"), 20), "1d2r chunked x20")
os.environ.pop("LORA_CHUNKS")

# fused 1-D sub-ranges, every depth
n = 9000
plan = ls.Plan("1d1r", (n,))
a = np.random.default_rng(3).integers(0, 10, (n + 8,)).astype(np.float64)
for tb in range(1, 16):
    src = torch.from_numpy(a).cuda()
    dst = torch.zeros(n + 8, dtype=torch.float64, device="cuda")
    lo, hi = 4 * tb + 1, n - 4 * tb - 3
    plan.step_fused(src, dst, None, lo, hi, tb, 0, False, False)
    torch.cuda.synchronize()
print("ok fused 1-D sub-ranges tb 1..15", flush=True)

# general weights: direct forms
w49 = np.random.default_rng(4).uniform(-1, 1, 49)
a = np.random.default_rng(5).uniform(-1, 1, (40 + 8, 136 + 8))
out = np.zeros_like(a)
ops.run_host("box2d1r", a, out, w49, 3, (40, 136), mode=ls.WEIGHTS_GENERAL)
check("box2d1r", out, oracle.run(2, a, w49, 3), "2-D general weights (direct49)")
w27 = np.random.default_rng(6).uniform(-1, 1, 27)
a = np.random.default_rng(7).uniform(-1, 1, (6 + 2, 20 + 4, 64 + 8))
out = np.zeros_like(a)
ops.run_host("box3d1r", a, out, w27, 2, (6, 20, 64), mode=ls.WEIGHTS_GENERAL)
check("box3d1r", out, oracle.run(3, a, w27, 2), "3-D general weights (direct27)")

# radius-2 3-D shapes: every form (13-point, fully separable, rank 1 along h, 125 taps) and every kernel variant
rng = np.random.default_rng(8)
dense = rng.uniform(-1, 1, 125)
hsep = np.einsum("i,jk->ijk", rng.uniform(-1, 1, 5), rng.uniform(-1, 1, (5, 5))).reshape(-1)
for variant in ("0", "1", "2"):
    os.environ["LORA_R2_VARIANT"] = variant
    for shape, w, dims in (("star3d2r", oracle.reference_params_r2("star3d2r"), (9, 7, 130)), ("box3d2r", oracle.reference_params_r2("box3d2r"), (20, 5, 113)),
                           ("box3d2r", hsep, (6, 9, 31)), ("box3d2r", dense, (17, 3, 64))):
        a = rng.uniform(-1, 1, oracle.padded_shape_r2(dims))
        out = np.zeros_like(a)
        ops.BY_SHAPE[shape](a, out, w, 3, *dims)
        ref = oracle.run_r2(a, w, 3)
        err = np.abs(out - ref).max() / np.abs(ref).max()
        assert err <= RTOL, (shape, dims, variant, err)
os.environ.pop("LORA_R2_VARIANT")
print("ok radius-2 shapes, variants 0..2", flush=True)

# periodic boundary: the halo refresh kernel on every layout
for shape, dims in (("1d2r", (5001,)), ("star2d3r", (33, 71)), ("box2d3r", (40, 66)), ("star3d1r", (5, 9, 31)), ("box3d1r", (6, 10, 64))):
    a = rng.uniform(-1, 1, oracle.padded_shape(shape, dims))
    plan = ls.Plan(shape, dims)
    plan.boundary = "periodic"
    got = plan.run(torch.from_numpy(a).cuda(), plan.new_buffer(), 4).cpu().numpy()
    ref = oracle.run_periodic(shape, a, oracle.effective_params(shape), 4)
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err <= RTOL, (shape, dims, err)
print("ok periodic boundary", flush=True)
print("sanitize_small OK")

"""Five-second sanity pass over the paths touched last (thread-local last-call figures, the radius-2 launch planner as a
host function, SEP5 on by default): drop-in box3d2r / 1d2r / box2d1r on small grids against the oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402

import lorastencil_b200 as ls  # noqa: E402
import oracle  # noqa: E402
from lorastencil_b200 import ops  # noqa: E402

ops.set_verbose(False)
rng = np.random.default_rng(1)
dims = (20, 9, 130)
a = rng.integers(0, 100, oracle.padded_shape_r2(dims)).astype(np.float64)
w = oracle.reference_params_r2("box3d2r")
out = np.zeros_like(a)
ops.gpu_box_3d2r(a, out, w, 3, *dims)
assert np.array_equal(out, oracle.run_r2(a, w, 3))
print("box3d2r default table (exact) ok; loop ms", ops.last_loop_ms(), "bands", ops.last_bands())
for shape, dims, times in (("1d2r", (70001,), 31), ("box2d1r", (64, 130), 4), ("star3d1r", (12, 8, 64), 5)):
    a = oracle.fill_rand(shape, dims)
    out = np.zeros_like(a)
    ops.BY_SHAPE[shape](a, out, oracle.reference_params(shape), times, *dims)
    ref = oracle.run(shape, a, oracle.effective_params(shape), times)
    if len(dims) == 1:
        out, ref = out[:-1], ref[:-1]
    assert np.abs(out - ref).max() <= 1e-12 * np.abs(ref).max(), shape
    print(shape, "ok; loop ms", ops.last_loop_ms(), "chunks", ops.last_chunks(), "gpus", ops.last_gpus())
print("final sanity OK")

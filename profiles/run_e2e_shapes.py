"""End-to-end time per call of the 2-D / 3-D drop-in operators (pinned host in -> GPU -> pinned host out, 100 launches at
the BASELINE sizes).   python profiles/run_e2e_shapes.py [--shapes star2d3r,box3d1r] [--launches 100]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import lorastencil_b200 as ls  # noqa: E402
from lorastencil_b200 import ops  # noqa: E402

SIZES = {"star2d1r": (10240, 10240), "box2d1r": (10240, 10240), "star2d3r": (10240, 10240), "box3d1r": (512, 512, 512),
         "star3d1r": (512, 512, 512)}
ap = argparse.ArgumentParser()
ap.add_argument("--shapes", default=",".join(SIZES))
ap.add_argument("--launches", type=int, default=100)
ap.add_argument("--reps", type=int, default=4)
args = ap.parse_args()
ops.set_verbose(False)
for shape in args.shapes.split(","):
    dims = SIZES[shape]
    plan = ls.Plan(shape, dims)
    hin = torch.randint(0, 100, plan.padded_shape).double().pin_memory()
    hout = torch.empty(plan.padded_shape, dtype=torch.float64).pin_memory()
    del plan
    params = ls.reference_table(shape)
    best = None
    for _ in range(args.reps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ops.BY_SHAPE[shape](hin, hout, params, args.launches, *dims)
        ms = (time.perf_counter() - t0) * 1e3
        best = ms if best is None else min(best, ms)
    cells = 1
    for d in dims:
        cells *= d
    print(f"{shape} {dims}: {best:.1f} ms per call, {cells * args.launches / best / 1e6:.1f} GStencil/s end to end, "
          f"bands {ops.last_bands()}, launch loop {ops.last_loop_ms():.1f} ms", flush=True)

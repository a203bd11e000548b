"""Short device-resident run of every shape at its BASELINE size -- the command profiled by ncu.

    python profiles/run_shapes.py [--launches 4] [--shapes 1d2r,box2d3r,...]
Prints one line per shape with the CUDA-event time per launch (not a bench number when run under ncu)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import lorastencil_b200 as ls  # noqa: E402

SIZES = {"1d1r": (1 << 28,), "1d2r": (1 << 28,), "star2d1r": (10240, 10240), "box2d1r": (10240, 10240),
         "star2d3r": (10240, 10240), "box2d3r": (10240, 10240), "box3d1r": (512, 512, 512), "star3d1r": (512, 512, 512)}

ap = argparse.ArgumentParser()
ap.add_argument("--launches", type=int, default=4)
ap.add_argument("--shapes", default=",".join(SIZES))
ap.add_argument("--tb", type=int, default=0, help="temporal block to request (0 = the plan's default)")
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--dims", default="", help="override the size, e.g. 1024x1024x1024 (applies to every shape listed)")
args = ap.parse_args()
for shape in args.shapes.split(","):
    dims = tuple(int(x) for x in args.dims.split("x")) if args.dims else SIZES[shape]
    plan = ls.Plan(shape, dims)
    if args.tb:
        plan.temporal_block = args.tb
    g = torch.Generator(device="cuda").manual_seed(1)
    b0 = torch.randint(0, 100, plan.padded_shape, generator=g, device="cuda").double()
    b1 = plan.new_buffer()
    torch.cuda.synchronize()
    best = None
    for _ in range(args.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.run(b0, b1, args.launches)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    cells = 1
    for d in dims:
        cells *= d
    print(f"{shape} {dims}: {best / args.launches * 1e3:.1f} us/time step, {cells * args.launches / best / 1e6:.1f} GStencil/s"
          f"  [tb {plan.temporal_block}; {plan.describe}]")
    del plan, b0, b1
    torch.cuda.empty_cache()

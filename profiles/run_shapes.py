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
args = ap.parse_args()
for shape in args.shapes.split(","):
    dims = SIZES[shape]
    plan = ls.Plan(shape, dims)
    g = torch.Generator(device="cuda").manual_seed(1)
    b0 = torch.randint(0, 100, plan.padded_shape, generator=g, device="cuda").double()
    b1 = plan.new_buffer()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    plan.run(b0, b1, args.launches)
    e1.record()
    torch.cuda.synchronize()
    print(f"{shape} {dims}: {e0.elapsed_time(e1) / args.launches * 1e3:.1f} us/launch  [{plan.describe}]")
    del plan, b0, b1
    torch.cuda.empty_cache()

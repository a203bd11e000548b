"""Device-resident timing of the section-8(f)-4 extensions: the radius-2 3-D shapes (stencil3d_r2.cu, three forms) at
512^3 and the periodic boundary mode against the reference halo semantics (one launch per step both times).

    python profiles/run_extensions.py [--launches 6] [--reps 3] [--variants 0,1,2] [--periodic 1] [--out FILE.json]
CUDA events on the launching stream, `--reps` repetitions after one warm-up run, best and median."""
import argparse
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import lorastencil_b200 as ls  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--launches", type=int, default=6)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--out", default="")
ap.add_argument("--variants", default="0,1,2", help="kernel variants of the radius-2 shapes to time (LORA_R2_VARIANT)")
ap.add_argument("--periodic", type=int, default=1, help="also time the periodic boundary mode against the reference one")
args = ap.parse_args()


def timed(plan, launches):
    g = torch.Generator(device="cuda").manual_seed(1)
    b0 = torch.randint(0, 100, plan.padded_shape, generator=g, device="cuda").double()
    b1 = plan.new_buffer()
    plan.run(b0, b1, 2)  # warm-up
    torch.cuda.synchronize()
    ms = []
    for _ in range(args.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.run(b0, b1, launches)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    cells = float(np.prod(plan.dims))
    rate = [cells * launches / m / 1e6 for m in ms]
    return {"gstencil_best": round(max(rate), 1), "gstencil_median": round(statistics.median(rate), 1),
            "us_per_launch_best": round(min(ms) / launches * 1e3, 1), "describe": plan.describe}


rows = []
rng = np.random.default_rng(0)
dense = rng.uniform(-1, 1, 125)
for name, shape, params in (("star3d2r default (13 taps)", "star3d2r", None), ("box3d2r default (rank 1 along h: 30 taps)", "box3d2r", None),
                            ("box3d2r dense table (125 taps)", "box3d2r", dense)):
    for variant in args.variants.split(","):
        os.environ["LORA_R2_VARIANT"] = variant  # csrc/stencil3d_r2.cu reads it at every launch
        plan = ls.Plan(shape, (512, 512, 512), params=params, mode=ls.WEIGHTS_GENERAL)
        r = {"case": name, "variant": int(variant), "dims": [512, 512, 512], "launches": args.launches, **timed(plan, args.launches)}
        r["fixed_16B_roofline_frac"] = round(r["gstencil_best"] / 403.5, 3)  # 16 B per cell and launch at 6455.6 GB/s
        rows.append(r)
        print(json.dumps(r), flush=True)
        del plan
        torch.cuda.empty_cache()
os.environ.pop("LORA_R2_VARIANT", None)
for sep5 in ("0", "1"):  # the default box table is fully separable: 25 + 5 taps (two cells per thread) vs 5 + 5 + 5
    os.environ["LORA_R2_SEP5"] = sep5
    plan = ls.Plan("box3d2r", (512, 512, 512))
    r = {"case": f"box3d2r default, LORA_R2_SEP5={sep5}", "dims": [512, 512, 512], "launches": args.launches, **timed(plan, args.launches)}
    r["fixed_16B_roofline_frac"] = round(r["gstencil_best"] / 403.5, 3)
    rows.append(r)
    print(json.dumps(r), flush=True)
    del plan
    torch.cuda.empty_cache()
os.environ.pop("LORA_R2_SEP5", None)
if args.periodic:
    for shape, dims in (("star2d3r", (10240, 10240)), ("box3d1r", (512, 512, 512)), ("1d2r", (1 << 26,))):
        for boundary in ("reference", "periodic"):
            plan = ls.Plan(shape, dims)
            plan.temporal_block = 1  # the periodic mode runs one launch per step: compare like with like
            plan.boundary = boundary
            r = {"case": f"{shape} unfused, boundary {boundary}", "dims": list(dims), "launches": args.launches,
                 **timed(plan, args.launches)}
            rows.append(r)
            print(json.dumps(r), flush=True)
            del plan
            torch.cuda.empty_cache()
if args.out:
    with open(args.out, "w") as f:
        json.dump({"device": torch.cuda.get_device_name(0), "rows": rows}, f, indent=1)

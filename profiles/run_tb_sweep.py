"""Temporal-block sweep of the 1-D fused kernel: GStencil/s (device-resident, CUDA events) for TB = 1..8.

    python profiles/run_tb_sweep.py [--n 268435456] [--times 96] [--tbs 1,2,4,8]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import lorastencil_b200 as ls  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 28)
ap.add_argument("--times", type=int, default=96)
ap.add_argument("--tbs", default="1,2,3,4,5,6,7,8")
ap.add_argument("--shape", default="1d2r")
args = ap.parse_args()
plan = ls.Plan(args.shape, (args.n,))
g = torch.Generator(device="cuda").manual_seed(1)
b0 = torch.randint(0, 10, plan.padded_shape, generator=g, device="cuda").double()
b1 = plan.new_buffer()
out = []
for tb in [int(t) for t in args.tbs.split(",")]:
    plan.temporal_block = tb
    plan.run(b0, b1, 2 * tb)
    torch.cuda.synchronize()
    best = None
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = plan.launches
        e0.record()
        plan.run(b0, b1, args.times)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
        nl = plan.launches - l0
    r = {"tb": tb, "gstencils": args.n * args.times / best / 1e6, "ms": best, "kernel_launches": nl,
         "us_per_kernel": best * 1e3 / nl}
    out.append(r)
    print(json.dumps(r), flush=True)

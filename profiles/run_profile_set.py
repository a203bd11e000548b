"""The launches `ncu --set full --profile-from-start off` captures for the round's kernel evidence: after a warm-up,
exactly one launch (or one fused sweep) of every production kernel at its BASELINE size, between
cudaProfilerStart / cudaProfilerStop.

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2_prof \
        python profiles/run_profile_set.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import lorastencil_b200 as ls  # noqa: E402

# (shape, dims, temporal block, launches in the captured run)
SET = [("1d2r", (1 << 28,), 15, 15), ("1d1r", (1 << 28,), 15, 15), ("star2d3r", (10240, 10240), 3, 3),
       ("star2d1r", (10240, 10240), 3, 3), ("star2d1r", (10240, 10240), 1, 1), ("box2d3r", (10240, 10240), 1, 1),
       ("box3d1r", (512, 512, 512), 1, 1), ("star3d1r", (512, 512, 512), 1, 1),
       ("star3d1r", (512, 512, 512), 2, 4), ("box3d1r", (512, 512, 512), 2, 4),
       ("box2d3r", (10240, 10240), 2, 4), ("star2d1r", (10240, 10240), 2, 4)]
only = sys.argv[1].split(",") if len(sys.argv) > 1 else None  # e.g. star3d1r:2,box3d1r:2 or 1d2r
for shape, dims, tb, launches in SET:
    if only and shape not in only and f"{shape}:{tb}" not in only:
        continue
    plan = ls.Plan(shape, dims)
    plan.temporal_block = tb
    g = torch.Generator(device="cuda").manual_seed(1)
    b0 = torch.randint(0, 100, plan.padded_shape, generator=g, device="cuda").double()
    b1 = plan.new_buffer()
    plan.run(b0, b1, 2 * launches)  # warm-up (not captured); even count: the captured run starts from buffer 0 again
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    plan.run(b0, b1, launches)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(f"captured {shape} {dims} tb={plan.temporal_block} launches={launches}: {plan.describe}", flush=True)
    del plan, b0, b1
    torch.cuda.empty_cache()

import os, sys
sys.path.insert(0, "/root/repo")
import torch, lorastencil_b200 as ls
for dims, launches in (((10240, 10240), 60), ((40960, 40960), 10)):
    for shape in ("box2d1r", "star2d3r"):
        plan = ls.Plan(shape, dims); plan.temporal_block = 1
        b0 = torch.rand(plan.padded_shape, dtype=torch.float64, device="cuda"); b1 = plan.new_buffer()
        plan.run(b0, b1, 2); torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); plan.run(b0, b1, launches); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"  {shape} {dims}: {dims[0]*dims[1]*launches/best/1e6:.1f} GStencil/s")
        del plan, b0, b1; torch.cuda.empty_cache()

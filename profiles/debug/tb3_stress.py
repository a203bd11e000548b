"""Run-to-run determinism of the 3-D kernels: the same job N times per (shape, temporal block); any run that differs
from the first one is a race.  python profiles/debug/tb3_stress.py [N]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

import lorastencil_b200 as ls  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for shape in ("box3d1r", "star3d1r"):
    for dims in ((70, 40, 250), (40, 50, 130), (64, 64, 64)):
        plan = ls.Plan(shape, dims)
        a = np.random.default_rng(1).integers(0, 100, plan.padded_shape).astype(np.float64)
        for tb in (1, 2):
            plan.temporal_block = tb
            first, bad = None, 0
            for _ in range(N):
                b0, b1 = torch.from_numpy(a).cuda(), plan.new_buffer()
                r = plan.run(b0, b1, 8).clone()
                torch.cuda.synchronize()
                if first is None:
                    first = r
                elif not torch.equal(first, r):
                    bad += 1
            print(f"{shape} {dims} tb {tb}: {bad} of {N - 1} repeats differ from the first run", flush=True)

"""Where do fused (tb = 2) and unfused 3-D runs differ?  python profiles/debug/tb3_mismatch.py [shape] [h m n] [times]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

import lorastencil_b200 as ls  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "box3d1r"
dims = tuple(int(x) for x in sys.argv[2:5]) if len(sys.argv) > 4 else (40, 50, 130)
times = int(sys.argv[5]) if len(sys.argv) > 5 else 4
plan = ls.Plan(shape, dims)
rng = np.random.default_rng(1)
a = rng.integers(0, 100, plan.padded_shape).astype(np.float64)
res = []
for tb in (1, 2):
    plan.temporal_block = tb
    b0, b1 = torch.from_numpy(a).cuda(), plan.new_buffer()
    r = plan.run(b0, b1, times)
    torch.cuda.synchronize()
    res.append(r.cpu().numpy())
bad = np.argwhere(res[0] != res[1])
print(shape, dims, times, "mismatches:", len(bad))
if len(bad):
    print("planes", np.unique(bad[:, 0])[:40])
    print("rows", np.unique(bad[:, 1])[:60])
    print("cols", np.unique(bad[:, 2])[:80])
    for idx in bad[:8]:
        print(tuple(idx), res[0][tuple(idx)], res[1][tuple(idx)])

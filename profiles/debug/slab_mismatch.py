"""Debug helper: where do slab runs (LORA_DEVICES) differ from the single-GPU operator?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import oracle
from lorastencil_b200 import ops
ops.set_verbose(False)
shape = sys.argv[1]; dims = tuple(int(x) for x in sys.argv[2].split(","))
rng = np.random.default_rng(1)
a = rng.integers(0, 100, size=oracle.padded_shape(shape, dims)).astype(np.float64)
p = oracle.reference_params(shape)
for times in [int(x) for x in sys.argv[3].split(",")]:
    for k in [int(x) for x in sys.argv[4].split(",")]:
        os.environ.pop("LORA_DEVICES", None)
        one = np.zeros_like(a); ops.BY_SHAPE[shape](a, one, p, times, *dims)
        for rep in range(3):
            os.environ["LORA_DEVICES"] = ",".join(["0"] * k)
            many = np.zeros_like(a); ops.BY_SHAPE[shape](a, many, p, times, *dims)
            if np.array_equal(one, many):
                print(f"{shape} {dims} x{times} k={k} rep{rep}: identical", flush=True)
            else:
                bad = np.argwhere(one != many)
                rows = sorted(set(bad[:, 0].tolist()))
                cols = sorted(set(bad[:, 1].tolist())) if a.ndim > 1 else []
                print(f"{shape} {dims} x{times} k={k} rep{rep}: {len(bad)} cells differ; rows {rows[:12]}..{rows[-12:]} ({len(rows)} rows); cols {cols[:6]}..{cols[-6:]} ({len(cols)})", flush=True)

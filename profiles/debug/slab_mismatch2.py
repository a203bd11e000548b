"""Debug helper: dump the neighbourhood of a slab-vs-single mismatch for offline analysis."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import oracle
from lorastencil_b200 import ops
ops.set_verbose(False)
shape = sys.argv[1]; dims = tuple(int(x) for x in sys.argv[2].split(","))
times = int(sys.argv[3]); k = int(sys.argv[4]); tries = int(sys.argv[5])
rng = np.random.default_rng(1)
a = rng.integers(0, 100, size=oracle.padded_shape(shape, dims)).astype(np.float64)
p = oracle.reference_params(shape)
os.environ.pop("LORA_DEVICES", None)
one = np.zeros_like(a); ops.BY_SHAPE[shape](a, one, p, times, *dims)
found = 0
for rep in range(tries):
    os.environ["LORA_DEVICES"] = ",".join(["0"] * k)
    many = np.zeros_like(a); ops.BY_SHAPE[shape](a, many, p, times, *dims)
    if np.array_equal(one, many):
        continue
    bad = np.argwhere(one != many)
    r0, r1 = bad[:, 0].min(), bad[:, 0].max(); c0, c1 = bad[:, 1].min(), bad[:, 1].max()
    print(f"rep {rep}: rows {r0}..{r1} cols {c0}..{c1} count {len(bad)}", flush=True)
    R0, R1 = max(r0 - 40, 0), r1 + 41; C0, C1 = max(c0 - 24, 0), c1 + 25
    np.savez_compressed(f"gpurun_out/r2_mismatch_{found}.npz", a=a[R0:R1, C0:C1], one=one[R0:R1, C0:C1], many=many[R0:R1, C0:C1],
                        origin=np.array([R0, C0]), box=np.array([r0, r1, c0, c1]), times=times, k=k)
    found += 1
    if found >= 4:
        break
print("events", found, "of", rep + 1)

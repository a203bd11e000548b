"""Regenerates tests/golden/single_step.json from the REFERENCE's own code.

Run in the authoring container (needs /root/reference to build oracle/_ref):
    python tests/golden/make_golden.py
For every case it feeds the reference's default input (unseeded glibc rand() fill, restated in
oracle.fill_rand) and the reference CLI's weight table to the reference's verbatim `test_cpu`
(oracle/_ref/libref_cpu_*.so, compiled from /root/reference/src/*/main.cu) and records probe values:
first / centre / last interior cell, the interior sum and a sha256 of the interior bytes.  The same
numbers appear in SURVEY.md section 8(c).
"""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle  # noqa: E402

CASES = [("1d1r", (1024,)), ("1d2r", (1024,)), ("box2d1r", (64, 64)), ("box2d3r", (64, 64)), ("star2d1r", (64, 64)),
         ("star2d3r", (64, 64)), ("box2d1r", (1024, 1024)), ("box3d1r", (16, 16, 64)), ("star3d1r", (16, 16, 64)),
         # ragged sizes the reference kernels cannot run (S4) but its test_cpu can
         ("1d2r", (1000,)), ("box2d3r", (50, 70)), ("star2d3r", (33, 130)), ("star2d1r", (40, 36)),
         ("box3d1r", (5, 9, 30)), ("star3d1r", (7, 33, 132))]


def main():
    assert oracle.build_ref(), "needs /root/reference to build oracle/_ref"
    out = []
    for shape, dims in CASES:
        d = oracle.dim_of(shape)
        a = oracle.fill_rand(shape, dims)
        p = oracle.reference_params(shape)
        r = oracle.ref_cpu_step(d, a, p)
        halo = oracle.HALO[d]
        sl = tuple(slice(h, h + x) for h, x in zip(halo, dims))
        interior = np.ascontiguousarray(r[sl])
        out.append({
            "shape": shape, "dims": list(dims), "in_first4": a.ravel()[:4].tolist(),
            "first": float(r[tuple(halo)]),
            "centre": float(r[tuple(h + x // 2 for h, x in zip(halo, dims))]),
            "last": float(r[tuple(h + x - 1 for h, x in zip(halo, dims))]),
            "interior_sum": float(interior.sum()),
            "interior_sha256": hashlib.sha256(interior.tobytes()).hexdigest(),
        })
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "single_step.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path, len(out), "cases")


if __name__ == "__main__":
    main()

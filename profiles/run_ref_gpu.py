"""The reference's own GPU operators, recompiled unmodified for sm_100a (oracle/_ref/libref_gpu_*.so = the "A100
artifact recompile"), timed on the same B200 beside this library's operators -- same padded host arrays, same launch
counts, each side's own timed region (the launch loop, as both print it in their banner).

    python profiles/run_ref_gpu.py [--launches 21] > profiles/r2_ref_gpu_recompile.json
Both sides' outputs are compared as well (bit-identical while the integers stay exact, max relative error otherwise).
Not part of bench.py: the reference binaries are test infrastructure (oracle/), this is a side-by-side report."""
import argparse
import json
import os
import re
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import oracle  # noqa: E402
from lorastencil_b200 import ops  # noqa: E402

SIZES = {"1d1r": (1 << 28,), "1d2r": (1 << 28,), "star2d1r": (10240, 10240), "box2d1r": (10240, 10240),
         "star2d3r": (10240, 10240), "box2d3r": (10240, 10240), "box3d1r": (512, 512, 512), "star3d1r": (512, 512, 512)}
K = oracle.ARTIFACT_K


def captured_stdout(fn):
    """Run fn() with file descriptor 1 redirected to a temp file (the operators printf their banner)."""
    sys.stdout.flush()
    saved = os.dup(1)
    with tempfile.TemporaryFile(mode="w+b") as tmp:
        os.dup2(tmp.fileno(), 1)
        try:
            fn()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
        tmp.seek(0)
        return tmp.read().decode(errors="replace")


def banner_gstencils(text, shape):
    m = re.findall(r"GStencil/s = ([0-9.eE+-]+|inf|nan)", text)
    return float(m[-1]) / K[shape] if m else None


ap = argparse.ArgumentParser()
ap.add_argument("--launches", type=int, default=21)
ap.add_argument("--shapes", default=",".join(SIZES))
args = ap.parse_args()
ops.set_verbose(True)
rows = []
for shape in args.shapes.split(","):
    dims = SIZES[shape]
    rng = np.random.default_rng(1)
    a = rng.integers(0, 10, size=oracle.padded_shape(shape, dims)).astype(np.float64)
    p = oracle.reference_params(shape)
    out = np.zeros_like(a)
    ref_txt, ref_out = "", [None]

    def run_ref():
        ref_out[0] = oracle.ref_gpu_run(shape, a, p, args.launches)
    for _ in range(2):  # second call = warm
        ref_txt = captured_stdout(run_ref)
    our_txt = ""
    for _ in range(2):
        our_txt = captured_stdout(lambda: ops.BY_SHAPE[shape](a, out, p, args.launches, *dims))
    # the two sides ran on the same input: compare what they returned (full padded output; 1-D leaves the last double)
    got, want = (out[:-1], ref_out[0][:-1]) if len(dims) == 1 else (out, ref_out[0])
    finite = np.isfinite(want)
    scale = float(np.abs(want[finite]).max()) if finite.any() else 0.0
    err = float(np.abs(got[finite] - want[finite]).max() / scale) if scale > 0 else 0.0
    r = {"shape": shape, "dims": list(dims), "launches": args.launches,
         "reference_sm100a_recompile_gstencils": banner_gstencils(ref_txt, shape),
         "this_library_gstencils": banner_gstencils(our_txt, shape),
         "outputs_bit_identical": bool(np.array_equal(got, want)), "max_rel_err_over_finite_cells": err,
         "same_non_finite_cells": bool(np.array_equal(np.isfinite(got), finite))}
    if r["reference_sm100a_recompile_gstencils"] and r["this_library_gstencils"]:
        r["speedup"] = r["this_library_gstencils"] / r["reference_sm100a_recompile_gstencils"]
    rows.append(r)
    print(json.dumps(r), flush=True, file=sys.stderr)
print(json.dumps({"what": "GStencil/s (cells x launches / s / 1e9, K = 1) of each side's launch loop, from its own banner",
                  "rows": rows}))

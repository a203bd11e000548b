"""One BASELINE.json-sized parity case in its own process (run by tests/test_parity_baseline_sizes.py and by
profiles/run_ref_gpu.py):

    python tests/baseline_size_check.py SHAPE DIMS TIMES[,TIMES...] [--band ROWS] [--devices 0,0]

For every launch count: the reference GPU operator (recompiled unmodified for sm_100a: oracle/_ref/libref_gpu_*.so)
and our drop-in operator run on the SAME padded host array (integer-valued, like the reference's fill); the two full
padded outputs are compared -- bitwise while every intermediate is an exact integer below 2^53, max relative error
<= 1e-12 afterwards.  `--band R`: additionally R interior rows / planes in the middle of the grid are recomputed by
the CPU oracle (one launch) from a cut-out of the input.  `--devices`: our side runs as slabs on these devices
(LORA_DEVICES; a repeated device = several slabs sharing one GPU) -- the multi-GPU path at the same size.
A separate process per case because the reference operators never free their device buffers
(src/2d/gpu.cu:392-421 has no cudaFree): at 40960^2 every reference call leaks 27 GB.
Prints one JSON line; exit code 0 = every comparison passed."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import oracle  # noqa: E402
from lorastencil_b200 import ops  # noqa: E402

EXACT_UPTO = {"1d1r": 8, "1d2r": 8, "box2d1r": 5, "box2d3r": 5, "star2d1r": 6, "star2d3r": 9, "box3d1r": 8, "star3d1r": 15}
RTOL = 1e-12


def main():
    import torch
    shape = sys.argv[1]
    dims = tuple(int(x) for x in sys.argv[2].split(","))
    times_list = [int(x) for x in sys.argv[3].split(",")]
    band = int(sys.argv[sys.argv.index("--band") + 1]) if "--band" in sys.argv else 0
    if "--devices" in sys.argv:
        os.environ["LORA_DEVICES"] = sys.argv[sys.argv.index("--devices") + 1]
    use_ref = "--no-ref" not in sys.argv and oracle.ref_available("gpu", oracle.dim_of(shape))
    ops.set_verbose(False)
    d = oracle.dim_of(shape)
    padded = oracle.padded_shape(shape, dims)
    # integer-valued input like the reference's fill (rand() % 10000 in 1-D, % 100 otherwise), generated on the GPU
    g = torch.Generator(device="cuda").manual_seed(20260)
    a = torch.randint(0, 10000 if d == 1 else 100, padded, generator=g, device="cuda").double().cpu().numpy()
    torch.cuda.empty_cache()
    p = oracle.reference_params(shape)
    eff = oracle.effective_params(shape, p)
    report = {"shape": shape, "dims": list(dims), "cases": [], "ok": True, "reference_gpu": use_ref,
              "devices": os.environ.get("LORA_DEVICES")}
    devnull = os.open(os.devnull, os.O_WRONLY)
    for times in times_list:
        t0 = time.time()
        out = np.full_like(a, -7.0)
        ops.BY_SHAPE[shape](a, out, p, times, *dims)
        case = {"times": times, "gpus": ops.last_gpus()}
        if d == 1:
            assert out[-1] == -7.0  # S3: 1-D copies back cols-1 doubles
        if use_ref:
            saved = os.dup(1)  # the reference prints its banner
            os.dup2(devnull, 1)
            try:
                ref = oracle.ref_gpu_run(shape, a, p, times)
            finally:
                os.dup2(saved, 1)
                os.close(saved)
            got, want = (out[:-1], ref[:-1]) if d == 1 else (out, ref)
            if times <= EXACT_UPTO[shape]:
                case["vs_reference_gpu"] = "bit-identical" if np.array_equal(got, want) else "MISMATCH"
                ok = case["vs_reference_gpu"] == "bit-identical"
            else:
                scale = float(np.abs(want).max())
                err = float(np.abs(got - want).max() / scale) if np.isfinite(scale) and scale > 0 else float("nan")
                case["vs_reference_gpu"] = f"max rel err {err:.3g}"
                ok = err <= RTOL
            report["ok"] &= bool(ok)
            del ref
        if band and times == 1:
            # CPU oracle on a cut-out: `band` interior indices of the outermost axis around the middle
            h0 = oracle.HALO[d][0]
            lo = dims[0] // 2
            cut = np.ascontiguousarray(a[lo:lo + band + 2 * h0])
            want = oracle.step(d, cut, eff)[h0:h0 + band]
            got = out[lo + h0:lo + h0 + band]
            inner = tuple(slice(h, -h) for h in oracle.HALO[d][1:])
            same = np.array_equal(got[(slice(None),) + inner], want[(slice(None),) + inner])
            case["vs_cpu_oracle_band"] = "bit-identical" if same else "MISMATCH"
            report["ok"] &= bool(same)
        case["seconds"] = round(time.time() - t0, 2)
        report["cases"].append(case)
        del out
    print(json.dumps(report))
    return 0 if report["ok"] else 1


if __name__ == "__main__":
    sys.exit(main())

"""End-to-end timing of the 1-D drop-in operator (pinned host in -> GPU -> pinned host out) per call, for a few
chunk counts.   python profiles/run_e2e.py [--n 268435456] [--times 1000] [--chunks 0,4,8,16]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lorastencil_b200 as ls
from lorastencil_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 28)
ap.add_argument("--times", type=int, default=1000)
ap.add_argument("--chunks", default="0,4,8,16")
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
ops.set_verbose(False)
n = args.n
hin = torch.randint(0, 10000, (n + 8,)).double().pin_memory()
hout = torch.empty(n + 8, dtype=torch.float64).pin_memory()
params = ls.reference_table("1d2r")
for ch in args.chunks.split(","):
    if ch == "auto":
        os.environ.pop("LORA_CHUNKS", None)
    else:
        os.environ["LORA_CHUNKS"] = ch
    calls = []
    for _ in range(args.reps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ops.gpu_1d2r(hin, hout, params, args.times, n)
        calls.append((time.perf_counter() - t0) * 1e3)
    best = min(calls[1:])
    print(json.dumps({"chunks": ch, "used": ops.last_chunks(), "calls_ms": [round(c, 1) for c in calls],
                      "loop_ms": round(ops.last_loop_ms(), 1), "gstencils_e2e": n * args.times / best / 1e6}), flush=True)

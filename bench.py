#!/usr/bin/env python
"""bench.py -- GStencil/s of the LoRAStencil hot path on N B200s (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W                 # our CUDA path
    python bench.py --impl reference --gpus 1 --steps K --warmup W # the reference's CPU stencil (test_cpu)
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # slab-decomposed, one rank per GPU

Workload (config.workload): BASELINE.json configs[1] -- `lorastencil_1d 1d2r 268435456 1000`: one STEP is
the whole job, 1000 launches of the 9-tap 1-D operator over 2^28 points (per GPU: weak scaling, the
global line is N x 2^28 points cut into slabs with a 4-element halo exchange per launch).

Numbers on the JSON line:
  value      cells x launches / second / 1e9 with the grid resident in HBM (CUDA events on the launch stream,
             max over ranks).  K = 1 convention; the artifact's own printout multiplies by 2 for 1d2r
             (src/1d/gpu_2r.cu:134) -- that figure is `value_artifact_units`.
  e2e        the same job through the reference-facing operator `gpu_1d2r(in, out, params, times, n)` with
             pinned HOST buffers: H2D of the padded line, 1000 launches, D2H, all inside the timed region.
  roofline   dominant kernel (k_stencil1d): 16 B per cell per launch (one FP64 read + one write,
             SURVEY.md section 8d) / its average launch duration, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the reference's own test_cpu (oracle/_ref, compiled from its main.cu) on the host cores,
             timed on a bounded sample (a few launches over the same 2^28-point line).
  shapes     all eight shapes at their BASELINE sizes on 1 GPU: device-resident GStencil/s (best and median of 5
             repetitions, clocks sampled per shape), e2e through the drop-in operator with pinned and with pageable host
             buffers, and the reference's test_cpu on 1 core and on all host cores; `worst_shape_frac` = the smallest
             fixed-16-B roofline fraction among them.
  parity     (every N) small grids of all shapes through the same slab driver the timed run uses, gathered and compared
             with the CPU oracle on rank 0 -- multi-GPU correctness travels with the scaling record.
  extensions (N = 1, with `shapes`) the section-8(f)-4 additions, which are not reference shapes and therefore not part
             of `shapes` / `worst_shape_frac`: the radius-2 3-D shapes star3d2r / box3d2r at 512^3 (device-resident, best
             and median of 3) and the periodic boundary mode against the reference halo at one launch per step, each with
             a small-grid comparison against the CPU oracle.  Failure-proof: an exception becomes {"error": ...}.
  scaling_extra  (every N) BASELINE.json configs[4]: box2d1r 40960^2 and box3d1r 1024^3 strong scaling (global grid
             fixed, N slabs), and per-GPU-constant 2-D / 3-D slabs (weak scaling), with the bytes exchanged per step.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

HALO = {1: (4,), 2: (4, 4), 3: (1, 2, 4)}
ARTIFACT_K = {"1d1r": 3, "1d2r": 2, "star2d1r": 3, "box2d1r": 3, "star2d3r": 1, "box2d3r": 3, "box3d1r": 1, "star3d1r": 1}
HEADLINE = ("1d2r", (1 << 28,), 1000)
# BASELINE.json configs (b)-(d) on one GPU; launches per measurement kept short, the per-launch cost is constant
SHAPE_TABLE = [("1d1r", (1 << 28,), 50), ("1d2r", (1 << 28,), 50), ("star2d1r", (10240, 10240), 100),
               ("box2d1r", (10240, 10240), 100), ("star2d3r", (10240, 10240), 100), ("box2d3r", (10240, 10240), 100),
               ("box3d1r", (512, 512, 512), 100), ("star3d1r", (512, 512, 512), 100)]


def headline_config():
    """config of the JSON line -- the SAME dict in both arms (ours and --impl reference), so the driver can match them."""
    shape, dims, times = HEADLINE
    return {"workload": f"lorastencil_1d {shape} {dims[0]} {times} (BASELINE.json configs[1])", "shape": shape,
            "points_per_gpu": dims[0], "launches_per_job": times,
            "step": "our arm: one step = the whole job (all launches); reference arm: one step = a bounded sample of "
                    "it (1 launch over the same line); both values are cells x launches / s",
            "l2": "inputs (2.1 GB per buffer) larger than L2; no flush needed"}


def measured_traffic(kernel_key):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/r2_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum), or None."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    try:
        return json.load(open(path)).get(kernel_key)
    except (OSError, ValueError):
        return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled from a thread every 20 ms (nvidia-smi -lms
    as a fallback -- it takes most of a second to start, too slow for a half-second region)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.sm, self.mx, self.reasons, self.thread, self.stop = [], [], set(), None, False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            # NVML indexes physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain list of indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [int(x) for x in vis.split(",")] if vis and all(x.strip().isdigit() for x in vis.split(",")) else None
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(ids[index] if ids and index < len(ids) else index)
        except Exception:  # noqa: BLE001
            self.nvml = None

    def _poll(self):
        n = self.nvml
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self.stop:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)))
                r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.thread is not None:
            self.stop = True
            self.thread.join(timeout=2)
            return
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
                out = ""
            self.lines = [l for l in out.splitlines() if l.strip()]

    def summary(self):
        if self.thread is not None:
            if not self.sm:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "nvml"}
            return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": max(self.mx), "reasons": sorted(self.reasons),
                    "samples": len(self.sm), "source": "nvml, 20 ms polling during the timed region"}
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[2:6]):
                if v == "Active":
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "nvidia-smi"}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvidia-smi -lms 100"}


# ------------------------------------------------------------------------------------------------
# reference CPU arm / cpu_baseline: the reference's verbatim test_cpu on all host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_runner(shape, dims, cores=None):
    """Returns (run_once, cores, kind): run_once() does ONE launch over the whole grid with the reference's
    test_cpu (oracle/_ref) split over `cores` host threads along the outermost axis; falls back to the
    OpenMP oracle port when oracle/_ref is absent."""
    import oracle
    d = oracle.dim_of(shape)
    padded = oracle.padded_shape(shape, dims)
    rng = np.random.default_rng(0)
    mod = 10000 if d == 1 else 100
    a = rng.integers(0, mod, size=padded).astype(np.float64)
    out = np.zeros_like(a)
    params = np.ascontiguousarray(oracle.reference_params(shape))
    if cores is None:
        try:
            cores = len(os.sched_getaffinity(0))
        except AttributeError:
            cores = os.cpu_count() or 1
    h0 = oracle.HALO[d][0]
    rest = int(np.prod(padded[1:])) if d > 1 else 1
    if oracle.ref_available("cpu", d):
        fn = oracle.ref_cpu_fn(d)
        n0 = dims[0]
        cuts = [n0 * i // cores for i in range(cores + 1)]

        def piece(i):
            lo, hi = cuts[i], cuts[i + 1]
            if hi <= lo:
                return
            off = lo * rest * 8
            fn(a.ctypes.data + off, out.ctypes.data + off, params.ctypes.data, hi - lo + 2 * h0,
               *[int(x) for x in padded[1:]])

        def run_once():
            ts = [threading.Thread(target=piece, args=(i,)) for i in range(cores)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
        return run_once, cores, "reference"
    L = oracle.lib()
    cores = L.oracle_max_threads()

    def run_once():
        oracle.step(shape, a, params)
    return run_once, cores, "port"


def time_cpu(shape, dims, min_seconds=15.0, max_launches=60, cores=None):
    run_once, cores, kind = cpu_reference_runner(shape, dims, cores)
    run_once()  # warm (page-faults the output)
    n, t0 = 0, time.perf_counter()
    while True:
        run_once()
        n += 1
        el = time.perf_counter() - t0
        if el * cores >= min_seconds or n >= max_launches:
            break
    cells = float(np.prod(dims))
    return {"value": cells * n / el / 1e9, "unit": "GStencil/s", "cores": cores, "kind": kind,
            "sample": f"{n} launch(es) of {shape} over the full {'x'.join(map(str, dims))} grid "
                      f"(of the job's launches), {el:.2f} s wall"}


def cpu_baselines_in_a_clean_process(jobs):
    """time_cpu for a list of (shape, dims, kwargs) in a child process that never touches CUDA or torch: inside the
    GPU process (CUDA context, pinned allocations, staging / NVML threads) the same loop measured 2x slower."""
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-jobs", json.dumps(jobs)], capture_output=True, text=True)
    lines = [l for l in r.stdout.splitlines() if l.startswith("[")]
    if r.returncode != 0 or not lines:
        raise RuntimeError("cpu baseline child failed: " + r.stdout[-500:] + r.stderr[-500:])
    return json.loads(lines[-1])


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    shape, dims, times = HEADLINE
    run_once, cores, kind = cpu_reference_runner(shape, dims)
    for _ in range(args.warmup):
        run_once()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run_once()
    el = time.perf_counter() - t0
    cells = float(np.prod(dims))
    v = cells * args.steps / el / 1e9
    sample = f"each step = 1 launch of {shape} over the full {dims[0]}-point line (the job is {times} such launches)"
    print(json.dumps({
        "impl": "reference", "metric": "GStencil/s", "value": v, "unit": "GStencil/s (cells x launches / s / 1e9)",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": headline_config(),
        "cpu_baseline": {"value": v, "unit": "GStencil/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "GStencil/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def device_fill(torch, shape_padded, mod, device, seed):
    g = torch.Generator(device=device).manual_seed(seed)
    return torch.randint(0, mod, shape_padded, generator=g, device=device).double()


def measure_shape(torch, ls, ops, shape, dims, launches, hbm_gbs, device_index, reps=5, e2e=True, cpu=False):
    """One row of the per-shape table: device-resident GStencil/s (CUDA events, best and median of `reps`, NVML clocks
    sampled during the repetitions), e2e through the reference-facing operator (pinned and pageable host buffers),
    the reference's test_cpu on 1 core and on all host cores."""
    plan = ls.Plan(shape, dims)
    d = len(dims)
    b0 = device_fill(torch, plan.padded_shape, 10000 if d == 1 else 100, "cuda", 1)
    b1 = plan.new_buffer()
    plan.run(b0, b1, 3)
    torch.cuda.synchronize()
    ms_all = []
    with ClockSampler(device_index) as clk:
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.run(b0, b1, launches)
            e1.record()
            torch.cuda.synchronize()
            ms_all.append(e0.elapsed_time(e1))
    cells = float(np.prod(dims))
    best, med = min(ms_all), statistics.median(ms_all)
    per_launch_s = best / 1e3 / launches
    gst = cells / per_launch_s / 1e9
    ach = cells * 16 / per_launch_s / 1e9
    row = {"shape": shape, "dims": list(dims), "launches": launches, "gstencils": gst,
           "gstencils_median": cells * launches / (med / 1e3) / 1e9, "reps": reps,
           "gstencils_artifact_units": gst * ARTIFACT_K[shape], "us_per_launch": per_launch_s * 1e6,
           "hbm_gbs_algorithmic": ach, "roofline_frac": ach / hbm_gbs, "roofline_frac_median": ach / hbm_gbs * best / med,
           "form": plan.describe, "temporal_block": plan.temporal_block, "clocks": clk.summary()}
    del plan, b0, b1
    torch.cuda.empty_cache()
    if e2e:
        # the drop-in operator gpu_X(in, out, params, times, dims...) with HOST buffers: H2D, all launches, D2H inside
        # the timed region.  Pinned buffers (what an application that cares would pass) and pageable ones (what the
        # reference's main.cu mallocs, src/2d/main.cu:224-225).
        padded = tuple(x + 2 * h for x, h in zip(dims, HALO[d]))
        params = ls.reference_table(shape)
        times = 1000 if d == 1 else 100
        nel = int(np.prod(padded))
        src = torch.randint(0, 10000 if d == 1 else 100, padded).double()
        for kind in ("pinned", "pageable"):
            if kind == "pinned":
                hin, hout = src.pin_memory(), torch.empty(padded, dtype=torch.float64).pin_memory()
            else:
                hin, hout = src.numpy(), np.empty(padded, dtype=np.float64)
                hout.fill(0.0)  # fault the pages in, as a caller that reuses its arrays would have
            for _ in range(2):  # warm: device workspace, page faults, pinned staging slots
                ops.BY_SHAPE[shape](hin, hout, params, times, *dims)
            ts = []
            for _ in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                ops.BY_SHAPE[shape](hin, hout, params, times, *dims)
                ts.append(time.perf_counter() - t0)
            el = min(ts)
            row["e2e" if kind == "pinned" else "e2e_pageable"] = {
                "value": cells * times / el / 1e9, "unit": "GStencil/s", "launches": times, "ms_per_call": el * 1e3,
                "launch_loop_ms": ops.last_loop_ms(), "bands": ops.last_bands(), "chunks": ops.last_chunks(), "h2d_bytes_per_step": nel * 8, "d2h_bytes_per_step": nel * 8 - (8 if d == 1 else 0),
                "host_buffers": kind, "api": f"lorastencil_b200.ops.{ops.BY_SHAPE[shape].__name__} -> lora_{ops.BY_SHAPE[shape].__name__} (C ABI)"}
            del hin, hout
        del src
        ops.release_workspace()
    if cpu:
        one = time_cpu(shape, dims, min_seconds=1.0, max_launches=1, cores=1)
        allc = time_cpu(shape, dims, min_seconds=4.0, max_launches=8)
        row["cpu_baseline"] = {"one_core": one, "all_cores": allc}
    return row


def measure_extensions(torch, ls, hbm_gbs):
    """Rows for the additions beyond the reference's shape list (DESIGN.md sections 2.5 and 4)."""
    import oracle
    rows = []

    def timed(plan, launches=6, reps=3):
        b0 = device_fill(torch, plan.padded_shape, 100, "cuda", 1)
        b1 = plan.new_buffer()
        plan.run(b0, b1, 2)
        torch.cuda.synchronize()
        ms = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.run(b0, b1, launches)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        cells = float(np.prod(plan.dims))
        gst = cells * launches / (min(ms) / 1e3) / 1e9
        return {"launches": launches, "reps": reps, "gstencils": gst,
                "gstencils_median": cells * launches / (statistics.median(ms) / 1e3) / 1e9,
                "roofline_frac": gst * 16 / hbm_gbs, "form": plan.describe, "temporal_block": plan.temporal_block}

    rng = np.random.default_rng(2)
    for shape in ("star3d2r", "box3d2r"):
        small = (9, 12, 70)
        a = rng.uniform(-1, 1, oracle.padded_shape_r2(small))
        w = oracle.reference_params_r2(shape)
        ps = ls.Plan(shape, small)
        got = ps.run(torch.from_numpy(a).cuda(), ps.new_buffer(), 3).cpu().numpy()
        ref = oracle.run_r2(a, w, 3)
        err = float(np.abs(got - ref).max() / np.abs(ref).max())
        row = {"what": f"{shape} (radius 2, 125 weights; not a reference shape)", "dims": [512, 512, 512],
               "parity": {"dims": list(small), "launches": 3, "max_rel_err": err, "ok": err <= 1e-12}}
        row.update(timed(ls.Plan(shape, (512, 512, 512))))
        rows.append(row)
        torch.cuda.empty_cache()
    for shape, dims, small in (("star2d3r", (10240, 10240), (40, 66)), ("box3d1r", (512, 512, 512), (6, 10, 64))):
        a = rng.uniform(-1, 1, oracle.padded_shape(shape, small))
        ps = ls.Plan(shape, small)
        ps.boundary = "periodic"
        got = ps.run(torch.from_numpy(a).cuda(), ps.new_buffer(), 3).cpu().numpy()
        ref = oracle.run_periodic(shape, a, oracle.effective_params(shape), 3)
        err = float(np.abs(got - ref).max() / np.abs(ref).max())
        row = {"what": f"{shape}, periodic boundary (one launch per step)", "dims": list(dims),
               "parity": {"dims": list(small), "launches": 3, "max_rel_err": err, "ok": err <= 1e-12}}
        for boundary in ("periodic", "reference"):
            plan = ls.Plan(shape, dims)
            plan.temporal_block = 1  # like with like: the periodic mode runs one launch per step
            plan.boundary = boundary
            t = timed(plan)
            if boundary == "periodic":
                row.update(t)
            else:
                row["gstencils_reference_halo_unfused"] = t["gstencils"]
            del plan
            torch.cuda.empty_cache()
        rows.append(row)
    return rows


def exchange_bytes_per_step(shape, dims, tb):
    """Bytes one slab sends to ONE neighbour per time step (its ghost zone, once per sweep of tb launches)."""
    d = len(dims)
    if d == 1:
        return 4 * tb * 8 / tb
    if d == 2:
        rows = 3 * tb if tb > 1 else 4  # fused: radius x tb rows per sweep; unfused: the 4-row storage halo per launch
        return rows * (dims[1] + 8) * 8 / tb
    return (dims[1] + 4) * (dims[2] + 8) * 8


def parity_probe(torch, dist, SlabRunner, dev, world, rank):
    """Small grids of every shape through the slab driver (the timed path), gathered on rank 0 and compared with the
    CPU oracle: bit-identical while the integers stay exact, <= 1e-12 relative afterwards."""
    cases = [("1d2r", (1 << 20,), 20), ("1d1r", (300000,), 7), ("box2d1r", (512, 640), 5), ("box2d3r", (384, 258), 4),
             ("star2d3r", (600, 516), 7), ("star2d1r", (512, 256), 6), ("box3d1r", (64, 64, 128), 5),
             ("star3d1r", (48, 40, 136), 4)]
    exact_upto = {"1d1r": 8, "1d2r": 8, "box2d1r": 5, "box2d3r": 5, "star2d1r": 6, "star2d3r": 9, "box3d1r": 8, "star3d1r": 15}
    out = []
    ok_all = True
    for shape, dims, times in cases:
        d = len(dims)
        padded = tuple(x + 2 * h for x, h in zip(dims, HALO[d]))
        a = np.random.default_rng(77).integers(0, 100, size=padded).astype(np.float64)
        r = SlabRunner(shape, dims, device=dev)
        r.load_global(a)
        r.run(times)
        torch.cuda.synchronize()
        got = r.gather_global(a.shape)
        mode = r.halo_mode
        r.close()
        if rank == 0:
            import oracle
            ref = oracle.run(shape, a, oracle.effective_params(shape), times)
            if d == 1:
                got, ref = got[:-1], ref[:-1]
            if times <= exact_upto[shape]:
                ok = bool(np.array_equal(got, ref))
                how = "bit-identical" if ok else "MISMATCH"
            else:
                err = float(np.abs(got - ref).max() / np.abs(ref).max())
                ok = err <= 1e-12
                how = f"max rel err {err:.2g}"
            ok_all &= ok
            out.append({"shape": shape, "dims": list(dims), "launches": times, "result": how})
    return {"ok": bool(ok_all), "against": "CPU oracle (oracle/oracle.c, pinned to the reference's test_cpu)", "n_gpus": world,
            "halo": mode, "cases": out}


def scaling_extra(torch, dist, SlabRunner, dev, world, rank):
    """BASELINE.json configs[4] (strong scaling of the two named grids) and per-GPU-constant 2-D / 3-D slabs (weak)."""
    table = [("box2d1r", (40960, 40960), 30, "strong"), ("box3d1r", (1024, 1024, 1024), 30, "strong"),
             ("box2d1r", (10240, 10240), 60, "weak"), ("star2d3r", (10240, 10240), 60, "weak"),
             ("box3d1r", (512, 512, 512), 60, "weak"), ("star3d1r", (512, 512, 512), 60, "weak")]
    rows = []
    for shape, dims, launches, mode in table:
        gdims = tuple(dims) if mode == "strong" else (dims[0] * world,) + tuple(dims[1:])
        r = SlabRunner(shape, gdims, device=dev)
        r.buf[0].copy_(device_fill(torch, r.geo.local_padded, 100, dev, 99 + rank))
        r.sync_ranks()
        r.run(4 if r.max_tb == 1 else 6)  # even number of sweeps: the timed run starts from a parity-consistent state
        r.sync_ranks()
        best = None
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r.sync_ranks()
            e0.record()
            r.run(launches)
            e1.record()
            r.sync_ranks()
            ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            best = ms if best is None else min(best, ms)
        cells = float(np.prod(gdims))
        tb = r.max_tb
        rows.append({"shape": shape, "mode": mode, "global_dims": list(gdims), "launches": launches,
                     "value": cells * launches / (best / 1e3) / 1e9, "unit": "GStencil/s", "ms_per_step": best / launches,
                     "temporal_block": tb, "halo": r.halo_mode,
                     "exchange_bytes_per_step": 0 if world == 1 else exchange_bytes_per_step(shape, gdims, tb) * 2 * (world - 1)})
        r.close()
        del r
        torch.cuda.empty_cache()
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--times", type=int, default=HEADLINE[2], help="launches per step (default: the job's 1000)")
    ap.add_argument("--no-shapes", action="store_true", help="skip the per-shape table")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the slab-vs-oracle parity probe")
    ap.add_argument("--no-extra", action="store_true", help="skip scaling_extra (config e strong / 2-D, 3-D weak scaling)")
    # other BASELINE.json configs (parity-test cases by default, measurable on request): e.g. config e
    #   --shape box2d1r --dims 40960,40960 --times 100 --scaling strong     (global grid cut into N slabs)
    ap.add_argument("--shape", default=None, help="measure this shape instead of the headline 1d2r job")
    ap.add_argument("--dims", default=None, help="comma-separated interior sizes (per GPU if weak, global if strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--cpu-jobs", default=None, help=argparse.SUPPRESS)  # internal: cpu_baselines_in_a_clean_process
    args = ap.parse_args()
    if args.cpu_jobs:
        print(json.dumps([time_cpu(s_, tuple(d_), **kw) for s_, d_, kw in json.loads(args.cpu_jobs)]))
        return
    if args.impl == "reference":
        return run_reference_arm(args)
    # libraries (NCCL's version banner, ...) write to fd 1; the contract is ONE JSON line on stdout, so everything
    # else goes to stderr and the JSON line is written to the saved descriptor
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    import lorastencil_b200 as ls
    from lorastencil_b200 import ops
    from lorastencil_b200.slab import SlabRunner

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}"
    ops.set_verbose(False)
    hbm_gbs, peak_src = peaks()
    shape, dims, _ = HEADLINE
    custom = args.shape is not None
    if custom:
        shape = args.shape
        dims = tuple(int(x) for x in args.dims.split(","))
        args.no_shapes = args.no_cpu = True
        args.no_e2e = args.no_parity = args.no_extra = True
    times = args.times
    n = dims[0]
    if args.scaling == "weak":
        global_dims = (n * world,) + tuple(dims[1:])
    else:
        global_dims = tuple(dims)
    cells_per_gpu = float(np.prod(global_dims)) / world

    # ---- device-resident arm: slab runner (world == 1: a plain plan) ----
    runner = SlabRunner(shape, global_dims, device=dev)
    runner.buf[0].copy_(device_fill(torch, runner.geo.local_padded, 10000 if len(dims) == 1 else 100, dev, 1234 + rank))
    runner.sync_ranks()
    plan = runner.plan

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        runner.run(times)

    for _ in range(args.warmup):
        one_step()
    barrier()
    launches0, sweeps0 = plan.launches, runner.launch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        e0.record()
        for _ in range(args.steps):
            one_step()
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    gpu_launches = plan.launches - launches0
    total_cells_launches = cells_per_gpu * world * times * args.steps
    value = total_cells_launches / (ms / 1e3) / 1e9
    # dominant kernel: the full-slab (world 1) or interior (world > 1) sweep; with temporal blocking one launch
    # advances `tb` time steps, so its algorithmic bytes are 16 B x cells x tb
    main_launches = gpu_launches if world == 1 else (runner.launch - sweeps0)
    steps_per_launch = times * args.steps / main_launches
    us_per_launch = ms * 1e3 / main_launches
    achieved = cells_per_gpu * 16 * steps_per_launch / (us_per_launch * 1e-6) / 1e9

    # ---- e2e: through the reference-facing operator with pinned host buffers ----
    e2e = None
    if not args.no_e2e:
        nloc = runner.geo.local_padded[0]  # n + 8 on one device; slab + halo / ghost cells on a rank of N
        hin = torch.empty(nloc, dtype=torch.float64).pin_memory()
        hout = torch.empty(nloc, dtype=torch.float64).pin_memory()
        hin.copy_(torch.randint(0, 10000, (nloc,)).double())
        params = ls.reference_table(shape)
        k_e2e = max(1, min(args.steps, 3))
        if world == 1:
            for _ in range(2):  # warm: allocates the operator's device workspace, faults the pinned pages in
                ops.gpu_1d2r(hin, hout, params, times, n)
            barrier()
            t0 = time.perf_counter()
            for _ in range(k_e2e):
                ops.gpu_1d2r(hin, hout, params, times, n)
            torch.cuda.synchronize()
            el = time.perf_counter() - t0
            e2e_detail = {"launch_loop_ms_last_call": ops.last_loop_ms(), "whole_call_ms_last_call": ops.last_total_ms(),
                          "chunks": ops.last_chunks(),
                          "overlap": "the operator cuts the line into ghost-margined chunks (4 x times cells) and overlaps "
                                     "chunk H2D / launches / D2H on separate streams; results bit-identical to the plain path"}
        else:
            # N GPUs, host-resident line: no exchange is needed at all -- a cell after `times` launches depends on
            # 4 x times cells either side, so every rank runs the reference-facing operator on its slab plus a margin of
            # that width (lorastencil_b200.slab.host_segment / run_host_segment), chunked and copy-overlapped as at N = 1
            from lorastencil_b200.slab import host_segment, run_host_segment
            lo_s, hi_s, gl_s, gr_s = host_segment(n * world, world, rank, times)
            nseg = (hi_s - lo_s) + gl_s + gr_s + 8
            del hin, hout
            hin = torch.empty(nseg, dtype=torch.float64).pin_memory()
            hout = torch.empty(nseg, dtype=torch.float64).pin_memory()
            hin.copy_(torch.randint(0, 10000, (nseg,)).double())
            nloc = nseg
            e2e_detail = {"margin_cells": max(gl_s, gr_s),
                          "how": "every rank: its slab + a margin of 4 x times cells through gpu_1d2r (no halo exchange "
                                 "needed for host-resident data: the margin covers the dependency cone)"}
            for _ in range(2):
                run_host_segment(shape, hin, hout, params, times)
            barrier()
            t0 = time.perf_counter()
            for _ in range(k_e2e):
                run_host_segment(shape, hin, hout, params, times)
            barrier()
            el = time.perf_counter() - t0
            t = torch.tensor([el], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el = float(t.item())
            # the limiter: all ranks moving their segments H2D and D2H at the same time, no launches at all -- what the
            # host side (PCIe root complexes, the NUMA node the pinned pages live on) can sustain for N GPUs at once
            dbuf = torch.empty(nseg, dtype=torch.float64, device=dev)
            dbuf2 = torch.empty(nseg, dtype=torch.float64, device=dev)
            s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
            barrier()
            t0 = time.perf_counter()
            with torch.cuda.stream(s_up):
                dbuf.copy_(hin, non_blocking=True)
            with torch.cuda.stream(s_dn):
                hout.copy_(dbuf2, non_blocking=True)
            barrier()
            tc = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
            copy_s = float(tc.item())
            del dbuf, dbuf2
            e2e_detail["host_copy_only_ms"] = copy_s * 1e3
            e2e_detail["host_copy_ceiling"] = cells_per_gpu * world * times / copy_s / 1e9
            e2e_detail["host_copy_note"] = ("all ranks copying their segment H2D and D2H concurrently with no launches: the "
                                            "e2e figure cannot exceed host_copy_ceiling on this box; the launch loop alone "
                                            "is the device-timed `value`")
            try:
                import pynvml
                pynvml.nvmlInit()
                hnd = pynvml.nvmlDeviceGetHandleByIndex(local)
                e2e_detail["gpu_numa_node"] = int(pynvml.nvmlDeviceGetNumaNodeId(hnd))
            except Exception:  # noqa: BLE001
                pass
        e2e = {"value": cells_per_gpu * world * times * k_e2e / el / 1e9, "unit": "GStencil/s",
               "h2d_bytes_per_step": nloc * 8 * world, "d2h_bytes_per_step": (nloc - 1) * 8 * world,
               "steps": k_e2e, "ms_per_step": el / k_e2e * 1e3,
               "api": "lorastencil_b200.ops.gpu_1d2r -> lora_gpu_1d2r (C ABI), pinned host buffers" if world == 1 else
                      "lorastencil_b200.slab.run_host_segment -> ops.gpu_1d2r -> lora_gpu_1d2r (C ABI) per rank, pinned host buffers", **e2e_detail}
        del hin, hout

    halo_mode_used, plan_describe, geo = runner.halo_mode, plan.describe, runner.geo
    runner_max_tb = runner.max_tb
    runner.close()
    del runner
    torch.cuda.empty_cache()
    parity = None if args.no_parity else parity_probe(torch, dist, SlabRunner, dev, world, rank)
    extra = None if args.no_extra else scaling_extra(torch, dist, SlabRunner, dev, world, rank)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    halo_how = {"p2p": "one kernel launch per sweep: its band tasks run first and store their cells a second time straight "
                       "into the neighbours' ghost zones over NVLink peer memory (CUDA IPC), the last band task raises a "
                       "64-bit flag there, the next sweep waits for the flags in stream order; no communication library "
                       "on the data path",
                "nccl": "NCCL send/recv of the edge bands on a side stream"}[halo_mode_used]
    dimname = {1: "1d", 2: "2d", 3: "3d"}[len(dims)]
    buf_gb = float(np.prod(geo.local_padded)) * 8 / 1e9
    if custom:
        config = {"workload": f"lorastencil_{dimname} {shape} {' '.join(map(str, dims))} {times} "
                              + ("per GPU" if args.scaling == "weak" else "global grid"), "shape": shape,
                  "points_per_gpu": int(cells_per_gpu), "launches_per_job": times}
    else:
        config = headline_config()
    detail = {"global_dims": list(global_dims), "scaling": args.scaling,
              "decomposition": "single device" if world == 1 else
              f"{world} slabs along the outermost axis, {geo.wl if geo.prev is not None else geo.wr}"
              f"-deep ghost zones exchanged once per sweep of {runner_max_tb} launch(es): " + halo_how,
              "buffer_gb": buf_gb,
              "values": "reference weights: FP64 overflows to inf after ~130-340 launches exactly as in the reference run; timing only",
              "kernel_form": plan_describe, "temporal_block": runner_max_tb}
    tkey = f"{shape}:{'x'.join(map(str, dims))}:tb{runner_max_tb}"
    traffic = (measured_traffic(tkey) or {}).get("bytes")
    kernel = ((f"k_stencil1d_tb (tb = {runner_max_tb})" if runner_max_tb > 1 else "k_stencil1d")
              if len(dims) == 1 else f"k_stencil{len(dims)}d")
    hbm = {"bound": "hbm", "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s", "frac": achieved / hbm_gbs,
           "traffic": traffic, "traffic_source": (measured_traffic(tkey) or {}).get("source"),
           "algorithmic_bytes_per_launch": cells_per_gpu * 16 * steps_per_launch,
           "time_steps_per_launch": steps_per_launch, "peak_source": peak_src,
           "frac_of_nominal_8TBs": achieved / 8000.0,
           "note": "SURVEY.md 8(d) convention: fixed 16 B per cell per time step; a temporally blocked sweep moves fewer "
                   "DRAM bytes than that, so this fraction can exceed 1 -- dram_frac is what the DRAM really does"}
    if traffic:
        dram_gbs = traffic / (us_per_launch * 1e-6) / 1e9
        hbm["dram_achieved_gbs"] = dram_gbs
        hbm["dram_frac"] = dram_gbs / hbm_gbs
    flop_per_cell = {1: 18.0}.get(len(dims))  # 9 taps = 9 FP64 FMA per cell per time step
    try:
        fp64_peak = float(json.load(open(os.path.join(ROOT, "profiles", "r1_fp64_pipes.json")))["dfma_tflops"])
    except (OSError, ValueError, KeyError):
        fp64_peak = None
    if flop_per_cell and fp64_peak and runner_max_tb > 1:
        # the fused sweep is bound by the FP64 pipe, not by HBM: that is the roofline reported first; the fixed-16-B
        # HBM figure of the SURVEY convention sits beside it
        tf = value / world * flop_per_cell / 1e3
        roofline = {"bound": "fp64", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak,
                    "traffic": traffic, "kernel": kernel, "us_per_launch": us_per_launch,
                    "peak_source": "measured DFMA stream on this pool (profiles/r1_fp64_pipes.json; MEASURED_PEAKS.json has no "
                                   "FP64 figure): FP64 FMA and FP64 DMMA share one pipe at 36.5-37.1 TFLOP/s",
                    "note": "9 FMA per cell per time step is the floor for general 9-tap weights; with "
                            f"{runner_max_tb} launches fused per sweep this pipe, not HBM, bounds the kernel",
                    "hbm_fixed16": hbm}
    else:
        roofline = dict(hbm, kernel=kernel, us_per_launch=us_per_launch)
    line = {
        "metric": "GStencil/s", "value": value, "unit": "GStencil/s (cells x launches / s / 1e9)",
        "value_artifact_units": value * ARTIFACT_K[shape],
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config, "detail": detail,
        "gpu_launches": gpu_launches,
        "clocks": clocks.summary(),
        "roofline": roofline,
    }
    if e2e:
        line["e2e"] = e2e
    if parity is not None:
        line["parity"] = parity
    if extra is not None:
        line["scaling_extra"] = extra
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baselines_in_a_clean_process([(shape, list(dims), {})])[0]
    if world == 1 and not args.no_shapes:
        line["shapes"] = [measure_shape(torch, ls, ops, s, d, l, hbm_gbs, local) for s, d, l in SHAPE_TABLE]
        if not args.no_cpu:  # the reference's test_cpu per shape: 1 core (as shipped) and all host cores
            jobs = []
            for s_, d_, _ in SHAPE_TABLE:
                jobs.append((s_, list(d_), {"min_seconds": 1.0, "max_launches": 1, "cores": 1}))
                jobs.append((s_, list(d_), {"min_seconds": 4.0, "max_launches": 8}))
            res = cpu_baselines_in_a_clean_process(jobs)
            for i, row in enumerate(line["shapes"]):
                row["cpu_baseline"] = {"one_core": res[2 * i], "all_cores": res[2 * i + 1]}
        worst = min(line["shapes"], key=lambda r: r["roofline_frac"])
        line["worst_shape_frac"] = worst["roofline_frac"]
        line["worst_shape"] = worst["shape"]
        try:  # additions beyond the reference's shape list: never allowed to cost the line
            line["extensions"] = measure_extensions(torch, ls, hbm_gbs)
        except Exception as e:  # noqa: BLE001
            line["extensions"] = {"error": f"{type(e).__name__}: {e}"}
    json_out.write(json.dumps(line) + "\n")
    json_out.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

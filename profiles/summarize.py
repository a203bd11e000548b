"""Turns ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python profiles/summarize.py full   gpurun_out/prof_r1_shapes.ncu-rep  profiles/r1_ncu_shapes.md
    python profiles/summarize.py list   gpurun_out/launches_r1.csv         profiles/r1_launches_bench.md
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of ncu peak"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor pipe inst"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/CTA"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "stall long_scoreboard %"),
    ("smsp__average_warp_latency_issue_stalled_short_scoreboard.pct", "stall short_scoreboard %"),
    ("smsp__average_warp_latency_issue_stalled_math_pipe_throttle.pct", "stall math_pipe_throttle %"),
    ("smsp__average_warp_latency_issue_stalled_mio_throttle.pct", "stall mio_throttle %"),
    ("smsp__average_warp_latency_issue_stalled_barrier.pct", "stall barrier %"),
    ("smsp__average_warp_latency_issue_stalled_wait.pct", "stall wait %"),
]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of `{rep}`\n\n")
        f.write("One launch per kernel, `--clock-control none`, caches flushed by ncu between replays (cold-cache numbers).\n")
        f.write("Algorithmic bytes = 16 B x interior cells.\n\n")
        for d in data:
            name = d[idx["Kernel Name"]]
            f.write(f"## {name}\n\n| metric | value |\n|---|---|\n")
            for m, label in METRICS:
                if m in idx:
                    f.write(f"| {label} (`{m}`) | {d[idx[m]]} {units[idx[m]]} |\n")
            f.write("\n")
    print("wrote", out)


def launches(csvfile, out):
    rows = [r for r in csv.reader(open(csvfile)) if len(r) > 5]
    hdr = rows[0]
    i_name, i_val = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        a = agg.setdefault(r[i_name], [0, 0.0])
        a[0] += 1
        a[1] += float(r[i_val].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary of `{csvfile}` (gpu__time_duration.sum, --clock-control none)\n\n")
        f.write("Per-launch times are cold-cache and serialised; compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{n[:110]}` | {c} | {t / 1e3:.1f} | {t / c / 1e3:.1f} | {100 * t / tot:.1f}% |\n")
    print("wrote", out)


if __name__ == "__main__":
    {"full": full, "list": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])

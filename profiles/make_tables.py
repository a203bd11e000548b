"""Markdown tables for DESIGN.md sections 5 and 6 from the committed bench records.
    python profiles/make_tables.py [profiles/r2_bench_n1.json profiles/r2_bench_n2.json profiles/r2_bench_n8.json]"""
import json
import os
import sys

here = os.path.dirname(os.path.abspath(__file__))
paths = sys.argv[1:] or [os.path.join(here, f"r2_bench_n{n}.json") for n in (1, 2, 8)]


def load(p):
    with open(p) as f:
        return json.loads(f.read().strip().splitlines()[-1])


recs = {}
for p in paths:
    if os.path.exists(p):
        d = load(p)
        recs[d["n_gpus"]] = d
d1 = recs.get(1)
if d1:
    print("| shape (size) | device-resident, best / median of 5 | fixed-16-B fraction | e2e pinned / pageable | `test_cpu` 1 core / all cores | launches per sweep |")
    print("|---|---|---|---|---|---|")
    print(f"| 1d2r 2^28 x 1000 (headline `value`) | {d1['value']:.0f} | {d1['roofline']['hbm_fixed16']['frac']:.2f} (FP64 pipe {d1['roofline']['frac']:.2f}) | "
          f"{d1['e2e']['value']:.0f} / -- | -- / {d1['cpu_baseline']['value']:.1f} ({d1['cpu_baseline']['cores']} cores) | 15 |")
    for v in d1["shapes"]:
        cb = v.get("cpu_baseline") or {}
        one, allc = cb.get("one_core") or {}, cb.get("all_cores") or {}
        cpu = f"{one.get('value', 0):.2f} / {allc.get('value', 0):.1f} ({allc.get('cores', '?')})" if one and allc else "--"
        print(f"| {v['shape']} | {v['gstencils']:.0f} / {v['gstencils_median']:.0f} | {v['roofline_frac']:.2f} | "
              f"{v['e2e']['value']:.0f} / {v['e2e_pageable']['value']:.0f} | {cpu} | {v['temporal_block']} |")
    print()
rows = {}
for n, d in sorted(recs.items()):
    rows.setdefault(("1d2r 2^28 per GPU x 1000, weak (headline)", ""), {})[n] = d["value"]
    rows.setdefault(("same, end to end with host buffers (`e2e`)", ""), {})[n] = d["e2e"]["value"]
    for e in d.get("scaling_extra", []):
        key = (f"{e['shape']} {'x'.join(map(str, e.get('dims', []))) if e.get('dims') else ''} {e['mode']}", f"tb {e.get('temporal_block')}")
        rows.setdefault(key, {})[n] = e["value"]
ns = sorted(recs)
print("| workload | " + " | ".join(f"N = {n}" + (" (eff.)" if n > 1 else "") for n in ns) + " |")
print("|---|" + "---|" * len(ns))
for (name, tb), vals in rows.items():
    base = vals.get(1)
    cells = []
    for n in ns:
        v = vals.get(n)
        if v is None:
            cells.append("--")
        elif n == 1 or not base:
            cells.append(f"{v:.0f}")
        else:
            cells.append(f"{v:.0f} ({v / (n * base):.3f})")
    print(f"| {name} {tb} | " + " | ".join(cells) + " |")

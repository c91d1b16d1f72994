#!/usr/bin/env python
"""Kernel shares of one decode: ncu launch list (gpu__time_duration, cold cache, serialised) vs bench.py's event-timed table.
    python tools/launch_shares.py gpurun_out/launches_r02b.csv gpurun_out/bench_r02b.json > profiles/r02b_launch_shares.txt"""
import csv
import json
import sys

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from tools.ncu_traffic import category  # noqa: E402

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = None, []
for r in rows:
    if r and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(dict(zip(hdr, r)))
# the last complete decode: from the last beam_init_kernel to the following beam_finalize_kernel
names = [d["Kernel Name"] for d in data]
ends = [i for i, n in enumerate(names) if "beam_finalize" in n]
starts = [i for i, n in enumerate(names) if "beam_init" in n]
end = ends[-1]
start = max(s for s in starts if s < end)
sel = data[start:end + 1]
agg = {}
for d in sel:
    c = category(d["Kernel Name"])
    agg[c] = agg.get(c, 0.0) + float(d["Metric Value"]) / 1e3
tot = sum(agg.values())
bench = json.load(open(sys.argv[2]))
kern = bench["kernels"]
btot = sum(v["ms_per_step"] for v in kern.values())
print(f"one decode (beam_init .. beam_finalize): {len(sel)} launches, {tot / 1e3:.3f} ms under ncu (cold cache, serialised, full clock); "
      f"bench: {bench['ms_per_step']:.3f} ms per step (graph replay), event-timed kernel sum {btot:.3f} ms")
print(f"{'category':14s} {'ncu us':>10s} {'ncu share':>10s} {'event ms':>10s} {'event share':>12s}")
for c in sorted(set(agg) | set(kern)):
    a = agg.get(c, 0.0)
    b = kern.get(c, {}).get("ms_per_step", 0.0)
    print(f"{c:14s} {a:10.1f} {100 * a / tot:9.1f}% {b:10.3f} {100 * b / btot:11.1f}%")

#!/usr/bin/env python
"""Where a captured kernel spends its issue slots and stall samples: per 50-instruction block of SASS, plus the opcode mix of
the hot range.  Usage: python tools/ncu_hot.py <file.ncu-rep> [launch index] [units per kernel, to print opcodes per unit]"""
import collections, csv, subprocess, sys, io

rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0; units = float(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw))); h = rr[0]
keys = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "smsp__average_warp_latency_per_inst_issued.ratio"]
for k in keys:
    if k in h: print(k, [r[h.index(k)] for r in rr[2:]])
for i, name in enumerate(h):
    if "issue_stalled" in name and name.endswith("per_issue_active.ratio"):
        val = float(rr[2 + which][i])
        if val > 0.1: print("  stall", name.split("issue_stalled_")[1].split("_per_issue")[0], round(val, 2))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
per = max(1, (len(starts) - 1) // max(1, len(rr) - 2))  # the source page may list every launch more than once
a, b = starts[which * per], starts[which * per + 1]
print(rows[a][1][:100])
hdr, body = rows[a + 1], rows[a + 2:b]
S, E = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
tot = sum(int(r[S]) for r in body); tote = sum(int(r[E]) for r in body)
print("samples", tot, "warp instructions", tote, "sass lines", len(body))
for k in range(0, len(body), 50):
    s = sum(int(r[S]) for r in body[k:k + 50]); e = sum(int(r[E]) for r in body[k:k + 50])
    if s > 0.02 * tot or e > 0.02 * tote: print(f"{k:5d} samples {100 * s / tot:5.1f}%  instr {100 * e / tote:5.1f}%  {body[k][1].strip()[:60]}")
if units:
    c = collections.Counter()
    for r in body:
        ins = r[1].strip().split()
        if not ins: continue
        op = ins[1] if ins[0].startswith("@") and len(ins) > 1 else ins[0]
        c[op.split(".")[0]] += int(r[E])
    print("per unit:", round(tote / units, 1), [(k, round(v / units, 1)) for k, v in c.most_common(16)])

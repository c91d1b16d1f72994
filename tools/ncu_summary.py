#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): key raw metrics per captured launch + top stalled source lines."""
import csv, subprocess, sys
rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 12
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed.sum',
        'sm__cycles_elapsed.max', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__block_size',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_barrier_per_warp_active.pct', 'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct']
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index('Kernel Name')][:90])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"   {w:75s} {r[i]:>16s} {units[i]}")
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {'name': r[1], 'rows': []}; secs.append(cur); continue
    if cur is not None: cur['rows'].append(r)
seen = set()
for s in secs:
    if s['name'] in seen or not s['rows']: continue
    seen.add(s['name'])
    hdr = s['rows'][0]; body = s['rows'][1:]
    i_src, i_s, i_ex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    ok = [b for b in body if len(b) > i_s and b[i_s].isdigit()]
    tot = sum(int(b[i_s]) for b in ok)
    print("== source:", s['name'][:80], "samples", tot, "sass lines", len(body))
    agg = {}
    for b in ok:
        for j in range(len(hdr)):
            if hdr[j].startswith('stall_') and 'Not Issued' not in hdr[j] and b[j].isdigit():
                agg[hdr[j]] = agg.get(hdr[j], 0) + int(b[j])
    print("   stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
    for b in sorted(ok, key=lambda b: -int(b[i_s]))[:ntop]:
        st = {hdr[j]: int(b[j]) for j in range(len(hdr)) if hdr[j].startswith('stall_') and 'Not Issued' not in hdr[j] and b[j].isdigit() and int(b[j]) > 0}
        dom = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        print(f"   {int(b[i_s]):6d} {b[i_ex]:>9s}  {b[i_src].strip()[:64]:64s} {dom}")

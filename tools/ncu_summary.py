"""Summarise an ncu report: headline metrics + dynamic SASS opcode histogram per draw (uses `ncu -i`)."""
import csv, re, collections, subprocess, sys, io
rep = sys.argv[1]; per = float(sys.argv[2]) if len(sys.argv) > 2 else 4096 * 1100
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__cycles_elapsed.avg', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.avg.per_second', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
for h, u, v in zip(hdr, units, vals):
    if h in want or (h.startswith('smsp__average_warps_issue_stalled') and float(v or 0) > 0.1):
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; isrc = hdr.index("Source"); iex = hdr.index("Instructions Executed")
byop = collections.Counter(); tot = 0
for r in rows[2:]:
    try: ex = int(r[iex])
    except Exception: continue
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[isrc]); op = m.group(2) if m else '?'
    byop[op] += ex; tot += ex
print("warp instructions per unit (draw):", round(tot / per, 1))
print("  ".join(f"{op}:{c/per:.1f}" for op, c in byop.most_common(28)))
print("fp64 per unit:", round(sum(c for op, c in byop.items() if op in ('DFMA', 'DADD', 'DMUL', 'DSETP', 'DMNMX')) / per, 1))

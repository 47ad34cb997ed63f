#!/usr/bin/env python
"""Print the key metrics of an .ncu-rep (run here, no GPU needed): python scripts/ncu_summary.py gpurun_out/x.ncu-rep"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_lsu.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum',
        'sm__cycles_elapsed.max', 'sm__cycles_active.avg']
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') or h.startswith('smsp__average_warp_latency_issue_stalled')]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('=' * 100)
    print(d.get('Kernel Name', '')[:110])
    for w in want:
        if w in d:
            print(f"  {w:72s} {d[w]:>18s} {units[hdr.index(w)]}")
    st = sorted(((float(d[h].replace(',', '')), h) for h in stall if d.get(h) not in (None, '', 'n/a')), reverse=True)[:8]
    for v, h in st:
        print(f"  stall {h.replace('smsp__average_', ''):66s} {v:18.3f}")

#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): python profiles/ncu_summary.py gpurun_out/x.ncu-rep"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed.sum', 'smsp__inst_executed_op_shared_atom.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum']
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
for r in rows[2:]:
    print('-' * 100)
    for w in want:
        if w in hdr:
            i = hdr.index(w); print(f"{w} = {r[i]} {units[i]}")
    st = sorted(((float(r[hdr.index(s)] or 0), s) for s in stall), reverse=True)[:8]
    for v, s in st:
        print(f"  stall {s.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')}: {v:.3f}")

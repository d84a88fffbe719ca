#!/usr/bin/env python
"""Key figures of every kernel in an .ncu-rep (ncu --set full): duration, tensor / tc pipe activity, issue slots, L1TEX / L2 /
DRAM throughput and bytes, instructions.  usage: ncu_summary.py rep [label ...]  (labels name the launches in order)"""
import csv, subprocess, sys
rep, labels = sys.argv[1], sys.argv[2:]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
WANT = [('gpu__time_duration.sum', 'duration'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor math pipe active'),
        ('sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active', 'tc pipe (math + operand fetch) active'),
        ('l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'shared-memory pipe: tensor operand wavefronts'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput'),
        ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput'),
        ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput'),
        ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM written'),
        ('smsp__inst_executed.sum', 'warp instructions'), ('launch__grid_size', 'grid'),
        ('launch__registers_per_thread', 'registers / thread')]
name_i = hdr.index('Kernel Name')
for k, r in enumerate(data):
    print(f'== launch {k}: {r[name_i][:60]}  {labels[k] if k < len(labels) else ""}')
    for m, txt in WANT:
        if m in hdr:
            i = hdr.index(m)
            print(f'   {txt:48s} {r[i]:>16s} {units[i]}')

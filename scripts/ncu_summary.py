"""Summarise an .ncu-rep (one or more captured launches) into a small CSV for profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_kernel_ncu_full_summary.csv
"""
import csv
import re
import subprocess
import sys

KEEP = [r'^gpu__time_duration\.sum$', r'^dram__bytes_(read|write)\.sum$', r'^dram__throughput\.avg\.pct_of_peak_sustained_elapsed$',
        r'^sm__cycles_elapsed\.avg\.per_second$', r'^launch__(registers_per_thread|grid_size|block_size|occupancy_limit_\w+|waves_per_multiprocessor)$',
        r'^sm__warps_active\.avg\.pct_of_peak_sustained_active$', r'^smsp__inst_executed\.sum$', r'^smsp__issue_active\.avg\.pct_of_peak_sustained_active$',
        r'^sm__inst_executed_pipe_(fma|alu|lsu|xu)\.avg\.pct_of_peak_sustained_active$', r'^sm__pipe_fma_cycles_active\.avg\.pct_of_peak_sustained_active$',
        r'^l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum(\.pct_of_peak_sustained_elapsed)?$',
        r'^l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_(ld|st|atom)\.sum$', r'^l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed$',
        r'^lts__throughput\.avg\.pct_of_peak_sustained_elapsed$', r'^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$',
        r'^smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio$', r'^smsp__cycles_active\.avg$']


def main():
    rep = sys.argv[1]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    w = csv.writer(sys.stdout)
    w.writerow(['kernel', 'metric', 'unit', 'value'])
    ki = head.index('Kernel Name')
    for r in rows[2:]:
        for i, k in enumerate(head):
            if any(re.search(p, k) for p in KEEP):
                w.writerow([r[ki], k, units[i], r[i]])


if __name__ == '__main__':
    main()

#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into a small text table for profiles/.

    python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep > profiles/rNN_x.txt
"""
import csv
import subprocess
import sys

WANT = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__cycles_active.avg']


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        idx = [hdr.index(w) for w in WANT if w in hdr]
        print('# %s' % rep)
        for r in rows[2:]:
            for i in idx:
                print('%-72s %s %s' % (hdr[i], r[i][:90], units[i]))
            print()


if __name__ == '__main__':
    main()

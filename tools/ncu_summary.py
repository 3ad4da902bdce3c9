#!/usr/bin/env python
"""Summarise an .ncu-rep (read here with `ncu -i`, no GPU needed) into a markdown table for profiles/.
usage: tools/ncu_summary.py <report.ncu-rep> "<title>" [max_rows] > profiles/<name>.md"""
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_bytes.sum']


def main():
    rep, title = sys.argv[1], sys.argv[2]
    maxr = int(sys.argv[3]) if len(sys.argv) > 3 else 12
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    cols = [w for w in WANT if w in hdr]
    print("# %s\n" % title)
    print("| kernel | " + " | ".join("%s [%s]" % (w, units[hdr.index(w)]) for w in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for r in rows[2:2 + maxr]:
        print("| %s | " % r[hdr.index('Kernel Name')].split('(')[0] + " | ".join(r[hdr.index(w)] for w in cols) + " |")


if __name__ == "__main__":
    main()

"""Summarise `ncu -i suite.ncu-rep --page raw --csv` into one row per kernel launch with the roofline-relevant metrics.
usage: ncu -i gpurun_out/suite.ncu-rep --page raw --csv > raw.csv ; python tools/summarize_ncu_suite.py raw.csv > profiles/<name>.csv"""
import csv
import re
import sys

KEEP = [('gpu__time_duration.sum', 'time_us'), ('dram__bytes_read.sum', 'dram_read_MB'), ('dram__bytes_write.sum', 'dram_write_MB'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor_pipe_pct'),
        ('sm__inst_executed_pipe_tensor.sum', 'tensor_inst'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_active_pct'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_pct'), ('lts__t_bytes.sum', 'l2_MB'),
        ('launch__registers_per_thread', 'regs'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
        ('launch__shared_mem_per_block_dynamic', 'dyn_smem')]


def num(v):
    try:
        return float(v.replace(',', ''))
    except Exception:
        return None


def main():
    rows = list(csv.reader(open(sys.argv[1], newline='')))
    hdr_i = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr, units = rows[hdr_i], rows[hdr_i + 1]
    col = {n: i for i, n in enumerate(hdr)}
    out = csv.writer(sys.stdout)
    out.writerow(['kernel'] + [k for _, k in KEEP] + ['achieved_GBs'])
    for r in rows[hdr_i + 2:]:
        if len(r) < len(hdr):
            continue
        name = re.sub(r'\(.*', '', r[col['Kernel Name']])
        name = re.sub(r'rih::|void |tc::', '', name)[:90]
        vals = []
        t_us = None
        rd = wr = None
        for m, k in KEEP:
            v = num(r[col[m]]) if m in col else None
            u = units[col[m]] if m in col else ''
            if v is not None:
                if k == 'time_us':
                    v = v * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(u, 1e-3)
                    t_us = v
                if k.endswith('_MB'):
                    v = v * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(u, 1e-6)
                if k == 'dram_read_MB':
                    rd = v
                if k == 'dram_write_MB':
                    wr = v
            vals.append('' if v is None else ('%.4g' % v))
        gbs = ''
        if t_us and rd is not None and wr is not None:
            gbs = '%.1f' % ((rd + wr) * 1e6 / (t_us * 1e-6) / 1e9)
        out.writerow([name] + vals + [gbs])


if __name__ == '__main__':
    main()

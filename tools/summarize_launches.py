"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel name.
usage: python tools/summarize_launches.py gpurun_out/launches.csv [top_n]"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    rows = []
    with open(path, newline='') as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r.get('Metric Unit', 'ns')
        scale = {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'nsecond': 1e-3, 'ms': 1e3, 'msecond': 1e3}.get(unit, 1e-3)
        rows.append((r['Kernel Name'], v * scale))
    agg = defaultdict(lambda: [0, 0.0])
    for name, us in rows:
        short = re.sub(r'\(.*', '', name)
        short = re.sub(r'rih::|void |at::native::|<unnamed>::', '', short)[:110]
        agg[short][0] += 1
        agg[short][1] += us
    total = sum(v[1] for v in agg.values())
    print('total kernel time %.3f ms over %d launches' % (total / 1e3, len(rows)))
    print('%-112s %7s %10s %6s' % ('kernel', 'count', 'ms', 'share'))
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print('%-112s %7d %10.3f %5.1f%%' % (name, n, us / 1e3, 100 * us / total))


if __name__ == '__main__':
    main()

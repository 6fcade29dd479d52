#!/usr/bin/env python
"""Per-kernel table (count, total ms, average us, share) of an ncu launch list:
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv <command>;  python profiles/launch_table.py X.csv"""
import collections, csv, sys

def table(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        k = r[ki].split('(')[0][-70:]; v = float(r[vi].replace(',', '')) / 1e6
        agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += v
    tot = sum(v for _, v in agg.values())
    out = ['%12s %6s %10s %8s  kernel' % ('total ms', 'count', 'avg us', 'share')]
    for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append('%12.3f %6d %10.1f %7.2f%%  %s' % (v, c, 1e3 * v / c, 100 * v / tot, k))
    return '\n'.join(out)

if __name__ == '__main__':
    print(table(sys.argv[1]))

#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: instructions executed per SASS region (between branch targets)."""
import csv, sys, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia = hdr.index('Instructions Executed'); isrc = hdr.index('Source'); ist = hdr.index('# Samples'); ith = hdr.index('Avg. Threads Executed')
rows = rows[:2] + [r for r in rows[2:] if len(r) > max(ia, isrc, ist, ith)]
tot = sum(int(r[ia]) for r in rows[2:] if r[ia].isdigit())
print('total warp instructions', tot)
# group consecutive instructions with identical execution count (basic-block proxy)
blocks = []
cur = None
for r in rows[2:]:
    if not r[ia].isdigit():
        continue
    n = int(r[ia]); op = r[isrc].split()[0] if not r[isrc].split()[0].startswith('@') else r[isrc].split()[1]
    if cur and cur['n'] == n:
        cur['ops'].append(op); cur['samples'] += int(r[ist] or 0); cur['thr'].append(float(r[ith] or 0))
    else:
        cur = {'n': n, 'ops': [op], 'samples': int(r[ist] or 0), 'thr': [float(r[ith] or 0)], 'first': r[isrc].strip()}
        blocks.append(cur)
for b in sorted(blocks, key=lambda b: -b['n'] * len(b['ops']))[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    from collections import Counter
    c = Counter(b['ops'])
    print('%5.1f%% of instr | exec %11d x %3d instr | thr %4.1f | samples %6d | %s' % (
        100.0 * b['n'] * len(b['ops']) / tot, b['n'], len(b['ops']), sum(b['thr']) / len(b['thr']), b['samples'],
        ' '.join('%s:%d' % kv for kv in c.most_common(8))))

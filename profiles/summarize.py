#!/usr/bin/env python
"""Turns the ncu artefacts of gpurun_out/ into the text summaries committed under profiles/ (run here, no GPU needed)."""
import collections, csv, subprocess, sys

def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        k = r[ki].split('(')[0][-60:]; v = float(r[vi].replace(',', '')) / 1e6
        agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += v
    tot = sum(v for _, v in agg.values())
    with open(out, 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 2 --warmup 3 --kernels-only (c3, 1M)\n')
        f.write('# per-launch times are cold-cache and serialised: compare SHARES\n')
        f.write('%12s %6s %8s  kernel\n' % ('total ms', 'count', 'share'))
        for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write('%12.3f %6d %7.2f%%  %s\n' % (v, c, 100 * v / tot, k))

WANT = ['gpu__time_duration.sum', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg.per_second']

def full(rep, out):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]; units = rows[1]
    with open(out, 'w') as f:
        f.write('# ncu --set full --clock-control none (%s); selected raw metrics per captured launch\n' % rep)
        for r in rows[2:]:
            f.write('\n== %s\n' % r[hdr.index('Kernel Name')].split('(')[0])
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w); f.write('  %-70s %-10s %s\n' % (w, units[i], r[i]))
            st = sorted(((float(r[i].replace(',', '')), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''))
                         for i, h in enumerate(hdr) if 'smsp__average_warps_issue_stalled' in h and 'per_issue_active' in h), reverse=True)[:7]
            f.write('  stall reasons per issue: ' + ', '.join('%s=%.2f' % (n, v) for v, n in st) + '\n')

if __name__ == '__main__':
    rnd = sys.argv[1] if len(sys.argv) > 1 else 'r02'
    launches('gpurun_out/%s_launches_c3.csv' % rnd, 'profiles/%s_launches_c3.txt' % rnd)
    full('gpurun_out/%s_allpairs.ncu-rep' % rnd, 'profiles/%s_allpairs_full.txt' % rnd)
    full('gpurun_out/%s_tree_sph.ncu-rep' % rnd, 'profiles/%s_tree_sph_full.txt' % rnd)

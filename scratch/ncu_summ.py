"""Summarise ncu outputs (run here, no GPU): launch-list csv -> per-kernel table; .ncu-rep -> key metrics."""
import csv, subprocess, sys, collections, re

def launches(path, top=30):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); gi = hdr.index('Grid Size'); bi = hdr.index('Block Size')
    agg = collections.OrderedDict(); seq = []
    for r in rows[1:]:
        name = re.sub(r'\(.*', '', r[ki]).replace('void ', '').replace('<unnamed>::', '')
        us = float(r[vi].replace(',', '')) / 1e3
        seq.append((name, r[gi], r[bi], us))
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
    tot = sum(v[1] for v in agg.values())
    print('| kernel | launches | total us | share |\n|---|---|---|---|')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print('| `%s` | %d | %.1f | %.1f %% |' % (k[:70], v[0], v[1], 100 * v[1] / tot))
    print('\ntotal %.1f us over %d launches' % (tot, len(seq)))
    return seq

WANT = ['gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
        'smsp__inst_executed.sum', 'launch__shared_mem_per_block_dynamic', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active']

def rep(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ni = hdr.index('Kernel Name')
    for r in rows[2:]:
        print('\n### `%s`' % r[ni][:100])
        for i, h in enumerate(hdr):
            if h in WANT:
                print('- %s = %s %s' % (h, r[i], units[i]))

if __name__ == '__main__':
    if sys.argv[1].endswith('.csv'):
        seq = launches(sys.argv[1])
        if len(sys.argv) > 2:
            n = int(sys.argv[2])
            print('\nlast %d launches in program order:' % n)
            for s in seq[-n:]:
                print('| `%s` | %s | %s | %.1f |' % (s[0][:60], s[1], s[2], s[3]))
    else:
        rep(sys.argv[1])

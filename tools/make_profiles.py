"""Turn ncu outputs under gpurun_out/ into the tracked text summaries under profiles/.
   python tools/make_profiles.py <tag> launches=<csv> [full=<ncu-rep>:<title>] ..."""
import collections, csv, os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, 'profiles')
os.makedirs(OUT, exist_ok=True)

def launches(path, dst, title):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = row['Kernel Name'].split('(')[0].replace('void ', '')
        v = float(row['Metric Value'].replace(',', '')) * {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(row['Metric Unit'], 1)
        agg[name][0] += 1; agg[name][1] += v; tot += v
    with open(dst, 'w') as f:
        f.write('# %s\n# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n' % title)
        f.write('# total %.3f ms over %d launches\n' % (tot / 1e6, sum(n for n, _ in agg.values())))
        f.write('%-72s %6s %10s %7s %9s\n' % ('kernel', 'n', 'ms', 'share', 'avg_us'))
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write('%-72s %6d %10.3f %6.1f%% %9.1f\n' % (k[:72], n, t / 1e6, 100 * t / tot, t / n / 1e3))
    print('wrote', dst)

WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg', 'lts__t_sector_hit_rate.pct']

def full(rep, dst, title):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    if 'prof_conv' in os.path.basename(rep) and 'dram__bytes_read.sum' in hdr:
        # the roofline probe's DRAM traffic per launch, read by bench.py (roofline.traffic) -- tied to THIS capture
        import json
        r = rows[2]

        def to_bytes(name):
            v = float(r[hdr.index(name)].replace(',', ''))
            return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(units[hdr.index(name)], 1)
        tr = to_bytes('dram__bytes_read.sum') + to_bytes('dram__bytes_write.sum')
        with open(os.path.join(OUT, 'roofline_traffic.json'), 'w') as f:
            json.dump({'conv_probe_dram_bytes': tr, 'kernel': r[hdr.index('Kernel Name')][:80],
                       'source': 'dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture (profiles/%s)' % os.path.basename(dst)}, f, indent=1)
        print('wrote roofline_traffic.json', tr)
    with open(dst, 'w') as f:
        f.write('# %s\n# ncu --set full --clock-control none --import-source on  (source: %s)\n' % (title, os.path.basename(rep)))
        for r in rows[2:]:
            f.write('---- %s\n' % r[hdr.index('Kernel Name')][:100])
            for w in WANT:
                if w in hdr:
                    f.write('%-72s %s %s\n' % (w, r[hdr.index(w)], units[hdr.index(w)]))
    print('wrote', dst)

if __name__ == '__main__':
    tag = sys.argv[1]
    for a in sys.argv[2:]:
        kind, val = a.split('=', 1)
        if kind == 'launches':
            path, title = val.split(':', 1)
            launches(path, os.path.join(OUT, '%s_launches_%s.txt' % (tag, os.path.basename(path).replace('.csv', ''))), title)
        else:
            path, title = val.split(':', 1)
            full(path, os.path.join(OUT, '%s_ncu_%s.txt' % (tag, os.path.basename(path).replace('.ncu-rep', ''))), title)

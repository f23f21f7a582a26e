"""Top stall sites of a kernel from `ncu -i <rep> --page source --csv` (SASS view): address, samples, executed, dominant stall, instruction.
   python tools/ncu_hot.py <rep> [topN]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(raw) if l.startswith('"Address"'))
ends = [i for i, l in enumerate(raw) if i > start and l.startswith('"Kernel Name"')]
rows = list(csv.DictReader(raw[start:(ends[0] if ends else len(raw))]))      # first kernel of the report
stalls = [k for k in rows[0] if k.startswith('stall_') and 'Not Issued' not in k]
tot = sum(int(r['# Samples'] or 0) for r in rows)
print('total samples', tot, 'instructions', len(rows))
agg = {}
for r in rows:
    for k in stalls:
        agg[k] = agg.get(k, 0) + int(r[k] or 0)
print('stall mix:', ', '.join('%s %.1f%%' % (k[6:], 100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
rows.sort(key=lambda r: -int(r['# Samples'] or 0))
for r in rows[:top]:
    n = int(r['# Samples'] or 0)
    dom = max(stalls, key=lambda k: int(r[k] or 0))
    print('%6s %6d %5.1f%% exec %9s  %-14s %s' % (r['Address'][-5:], n, 100.0 * n / max(tot, 1), r['Instructions Executed'], dom[6:], r['Source'][:90]))

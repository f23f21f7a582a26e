#!/usr/bin/env python
"""Turn the per-tensor errors a `pytest -m gpu` run recorded on a B200 (gpurun_out/parity_errors.json, written by
tests/conftest.py) into tests/golden/parity_measured.json -- the table the bf16 gates of tests/test_parity_gpu.py are derived
from (gate = 3 x measured) -- and keep the readable copy under profiles/.

    python tools/update_parity_gates.py [round-tag]          # e.g. r02a
"""
import json
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else 'r02'
    rows = json.load(open(os.path.join(REPO, 'gpurun_out', 'parity_errors.json')))
    table = {}
    for r in rows:
        d = table.setdefault(r['prec'], {}).setdefault(r['case'], {})
        d[r['key']] = max(d.get(r['key'], 0.0), r['err'])            # a tensor checked twice keeps its worst error
    table['_meta'] = {'what': 'per-tensor parity error of libsg_b200 vs the goldens of the unmodified reference, measured on a B200 by '
                              'tests/test_parity_gpu.py (rel-L2; scalars: relative error; analytically-zero tensors: max |value|)',
                      'source': 'profiles/%s_parity_errors.txt' % tag}
    with open(os.path.join(REPO, 'tests', 'golden', 'parity_measured.json'), 'w') as f:
        json.dump(table, f, indent=1, sort_keys=True)
    shutil.copy(os.path.join(REPO, 'gpurun_out', 'parity_errors.txt'), os.path.join(REPO, 'profiles', '%s_parity_errors.txt' % tag))
    over = [(r['prec'], r['case'], r['key'], r['err']) for r in rows if r['prec'] == 'fp32x' and not r['note'].startswith('zero') and r['err'] > 1e-3]
    print('%d rows; fp32x tensors above 1e-3: %d' % (len(rows), len(over)))
    for o in sorted(over, key=lambda t: -t[3])[:12]:
        print('   %-6s %-28s %-44s %.2e' % o)
    outs = [r['err'] for r in rows if r['prec'] == 'fp32x' and r['note'].split()[0] in ('out', 'scalar')]
    print('fp32x outputs / scalars: max %.2e over %d' % (max(outs), len(outs)))


if __name__ == '__main__':
    main()

"""Aggregate an `ncu --metrics gpu__time_duration.sum,... --csv --log-file F` launch list by kernel name (developer tool).
usage: python tools/launch_summary.py F [skip_first_launches]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr, data = rows[hi], rows[hi + 1:]
iK, iM, iV, iU, iID = (hdr.index(k) for k in ('Kernel Name', 'Metric Name', 'Metric Value', 'Metric Unit', 'ID'))
t = collections.defaultdict(float)
n = collections.Counter()
for r in data:
    if len(r) <= iV or r[iM] != 'gpu__time_duration.sum' or int(r[iID]) < skip:
        continue
    v = float(r[iV].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[iU], 1.0)
    name = re.sub(r'\(.*', '', r[iK]).replace('void ', '').replace('impb::', '').replace('<unnamed>::', '')
    name = re.sub(r'Radix3<\(int\)(\d+), \(int\)(\d+), \(int\)(\d+)>', r'R\1.\2.\3', name)
    name = re.sub(r'\((int|bool)\)', '', name)
    t[name[:100]] += v
    n[name[:100]] += 1
tot = sum(t.values())
for k, v in sorted(t.items(), key=lambda kv: -kv[1])[:30]:
    print(f'{v:10.1f} us {100 * v / tot:5.1f}%  x{n[k]:<4d} {k}')
print(f'{tot:10.1f} us total')

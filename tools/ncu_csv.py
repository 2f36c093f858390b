"""Pretty-print an `ncu --csv` metrics log: one block per launch."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
data = {}
for r in rows:
    if len(r) > 10 and r[0] == 'ID':
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        key = (int(d['ID']), d['Kernel Name'][:70])
        data.setdefault(key, {})[d['Metric Name']] = d['Metric Value']
for k, v in sorted(data.items()):
    print(k)
    for m, val in v.items():
        print('   ', m, val)

"""Summarise an `ncu --page source --csv` dump: stall totals and the hottest SASS lines."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if r and r[0] == 'Address')
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows if len(r) == len(hdr) and r[0].startswith('0x')]
tot = sum(int(r[ix['# Samples']]) for r in data)
print('total samples', tot, 'instructions', len(data))
keys = ['stall_barrier', 'stall_lg', 'stall_long_sb', 'stall_math', 'stall_mio', 'stall_short_sb', 'stall_wait',
        'stall_not_selected', 'stall_selected', 'stall_no_inst', 'stall_dispatch', 'stall_branch_resolving']
for k in keys:
    print(f'  {k:24s}', sum(int(r[ix[k]]) for r in data))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:n]:
    print(r[ix['# Samples']].rjust(6), r[ix['Source']].strip()[:80].ljust(80), '| lg', r[ix['stall_lg']], 'long',
          r[ix['stall_long_sb']], 'mio', r[ix['stall_mio']], 'short', r[ix['stall_short_sb']], 'bar',
          r[ix['stall_barrier']], 'math', r[ix['stall_math']])

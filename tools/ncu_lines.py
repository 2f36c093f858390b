"""Attribute the executed instructions of one kernel of an ncu report to source lines and opcodes (developer tool).

The ncu CLI exports the SASS view with per-instruction counters (`--page source --csv`) but not the CUDA-source view;
this joins that export with the line table of the same kernel from `nvdisasm -g` (build with -lineinfo).

usage: python tools/ncu_lines.py <report.ncu-rep> <kernel regex for ncu> <mangled-name substring> <cubin> [depth]
  cubin: one of the files `cuobjdump -xelf all impulcifer_pip313_b200/libimpulcifer_b200.so` writes
  depth: 0 = innermost inlined location (default), -1 = outermost
"""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    rep, kernel, mangled, cubin = sys.argv[1:5]
    depth = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    text = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{kernel}'],
                          capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    h = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
    hdr = rows[h]
    ie, ismp, isrc = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Source')
    execd = []
    for r in rows[h + 1:]:
        if len(r) > ie and r[ie]:
            try:
                execd.append((r[isrc].strip(), int(r[ie]), int(r[ismp] or 0)))
            except ValueError:
                pass
    dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout.split('\n')
    start = [i for i, l in enumerate(dis) if l.startswith('.text.') and mangled in l][0]
    cur, seq = None, []
    for l in dis[start + 1:]:
        if (l.startswith('.text.') or l.startswith('.section')) and seq:
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            cur = [(m.group(1).split('/')[-1], int(m.group(2)))]
            cur += [(a.split('/')[-1], int(b)) for a, b in re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))]
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+.*?;', l):
            seq.append(cur)
    n = min(len(execd), len(seq))  # a report with several launches of the kernel repeats the listing
    by_line, by_op, samples, total = collections.Counter(), collections.Counter(), collections.Counter(), 0
    for (src, e, s), chain in zip(execd[:n], seq[:n]):
        key = (chain[depth] if depth >= 0 and depth < len(chain) else chain[-1]) if chain else ('?', 0)
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', src)
        by_line[key] += e
        by_op[m.group(2) if m else '?'] += e
        samples[key] += s
        total += e
    print(f'{total} warp instructions over {n} SASS lines')
    for k, v in by_line.most_common(25):
        print(f'  {k[0]}:{k[1]:<5d} {100 * v / total:5.1f} %   stall samples {samples[k]}')
    print('opcodes:')
    for op, v in by_op.most_common(16):
        print(f'  {op:10s} {100 * v / total:5.1f} %')


if __name__ == '__main__':
    main()

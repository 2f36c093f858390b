"""Turn the two ncu CSV logs (gpurun_out/all_kernels.csv, gpurun_out/launches_r01b.csv) into the tables committed
under profiles/ (r01_all_kernels.md, r01_launches.md)."""
import collections
import csv
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)


def short(n):
    if n.startswith('void at::') or 'at::native' in n or 'elementwise_kernel' in n or 'CatArray' in n or 'indexFunc' in n:
        return '(torch index / elementwise plumbing)'
    m = re.match(r'(void )?(impb::|<unnamed>::)?(\w+)', n)
    key = m.group(3)
    if key.startswith(('k_tile_', 'k_fxg_')):
        fun = re.findall(r'<unnamed>::(\w+)', n)
        rx = re.search(r'Radix3<(\d+), (\d+), (\d+)>', n)
        row = re.search(r'(?:Radix3<[^>]*>|, \d+), (true|false|1|0),', n)
        key += '<' + ('S=%d' % (int(rx.group(1)) * int(rx.group(2)) * int(rx.group(3))) if rx else 'generic')
        key += ',' + (fun[0] if fun else 'RowPlain') + '>'
    if key.startswith('k_fx_'):
        key += '<double>' if '<double' in n else '<float>'
    return key


def read(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    per = {}
    for x in csv.DictReader(lines):
        k = int(x['ID'])
        per.setdefault(k, {'name': x['Kernel Name']})
        v = float(x['Metric Value'].replace(',', ''))
        u = x['Metric Unit']
        if x['Metric Name'].startswith('dram__bytes'):
            v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
        if x['Metric Name'] == 'gpu__time_duration.sum':
            v *= {'ns': 1e-3, 'us': 1, 'ms': 1e3}[u]
        per[k][x['Metric Name']] = v
    return per


hbm = json.load(open('MEASURED_PEAKS.json')).get('hbm_gbs', 6535.7)
per = read('gpurun_out/all_kernels.csv')
agg = collections.OrderedDict()
for k, d in sorted(per.items()):
    a = agg.setdefault(short(d['name']), [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get('gpu__time_duration.sum', 0)
    a[2] += d.get('dram__bytes_read.sum', 0)
    a[3] += d.get('dram__bytes_write.sum', 0)
tot = sum(a[1] for a in agg.values())
out = ["# Round 1 — every kernel of the hot path: device time and achieved HBM traffic (ncu, B200, final build)", "",
       "Command: `ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
       "--clock-control none --csv python tools/all_kernels.py 32`",
       "(after the same command exited 0 without ncu).  One pass = `BatchDispatcher.process` + `postprocess` on 32 synthetic 7.1 "
       "sets (448 IRs) followed by the", "`HRIR` chain on one set (ingest, crop, align, decay times, reflection levels, "
       "room correction, headphone compensation on the uncropped length (chirp-z), channel balance, virtual bass, "
       "normalize, resample, PCM packing).  Times under ncu are",
       f"serialised and cold-cache.  HBM % is against the measured copy peak {hbm:.1f} GB/s (`MEASURED_PEAKS.json`).", "",
       "| kernel | launches | time us | share | DRAM read MB | DRAM write MB | achieved GB/s | % of HBM peak |",
       "|---|---|---|---|---|---|---|---|"]
for key, (c, t, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    gbs = (rd + wr) / (t * 1e-6) / 1e9 if t > 0 else 0
    out.append(f"| `{key}` | {c} | {t:.1f} | {100 * t / tot:.1f} % | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {gbs:.0f} | {100 * gbs / hbm:.1f} % |")
out += ["", f"Total {tot / 1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches."]
open('profiles/r01_all_kernels.md', 'w').write("\n".join(out) + "\n")
shutil.copy('gpurun_out/all_kernels.csv', 'profiles/r01_all_kernels.csv')

per = read('gpurun_out/launches_r01b.csv')
recs = [(k, d['name'], d['gpu__time_duration.sum']) for k, d in sorted(per.items())]
starts = [i for i, (_, n, _) in enumerate(recs) if 'k_fx_fwd' in n]
# a bench step = the launches between two consecutive K1 column passes of full-batch process() calls (the warm-up /
# timed steps have ~90 launches in between; the chunked e2e calls that follow are shorter): take the last full one
# (a step may hold several K1 chunks: its first column pass is the one that follows a long gap)
begins = [b for a, b in zip(starts[:-1], starts[1:]) if b - a >= 60]
lo, hi = begins[-2], begins[-1]
step = recs[lo:hi]
tot = sum(v for _, _, v in step)
agg = collections.OrderedDict()
for _, n, v in step:
    a = agg.setdefault(short(n), [0, 0.0])
    a[0] += 1
    a[1] += v
out = ["# Round 1 — launch list of one bench step (128 recording sets = 1792 IRs), B200, final build of the round", "",
       "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 700 --csv python bench.py --steps 1 --warmup 3 "
       "--no-cpu-baseline --no-fp32`",
       "(taken right after the same command exited 0 without ncu; per-launch times under ncu are cold-cache and serialised: "
       "compare shares).",
       f"One step = launches {recs[lo][0]}..{recs[hi - 1][0]} of the capture (`r01_launches.csv`): {len(step)} launches, "
       f"{tot / 1e3:.3f} ms of device time.", "", "| launches | device time us | share | kernel |", "|---|---|---|---|"]
for key, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| {c} | {v:.1f} | {100 * v / tot:.1f} % | `{key}` |")
k1 = sum(v for k, (c, v) in agg.items() if k.startswith('k_fx_'))
out += ["", f"K1 (k_fx_fwd_async + k_fx_conv + k_fx_inv) = {100 * k1 / tot:.1f} % of the step's device time under ncu."]
open('profiles/r01_launches.md', 'w').write("\n".join(out) + "\n")
shutil.copy('gpurun_out/launches_r01b.csv', 'profiles/r01_launches.csv')
print(open('profiles/r01_all_kernels.md').read()[-3500:])
print(open('profiles/r01_launches.md').read()[-2200:])

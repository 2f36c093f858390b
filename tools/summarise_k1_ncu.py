"""profiles/<tag>_k1_ncu_full.md (+ k1_traffic.json for the FP64 capture) from an `ncu --set full` report of the K1
launch triple.  usage: python tools/summarise_k1_ncu.py gpurun_out/k1_full.ncu-rep [tag=r02] [precision=64]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
rep = sys.argv[1]
tag = sys.argv[2] if len(sys.argv) > 2 else 'r02'
prec = int(sys.argv[3]) if len(sys.argv) > 3 else 64
suffix = '' if prec == 64 else '_f32'
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
open(f'profiles/{tag}_k1_ncu_raw{suffix}.csv', 'w').write(raw)
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]


def col(name):
    return hdr.index(name)


def val(r, name):
    i = col(name)
    v = float(r[i].replace(',', ''))
    return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1, 'ms': 1e3, 'us': 1, 'ns': 1e-3}.get(units[i], 1)


hbm = json.load(open('MEASURED_PEAKS.json')).get('hbm_gbs', 6535.7) if os.path.exists('MEASURED_PEAKS.json') else 6535.7
pipe = 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active' if prec == 64 else \
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'
out = [f"# {tag} — K1 (sweep deconvolution, FP{prec}), `ncu --set full --clock-control none`, B200", "",
       "Command (after the same command exited 0 without ncu):",
       f"`ncu --set full --clock-control none --import-source on -k regex:k_fx -s 6 -c 3 python tools/quick_bench.py 224 {prec} 2`",
       f"(224 IRs = 112 ear pairs per launch, N = 768 x 768).  Raw metric dump: `{tag}_k1_ncu_raw{suffix}.csv`.", "",
       f"| kernel | time us | DRAM rd MB | DRAM wr MB | DRAM GB/s | DRAM % | LSU wavefronts % | {'FP64' if prec == 64 else 'FMA'} pipe % | issue % | warps % | regs "
       "| long_sb | mio | lg | short_sb | barrier |", "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
tot_b = tot_t = 0.0
for r in data:
    n = r[col('Kernel Name')].split('<')[0].replace('void ', '').replace('impb::', '')
    t, rd, wr = val(r, 'gpu__time_duration.sum'), val(r, 'dram__bytes_read.sum'), val(r, 'dram__bytes_write.sum')
    tot_b += rd + wr
    tot_t += t
    g = lambda k: float(r[col(k)].replace(',', ''))
    out.append(f"| {n} | {t:.1f} | {rd / 1e6:.0f} | {wr / 1e6:.0f} | {(rd + wr) / t / 1e3:.0f} | "
               f"{g('dram__cycles_active.avg.pct_of_peak_sustained_elapsed'):.1f} | "
               f"{g('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'):.1f} | "
               f"{g(pipe):.1f} | "
               f"{g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | "
               f"{g('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {g('launch__registers_per_thread'):.0f} | "
               f"{g('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio'):.2f} | "
               f"{g('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio'):.2f} | "
               f"{g('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio'):.2f} | "
               f"{g('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio'):.2f} | "
               f"{g('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio'):.2f} |")
per_ir = tot_b / 224
alg = 6260320 if prec == 64 else 3130160
out += ["", "(stall columns: warps stalled per issued instruction.)", "",
        f"DRAM traffic of the launch triple: {tot_b / 1e9:.2f} GB for 224 IRs = **{per_ir / 1e6:.2f} MB per IR** against "
        f"{alg / 1e6:.2f} MB algorithmic ({tot_b / tot_t / 1e3:.0f} GB/s = {100 * tot_b / tot_t / 1e3 / hbm:.0f} % of the measured "
        f"{hbm:.1f} GB/s copy peak while the kernels run).  Sum of the three kernels: {tot_t:.1f} us for 224 IRs."]
open(f'profiles/{tag}_k1_ncu_full{suffix}.md', 'w').write("\n".join(out) + "\n")
if prec == 64:
    json.dump({"dram_bytes_per_ir": per_ir,
               "source": f"profiles/{tag}_k1_ncu_full.md (ncu --set full, 224 IRs per launch triple)"},
              open('profiles/k1_traffic.json', 'w'))
print("\n".join(out[6:12]))
print(per_ir)

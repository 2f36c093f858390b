#!/usr/bin/env python
"""Headline benchmark: IR deconvolutions/s for the 6.15 s @ 48 kHz sweep (BASELINE.json), 7.1 layout.

    python bench.py --gpus N --steps K --warmup W              # this framework (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's OWN CPU code (oracle/_ref) under
                                                                     # its own pool; the oracle port if _ref is absent

A "step" is one pass of the hot path over one batch of synthetic recording sets per GPU:
  deconvolution of all 14 (speaker, ear) columns of every set (K1) -> first-peak search (K2) -> pairwise
  head crop + fade-in (K3) -> ipsilateral / onset alignment (K6, K7) -> Lundeby knee (K4) -> tail crop (K5)
i.e. ``HRIR.open_recording -> crop_heads -> align_ipsilateral_all -> align_onset_groups_peak_leftref ->
crop_tails`` of the reference, through ``BatchDispatcher.process``.  ``value`` counts deconvolved impulse
responses per second over all ranks with the recordings already resident in HBM; ``e2e`` is the same
call fed from pinned HOST buffers with the H2D copy of the recordings and the D2H read of the cropped
responses inside the timed region.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

FS = 48000
LAYOUT = ['FL', 'FR', 'FC', 'SL', 'SR', 'BL', 'BR']
N_IRS = 2 * len(LAYOUT)
N_BASE_SETS = 2          # distinct host-synthesised recording sets; device batches are seeded variants


def _host_sets(test_signal, n):
    from impulcifer_pip313_b200 import synth
    return [synth.recording_set(test_signal, FS, LAYOUT, seed=1234 + i) for i in range(n)]


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path, one recording set per task, all host cores
# ---------------------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_init(inv, est_len, n_oct):
    _CPU.update(inv=inv, est_len=est_len, n_oct=n_oct)
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(1)
    except Exception:
        pass


def _cpu_one_set(rec):
    import impulcifer_oracle as O
    irs = O.open_recording(rec, FS, LAYOUT, _CPU['inv'], FS)
    O.crop_heads(irs, FS)
    O.align_ipsilateral_all(irs, FS)
    O.align_onset_groups_peak_leftref(irs)
    tail = O.crop_tails(irs, FS, _CPU['est_len'], _CPU['n_oct'], FS)
    return tail


def cpu_throughput(test_signal, inv, n_oct, base_sets, sets_per_round, rounds, warm=1):
    """IRs/s of the CPU path: process pool over all host cores (the reference's own parallel_map is a
    ProcessPoolExecutor with max_workers=os.cpu_count(), core/parallel_utils.py:97-152)."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    workers = min(cores, sets_per_round)
    ctx = mp.get_context("fork")
    with ctx.Pool(workers, initializer=_cpu_init, initargs=(inv, len(test_signal), n_oct)) as pool:
        work = [base_sets[i % len(base_sets)] for i in range(sets_per_round)]
        for _ in range(warm):
            pool.map(_cpu_one_set, work[:workers])
        times = []
        for _ in range(rounds):
            t0 = time.perf_counter()
            pool.map(_cpu_one_set, work)
            times.append(time.perf_counter() - t0)
    return sets_per_round * N_IRS / min(times), workers, times


# ---------------------------------------------------------------------------------------------------------
# CPU arm, kind "reference": the UNMODIFIED reference (oracle/_ref, staged by oracle/make_ref.py) running the same
# per-set path under its own pool, core.parallel_utils.parallel_map (a ProcessPoolExecutor per call with
# max_workers = min(cpu_count, items), core/parallel_utils.py:97-152) — SURVEY.md 8(d).
# ---------------------------------------------------------------------------------------------------------
_REF = {}


def reference_available():
    import ref_import
    return ref_import.available()


def _ref_init(base_sets, est):
    import ref_import
    ref_import.load()
    import core.hrir as ref_hrir
    _REF['est'] = est          # built once by the parent (fork: inherited, not re-derived per worker)
    _REF['base'] = base_sets
    _REF['hrir'] = ref_hrir
    # the WAV container is not part of the path: hand the in-memory float64 tracks to open_recording
    ref_hrir.read_wav = lambda file_path, expand=False: (FS, _REF['base'][int(file_path)])


def _ref_one_set(i):
    """HRIR.open_recording -> crop_heads -> align_ipsilateral_all -> align_onset_groups_peak_leftref -> crop_tails
    (core/pipeline.py:565-601), the reference's own methods."""
    import warnings
    h = _REF['hrir'].HRIR(_REF['est'])
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        h.open_recording(str(i % len(_REF['base'])), LAYOUT)
        h.crop_heads()
        h.align_ipsilateral_all()
        h.align_onset_groups_peak_leftref()
        return h.crop_tails()


def reference_throughput(base_sets, sets_per_round, rounds, warm=1, est=None):
    """IRs/s of the reference's own code under its own process pool; one parallel_map call per round (the pool is
    created inside the call, as the reference does for every stage)."""
    import ref_import
    ref_import.load()
    from core.parallel_utils import parallel_map
    if est is None:
        from core.impulse_response_estimator import ImpulseResponseEstimator as RefEstimator
        est = RefEstimator(min_duration=5.0, fs=FS)
    workers = min(os.cpu_count() or 4, sets_per_round)
    items = list(range(sets_per_round))
    for _ in range(warm):
        parallel_map(_ref_one_set, items[:workers], initializer=_ref_init, initargs=(base_sets, est))
    times = []
    for _ in range(rounds):
        t0 = time.perf_counter()
        tails = parallel_map(_ref_one_set, items, initializer=_ref_init, initargs=(base_sets, est))
        times.append(time.perf_counter() - t0)
    assert len(tails) == sets_per_round and all(int(t) > 0 for t in tails)
    return sets_per_round * N_IRS / min(times), workers, times


WORKLOAD = ("synthetic 7.1 layout (FL,FR,FC,SL,SR,BL,BR x 2 ears) 6.15 s@48 kHz sweep deconvolution + crop "
            "(BASELINE.json configs[1])")


# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t_first_load = None   # set by the caller when the GPU starts doing the benchmark's work
        self.t_timed = (None, None)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        t0, t1 = self.t_timed
        in_timed = 0
        for ts, r in self.rows:
            # samples taken while this process kept the GPU busy with the benchmark workload (warm-up steps
            # run the same work as the timed steps); the short timed region alone can fall between two samples
            if self.t_first_load is not None and ts < self.t_first_load:
                continue
            if t1 is not None and ts > t1 + 0.05:
                continue
            if t0 is not None and t0 <= ts <= t1 + 0.05:
                in_timed += 1
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "samples_in_timed_region": in_timed,
                "window": "warm-up + timed steps (same workload), nvidia-smi every 50 ms"}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------------------
# BASELINE.json configs[2], [3], [4] as extra keys of the JSON line (device resident, one rank's GPU each;
# `value` / `config` stay configs[1]).  Synthetic inputs of the named shapes, CUDA-event timing.
# ---------------------------------------------------------------------------------------------------------
def _time_ms(fn, reps=3, warm=1):
    import torch
    for _ in range(warm):
        out = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.mean(ts)), out


def named_config3(dev, local_rank, peak, sets=8):
    """configs[2]: synthetic 192 kHz 32-bit sweep recordings (7.1 layout) -> deconvolution + crop chain -> resample
    to 48 kHz (1:4 decimation through the FFT plan) -> fractional-octave smoothing -> 9 600-tap minimum-phase FIR."""
    import torch
    from impulcifer_pip313_b200 import ops, synth
    from impulcifer_pip313_b200.dispatcher import BatchDispatcher
    from impulcifer_pip313_b200.estimator import ImpulseResponseEstimator
    from impulcifer_pip313_b200.frequency_response import FrequencyResponse, minimum_phase_batch, smoothen_batch
    from impulcifer_pip313_b200.impulse_response import frequency_response_batch, resample_taps
    fs = 192000
    est = ImpulseResponseEstimator(min_duration=5.0, fs=fs, device=local_rank)
    rec = synth.recording_set(est.test_signal, fs, LAYOUT, seed=7)
    base = torch.from_numpy(rec).to(dev)
    batch = torch.stack([base * (0.6 + 0.05 * i) for i in range(sets)])
    del base
    disp = BatchDispatcher(est, LAYOUT)
    up, down, taps = resample_taps(48000, fs)
    grid = FrequencyResponse.generate_frequencies(f_step=1.01, f_min=10, f_max=24000)
    L = 2 * fs + len(est)

    def chain():
        res = disp.process(batch)
        by_tail = {}
        for r in res:
            by_tail.setdefault(r.tail_ind, []).append(r.data)
        firs = []
        for tail, rows in by_tail.items():
            x = torch.cat(rows)
            y = ops.window_resample(x, taps, up, down)
            raw = frequency_response_batch(y, 48000)
            sm = smoothen_batch(grid, raw, 1 / 3, 1, 1 / 3, 1, 100.0, 10000.0)
            firs.append(minimum_phase_batch(grid, -(sm - sm.mean(dim=1, keepdim=True)), fs=48000, f_res=5,
                                            normalize=False))
        return firs

    ms, firs = _time_ms(chain)
    plan = est.plan(L)
    B = sets * N_IRS
    out = torch.empty((B, L), dtype=torch.float64, device=dev)
    lay = disp._layout
    offs = (torch.arange(sets, device=dev, dtype=torch.int64)[:, None] * (2 * batch.shape[2]) + lay['offs'][None, :]).reshape(-1)
    k1_ms, _ = _time_ms(lambda: plan.execute(batch.reshape(-1), out=out, length=L, row_offsets=offs))
    achieved = B * 16 * L / (k1_ms / 1e3) / 1e9
    return {"workload": "synthetic 192 kHz 32-bit sweep recordings (7.1 layout) -> deconvolution + crop -> resample to "
                        "48 kHz -> smoothen_fractional_octave(1/3) -> minimum_phase (9 600 taps) (BASELINE.json configs[2])",
            "sets": sets, "irs": B, "samples_per_ir": L, "fft_plan": f"{plan.n1}x{plan.n2}", "ms": ms,
            "value": B / (ms / 1e3), "unit": "IR/s", "fir_taps": int(firs[0].shape[1]),
            "roofline": {"bound": "hbm", "kernel": "K1 deconvolution at 192 kHz (half-width tiles)", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak, "algorithmic_bytes_per_ir": 16 * L,
                         "irs_per_launch": B, "ms_per_launch": k1_ms}}


def named_config4(dev, local_rank, peak, sets=8):
    """configs[3]: TrueHD 11-channel layout (22 responses per set) -> deconvolution + crop chain, then per set the
    specific room corrections against a target and the headphone compensation curves, batched over `sets` sets."""
    import torch
    from impulcifer_pip313_b200 import HRIR, ImpulseResponse, synth
    from impulcifer_pip313_b200 import room_correction as rc
    from impulcifer_pip313_b200.dispatcher import BatchDispatcher
    from impulcifer_pip313_b200.estimator import ImpulseResponseEstimator
    from impulcifer_pip313_b200.frequency_response import FrequencyResponse
    est = ImpulseResponseEstimator(min_duration=5.0, fs=FS, device=local_rank)
    layout = synth.LAYOUT_TRUEHD_11
    rec = torch.from_numpy(synth.recording_set(est.test_signal, FS, layout, seed=5)).to(dev)
    batch = torch.stack([rec * (0.6 + 0.05 * i) for i in range(sets)])
    hp = synth.recording_set(est.test_signal, FS, ['FL', 'FR'], seed=6)
    disp = BatchDispatcher(est, layout)
    g48 = FrequencyResponse.generate_frequencies(f_step=1.01, f_min=10, f_max=FS / 2)
    target = FrequencyResponse(name='t', frequency=g48.copy(), raw=-1.2 * (np.log10(g48) - 3.0))

    def chain():
        out = []
        hh = HRIR(est)
        hh.ingest(hp, FS, ['FL', 'FR'])
        left, right = rc.headphone_compensation(hh)
        for res in disp.process(batch):
            h = HRIR(est)
            h.irs = {}
            for name, row in zip(res.names, res.data):
                spk, side = name.split('-')
                ir = ImpulseResponse(np.zeros(0), FS)
                ir._set_device(row)
                h.irs.setdefault(spk, {})[side] = ir
            out.append(rc.calculate_specific_room_corrections(h, target, limit=400))
        return out, left, right

    ms, (frs, left, right) = _time_ms(chain, reps=2)
    n_irs = sets * 2 * len(layout)
    return {"workload": "TrueHD 11-channel layout (22 responses per set) -> deconvolution + crop -> specific room "
                        "corrections + headphone compensation, batched (BASELINE.json configs[3])",
            "sets": sets, "irs": n_irs, "ms": ms, "value": n_irs / (ms / 1e3), "unit": "IR/s",
            "room_curves": int(sum(len(p) for f in frs for p in f.values()))}


def throughput_sweep(dev, disp, base_d, n_samples, world, rank, S, totals=(1024, 4096, 16384, 65536)):
    """configs[4]: `total` synthetic recording sets sharded over the ranks (set s -> rank s mod G), generated ON THE
    DEVICE from seeds in batches of S sets (never staged from the host); only the hot-path calls are timed (CUDA events
    around each process(), summed), max over ranks."""
    import torch
    import torch.distributed as dist
    gen = torch.Generator(device=dev)
    out = []
    buf = torch.empty((S, 2, n_samples), dtype=torch.float64, device=dev)
    for total in totals:
        mine = len(range(rank, total, world))
        gen.manual_seed(977 * total + rank)
        done, ms = 0, 0.0
        while done < mine:
            nb = min(S, mine - done)
            g = 0.5 + 0.5 * torch.rand((nb, 1, 1), generator=gen, device=dev, dtype=torch.float64)
            idx = (torch.arange(nb, device=dev) + done) % base_d.shape[0]
            torch.mul(base_d[idx], g, out=buf[:nb])
            buf[:nb] += 2e-6 * torch.randn((nb, 2, n_samples), generator=gen, device=dev, dtype=torch.float64)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            disp.process(buf[:nb])
            b.record()
            torch.cuda.synchronize()
            ms += a.elapsed_time(b)
            done += nb
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        out.append({"recording_sets": total, "irs": total * N_IRS, "ms": ms, "value": total * N_IRS / (ms / 1e3),
                    "unit": "IR/s"})
    return {"workload": "throughput sweep: 1k-64k synthetic 7.1 recording sets generated on the device from seeds, "
                        "sharded set s -> rank s mod G (BASELINE.json configs[4])", "n_gpus": world,
            "sets_per_launch": S, "points": out}


def run_reference(args, rank, world):
    """Reference arm: rank 0 times the reference's own CPU implementation of the path on the box's host cores.
    Each step is a bounded sample of the GPU arm's workload (same recording sets, same per-set path), sized so
    --steps K ends within minutes; IR/s is per-set work, so the arms compare on the identical workload string."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = max(cores, 4)
    if reference_available():
        import ref_import
        ref_import.load()
        from core.impulse_response_estimator import ImpulseResponseEstimator as RefEstimator
        est = RefEstimator(min_duration=5.0, fs=FS)
        ts = est.test_signal
        base = _host_sets(ts, N_BASE_SETS)
        _, workers, t_all = reference_throughput(base, per_step, rounds=args.steps, warm=min(args.warmup, 1), est=est)
        kind = "reference"
        how = (f"{per_step} recording sets ({per_step * N_IRS} IRs) per step through the unmodified reference "
               f"(oracle/_ref: HRIR.open_recording + crop_heads + align_* + crop_tails) under its own "
               f"core.parallel_utils.parallel_map, ProcessPoolExecutor of {workers} workers created per call")
    else:
        import impulcifer_oracle as O
        ts = O.generate_test_signal(5.0, FS)
        P, _, _ = O.sweep_parameters(FS)
        inv = O.generate_inverse_filter(ts, P)
        base = _host_sets(ts, N_BASE_SETS)
        _, workers, t_all = cpu_throughput(ts, inv, P, base, per_step, rounds=args.steps, warm=min(args.warmup, 1))
        kind = "port"
        how = (f"{per_step} recording sets ({per_step * N_IRS} IRs) per step through the oracle port of the "
               f"reference path (oracle/_ref not staged), fork pool of {workers} processes")
    value = per_step * N_IRS / (sum(t_all) / len(t_all))
    line = {
        "impl": "reference", "metric": "IR deconvolutions/s (6.15s@48kHz sweep)", "value": value, "unit": "IR/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(t_all) / len(t_all), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sets_per_step_sample": per_step, "irs_per_set": N_IRS,
                   "samples_per_ir": 2 * FS + len(ts)},
        "cpu_baseline": {"value": value, "unit": "IR/s", "cores": workers, "kind": kind, "sample": how},
        "e2e": {"value": value, "unit": "IR/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sets", type=int, default=int(os.environ.get("BENCH_SETS", "128")),
                    help="recording sets per GPU per step")
    ap.add_argument("--precision", type=int, default=64, choices=[64, 32])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fp32", action="store_true")
    ap.add_argument("--no-named-configs", action="store_true", help="skip configs[2..4] (192 kHz, TrueHD, sweep)")
    ap.add_argument("--sweep", default="1024,4096,16384,65536", help="total recording sets of the throughput sweep")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    from impulcifer_pip313_b200.estimator import ImpulseResponseEstimator
    est = ImpulseResponseEstimator(min_duration=5.0, fs=FS, device=local_rank, precision=args.precision)
    L = 2 * FS + len(est)
    base = _host_sets(est.test_signal, N_BASE_SETS)

    # ---- CPU baseline first: the fork pool must exist before this process touches CUDA --------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_sets = max(cores, 4)
        if reference_available():
            v, workers, times = reference_throughput(base, n_sets, rounds=2)
            cpu = {"value": v, "unit": "IR/s", "cores": workers, "kind": "reference",
                   "sample": f"{n_sets} recording sets ({n_sets * N_IRS} IRs), best of 2 rounds, the unmodified reference "
                             f"(oracle/_ref) open_recording+crop_heads+align+crop_tails under its own "
                             f"core.parallel_utils.parallel_map ({workers} worker processes)"}
        else:
            v, workers, times = cpu_throughput(est.test_signal, est.inverse_filter, est.n_octaves, base, n_sets, rounds=2)
            cpu = {"value": v, "unit": "IR/s", "cores": workers, "kind": "port",
                   "sample": f"{n_sets} recording sets ({n_sets * N_IRS} IRs), best of 2 rounds, oracle port of "
                             f"open_recording+crop_heads+align+crop_tails, fork pool of {workers} processes"}

    import torch
    import torch.distributed as dist
    from impulcifer_pip313_b200.dispatcher import BatchDispatcher
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- synthetic batch, resident in HBM: seeded variants of the host-synthesised base sets ------------------
    S = args.sets
    n_samples = base[0].shape[1]
    gen = torch.Generator(device=dev)
    gen.manual_seed(4321 + rank)
    base_d = torch.from_numpy(np.stack(base)).to(dev)
    batch = torch.empty((S, 2, n_samples), dtype=torch.float64, device=dev)
    for s in range(S):
        g = 0.5 + 0.5 * torch.rand(1, generator=gen, device=dev, dtype=torch.float64)
        batch[s] = base_d[s % N_BASE_SETS] * g
        batch[s] += 2e-6 * torch.randn((2, n_samples), generator=gen, device=dev, dtype=torch.float64)
    disp = BatchDispatcher(est, LAYOUT)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step():
        return disp.process(batch)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)  # let nvidia-smi come up before the GPU gets busy
    sampler.t_first_load = time.time()
    for _ in range(max(args.warmup, 3)):
        res = step()
    torch.cuda.synchronize()
    # keep the device under the same load until the sampler has seen it for a while
    t_hold = time.time()
    while rank == 0 and time.time() - t_hold < 0.4:
        res = step()
    tails = sorted({r.tail_ind for r in res})
    out_bytes = sum(r.data.numel() * 8 for r in res)
    launches_per_step = disp.launches

    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_a = time.time()
    e0.record()
    for _ in range(args.steps):
        res = step()
    e1.record()
    barrier()
    sampler.t_timed = (t_a, time.time())
    elapsed_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = world * S * N_IRS * args.steps / (elapsed_ms / 1e3)

    # ---- dominant kernel: the K1 launch triple, timed alone with events on its own stream ---------------------
    plan = est.plan(L)
    B = S * N_IRS
    set_base = torch.arange(S, device=dev, dtype=torch.int64)[:, None]
    lay = disp._layout
    row_offs = (set_base * (2 * n_samples) + lay['offs'][None, :]).reshape(-1)
    k1_out = torch.empty((B, L), dtype=torch.float64 if args.precision == 64 else torch.float32, device=dev)
    for _ in range(2):
        plan.execute(batch.reshape(-1), out=k1_out, length=L, row_offsets=row_offs)
    torch.cuda.synchronize()
    k1 = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        plan.execute(batch.reshape(-1), out=k1_out, length=L, row_offsets=row_offs)
        b.record()
        torch.cuda.synchronize()
        k1.append(a.elapsed_time(b))
    k1_ms = float(np.mean(k1))
    del k1_out
    # FP32 fast path of the same kernel (f32 samples in, f32 out; -90 dB parity class), for reference
    fp32 = None
    if args.precision == 64 and not args.no_fp32:
        plan32 = est.plan(L, precision=32)
        x32 = batch.to(torch.float32)
        o32 = torch.empty((B, L), dtype=torch.float32, device=dev)
        for _ in range(2):
            plan32.execute(x32.reshape(-1), out=o32, length=L, row_offsets=row_offs)
        torch.cuda.synchronize()
        t32 = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            plan32.execute(x32.reshape(-1), out=o32, length=L, row_offsets=row_offs)
            b.record()
            torch.cuda.synchronize()
            t32.append(a.elapsed_time(b))
        ms32 = float(np.mean(t32))
        fp32 = {"k1_ms_per_launch": ms32, "k1_ir_per_s": B / (ms32 / 1e3), "algorithmic_bytes_per_ir": 8 * L,
                "achieved_gbs": B * 8 * L / (ms32 / 1e3) / 1e9}
        del x32, o32, plan32
        # the whole hot path (same float64 recordings, FP32 deconvolution, float64 crop chain), 3 steps
        for _ in range(2):
            disp.process(batch, precision=32)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            disp.process(batch, precision=32)
        b.record()
        torch.cuda.synchronize()
        fp32["whole_path_ir_per_s"] = 3 * B / (a.elapsed_time(b) / 1e3)
    # secondary: per-IR frequency-response post-processing on the cropped responses of one step
    def time_post(**kw):
        for _ in range(2):
            post = disp.postprocess(res, **kw)
        torch.cuda.synchronize()
        tp = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            post = disp.postprocess(res, **kw)
            b.record()
            torch.cuda.synchronize()
            tp.append(a.elapsed_time(b))
        return float(np.mean(tp)), int(post[0][3].shape[1])

    fr_ms, fir_taps = time_post()
    eqw_ms, _ = time_post(worker=True)
    fr_post = {"value": S * N_IRS / (fr_ms / 1e3), "unit": "IR/s", "ms_per_step": fr_ms, "fir_taps": fir_taps,
               "what": "frequency_response + smoothen_fractional_octave(1/3) + minimum_phase_impulse_response(f_res=5) "
                       "per cropped IR (BatchDispatcher.postprocess), device resident, one GPU",
               "equalization_worker": {
                   "value": S * N_IRS / (eqw_ms / 1e3), "unit": "IR/s", "ms_per_step": eqw_ms,
                   "what": "frequency_response + the whole process_equalization_worker arithmetic per cropped IR: "
                           "smoothen_heavy_light (4 SG blends) + equalize(max_gain=40) on the device (K21) + "
                           "minimum_phase_impulse_response(f_res=5)"}}
    bytes_per_ir = (8 + (8 if args.precision == 64 else 4)) * L
    peak, peak_src = measured_peak()
    # ---- the other named configs (extra keys): sweep on every rank (it is sharded), configs 3 / 4 on rank 0 at N=1 ----
    extra = {}
    if not args.no_named_configs and args.precision == 64:
        extra["throughput_sweep"] = throughput_sweep(dev, disp, base_d, n_samples, world, rank, S,
                                                     totals=tuple(int(v) for v in args.sweep.split(',')))
        if world == 1:
            extra["config3_192khz"] = named_config3(dev, local_rank, peak)
            extra["config4_truehd"] = named_config4(dev, local_rank, peak)
    del base_d
    achieved = B * bytes_per_ir / (k1_ms / 1e3) / 1e9

    # ---- end to end: pinned host recordings in, cropped responses out, copies inside the timed region -------------
    # Host buffers hold the samples the way a PCM_32 WAV stores them (int32; K1 applies libsndfile's 2^-31 scaling
    # on load, bit-identical to the float64 the reference's reader produces).  Also timed from float64 buffers.
    def e2e_run(host_in):
        host_out = torch.empty(2 * max(out_bytes // 8, 1), dtype=torch.float64).pin_memory()
        for _ in range(2):
            disp.process_host(host_in, out=host_out)
        barrier()
        steps = args.steps  # the same step count as the device-resident `value`
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        for _ in range(steps):
            disp.process_host(host_in, out=host_out)
        b.record()
        wall_ms = (time.perf_counter() - t0) * 1e3  # process_host returns after its D2H stream has drained
        barrier()
        ms = max(a.elapsed_time(b), wall_ms)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        return world * S * N_IRS * steps / (ms / 1e3), steps, int(disp.host_bytes_out), int(host_in.numel() * host_in.element_size())

    # The PCM path is the designed multi-GPU step: ONE shared page-locked input arena holding all
    # world x S recording sets in set order, ONE shared result arena; rank r owns sets s = r (mod world)
    # (DeviceShard), DMAs them out of the input arena and lands their responses in its result slots.  No collective:
    # the step ends with the barrier after which rank 0 reads every set in set order.
    from impulcifer_pip313_b200.arena import HostArena, ResultArena
    from impulcifer_pip313_b200.dispatcher import DeviceShard
    tag = os.environ.get("MASTER_PORT", str(os.getpid()))
    total_sets = world * S
    max_tail = int(max(tails)) if tails else 96000
    max_tail = max(max_tail, 2 * FS)
    in_bytes = total_sets * 2 * n_samples * 4
    out_bytes_arena = ResultArena.nbytes(total_sets, N_IRS, max_tail)
    if rank == 0:
        a_in, a_out = HostArena(f"impb_in_{tag}", in_bytes, True), HostArena(f"impb_out_{tag}", out_bytes_arena, True)
    barrier()
    if rank != 0:
        a_in, a_out = HostArena(f"impb_in_{tag}", in_bytes, False), HostArena(f"impb_out_{tag}", out_bytes_arena, False)
    rec_arena = a_in.view(np.int32, 0, (total_sets, 2, n_samples))
    res_arena = ResultArena(a_out, total_sets, N_IRS, max_tail)
    shard = DeviceShard(world, rank)
    # every rank page-locks the slices it DMAs: the recording sets it owns and their result slots
    set_bytes = 2 * n_samples * 4
    pinned_in = a_in.register_ranges([(s * set_bytes, set_bytes) for s in shard.owned(total_sets)])
    pinned_out = a_out.register_ranges(res_arena.slot_ranges(shard.owned(total_sets)))
    # every rank writes (first-touches) the PCM samples of the sets it owns
    rec_arena[rank::world].copy_(torch.round(batch * 2.0 ** 31).clamp(-2.0 ** 31, 2.0 ** 31 - 1).to(torch.int32).cpu())
    barrier()

    def e2e_sharded():
        for _ in range(2):
            shard.run(disp, rec_arena, res_arena)
        barrier()
        steps = args.steps  # the same step count as the device-resident `value`
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        for _ in range(steps):
            shard.run(disp, rec_arena, res_arena)
        b.record()
        wall_ms = (time.perf_counter() - t0) * 1e3  # run() returns after its D2H stream has drained
        barrier()
        ms = max(a.elapsed_time(b), wall_ms)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        return world * S * N_IRS * steps / (ms / 1e3), steps, int(disp.host_bytes_out), S * 2 * n_samples * 4

    e2e_value, e2e_steps, n_out_bytes, h2d_bytes = e2e_sharded()
    arena_check = None
    if rank == 0:  # the gather is complete and in set order: every slot holds a set of the right shape
        tails_seen = [res_arena.get(s)[0] for s in range(total_sets)]
        arena_check = {"sets_in_arena": total_sets, "all_present": bool(all(t > 0 for t in tails_seen)),
                       "input_arena_bytes": in_bytes, "result_arena_bytes": out_bytes_arena,
                       "page_locked_per_rank_bytes": [int(pinned_in), int(pinned_out)],
                       "owner_of_set": "s mod n_gpus (DeviceShard)"}
    barrier()
    del rec_arena, res_arena
    a_in.close()
    a_out.close()
    host64 = torch.empty((S, 2, n_samples), dtype=torch.float64).pin_memory()
    host64.copy_(batch.cpu())
    e2e64_value, _, _, h2d64_bytes = e2e_run(host64)
    del host64

    if rank == 0:
        line = {
            "metric": "IR deconvolutions/s (6.15s@48kHz sweep)", "value": value, "unit": "IR/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if args.precision == 64 else "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "sets_per_gpu_per_step": S, "irs_per_set": N_IRS, "samples_per_ir": L,
                       "sweep_samples": len(est), "fft_plan": f"{plan.n1}x{plan.n2}",
                       "tail_samples": tails, "l2_policy": "inputs larger than L2 "
                       f"({batch.numel() * 8 / 2**20:.0f} MiB of recordings per step)", "parallelism": f"sets x{world}"},
            "e2e": {"value": e2e_value, "unit": "IR/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": n_out_bytes, "steps": e2e_steps,
                    "api": "DeviceShard.run -> BatchDispatcher.process_host: int32 PCM recordings (PCM_32 WAV samples) in ONE "
                           "shared page-locked input arena, results into ONE shared result arena in set order "
                           "(set s -> rank s mod n_gpus), chunked H2D / compute / D2H overlap, no collective",
                    "result_arena": arena_check,
                    "float64_host_buffers": {"value": e2e64_value, "h2d_bytes_per_step": h2d64_bytes}},
            "gpu_launches": int(launches_per_step * args.steps),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "K1 deconvolution (k_fx_fwd + k_fx_conv + k_fx_inv launch triple)",
                         "algorithmic_bytes_per_ir": bytes_per_ir, "irs_per_launch": B, "ms_per_launch": k1_ms,
                         "peak_source": peak_src,
                         "note": "FP64 FFT convolution is FP64-ALU/LSU bound, not HBM bound (DESIGN.md)"},
            "clocks": clocks,
        }
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))
            line["roofline"]["traffic"] = tr.get("dram_bytes_per_ir", 0) * B
        except Exception:
            pass
        line["fr_postprocess"] = fr_post
        if fp32 is not None:
            fp32["frac_of_peak"] = fp32["achieved_gbs"] / peak
            line["fp32_fast_path"] = fp32
        line.update(extra)
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

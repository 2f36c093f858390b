"""Host-side logic of the shim (no GPU): integer ingest layout, window arithmetic, sharding — against the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

import impulcifer_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_estimator_constants_match_oracle():
    from impulcifer_pip313_b200 import ImpulseResponseEstimator
    for dur, fs in ((0.05, 2000), (1.0, 48000)):
        est = ImpulseResponseEstimator(min_duration=dur, fs=fs)
        ts = O.generate_test_signal(dur, fs)
        P, low, high = O.sweep_parameters(fs)
        assert est.n_octaves == P and est.low == low and est.high == high
        assert np.array_equal(est.test_signal, ts)
        inv = O.generate_inverse_filter(ts, P)
        assert np.max(np.abs(est.inverse_filter - inv)) <= 1e-12 * np.max(np.abs(inv))
        assert len(est) == len(ts) and est.duration == len(ts) / fs
    with pytest.raises(ValueError):
        ImpulseResponseEstimator(min_duration=1.0, fs=44100.5)


@pytest.mark.parametrize("n_tracks,n_samples,speakers,side", [
    (2, 96000 + 7 * 391270, ['FL', 'FR', 'FC', 'SL', 'SR', 'BL', 'BR'], None),
    (1, 878540, ['FL', 'FR'], 'left'),
    (1, 487270, ['FC'], 'right'),
    (2, 487270, ['FC'], None),
    (1, 878540 - 100000, ['FL', 'FR'], 'left'),       # clipped last column
    (2, 96000 + 300000, ['FL'], None),                # short recording: fallbacks
    (2, 96000 + 250000, ['FL'], None),
    (4, 96000 + 2 * 391270, ['FL', 'FR', 'SL', 'SR'], None),
    (2, 96000 + 2 * 391270, ['FL', 'X'], None),       # non-standard speaker name is skipped
])
def test_ingest_layout_matches_oracle(n_tracks, n_samples, speakers, side):
    from impulcifer_pip313_b200.hrir import ingest_layout
    _, cols_o, plan_o = O.ingest_layout(n_tracks, n_samples, speakers, 295270, 48000, side)
    cols, plan = ingest_layout(n_tracks, n_samples, speakers, 295270, 48000, side)
    assert cols == cols_o and plan == plan_o


def test_ingest_layout_errors():
    from impulcifer_pip313_b200.hrir import ingest_layout
    with pytest.raises(ValueError, match="No valid columns"):
        ingest_layout(2, 96000 + 1000, ['FL'], 295270, 48000)
    with pytest.raises(ValueError, match="Silence length"):
        ingest_layout(2, 10 ** 6, ['FL'], 295270, 44100, silence_length=2.00001)


def test_window_arithmetic_and_tables():
    from impulcifer_pip313_b200 import frequency_response as FR
    from impulcifer_pip313_b200 import ops
    f = FR.FrequencyResponse.generate_frequencies(f_min=10, f_max=24000, f_step=1.01)
    assert np.array_equal(f, O.generate_frequencies(10, 24000, 1.01)) and len(f) == 783
    for octv in (1 / 3, 1 / 6, 1 / 5, 1 / 12, 1.3, 1.4, 2):
        assert FR.window_size(f, octv) == O.window_size(f, octv)
    assert np.allclose(FR.sigmoid(f, 100.0, 10000.0), O.sigmoid(f, 100.0, 10000.0), rtol=0, atol=1e-15)
    for t in (1, 7, 97, 21555, 93949, 686539):
        assert FR.next_fast_len(t) == O.next_fast_len_5smooth(t)
    for w in (7, 23, 139):
        c, e = ops.savgol_tables(w)
        assert np.allclose(c, O.savgol_coeffs(w), atol=1e-15) and np.allclose(e, O.savgol_edge_matrix(w), atol=1e-13)


def test_window_size_is_the_reference_loop_and_grids_are_fresh_copies():
    """window_size is vectorised and cached: it must still be the reference's loop (ratios added one by one,
    autoeq/frequency_response.py:1033-1049) on any grid, and a caller that edits a generated grid must not poison the
    cache behind generate_frequencies."""
    import math
    from impulcifer_pip313_b200 import frequency_response as FR

    def loop(frequency, octaves):
        steps = [frequency[i] / frequency[i - 1] for i in range(1, len(frequency))]
        step_size = sum(steps) / len(steps)
        w = round(math.log(2 ** octaves) / math.log(step_size))
        return w if w % 2 else w + 1

    rng = np.random.default_rng(3)
    grids = [FR.FrequencyResponse.generate_frequencies(f_min=10, f_max=fm, f_step=st)
             for fm in (22050, 24000, 96000) for st in (1.01, 1.007, 1.03)]
    grids.append(np.sort(rng.uniform(10, 20000, 400)))
    for f in grids:
        for octv in (1 / 3, 1 / 6, 1 / 12, 1.3, 2):
            assert FR.window_size(f, octv) == loop(f, octv)
            assert FR.window_size(list(f), octv) == loop(f, octv)
    a = FR.FrequencyResponse.generate_frequencies(f_min=10, f_max=24000, f_step=1.01)
    a[:] = 0.0
    b = FR.FrequencyResponse.generate_frequencies(f_min=10, f_max=24000, f_step=1.01)
    assert b[0] == 10 and np.array_equal(b, O.generate_frequencies(10, 24000, 1.01))


def test_resample_taps_match_port():
    from impulcifer_pip313_b200.impulse_response import resample_taps
    for new, old in ((44100, 48000), (48000, 192000), (96000, 48000)):
        up, down, taps = resample_taps(new, old)
        u, d, t = O.resample_filter(new, old)
        assert (up, down) == (u, d) and np.array_equal(taps, t * u)


def test_frequency_response_host_contract():
    from impulcifer_pip313_b200 import FrequencyResponse
    f = FrequencyResponse.generate_frequencies(f_min=10, f_max=24000, f_step=1.01)
    fr = FrequencyResponse(name='t', frequency=f, raw=np.zeros(len(f)), error=0)
    assert len(fr.error) == len(f) and fr.smoothed.size == 0
    with pytest.raises(TypeError):
        FrequencyResponse(name='')
    with pytest.raises(ValueError, match="Upper transition boundary"):
        fr.smoothen_fractional_octave(treble_f_lower=1000, treble_f_upper=1000)
    fr.raw[3] = np.nan
    with pytest.raises(ValueError, match="NaN values present"):
        fr.smoothen_fractional_octave()
    fr2 = FrequencyResponse(name='t', frequency=f, raw=np.zeros(len(f)))
    with pytest.raises(ValueError, match="Error data is missing"):
        fr2.equalize()
    # equalize() runs on the device since round 2 (K21): without one it must fail loudly, never fall back
    import torch
    if not torch.cuda.is_available():
        fr3 = FrequencyResponse(name='t', frequency=f, raw=np.zeros(len(f)), error=np.full(len(f), -10.0))
        with pytest.raises(Exception):
            fr3.equalize(max_gain=6.0, smoothen=False)


def test_device_shard_partition():
    from impulcifer_pip313_b200 import DeviceShard
    n = 37
    seen = []
    for r in range(8):
        sh = DeviceShard(8, r)
        own = sh.owned(n)
        assert all(sh.owner(i) == r for i in own) and sh.count(n) == len(own)
        seen += own
    assert sorted(seen) == list(range(n))
    assert DeviceShard(1, 0).gather({0: 'a', 1: 'b'}, 2) == ['a', 'b']
    with pytest.raises(ValueError):
        DeviceShard(2, 2)


def test_device_shard_gather_two_ranks_gloo(tmp_path):
    """N > 1 path on CPU: two gloo ranks shard 5 recording sets and rank 0 gathers them in set order."""
    script = tmp_path / "shard_worker.py"
    script.write_text(f'''
import os, sys, json
sys.path.insert(0, {ROOT!r})
import torch.distributed as dist
from impulcifer_pip313_b200 import DeviceShard
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
sh = DeviceShard(world, rank)
local = {{i: {{"set": i, "rank": rank, "tail": 1000 + i}} for i in sh.owned(5)}}
out = sh.gather(local, 5)
if rank == 0:
    assert [o["set"] for o in out] == [0, 1, 2, 3, 4]
    assert [o["rank"] for o in out] == [0, 1, 0, 1, 0]
    print("GATHER_OK")
else:
    assert out is None
dist.destroy_process_group()
''')
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29571", str(script)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "GATHER_OK" in r.stdout


def test_device_shard_arena_two_ranks_gloo(tmp_path):
    """The designed N > 1 step on CPU (gloo, world_size 2): one shared input arena, one shared result arena, rank r
    owns sets s = r (mod 2), every set's result lands in its slot in set order with no collective on the data path
    (the GPU compute is replaced by a stand-in dispatcher; the partition, slot and metadata logic is the product's)."""
    script = tmp_path / "arena_worker.py"
    script.write_text(f'''
import os, sys
sys.path.insert(0, {ROOT!r})
import numpy as np, torch
import torch.distributed as dist
from impulcifer_pip313_b200 import DeviceShard
from impulcifer_pip313_b200.arena import HostArena, ResultArena
from impulcifer_pip313_b200.dispatcher import SetResult
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n_sets, n_irs, max_tail, n_samples = 5, 4, 64, 100
in_bytes = n_sets * 2 * n_samples * 4
out_bytes = ResultArena.nbytes(n_sets, n_irs, max_tail)
if rank == 0:
    a_in, a_out = HostArena("impb_test_in", in_bytes, True), HostArena("impb_test_out", out_bytes, True)
dist.barrier()
if rank != 0:
    a_in, a_out = HostArena("impb_test_in", in_bytes, False), HostArena("impb_test_out", out_bytes, False)
rec = a_in.view(np.int32, 0, (n_sets, 2, n_samples))
res = ResultArena(a_out, n_sets, n_irs, max_tail)
sh = DeviceShard(world, rank)
for s in sh.owned(n_sets):
    rec[s] = s + 1          # every rank fills the sets it owns
dist.barrier()

class StandIn:              # process_host of the real dispatcher, without a GPU: tail depends on the set's content
    def process_host(self, recordings, chunk_sets=None, precision=None, out_sets=None):
        out = []
        for k in range(recordings.shape[0]):
            v = int(recordings[k, 0, 0])
            tail = 10 + v
            data = torch.full((n_irs, tail), float(v), dtype=torch.float64)
            out_sets[k, :n_irs * tail].view(n_irs, tail).copy_(data)
            out.append(SetResult([f"S{{i}}-left" for i in range(n_irs)], out_sets[k, :n_irs * tail].view(n_irs, tail), tail,
                                 np.arange(n_irs) + v, np.arange(n_irs), np.arange(n_irs) + 100 * v))
        return out

done = sh.run(StandIn(), rec, res)
assert done == len(sh.owned(n_sets))
dist.barrier()
if rank == 0:
    for s in range(n_sets):
        tail, data, peaks, crop, knees = res.get(s)
        assert tail == 11 + s and tuple(data.shape) == (n_irs, tail)
        assert float(data.min()) == float(data.max()) == s + 1
        assert peaks.tolist() == [s + 1 + i for i in range(n_irs)] and knees.tolist() == [100 * (s + 1) + i for i in range(n_irs)]
    print("ARENA_OK")
dist.barrier()
del rec, res
a_in.close(); a_out.close()
dist.destroy_process_group()
''')
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29572", str(script)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "ARENA_OK" in r.stdout


def test_next_row_host_constants_match_reference_formulas():
    """Host-side pieces of the next-row chains (no device work): the correction-limit mask, the microphone band taper,
    the virtual-bass filter constants and the golden fixtures they were pinned with."""
    from scipy import signal
    from impulcifer_pip313_b200 import room_correction as rc
    from impulcifer_pip313_b200.frequency_response import FrequencyResponse
    from impulcifer_pip313_b200.microphone_deviation import MicrophoneMatchingCorrector
    from impulcifer_pip313_b200.virtual_bass import bass_constants, rbj_high_shelf
    from _util import load
    f = FrequencyResponse.generate_frequencies(f_step=1.01, f_min=10, f_max=24000)
    for limit in (400, 1000, 63.0):
        start, end = np.argmax(f > limit / 2), np.argmax(f > limit)
        want = np.concatenate([np.ones(start), signal.windows.hann(end - start), np.zeros(len(f) - end)])
        np.testing.assert_allclose(rc.correction_limit_mask(f, limit), want, rtol=0, atol=1e-15)
    # the golden mismatch curve is zero outside the taper and clamped inside: same band weight as the reference used
    g = load('room_chain_golden.npz')
    c = MicrophoneMatchingCorrector(48000)
    w = c._band_weight()
    assert np.array_equal(c.frequency, g['mic_frequency'])
    assert np.all(g['mic_auto_mismatch'][w == 0] == 0) and np.all(np.abs(g['mic_auto_mismatch']) <= 12.0)
    assert w[np.searchsorted(c.frequency, 1000)] == 1.0 and w[0] == 0.0 and 0.0 <= w[-1] < 1e-3
    assert (c.win_samples, c.pre_samples) == (240, 24)
    # virtual-bass constants: causal impulse responses of the reference's own filter designs
    h_hp8, mpbass, sos_ild = bass_constants(4096, 48000, 250, 15.0)
    hp4 = signal.butter(4, 250 / 24000, btype="high", output="sos")
    imp = np.zeros(4096)
    imp[0] = 1.0
    assert np.array_equal(h_hp8, signal.sosfilt(np.vstack([hp4, hp4]), imp))
    x = np.random.default_rng(0).standard_normal(4096)
    direct = signal.sosfilt(np.vstack([hp4, hp4]), x)
    via_ir = np.convolve(x, h_hp8)[:4096]
    assert np.max(np.abs(direct - via_ir)) < 1e-12 * np.max(np.abs(direct))   # what virtual_bass.py relies on
    assert sos_ild.shape == (3, 6) and mpbass.shape == (4096,)
    sh = rbj_high_shelf(400.0, 48000, -3.0, 0.66)
    wz, hz = signal.sosfreqz(sh, worN=[1e-3, np.pi * 0.999])
    assert abs(20 * np.log10(abs(hz[0]))) < 1e-3 and abs(20 * np.log10(abs(hz[1])) + 3.0) < 1e-3

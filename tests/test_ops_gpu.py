"""K2..K13 parity: every device entry point against the CPU oracle (and through it the reference goldens)."""
import numpy as np
import pytest

import impulcifer_oracle as O
from _util import MINPHASE_TOL_F5, MINPHASE_TOL_F10, load, rel_peak_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _pack(arrs, T):
    """list of 1-D arrays -> (flat CUDA tensor, offsets, lengths)"""
    lens = np.array([len(a) for a in arrs], dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    flat = np.concatenate([np.asarray(a, np.float64) for a in arrs]) if len(arrs) else np.zeros(0)
    return T.from_numpy(flat).cuda(), offs, lens


def test_peak_index_edge_cases_bit_exact(T):
    from impulcifer_pip313_b200 import ops
    g = load('small_cases.npz')
    keys = [k[3:-2] for k in g.files if k.startswith('pk_') and k.endswith('_x')]
    arrs = [g[f'pk_{k}_x'] for k in keys]
    flat, offs, lens = _pack(arrs, T)
    got = ops.peak_index(flat, offs, lens).cpu().numpy()
    want = [int(g[f'pk_{k}_idx']) for k in keys]
    assert list(got) == want, dict(zip(keys, zip(got, want)))


def test_peak_index_fuzz_bit_exact(T):
    from impulcifer_pip313_b200 import ops
    rng = np.random.default_rng(42)
    arrs = []
    for i in range(200):
        n = int(rng.integers(3, 5000))
        kind = i % 5
        if kind == 0:
            x = rng.standard_normal(n)
        elif kind == 1:  # quantised -> many plateaus and ties
            x = np.round(rng.standard_normal(n) * 3) / 3
        elif kind == 2:  # decaying IR after a delay
            d = int(rng.integers(0, n - 1))
            x = np.zeros(n)
            x[d:] = rng.standard_normal(n - d) * np.exp(-np.arange(n - d) / 50.0)
        elif kind == 3:  # negative dominant
            x = -np.abs(rng.standard_normal(n)) * np.hanning(n)
        else:  # everything below threshold except a ramp
            x = np.linspace(0, 1, n) * rng.choice([-1.0, 1.0])
        arrs.append(x)
    flat, offs, lens = _pack(arrs, T)
    got = ops.peak_index(flat, offs, lens).cpu().numpy()
    want = [O.peak_index(a) for a in arrs]
    assert list(got) == want
    # a different height
    got2 = ops.peak_index(flat, offs, lens, height=0.5).cpu().numpy()
    assert list(got2) == [O.peak_index(a, peak_height=0.5) for a in arrs]


def test_lundeby_degenerate_and_golden(T):
    from impulcifer_pip313_b200 import ops
    g = load('small_cases.npz')
    keys = [k[3:-2] for k in g.files if k.startswith('dc_') and k.endswith('_x')]
    arrs = [g[f'dc_{k}_x'] for k in keys]
    flat, offs, lens = _pack(arrs, T)
    pk, knee, noise, win = [t.cpu().numpy() for t in ops.lundeby(flat, offs, lens, 48000)]
    for i, k in enumerate(keys):
        ref = g[f'dc_{k}_out']
        assert (pk[i], knee[i], win[i]) == (int(ref[0]), int(ref[1]), int(ref[3])), k
        assert abs(noise[i] - ref[2]) < 1e-9, k


def test_lundeby_fuzz(T):
    from impulcifer_pip313_b200 import ops
    rng = np.random.default_rng(9)
    fs = 48000
    arrs = []
    for i in range(48):
        n = int(rng.integers(5000, 150000))
        d = int(rng.integers(10, 4000))
        rt = rng.uniform(0.05, 0.9)
        x = np.zeros(n)
        m = n - d
        x[d:] = rng.standard_normal(m) * 10 ** (-3 * np.arange(m) / (fs * rt))
        x[d] = 2.0
        x += 10 ** rng.uniform(-6, -3) * rng.standard_normal(n)
        arrs.append(x)
    flat, offs, lens = _pack(arrs, T)
    pk, knee, noise, win = [t.cpu().numpy() for t in ops.lundeby(flat, offs, lens, fs)]
    for i, a in enumerate(arrs):
        rp, rk, rn, rw = O.decay_params(a, fs)
        assert (pk[i], knee[i], win[i]) == (rp, rk, rw), i
        assert abs(noise[i] - rn) < 1e-9


def test_lundeby_steepest_possible_decay(T):
    """The refined pass of decay_params builds 3 windows per 10 dB of the first-pass slope.  Window levels are clamped
    to [-200 dB, 0 dB] 30 ms apart, so the slope is at most 200 dB / 30 ms and the list never exceeds 4000 windows
    (k_lundeby holds 4096 and traps beyond).  These rows drive the reference to the steepest slopes it can measure
    (3840 windows at 48 kHz, same at 192 kHz): integers must still be the reference's (golden st_steep*)."""
    from impulcifer_pip313_b200 import ops
    g = load('small_cases.npz')
    for tag, fs in (('steep48', 48000), ('steep48b', 48000), ('steep192', 192000)):
        x = g[f'st_{tag}_x']
        ref = g[f'st_{tag}_out']
        flat, offs, lens = _pack([x], T)
        pk, knee, noise, win = [t.cpu().numpy() for t in ops.lundeby(flat, offs, lens, fs)]
        assert (pk[0], knee[0], win[0]) == (int(ref[0]), int(ref[1]), int(ref[3])), tag
        assert abs(noise[0] - ref[2]) < 1e-9, tag
        assert 2 * fs / win[0] > 2900, (tag, win[0])  # really in the thousands-of-windows regime
        assert O.decay_params(x, fs)[1] == int(ref[1])


def test_view_max_hints_never_change_results(T):
    """impb_view.hint_max / hint_argmax only skip scans: peak_index and lundeby of a view must be identical with
    hints, without hints and on the materialised array (oracle), also when the hinted sample is faded or shifted
    out of the view."""
    from impulcifer_pip313_b200 import ops
    rng = np.random.default_rng(17)
    fs, L, head = 48000, 120000, 48
    fade_in = O.hann(2 * head)[:head]
    rows, crops, shifts = [], [], []
    for i in range(12):
        d = int(rng.integers(300, 3000))
        x = 1e-4 * rng.standard_normal(L)
        x[d:] += rng.standard_normal(L - d) * 10 ** (-3 * np.arange(L - d) / (fs * rng.uniform(0.1, 0.5)))
        x[d] = 3.0 if i % 2 else -3.0
        rows.append(x)
        crops.append(d - int(rng.integers(head, 200)))
        shifts.append(int(rng.integers(-40, 40)))
    # rows whose maximum is NOT usable: inside the faded head, shifted out on the left, cropped away entirely
    crops[3] = int(np.argmax(np.abs(rows[3]))) - 5
    shifts[4] = -(int(np.argmax(np.abs(rows[4]))) - crops[4] + 7)
    crops[5] = int(np.argmax(np.abs(rows[5]))) + 100
    src = np.stack(rows)
    flat = T.from_numpy(src.reshape(-1)).cuda()
    crops, shifts = np.array(crops), np.array(shifts)
    n1 = L - crops
    lo = np.maximum(shifts, 0)
    hi = np.maximum(np.minimum(n1 + shifts, n1), lo)
    offs = np.arange(12) * L + crops
    hint_max = T.from_numpy(np.abs(src).max(axis=1)).cuda()
    hint_arg = T.from_numpy(np.argmax(np.abs(src), axis=1) - crops).cuda()
    base = dict(shift=shifts, valid_lo=lo, valid_hi=hi, fade_in=fade_in)
    hinted = dict(base, hint_max=hint_max, hint_argmax=hint_arg)
    p0 = ops.peak_index(flat, offs, n1, max_len=L, view=base).cpu().numpy()
    p1 = ops.peak_index(flat, offs, n1, max_len=L, view=hinted).cpu().numpy()
    l0 = [t.cpu().numpy() for t in ops.lundeby(flat, offs, n1, fs, max_len=L, view=base)]
    l1 = [t.cpu().numpy() for t in ops.lundeby(flat, offs, n1, fs, max_len=L, view=hinted)]
    assert np.array_equal(p0, p1)
    for a, b in zip(l0, l1):
        assert np.array_equal(a, b)
    for r in range(12):
        d = src[r, crops[r]:].copy()
        d[:head] *= fade_in
        d = O.shift(d, int(shifts[r]))
        assert p1[r] == O.peak_index(d), r
        rp, rk, rn, rw = O.decay_params(d, fs)
        assert (l1[0][r], l1[1][r], l1[3][r]) == (rp, rk, rw), r
        assert abs(l1[2][r] - rn) < 1e-9


def test_xcorr_lag(T):
    from impulcifer_pip313_b200 import ops
    rng = np.random.default_rng(2)
    pairs = []
    for i in range(40):
        n = 1440 if i % 3 else int(rng.integers(20, 1440))
        base = rng.standard_normal(n + 200) * np.exp(-np.arange(n + 200) / 150.0)
        s = int(rng.integers(-60, 60))
        a = base[100:100 + n] + 0.01 * rng.standard_normal(n)
        b = base[100 + s:100 + s + n] + 0.01 * rng.standard_normal(n)
        pairs.append((a, b))
    flat, offs, lens = _pack([p[0] for p in pairs] + [p[1] for p in pairs], T)
    k = len(pairs)
    got = ops.xcorr_lag(flat, offs[:k], lens[:k], offs[k:], lens[k:]).cpu().numpy()
    want = [O.xcorr_lag(a, b) for a, b in pairs]
    assert list(got) == want


def test_gather_window_matches_reference_sequence(T):
    """crop_heads slice + fade-in, two shifts, crop_tails slice + fade-out, gain — bit-exact vs numpy."""
    from impulcifer_pip313_b200 import ops
    rng = np.random.default_rng(3)
    fs = 48000
    head = 48
    fade_in = O.hann(2 * head)[:head]
    fo_len = 1892
    fade_out = O.hann(fo_len)[fo_len // 2:]
    L = 60000
    src = rng.standard_normal((6, L))
    crops = np.array([100, 0, 2500, 31, 17, 999])
    shifts1 = [0, 5, -7, 12, -3, 0]
    shifts2 = [0, -9, 4, 3, -20, 11]
    tail = 21600
    want = []
    los, his, tot = [], [], []
    for r in range(6):
        d = src[r, crops[r]:].copy()
        d[:head] *= fade_in
        n = len(d)
        d = O.shift(d, shifts1[r])
        d = O.shift(d, shifts2[r])
        d = d[:tail].copy()
        d *= np.concatenate([np.ones(len(d) - len(fade_out)), fade_out])
        d *= 0.731
        want.append(d)
        # composite shift and the surviving index range (host-side bookkeeping, as in the shim)
        lo, hi, s = 0, n, 0
        for sft in (shifts1[r], shifts2[r]):
            lo, hi = max(lo + sft, 0), min(hi + sft, n)
            s += sft
        los.append(lo), his.append(min(hi, tail)), tot.append(s)
    flat = T.from_numpy(src.reshape(-1)).cuda()
    starts = np.arange(6) * L + crops
    got = ops.gather_window(flat, starts, tail, shift=tot, valid_lo=los, valid_hi=his, fade_in=fade_in,
                            fade_out=fade_out, gain=0.731).cpu().numpy()
    assert np.array_equal(got, np.stack(want))


@pytest.mark.parametrize("n", [21600, 22500, 57600, 96000, 4096, 1000])
def test_magnitude_db(T, n):
    from impulcifer_pip313_b200 import ops
    rng = np.random.default_rng(n)
    x = rng.standard_normal((3, n)) * np.exp(-np.arange(n) / (n / 8.0))
    got = ops.magnitude_db(T.from_numpy(x).cuda()).cpu().numpy()
    for r in range(3):
        f, m = O.magnitude_response(x[r], 48000)
        assert got[r].shape == m.shape
        assert np.max(np.abs(got[r] - m)) < 1e-8


@pytest.mark.parametrize("n", [391270, 487270, 1009, 2 * 3 * 7 * 11 * 13, 30011])
def test_magnitude_db_lengths_with_large_prime_factors(T, n):
    """Lengths that are not 5-smooth go through the chirp-z (Bluestein) plan: 391 270 is the uncropped 48 kHz response
    the reference hands to frequency_response() in its headphone-compensation stage, 487 270 is data/demo/FC.wav."""
    from impulcifer_pip313_b200 import ImpulseResponse, ops
    rng = np.random.default_rng(n)
    x = rng.standard_normal((2, n)) * np.exp(-np.arange(n) / (n / 8.0))
    x[1, n // 3] = 25.0
    got = ops.magnitude_db(T.from_numpy(x).cuda()).cpu().numpy()
    for r in range(2):
        f, m = O.magnitude_response(x[r], 48000)
        assert got[r].shape == m.shape
        # same absolute-error regime as any FFT: compare linear magnitudes against the spectrum's peak
        lin_g, lin_m = 10 ** (got[r] / 20), 10 ** (m / 20)
        assert np.max(np.abs(lin_g - lin_m)) < 1e-11 * np.max(lin_m)
    if n == 391270:
        fr = ImpulseResponse(x[1].copy(), 48000).frequency_response()
        f_o, raw_o = O.frequency_response(x[1], 48000)
        assert np.array_equal(fr.frequency, f_o) and np.max(np.abs(fr.raw - raw_o)) < 1e-7


@pytest.mark.parametrize("n,step", [(21600, 2), (57600, 5), (84375, 7), (96000, 8), (90000, 8), (1000, 3), (1009, 4),
                                    (2 * 7 * 11 * 13, 2)])
def test_magnitude_db_decimated(T, n, step):
    """The kept bins of frequency_response() (`m[1::step]`, core/impulse_response.py:170-178): the folded transform
    when step divides n (21 600 / 2, 57 600 / 5, 96 000 / 8, 90 000 / 8), the paired transform otherwise (84 375 / 7,
    1000 / 3, and 2002 / 2 whose folded length 1001 is not 5-smooth), the full chirp-z transform for 1009; odd row
    counts and rows 30 dB apart."""
    from impulcifer_pip313_b200 import ops
    rng = np.random.default_rng(n + step)
    x = rng.standard_normal((5, n)) * np.exp(-np.arange(n) / (n / 8.0))
    x[1] *= 0.03
    x[4] *= 7.0
    for rows in (5, 2, 1):
        got = ops.magnitude_db_decimated(T.from_numpy(x[:rows]).cuda(), 1, step).cpu().numpy()
        for r in range(rows):
            _, m = O.magnitude_response(x[r], 48000)
            want = m[1::step]
            assert got[r].shape == want.shape
            assert np.max(np.abs(got[r] - want)) < 1e-8
    full = ops.magnitude_db(T.from_numpy(x).cuda()).cpu().numpy()
    got0 = ops.magnitude_db_decimated(T.from_numpy(x).cuda(), 0, step).cpu().numpy()
    assert got0.shape == full[:, 0::step].shape and np.max(np.abs(got0 - full[:, 0::step])) < 1e-8
    # rows as slices of one flat buffer (the dispatcher's cropped responses): same bits as the matrix form
    flat = np.concatenate([np.zeros(3), x[2], np.zeros(5), x[0], x[4], np.zeros(1)])
    offs = np.array([3 + n + 5, 3, 3 + n + 5 + n], dtype=np.int64)
    sl = ops.magnitude_db_decimated(T.from_numpy(flat).cuda(), 1, step, T.from_numpy(offs).cuda(), n).cpu().numpy()
    mat = ops.magnitude_db_decimated(T.from_numpy(x[[0, 2, 4]]).cuda(), 1, step).cpu().numpy()
    assert np.array_equal(sl, mat)


@pytest.mark.parametrize("normalize", [True, False])
@pytest.mark.parametrize("f_res", [5, 10])
def test_minphase_gain_grid_matches_reference_arithmetic(T, normalize, f_res):
    """The host part of minimum_phase_impulse_response (autoeq/frequency_response.py:653-676) as two launches:
    k=1 spline on log10 f, f <= f_min prefix, [normalize], * 2, 10^(./20), zero at Nyquist."""
    from scipy.fft import next_fast_len
    from scipy.interpolate import InterpolatedUnivariateSpline
    from impulcifer_pip313_b200.frequency_response import minphase_gain_grid
    fs = 48000
    freq = O.generate_frequencies(10, fs / 2, 1.01)
    rng = np.random.default_rng(5 + f_res)
    eq = np.cumsum(rng.standard_normal((37, len(freq))), axis=1) * 0.3
    eq[3] = 0.0
    eq[4] -= 40.0
    got = minphase_gain_grid(freq, T.from_numpy(eq).cuda(), fs, f_res, normalize).cpu().numpy()
    fr2 = f_res / 2
    f_min = np.max([freq[0], fr2])
    n = next_fast_len(round(fs // 2 / fr2))
    f = np.linspace(0.0, fs // 2, n)
    fq = f.copy()
    fq[0] = 0.001
    assert got.shape == (37, n)
    for r in range(37):
        sp = InterpolatedUnivariateSpline(np.log10(freq), eq[r], k=1)
        raw = sp(np.log10(fq))
        raw[f <= f_min] = sp(np.log10(f_min))
        if normalize:
            raw -= np.max(raw)
            raw -= 0.5
        raw *= 2
        raw = 10 ** (raw / 20)
        raw[-1] = 0.0
        assert np.max(np.abs(got[r] - raw) / np.maximum(raw, 1e-300)) < 1e-12


def test_frequency_response_chain_golden(T):
    """magnitude -> decimate -> log-f lerp onto the 783-point grid == reference frequency_response()."""
    from impulcifer_pip313_b200 import ops
    g = load('synth71_golden.npz')
    gold = load('demo_room_golden.npz')
    ir = gold['fc_left']
    fs = 48000
    mag = ops.magnitude_db(T.from_numpy(ir).cuda())
    n = len(ir)
    half = (n + 1) // 2
    f = np.arange(half) * (fs / n)
    step = int(round(len(f) / ((fs / 2) / 4.0))) or 1
    grid = O.generate_frequencies(10, fs / 2, 1.01)
    fd = f[1::step]
    md = mag[:, 1::step].contiguous()
    got = ops.loglerp(np.log10(fd), md, np.log10(grid)).cpu().numpy()[0]
    assert np.array_equal(grid, gold['fc_fr_frequency'])
    assert np.max(np.abs(got - gold['fc_fr_raw'])) < 1e-8
    del g


def test_savgol_blend_golden(T):
    from impulcifer_pip313_b200 import ops
    g = load('synth71_golden.npz')
    f = g['fr_frequency']
    raw = T.from_numpy(g['fr_raw']).cuda()
    k = O.sigmoid(f, 100.0, 10000.0)
    w = O.window_size(f, 1 / 3)
    got = ops.savgol_blend(raw, w, 1, w, 1, k).cpu().numpy()
    assert np.max(np.abs(got - g['fr_smoothed'])) < 1e-9
    # custom: 2-octave x2 normal, 1/12-octave x3 treble, 2k..12k transition
    k2 = O.sigmoid(f, 2000.0, 12000.0)
    got2 = ops.savgol_blend(raw, O.window_size(f, 2), 2, O.window_size(f, 1 / 12), 3, k2).cpu().numpy()
    assert np.max(np.abs(got2 - g['fr_custom'])) < 1e-9
    # window larger than the curve is rejected like scipy does
    with pytest.raises(ValueError):
        ops.savgol_blend(raw[:, :50].contiguous(), 99, 1, 7, 1, k[:50])


def test_minphase_fir_golden(T):
    from impulcifer_pip313_b200 import ops
    g = load('synth71_golden.npz')
    f = g['fr_frequency']
    fs = 48000
    gains5, gains10 = [], []
    for i in range(len(g['fr_names'])):
        eq = -(g['fr_smoothed'][i] - np.mean(g['fr_smoothed'][i]))
        gains5.append(O.minphase_grid(f, eq, fs, 5, False)[1])
        gains10.append(O.minphase_grid(f, eq, fs, 10, True)[1])
    got5 = ops.minphase_fir(T.from_numpy(np.stack(gains5)).cuda(), fs).cpu().numpy()
    got10 = ops.minphase_fir(T.from_numpy(np.stack(gains10)).cuda(), fs).cpu().numpy()
    assert got5.shape == (4, 9600) and got10.shape == (4, 4800)
    for i in range(4):
        # gates: tests/_util.py MINPHASE_TOL_* (the reference's own reproducibility floor, measured)
        assert rel_peak_err(got5[i], g['minphase_f5'][i]) < MINPHASE_TOL_F5
        assert rel_peak_err(got10[i], g['minphase_f10'][i]) < MINPHASE_TOL_F10
        # magnitude response of the FIR (what the EQ stage actually relies on) is tight
        H_ref = np.abs(np.fft.rfft(g['minphase_f5'][i]))
        H_got = np.abs(np.fft.rfft(got5[i]))
        # ... everywhere except the Nyquist end, where the reference itself takes the log of rounding noise
        assert np.max(np.abs(H_got - H_ref)[:-8]) / np.max(H_ref) < 1e-7
        assert np.max(np.abs(H_got - H_ref)) / np.max(H_ref) < 2e-5


def test_resample_matches_port(T):
    from impulcifer_pip313_b200 import ops
    g = load('synth71_golden.npz')
    ir = g['final_FL_left']
    up, down, taps = O.resample_filter(44100, 48000)
    got = ops.window_resample(T.from_numpy(ir).cuda(), taps * up, up, down).cpu().numpy()[0]
    ref = g['resampled_FL_left_44100']
    assert got.shape == ref.shape
    assert rel_peak_err(got, ref) < 1e-12
    # 4:1 decimation (192 kHz -> 48 kHz config) on a short signal
    x = np.random.default_rng(0).standard_normal(20000)
    up, down, taps = O.resample_filter(48000, 192000)
    got = ops.window_resample(T.from_numpy(x).cuda(), taps * up, up, down).cpu().numpy()[0]
    ref = O.resample(x, 48000, 192000)
    assert got.shape == ref.shape and rel_peak_err(got, ref) < 1e-12
    # decay window fused in front
    params = (100.0, 3000.0, 3100.0, -12.0)
    ref = O.resample(O.apply_decay_window(x.copy(), (100, 3000, 3100, -12.0)), 44100, 48000)
    up, down, taps = O.resample_filter(44100, 48000)
    got = ops.window_resample(T.from_numpy(x).cuda(), taps * up, up, down, decay=np.array(params)).cpu().numpy()[0]
    assert rel_peak_err(got, ref) < 1e-12


@pytest.mark.parametrize("fs_new,fs_old", [(44100, 48000), (48000, 44100), (88200, 96000), (32000, 48000), (96000, 48000)])
def test_resample_batched_ratios_match_port(T, fs_new, fs_old):
    """K9 on a ragged batch for the ratios the pipeline meets: 147/160 and 160/147 take the run-structured kernel
    (k_resample_runs: 12- / 14-output runs), 2/3 and 2/1 the generic kernels; every row against the nnresample port,
    including the first / last outputs whose windows hang over the ends of the row, with and without decay windows."""
    from impulcifer_pip313_b200 import ops
    rng = np.random.default_rng(fs_new + fs_old)
    rows, n = 5, 20011
    x = rng.standard_normal((rows, n))
    x[1, :300] = 0.0
    x[2, -5000:] *= 1e-3
    up, down, taps = O.resample_filter(fs_new, fs_old)
    got = ops.window_resample(T.from_numpy(x).cuda(), taps * up, up, down).cpu().numpy()
    for r in range(rows):
        ref = O.resample(x[r], fs_new, fs_old)
        assert got[r].shape == ref.shape
        assert rel_peak_err(got[r], ref) < 1e-12, r
    # device-resident taps take impb_window_resample (tap table in shared memory / generic kernels): every output is
    # the same single FMA chain in ascending tap order, so the two entry points must agree bit for bit
    got_dev = ops.window_resample(T.from_numpy(x).cuda(), T.from_numpy(taps * up).cuda(), up, down).cpu().numpy()
    assert np.array_equal(got, got_dev)
    prm = np.array([[100.0, 3000.0, 3100.0, -12.0], [np.nan] * 4, [0.0, 9000.0, 9000.0, -30.0], [50.0, 4.0, 54.0, -6.0],
                    [7.0, 19000.0, 19007.0, -20.0]])
    got = ops.window_resample(T.from_numpy(x).cuda(), taps * up, up, down, decay=prm).cpu().numpy()
    for r in range(rows):
        xr = x[r].copy() if np.isnan(prm[r, 0]) else O.apply_decay_window(
            x[r].copy(), (int(prm[r, 0]), int(prm[r, 1]), int(prm[r, 2]), float(prm[r, 3])))
        assert rel_peak_err(got[r], O.resample(xr, fs_new, fs_old)) < 1e-12, r


def test_equalize_matches_reference(T):
    """K21 against the reference's own FrequencyResponse.equalize (tests/golden/make_golden.py::equalize_cases):
    unclipped curves, one / many clipped regions, clipping from the first and up to the last point, the EQ worker's
    arguments, smoothen=False and a treble-limited gain; object API and batched API."""
    from impulcifer_pip313_b200.frequency_response import FrequencyResponse, equalize_batch
    g = load('equalize_golden.npz')
    f, err = g['frequency'], g['error_smoothed']
    cases = (('default', {}), ('worker', dict(max_gain=40, treble_f_lower=10000, treble_f_upper=24000.0)),
             ('nosmooth', dict(smoothen=False)), ('treble', dict(treble_max_gain=0.0, treble_gain_k=0.5)))
    for tag, kw in cases:
        got = equalize_batch(f, err, **kw)
        assert got.shape == g[f'eq_{tag}'].shape
        assert np.max(np.abs(got - g[f'eq_{tag}'])) < 1e-9, tag
        dev = equalize_batch(f, T.from_numpy(err).cuda(), **kw)
        assert np.array_equal(dev.cpu().numpy(), got)
    fr = FrequencyResponse(name='eq', frequency=f.copy(), raw=np.zeros(len(f)), error=err[1].copy(),
                           error_smoothed=err[1].copy(), smoothed=np.ones(len(f)))
    fr.equalize()
    assert np.max(np.abs(fr.equalization - g['eq_default'][1])) < 1e-9
    assert np.allclose(fr.equalized_raw, fr.equalization) and np.allclose(fr.equalized_smoothed, 1.0 + fr.equalization)
    with pytest.raises(ValueError, match='NaN values detected'):
        bad = err[0].copy()
        bad[5] = np.nan
        FrequencyResponse(name='eq', frequency=f.copy(), raw=np.zeros(len(f)), error=bad).equalize()
    with pytest.raises(ValueError, match='Error data is missing'):
        FrequencyResponse(name='eq', frequency=f.copy(), raw=np.zeros(len(f))).equalize()
    # clamp / invert behaviour of the reference's tests/test_frequency_response_core.py:73-96
    fr3 = FrequencyResponse(name='t', frequency=f, raw=np.zeros(len(f)), error=np.full(len(f), -10.0))
    fr3.equalize(max_gain=6.0, smoothen=False)
    assert np.allclose(fr3.equalization, 6.0, atol=1e-9)
    fr4 = FrequencyResponse(name='t', frequency=f, raw=np.zeros(len(f)), error=np.full(len(f), 3.0))
    fr4.equalize(max_gain=6.0)
    assert np.allclose(fr4.equalization, -3.0, atol=1e-9) and np.allclose(fr4.equalized_raw, -3.0, atol=1e-9)


def test_set_bookkeeping_kernels_match_host_arithmetic(T):
    """K22 (impb_set_heads / _shift / _shift2 / _tails) against the reference's per-set integer arithmetic written out
    in numpy (core/hrir.py:573-606, 936-955, 981-1001, 627-646) on random integers for the 7.1 layout, including the
    flag bits (response shorter than the fade-in, tail shorter than the fade-out, tail above the preallocated bound)."""
    from scipy.fftpack import next_fast_len  # 5-smooth, what core/hrir.py:640 calls (scipy.fft's is 11-smooth)
    from impulcifer_pip313_b200 import ops
    from impulcifer_pip313_b200.dispatcher import BatchDispatcher
    from impulcifer_pip313_b200.estimator import ImpulseResponseEstimator
    est = ImpulseResponseEstimator(min_duration=0.5, fs=48000)
    disp = BatchDispatcher(est, ['FL', 'FR', 'FC', 'SL', 'SR', 'BL', 'BR'])
    L = 2 * 48000 + len(est)
    n_samples = 96000 + 7 * L
    lay = disp._prepare(2, n_samples, T.device('cuda'))
    ld, n = lay['dev'], lay['n']
    S, seg = 9, 1440
    rng = np.random.default_rng(3)
    peaks = rng.integers(0, L - 5000, size=(S, n))
    peaks[0, :2] = 3            # crop clamps at 0
    e_arg = rng.integers(0, L, size=(S, n))
    left, right = np.array(lay['left_h']), np.array(lay['right_h'])
    delay = lay['delay'].cpu().numpy()
    crop_sp = np.maximum(np.minimum(peaks[:, left], peaks[:, right]) - delay[None, :], 0)
    crop = np.zeros((S, n), dtype=np.int64)
    crop[:, left] = crop_sp
    crop[:, right] = crop_sp
    n1 = L - crop
    rows = np.arange(S * n).reshape(S, n)
    off1 = rows * L + crop
    dv = lambda a: T.from_numpy(np.ascontiguousarray(a, dtype=np.int64)).cuda()
    g_crop, g_n1, g_off1, g_hint, xs = ops.set_heads(ld, dv(peaks.reshape(-1)), dv(e_arg.reshape(-1)), L, seg, S)
    assert np.array_equal(g_crop.cpu().numpy().reshape(S, n), crop)
    assert np.array_equal(g_n1.cpu().numpy().reshape(S, n), n1)
    assert np.array_equal(g_off1.cpu().numpy().reshape(S, n), off1)
    assert np.array_equal(g_hint.cpu().numpy().reshape(S, n), e_arg - crop)
    pa, pb = lay['pa'].cpu().numpy(), lay['pb'].cpu().numpy()
    assert np.array_equal(xs[0].cpu().numpy().reshape(S, -1), off1[:, pa])
    assert np.array_equal(xs[1].cpu().numpy().reshape(S, -1), np.minimum(n1[:, pa], seg))
    assert np.array_equal(xs[2].cpu().numpy().reshape(S, -1), off1[:, pb])
    assert np.array_equal(xs[3].cpu().numpy().reshape(S, -1), np.minimum(n1[:, pb], seg))
    # lags -> shift, surviving range, gathered group-leader views
    lags = rng.integers(-40, 40, size=(S, len(pa)))
    shift = np.zeros((S, n), dtype=np.int64)
    for r, k in zip(lay['pos_rows'].cpu().numpy(), lay['pos_k'].cpu().numpy()):
        shift[:, r] += np.maximum(lags[:, k], 0)
    for r, k in zip(lay['neg_rows'].cpu().numpy(), lay['neg_k'].cpu().numpy()):
        shift[:, r] += np.maximum(-lags[:, k], 0)
    lo = np.maximum(shift, 0)
    hi = np.maximum(np.minimum(n1 + shift, n1), lo)
    e_max = T.rand(S * n, dtype=T.float64, device='cuda')
    g_shift, g_lo, g_hi, gi, gm = ops.set_shift(ld, dv(lags.reshape(-1)), g_n1, g_off1, g_hint, e_max, S, True)
    assert np.array_equal(g_shift.cpu().numpy().reshape(S, n), shift)
    assert np.array_equal(g_lo.cpu().numpy().reshape(S, n), lo) and np.array_equal(g_hi.cpu().numpy().reshape(S, n), hi)
    gl = lay['gl'].cpu().numpy()
    for i, want in enumerate((off1[:, gl], n1[:, gl], shift[:, gl], lo[:, gl], hi[:, gl], (e_arg - crop)[:, gl])):
        assert np.array_equal(gi[i].cpu().numpy().reshape(S, -1), want), i
    assert np.array_equal(gm.cpu().numpy().reshape(S, -1), e_max.cpu().numpy().reshape(S, n)[:, gl])
    # group onset peaks -> composite shift
    gp = rng.integers(0, 300, size=(S, len(gl)))
    d = gp - gp[:, lay['g_ref']:lay['g_ref'] + 1]
    shift2 = np.zeros((S, n), dtype=np.int64)
    for r, g in zip(lay['g_rows'].cpu().numpy(), lay['g_of_row'].cpu().numpy()):
        shift2[:, r] += -d[:, g]
    lo2 = np.maximum(lo + shift2, 0)
    hi2 = np.maximum(np.minimum(hi + shift2, n1), lo2)
    ops.set_shift2(ld, dv(gp.reshape(-1)), g_n1, S, g_shift, g_lo, g_hi)
    assert np.array_equal(g_shift.cpu().numpy().reshape(S, n), shift + shift2)
    assert np.array_equal(g_lo.cpu().numpy().reshape(S, n), lo2) and np.array_equal(g_hi.cpu().numpy().reshape(S, n), hi2)
    # knees -> per-set tails and the ragged geometry
    knee = rng.integers(3000, 90000, size=(S, n))
    tails = np.minimum(n1.min(axis=1), np.array([next_fast_len(int(k)) for k in knee.max(axis=1)]))
    t_d, off_d, len_d, dst_d, flags = ops.set_tails(ld, dv(knee.reshape(-1)), g_n1, S, 48, 1892, 104976, L)
    assert int(flags.item()) == 0
    assert np.array_equal(t_d.cpu().numpy(), tails)
    set_off = np.concatenate([[0], np.cumsum(tails * n)])
    assert np.array_equal(off_d.cpu().numpy(), set_off)
    assert np.array_equal(len_d.cpu().numpy().reshape(S, n), np.repeat(tails[:, None], n, axis=1))
    assert np.array_equal(dst_d.cpu().numpy().reshape(S, n), set_off[:-1, None] + np.arange(n)[None, :] * tails[:, None])
    # flags: fade-out longer than a tail (2), tail above the bound (4), response shorter than the fade-in (1)
    assert int(ops.set_tails(ld, dv(knee.reshape(-1)), g_n1, S, 48, 200000, 104976, L)[4].item()) & 2
    assert int(ops.set_tails(ld, dv(knee.reshape(-1)), g_n1, S, 48, 1892, 4096, L)[4].item()) & 4
    assert int(ops.set_tails(ld, dv(knee.reshape(-1)), g_n1, S, int(n1.min()) + 1, 1892, 104976, L)[4].item()) & 1

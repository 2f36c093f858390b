"""Pin the CPU oracle (oracle/impulcifer_oracle.py) against vectors recorded from the UNMODIFIED reference.

The fixtures under tests/golden/ were produced by tests/golden/make_golden.py, which imports
115dkk/Impulcifer-pip313 itself.  ``room_responses`` additionally is the reference's own committed golden
file (data/demo/room-responses.wav), compared bit-exactly at int32 PCM.
"""
import numpy as np
import pytest

import impulcifer_oracle as O
from impulcifer_pip313_b200 import synth
from _util import MINPHASE_TOL_F5, MINPHASE_TOL_F10, assert_ck, load, rel_peak_err


@pytest.fixture(scope="module")
def demo():
    return load('demo_room_inputs.npz'), load('demo_room_golden.npz')


@pytest.fixture(scope="module")
def bundled_estimator(demo):
    inp, _ = demo
    sweep = inp['sweep'].astype(np.float64) / 2 ** 31
    P, _, _ = O.sweep_parameters(48000)
    # from_wav keeps the generated signal when it matches the WAV within 1e-4
    # (core/impulse_response_estimator.py:234-262)
    ts = O.generate_test_signal((len(sweep) - 1) / 48000, 48000)
    assert len(ts) == len(sweep) == 295270
    assert np.max(np.abs(ts - sweep)) < 1e-4
    return ts, O.generate_inverse_filter(ts, P), P


ROOM_FILES = [('room-FL,FR-left.wav', ['FL', 'FR'], 'left'), ('room-FL,FR-right.wav', ['FL', 'FR'], 'right'),
              ('room-FC-left.wav', ['FC'], 'left'), ('room-FC-right.wav', ['FC'], 'right'),
              ('room-BL,SL-left.wav', ['BL', 'SL'], 'left'), ('room-BL,SL-right.wav', ['BL', 'SL'], 'right'),
              ('room-SR,BR-left.wav', ['SR', 'BR'], 'left'), ('room-SR,BR-right.wav', ['SR', 'BR'], 'right')]


def test_room_responses_bit_exact(demo, bundled_estimator):
    """estimate -> peak_index -> crop_head -> decay_params -> crop_tails -> PCM_32 == room-responses.wav."""
    inp, gold = demo
    ts, inv, P = bundled_estimator
    irs = {}
    for name, speakers, side in ROOM_FILES:
        rec = inp[name].astype(np.float64)[None, :] / 2 ** 31
        for spk, pair in O.open_recording(rec, 48000, speakers, inv, 48000, side=side).items():
            irs.setdefault(spk, {}).update(pair)
    names = [f'{s}-{e}' for s, pair in irs.items() for e in pair]
    assert names == list(gold['names'])
    peaks, knees = [], []
    for spk, pair in irs.items():
        for side in pair:
            assert_ck(pair[side], gold[f'est_ck_{spk}-{side}'], what=f'{spk}-{side}')
            peaks.append(O.peak_index(pair[side]))
            pair[side] = O.crop_head(pair[side], 48000)
            pk, knee, noise, w = O.decay_params(pair[side], 48000)
            knees.append(knee)
    assert peaks == list(gold['peaks'])
    assert knees == list(gold['knees'])
    tail = O.crop_tails(irs, 48000, len(ts), P, 48000)
    assert tail == int(gold['tail']) == 21600
    q = O.pcm32(O.track_matrix(irs))
    assert q.shape == gold['room_responses'].shape == (32, 21600)
    assert np.array_equal(q, gold['room_responses'])


def test_fc_binaural_chain(demo, bundled_estimator):
    inp, gold = demo
    ts, inv, P = bundled_estimator
    rec = inp['FC.wav'].astype(np.float64) / 2 ** 31
    irs = O.open_recording(rec, 48000, ['FC'], inv, 48000)
    assert [O.peak_index(irs['FC'][s]) for s in ('left', 'right')] == list(gold['fc_peaks']) == [172160, 172160]
    O.crop_heads(irs, 48000)
    assert [len(irs['FC'][s]) for s in ('left', 'right')] == list(gold['fc_len_after_heads'])
    O.align_ipsilateral_all(irs, 48000)
    assert O.crop_tails(irs, 48000, len(ts), P, 48000) == int(gold['fc_tail'])
    assert rel_peak_err(irs['FC']['left'], gold['fc_left']) < 1e-12
    assert rel_peak_err(irs['FC']['right'], gold['fc_right']) < 1e-12
    f, raw = O.frequency_response(irs['FC']['left'], 48000)
    assert np.array_equal(f, gold['fc_fr_frequency']) and len(f) == 783
    assert np.max(np.abs(raw - gold['fc_fr_raw'])) < 1e-9
    sm = O.smoothen_fractional_octave(f, raw, 1 / 3)
    assert np.max(np.abs(sm - gold['fc_fr_smoothed'])) < 1e-9


@pytest.fixture(scope="module")
def synth_set():
    g = load('synth71_golden.npz')
    ts = O.generate_test_signal(5.0, 48000)
    P, _, _ = O.sweep_parameters(48000)
    inv = O.generate_inverse_filter(ts, P)
    rec = synth.recording_set(ts, 48000, synth.LAYOUT_71, seed=1234)
    return g, ts, inv, P, rec


def test_synth71_estimator_and_inputs(synth_set):
    g, ts, inv, P, rec = synth_set
    assert len(ts) == int(g['sweep_len']) == 295270 and P == float(g['n_octaves'])
    assert_ck(ts, g['sweep_ck'], rel=1e-12)
    assert_ck(inv, g['inv_ck'], rel=1e-12)
    assert_ck(rec[0], g['rec_ck'][0], rel=1e-12)
    assert_ck(rec[1], g['rec_ck'][1], rel=1e-12)


def test_synth71_full_chain(synth_set):
    g, ts, inv, P, rec = synth_set
    fs = 48000
    irs = O.open_recording(rec, fs, synth.LAYOUT_71, inv, fs)
    order = [(s, e) for s in irs for e in ('left', 'right')]
    assert [f'{s}-{e}' for s, e in order] == list(g['names'])
    peaks = [O.peak_index(irs[s][e]) for s, e in order]
    assert peaks == list(g['est_peaks'])
    for i, (s, e) in enumerate(order):
        assert_ck(irs[s][e], g['est_ck'][i], what=f'{s}-{e}')
        sl = irs[s][e][peaks[i] - 64:peaks[i] + 1984]
        assert rel_peak_err(sl, g['est_slices'][i]) < 1e-12
    dp = [O.decay_params(irs[s][e], fs) for s, e in order]
    assert [d[1] for d in dp] == list(g['est_knees'])
    assert [d[3] for d in dp] == list(g['est_wins'])
    assert np.allclose([d[2] for d in dp], g['est_noise'], rtol=0, atol=1e-9)
    O.crop_heads(irs, fs)
    assert [len(irs[s][e]) for s, e in order] == list(g['len_after_heads'])
    O.align_ipsilateral_all(irs, fs)
    for i, (s, e) in enumerate(order):
        assert_ck(irs[s][e], g['ck_after_ipsi'][i], what=f'ipsi {s}-{e}')
    O.align_onset_groups_peak_leftref(irs)
    assert [O.peak_index(irs[s][e]) for s, e in order] == list(g['peaks_after_align'])
    assert [O.decay_params(irs[s][e], fs)[1] for s, e in order] == list(g['knees_after_align'])
    assert O.crop_tails(irs, fs, len(ts), P, fs) == int(g['tail'])
    for i, (s, e) in enumerate(order):
        assert_ck(irs[s][e], g['final_ck'][i], what=f'final {s}-{e}')
    assert rel_peak_err(irs['FL']['left'], g['final_FL_left']) < 1e-12
    assert rel_peak_err(irs['BR']['left'], g['final_BR_left']) < 1e-12

    # FR post-processing
    for i, nm in enumerate(g['fr_names']):
        s, e = str(nm).split('-')
        f, raw = O.frequency_response(irs[s][e], fs)
        assert np.array_equal(f, g['fr_frequency'])
        assert np.max(np.abs(raw - g['fr_raw'][i])) < 1e-9
        raw = g['fr_raw'][i]
        sm = O.smoothen_fractional_octave(f, raw, 1 / 3)
        assert np.max(np.abs(sm - g['fr_smoothed'][i])) < 1e-10
        hs, he = O.smoothen_heavy_light(f, raw, raw - np.mean(raw))
        assert np.max(np.abs(hs - g['fr_hl_smoothed'][i])) < 1e-10
        assert np.max(np.abs(he - g['fr_hl_error'][i])) < 1e-10
        cu = O.smoothen_fractional_octave(f, raw, 2, 2, 1 / 12, 3, 2000, 12000)
        assert np.max(np.abs(cu - g['fr_custom'][i])) < 1e-10
        eq = -(g['fr_smoothed'][i] - np.mean(g['fr_smoothed'][i]))
        # the homomorphic transform is ill-conditioned at the type-II Nyquist null (log of rounding noise):
        # two correct FFT libraries already differ by ~1e-7 of peak here, see DESIGN.md "min-phase tolerance"
        mp5 = O.minimum_phase_impulse_response(f, eq, fs=fs, f_res=5, normalize=False)
        assert mp5.shape == g['minphase_f5'][i].shape == (9600,)
        assert rel_peak_err(mp5, g['minphase_f5'][i]) < MINPHASE_TOL_F5
        mp10 = O.minimum_phase_impulse_response(f, eq, fs=fs, f_res=10, normalize=True)
        assert mp10.shape == (4800,)
        assert rel_peak_err(mp10, g['minphase_f10'][i]) < MINPHASE_TOL_F10
    rs = O.resample(irs['FL']['left'], 44100, 48000)
    assert len(rs) == len(g['resampled_FL_left_44100']) == int(np.ceil(int(g['tail']) * 147 / 160))
    assert rel_peak_err(rs, g['resampled_FL_left_44100']) < 1e-12


def test_small_estimators():
    g = load('small_cases.npz')
    for tag, (dur, fs) in {'fs2000': (0.05, 2000), 'fs48000_1s': (1.0, 48000), 'fs44100_2s': (2.0, 44100)}.items():
        ts = O.generate_test_signal(dur, fs)
        P, _, _ = O.sweep_parameters(fs)
        inv = O.generate_inverse_filter(ts, P)
        assert len(ts) == int(g[f'{tag}_len']) and P == float(g[f'{tag}_n_octaves'])
        assert_ck(ts, g[f'{tag}_test_signal_ck'], rel=1e-12)
        assert_ck(inv, g[f'{tag}_inverse_filter_ck'], rel=1e-11)
        for delay in (0, 1, 127, 1000):
            if delay >= fs:
                continue
            rec = np.concatenate([np.zeros(delay), ts, np.zeros(int(0.2 * fs))])
            y = O.estimate(rec, inv)
            assert len(y) == len(rec)
            assert_ck(y, g[f'{tag}_d{delay}_est_ck'], what=f'{tag} d{delay}')
            assert O.peak_index(y) == int(g[f'{tag}_d{delay}_peak'])
            if f'{tag}_d{delay}_est' in g:
                assert rel_peak_err(y, g[f'{tag}_d{delay}_est']) < 1e-12
        noise = np.random.default_rng(11).standard_normal(len(ts) + fs // 2) * 0.1
        assert_ck(O.estimate(noise, inv), g[f'{tag}_noise_est_ck'], what=f'{tag} noise')


def test_peak_index_edge_cases():
    g = load('small_cases.npz')
    for key in [k[3:-2] for k in g.files if k.startswith('pk_') and k.endswith('_x')]:
        assert O.peak_index(g[f'pk_{key}_x']) == int(g[f'pk_{key}_idx']), key
    assert O.peak_index(np.zeros(0)) == 0


def test_decay_params_degenerate():
    g = load('small_cases.npz')
    for key in [k[3:-2] for k in g.files if k.startswith('dc_') and k.endswith('_x')]:
        pk, knee, noise, w = O.decay_params(g[f'dc_{key}_x'], 48000)
        ref = g[f'dc_{key}_out']
        assert (pk, knee, w) == (int(ref[0]), int(ref[1]), int(ref[3])), key
        assert abs(noise - ref[2]) < 1e-9, key


def test_ingest_layout_integer_logic():
    # 7.1 layout: 2 tracks, 7 speakers -> 7 columns, left = track 0, right = track 1
    L, fs = 295270, 48000
    total = 96000 + 7 * (96000 + L)
    sil, cols, plan = O.ingest_layout(2, total, synth.LAYOUT_71, L, fs)
    assert sil == 96000 and len(cols) == 7
    assert cols[0] == (96000, 96000 + 391270) and cols[6][1] == total
    assert plan[:2] == [('FL', 'left', 0, 0), ('FL', 'right', 1, 0)] and len(plan) == 14
    # single-side room file, two speakers on one track
    sil, cols, plan = O.ingest_layout(1, 878540, ['FL', 'FR'], L, fs, side='left')
    assert len(cols) == 2 and plan == [('FL', 'left', 0, 0), ('FR', 'left', 0, 1)]
    # clipped last column is kept only if it still holds a full sweep
    sil, cols, plan = O.ingest_layout(1, 878540 - 100000, ['FL', 'FR'], L, fs, side='left')
    assert len(cols) == 1
    with pytest.raises(ValueError):
        O.ingest_layout(2, 96000 + 1000, ['FL'], L, fs)
    with pytest.raises(ValueError):
        O.ingest_layout(2, total, synth.LAYOUT_71, L, 44100, silence_length=2.00001)


def test_next_fast_len_matches_scipy():
    from scipy import fftpack
    for t in [1, 2, 7, 11, 97, 1000, 21555, 686539, 93949, 100000, 5 ** 7 + 1]:
        assert O.next_fast_len_5smooth(t) == fftpack.next_fast_len(t), t


def test_savgol_matches_scipy():
    from scipy.signal import savgol_filter
    rng = np.random.default_rng(3)
    y = np.cumsum(rng.standard_normal(783))
    for w in (7, 13, 15, 23, 91, 99, 139):
        assert np.max(np.abs(O.savgol_filter(y, w, 2) - savgol_filter(y, w, 2))) < 1e-10
    f = O.generate_frequencies(10, 24000, 1.01)
    assert len(f) == 783
    assert [O.window_size(f, o) for o in (1 / 3, 1 / 6, 1 / 5, 1 / 12, 1.3, 1.4, 2)] == [23, 13, 15, 7, 91, 99, 139]


def test_apply_decay_window_matches_reference():
    """a11 pin: oracle.apply_decay_window against the reference's own core/decay.py:383-403 output
    (tests/golden/make_golden.py::decay_window)."""
    g = load('decay_window_golden.npz')
    x = np.random.default_rng(0).standard_normal(20000)
    for i, prm in enumerate(g['fixed_params']):
        got = O.apply_decay_window(x.copy(), (int(prm[0]), int(prm[1]), int(prm[2]), float(prm[3])))
        assert got.shape == g[f'fixed_out_{i}'].shape
        assert np.max(np.abs(got - g[f'fixed_out_{i}'])) <= 1e-14 * np.max(np.abs(x))
    assert O.apply_decay_window(x, None) is x

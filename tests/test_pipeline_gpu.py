"""Drop-in API parity on the GPU: the shim classes (HRIR / ImpulseResponse / FrequencyResponse) and the batch
dispatcher against the CPU oracle and the reference goldens — written the way the reference's own tests
read (tests/test_dsp_characterization.py, tests/test_pipeline_stages_helpers.py, tests/test_hrir_outputs.py)."""
import os
import warnings

import numpy as np
import pytest

import impulcifer_oracle as O
from _util import MINPHASE_TOL_F5, MINPHASE_TOL_F10, assert_ck, load, rel_peak_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


@pytest.fixture(scope="module")
def est48(T):
    from impulcifer_pip313_b200 import ImpulseResponseEstimator
    return ImpulseResponseEstimator(min_duration=5.0, fs=48000)


@pytest.fixture(scope="module")
def synth_rec(est48):
    from impulcifer_pip313_b200 import synth
    return synth.recording_set(est48.test_signal, est48.fs, synth.LAYOUT_71, seed=1234)


ROOM_FILES = [('room-FL,FR-left.wav', ['FL', 'FR'], 'left'), ('room-FL,FR-right.wav', ['FL', 'FR'], 'right'),
              ('room-FC-left.wav', ['FC'], 'left'), ('room-FC-right.wav', ['FC'], 'right'),
              ('room-BL,SL-left.wav', ['BL', 'SL'], 'left'), ('room-BL,SL-right.wav', ['BL', 'SL'], 'right'),
              ('room-SR,BR-left.wav', ['SR', 'BR'], 'left'), ('room-SR,BR-right.wav', ['SR', 'BR'], 'right')]


def test_room_responses_wav_bit_exact(T, tmp_path):
    """The reference's own committed golden: open_recording -> crop_head -> crop_tails -> write_wav must give
    the PCM_32 samples of data/demo/room-responses.wav (reference core/room_correction.py:109-132)."""
    from scipy.io import wavfile
    from impulcifer_pip313_b200 import HRIR, ImpulseResponseEstimator
    inp, gold = load('demo_room_inputs.npz'), load('demo_room_golden.npz')
    sweep_path = tmp_path / 'sweep.wav'
    wavfile.write(sweep_path, 48000, inp['sweep'])
    est = ImpulseResponseEstimator.from_wav(str(sweep_path))
    assert len(est) == 295270
    rir = HRIR(est)
    for name, speakers, side in ROOM_FILES:
        p = tmp_path / name
        wavfile.write(p, 48000, inp[name])
        rir.open_recording(str(p), speakers, side=side)
    assert [f'{s}-{e}' for s, pair in rir.irs.items() for e in pair] == list(gold['names'])
    peaks = []
    for speaker, pair in rir.irs.items():
        for side, ir in pair.items():
            peaks.append(ir.peak_index())
            ir.crop_head()
    assert peaks == list(gold['peaks'])
    knees = [ir.decay_params()[1] for pair in rir.irs.values() for ir in pair.values()]
    assert knees == list(gold['knees'])
    assert rir.crop_tails() == 21600
    out = tmp_path / 'room-responses.wav'
    rir.write_wav(str(out))
    fs, got = wavfile.read(out)
    assert fs == 48000 and got.shape == (21600, 32)
    diff = np.abs(got.T.astype(np.int64) - gold['room_responses'].astype(np.int64))
    assert diff.max() == 0, f"{np.count_nonzero(diff)} samples differ, max {diff.max()} LSB"


def test_hrir_full_chain_matches_reference_golden(T, est48, synth_rec):
    from impulcifer_pip313_b200 import HRIR
    g = load('synth71_golden.npz')
    fs = 48000
    h = HRIR(est48)
    h.ingest(synth_rec, fs, ['FL', 'FR', 'FC', 'SL', 'SR', 'BL', 'BR'])
    order = [(s, e) for s in h.irs for e in ('left', 'right')]
    assert [f'{s}-{e}' for s, e in order] == list(g['names'])
    assert [h.irs[s][e].peak_index() for s, e in order] == list(g['est_peaks'])
    dp = [h.irs[s][e].decay_params() for s, e in order]
    assert [d[1] for d in dp] == list(g['est_knees'])
    assert [d[3] for d in dp] == list(g['est_wins'])
    assert np.allclose([d[2] for d in dp], g['est_noise'], rtol=0, atol=1e-8)
    with warnings.catch_warnings():
        warnings.simplefilter("error")  # the synthetic set has every speaker on its proper side
        h.crop_heads()
    assert [len(h.irs[s][e]) for s, e in order] == list(g['len_after_heads'])
    h.align_ipsilateral_all()
    for i, (s, e) in enumerate(order):
        assert_ck(h.irs[s][e].data, g['ck_after_ipsi'][i], rel=1e-8, what=f'ipsi {s}-{e}')
    h.align_onset_groups_peak_leftref()
    assert [h.irs[s][e].peak_index() for s, e in order] == list(g['peaks_after_align'])
    assert h.crop_tails() == int(g['tail'])
    for i, (s, e) in enumerate(order):
        assert_ck(h.irs[s][e].data, g['final_ck'][i], rel=1e-8, what=f'final {s}-{e}')
    assert rel_peak_err(h.irs['FL']['left'].data, g['final_FL_left']) < 1e-9
    assert rel_peak_err(h.irs['BR']['left'].data, g['final_BR_left']) < 1e-9
    assert abs(h.irs['FL']['left'].data[-1]) < 1e-6

    # per-IR frequency-response post-processing through the class API
    for i, nm in enumerate(g['fr_names']):
        s, e = str(nm).split('-')
        fr = h.irs[s][e].frequency_response()
        assert np.array_equal(fr.frequency, g['fr_frequency'])
        assert np.max(np.abs(fr.raw - g['fr_raw'][i])) < 1e-7
        fr.raw = g['fr_raw'][i].copy()
        fr.smoothen_fractional_octave(window_size=1 / 3)
        assert np.max(np.abs(fr.smoothed - g['fr_smoothed'][i])) < 1e-9
        assert len(fr.equalization) == 0  # reset() side effect
        fr2 = fr.copy()
        fr2.error = fr2.raw - np.mean(fr2.raw)
        fr2.smoothen_heavy_light()
        assert np.max(np.abs(fr2.smoothed - g['fr_hl_smoothed'][i])) < 1e-9
        assert np.max(np.abs(fr2.error_smoothed - g['fr_hl_error'][i])) < 1e-9
        fr2.equalize(max_gain=40, treble_f_lower=10000, treble_f_upper=fs / 2)
        assert np.max(np.abs(fr2.equalization - g['eqw_equalization'][i])) < 1e-8
        fir = fr2.minimum_phase_impulse_response(fs=fs, normalize=False, f_res=5)
        assert fir.shape == (9600,) and rel_peak_err(fir, g['eqw_fir'][i]) < MINPHASE_TOL_F5
        fr.equalization = -(fr.smoothed - np.mean(fr.smoothed))
        assert rel_peak_err(fr.minimum_phase_impulse_response(fs=fs, f_res=10, normalize=True), g['minphase_f10'][i]) < MINPHASE_TOL_F10

    # next-row components: EQ-FIR application and set normalisation
    ir = h.irs['FL']['left'].copy()
    ir.equalize(g['eqw_fir'][0])
    assert len(ir) == int(g['tail']) + 9600 - 1
    assert_ck(ir.data, g['equalized_FL_left_ck'], rel=1e-8)
    assert rel_peak_err(ir.data[:4096], g['equalized_FL_left_head']) < 1e-9
    hn = h.copy()
    gain = hn.normalize(peak_target=-0.1, avg_target=None)
    assert abs(gain - float(g['normalize_gain'])) < 1e-8
    assert_ck(hn.irs['FC']['right'].data, g['normalized_FC_right_ck'], rel=1e-8)

    # resample (parity unpinned: compared with the restated nnresample port)
    ir = h.irs['FL']['left'].copy()
    ir.resample(44100)
    assert ir.fs == 44100 and rel_peak_err(ir.data, g['resampled_FL_left_44100']) < 1e-9
    h.resample(44100)
    assert h.fs == 44100 and abs(len(h.irs['FC']['left']) - int(g['tail']) * 44100 / 48000) <= 2


def test_track_layout_and_missing_rows(T, est48):
    """tests/test_hrir_outputs.py: requested track order, missing tracks are zero rows."""
    from impulcifer_pip313_b200 import HRIR, ImpulseResponse
    from impulcifer_pip313_b200.constants import HESUVI_TRACK_ORDER
    h = HRIR(est48)
    h.irs = {'FL': {'left': ImpulseResponse(np.full(8, 0.25), 48000), 'right': ImpulseResponse(np.full(8, -0.5), 48000)}}
    m = h.track_matrix(HESUVI_TRACK_ORDER)
    assert m.shape == (30, 8)
    assert np.all(m[0] == 0.25) and np.all(m[1] == -0.5) and not m[2:].any()
    assert np.array_equal(O.track_matrix({'FL': {'left': np.full(8, 0.25), 'right': np.full(8, -0.5)}},
                                         O.HESUVI_TRACK_ORDER), m)


def test_crop_heads_itd_and_side_warning(T):
    """tests/test_dsp_characterization.py:140-187: peak lands at `head`, ITD preserved, wrong-side warning."""
    from impulcifer_pip313_b200 import HRIR, ImpulseResponse
    fs = 48000

    class Est:
        def __init__(self):
            self.fs, self.n_octaves, self.device = fs, 13.0, 0

        def __len__(self):
            return 295270

    def imp(pos, n=5000):
        x = np.zeros(n)
        x[pos] = 1.0
        return x

    h = HRIR(Est())
    h.irs = {'FL': {'left': ImpulseResponse(imp(1000), fs), 'right': ImpulseResponse(imp(1048), fs)},
             'TFR': {'left': ImpulseResponse(imp(700), fs), 'right': ImpulseResponse(imp(720), fs)}}
    with pytest.warns(UserWarning, match="TFR measurement has lower delay to left ear"):
        h.crop_heads(head_ms=1)
    assert h.irs['FL']['left'].peak_index() == 48 and h.irs['FL']['right'].peak_index() == 96
    assert len(h.irs['FL']['left']) == 5000 - 952 == len(h.irs['FL']['right'])
    h2 = HRIR(Est())
    h2.fs = 44100
    with pytest.raises(ValueError):
        h2.crop_heads()
    with pytest.raises(ValueError):
        h2.crop_tails()


def test_shift_semantics(T):
    from impulcifer_pip313_b200 import ImpulseResponse
    x = np.arange(1.0, 11.0)
    for s in (0, 3, -4, 10, -10, 25):
        ir = ImpulseResponse(x.copy(), 48000)
        ir.shift(s)
        assert np.array_equal(ir.data, O.shift(x, s)), s


def test_batch_dispatcher_matches_oracle(T, est48, synth_rec):
    from impulcifer_pip313_b200 import BatchDispatcher, synth
    g = load('synth71_golden.npz')
    fs = 48000
    recs = [synth_rec] + [synth.recording_set(est48.test_signal, fs, synth.LAYOUT_71, seed=s) for s in (77, 78)]
    disp = BatchDispatcher(est48, synth.LAYOUT_71)
    batch = T.from_numpy(np.stack(recs)).cuda()
    results = disp.process(batch)
    assert len(results) == 3 and disp.launches > 0
    # set 0 against the reference golden
    r0 = results[0]
    assert r0.names == list(g['names']) and r0.tail_ind == int(g['tail'])
    assert list(r0.peaks) == list(g['est_peaks'])
    assert list(r0.knees) == list(g['knees_after_align'])
    d0 = r0.data.cpu().numpy()
    for i in range(14):
        assert_ck(d0[i], g['final_ck'][i], rel=1e-8, what=r0.names[i])
    # the other sets against the oracle chain
    P, _, _ = O.sweep_parameters(fs)
    for k in (1, 2):
        irs = O.open_recording(recs[k], fs, synth.LAYOUT_71, est48.inverse_filter, fs)
        O.crop_heads(irs, fs)
        O.align_ipsilateral_all(irs, fs)
        O.align_onset_groups_peak_leftref(irs)
        tail = O.crop_tails(irs, fs, len(est48), P, fs)
        assert results[k].tail_ind == tail
        got = results[k].as_dict()
        for s in irs:
            for e in irs[s]:
                assert rel_peak_err(got[s][e], irs[s][e]) < 1e-9, (k, s, e)
    # int32 PCM input gives the same result as float64 input of the same samples
    pcm = T.from_numpy(np.rint(np.stack(recs[:1]) * 2 ** 31).astype(np.int32)).cuda()
    rp = disp.process(pcm)[0]
    assert rp.tail_ind == r0.tail_ind and rel_peak_err(rp.data.cpu().numpy(), d0) < 1e-12
    # FP32 fast path: integers must still agree, values within -90 dB of peak
    r32 = disp.process(batch[:1], precision=32)[0]
    assert list(r32.peaks) == list(r0.peaks)
    err_db = 20 * np.log10(np.max(np.abs(r32.data.cpu().numpy()[:, :r0.tail_ind] - d0)) / np.max(np.abs(d0))) \
        if r32.tail_ind == r0.tail_ind else None
    assert err_db is None or err_db < -90.0
    # frequency-response post-processing of every row in one go
    post = disp.postprocess(results[:1])[0]
    grid, raw, sm, fir = post
    assert raw.shape == (14, 783) and fir.shape == (14, 9600)
    f, raw_o = O.frequency_response(d0[0], fs)
    assert np.max(np.abs(raw[0].cpu().numpy() - raw_o)) < 1e-7
    assert np.max(np.abs(sm[0].cpu().numpy() - O.smoothen_fractional_octave(f, raw[0].cpu().numpy(), 1 / 3))) < 1e-9


def test_batch_dispatcher_seed_fuzz_integers_and_values(T, est48):
    """Eight more seeded recording sets in one batch (different onset latencies, decay times, noise floors): every
    integer the path decides (first peaks, knees, tail length) equals the oracle chain, samples within 1e-9 of peak."""
    from impulcifer_pip313_b200 import BatchDispatcher, synth
    fs = 48000
    seeds = (5, 11, 23, 42, 97, 101, 314, 2718)
    recs = [synth.recording_set(est48.test_signal, fs, synth.LAYOUT_71, seed=s) for s in seeds]
    disp = BatchDispatcher(est48, synth.LAYOUT_71)
    results = disp.process(T.from_numpy(np.stack(recs)).cuda())
    P, _, _ = O.sweep_parameters(fs)
    for k, seed in enumerate(seeds):
        irs = O.open_recording(recs[k], fs, synth.LAYOUT_71, est48.inverse_filter, fs)
        peaks = [O.peak_index(irs[s][e]) for s in irs for e in ('left', 'right')]
        assert list(results[k].peaks) == peaks, seed
        O.crop_heads(irs, fs)
        O.align_ipsilateral_all(irs, fs)
        O.align_onset_groups_peak_leftref(irs)
        knees = [O.decay_params(irs[s][e], fs)[1] for s in irs for e in ('left', 'right')]
        assert list(results[k].knees) == knees, seed
        tail = O.crop_tails(irs, fs, len(est48), P, fs)
        assert results[k].tail_ind == tail, seed
        got = results[k].as_dict()
        for s in irs:
            for e in irs[s]:
                assert rel_peak_err(got[s][e], irs[s][e]) < 1e-9, (seed, s, e)


def test_repeat_runs_bitwise_deterministic(T, est48, synth_rec):
    """tests/test_brir_thread_determinism.py analogue: repeated runs give byte-identical results."""
    from impulcifer_pip313_b200 import BatchDispatcher, synth
    disp = BatchDispatcher(est48, synth.LAYOUT_71)
    batch = T.from_numpy(synth_rec[None]).cuda()
    a = disp.process(batch)[0].data.cpu().numpy()
    for _ in range(2):
        assert np.array_equal(a, disp.process(batch)[0].data.cpu().numpy())


def test_process_host_matches_process(T, est48, synth_rec):
    """Host-buffer entry point (chunked H2D / compute / D2H overlap) == device-resident entry point, bit for bit."""
    from impulcifer_pip313_b200 import BatchDispatcher, synth
    disp = BatchDispatcher(est48, synth.LAYOUT_71)
    recs = np.stack([synth_rec, synth_rec[:, ::-1].copy() * 0 + synth_rec * 0.5, synth_rec * 0.25])
    dev = disp.process(T.from_numpy(recs).cuda())
    host_in = T.from_numpy(recs).pin_memory()
    out = T.empty(sum(r.data.numel() for r in dev) + 16, dtype=T.float64).pin_memory()
    got = disp.process_host(host_in, out=out, chunk_sets=2)
    assert [r.tail_ind for r in got] == [r.tail_ind for r in dev]
    for a, b in zip(got, dev):
        assert not a.data.is_cuda and np.array_equal(a.data.numpy(), b.data.cpu().numpy())
    pcm = T.from_numpy(np.rint(recs * 2 ** 31).astype(np.int32)).pin_memory()
    got_pcm = disp.process_host(pcm, chunk_sets=1)
    ref_pcm = disp.process(T.from_numpy(np.rint(recs * 2 ** 31) / 2 ** 31).cuda())
    for a, b in zip(got_pcm, ref_pcm):
        assert np.array_equal(a.data.numpy(), b.data.cpu().numpy())


def test_truehd_11ch_layout_batched(T, est48):
    """BASELINE config 4 shape: TrueHD 11-channel layout (22 responses per set) through the dispatcher."""
    from impulcifer_pip313_b200 import BatchDispatcher, synth
    fs = 48000
    rec = synth.recording_set(est48.test_signal, fs, synth.LAYOUT_TRUEHD_11, seed=5)
    assert rec.shape == (2, 4399970)
    disp = BatchDispatcher(est48, synth.LAYOUT_TRUEHD_11)
    res = disp.process(T.from_numpy(rec[None]).cuda())[0]
    assert len(res.names) == 22 and res.names[:2] == ['FL-left', 'FL-right'] and res.names[-1] == 'TBR-right'
    P, _, _ = O.sweep_parameters(fs)
    irs = O.open_recording(rec, fs, synth.LAYOUT_TRUEHD_11, est48.inverse_filter, fs)
    O.crop_heads(irs, fs)
    O.align_ipsilateral_all(irs, fs)
    O.align_onset_groups_peak_leftref(irs)
    assert res.tail_ind == O.crop_tails(irs, fs, len(est48), P, fs)
    got = res.as_dict()
    for s in irs:
        for e in irs[s]:
            assert rel_peak_err(got[s][e], irs[s][e]) < 1e-9, (s, e)
    # TrueHD output layout: "<speaker>-left, <speaker>-right" rows in TRUEHD_11CH_ORDER
    order = [f'{s}-{e}' for s in synth.LAYOUT_TRUEHD_11 for e in ('left', 'right')]
    m = O.track_matrix(got, order)
    assert m.shape == (22, res.tail_ind) and np.array_equal(m[3], got['FR']['right'])


def test_192khz_resample_smooth_minphase(T):
    """BASELINE config 3 shape: 192 kHz sweep (L = 1 746 783, half-width tile plan) -> deconvolution -> crop ->
    resample to 48 kHz -> 1/3-octave smoothing -> minimum-phase FIR, each step against the oracle."""
    from impulcifer_pip313_b200 import ImpulseResponse, ImpulseResponseEstimator
    fs = 192000
    est = ImpulseResponseEstimator(min_duration=5.0, fs=fs)
    assert len(est) == 1362783
    rng = np.random.default_rng(4)
    L = 2 * fs + len(est)
    h = np.zeros(60000)
    h[500] = 1.0
    h[501:] = 0.05 * rng.standard_normal(59499) * 10 ** (-3 * np.arange(59499) / (fs * 0.2))
    rec = np.zeros((2, L))
    nfft = 1 << 21
    seg = np.fft.irfft(np.fft.rfft(est.test_signal, nfft) * np.fft.rfft(h, nfft), nfft)[:L]
    rec[0], rec[1] = 0.3 * seg, 0.2 * np.roll(seg, 40)
    rec += 1e-5 * rng.standard_normal(rec.shape)
    plan = est.plan(L)
    assert plan.nfft >= 2428174
    got = est.estimate_batch(rec)
    ref = np.stack([O.estimate(r, est.inverse_filter) for r in rec])
    assert rel_peak_err(got, ref) < 1e-9
    ir = ImpulseResponse(got[0], fs)
    pk = ir.peak_index()
    assert pk == O.peak_index(ref[0])
    ir.crop_head()
    knee = ir.decay_params()[1]
    assert knee == O.decay_params(O.crop_head(ref[0], fs), fs)[1]
    n_keep = 4 * 21600
    ir.data = ir.data[:n_keep]
    ref_c = O.crop_head(ref[0], fs)[:n_keep]
    ir.resample(48000)
    ref_r = O.resample(ref_c, 48000, fs)
    assert ir.fs == 48000 and len(ir) == len(ref_r) == 21600
    assert rel_peak_err(ir.data, ref_r) < 1e-9
    fr = ir.frequency_response()
    f, raw = O.frequency_response(ref_r, 48000)
    assert np.max(np.abs(fr.raw - raw)) < 1e-6
    fr.smoothen_fractional_octave(window_size=1 / 3)
    assert np.max(np.abs(fr.smoothed - O.smoothen_fractional_octave(f, fr.raw, 1 / 3))) < 1e-9
    fr.equalization = -(fr.smoothed - np.mean(fr.smoothed))
    fir = fr.minimum_phase_impulse_response(fs=48000, f_res=5, normalize=False)
    ref_fir = O.minimum_phase_impulse_response(f, fr.equalization, fs=48000, f_res=5, normalize=False)
    assert fir.shape == (9600,) and rel_peak_err(fir, ref_fir) < MINPHASE_TOL_F5


def test_equalization_firs_batch_matches_worker_golden(T):
    """The batched replacement of process_equalization_worker (core/parallel_workers.py:69-131): same FIRs, same
    input order as parallel_map, one launch per stage for all (speaker, side) items."""
    from impulcifer_pip313_b200 import FrequencyResponse
    from impulcifer_pip313_b200.dispatcher import equalization_firs
    g = load('synth71_golden.npz')
    f = g['fr_frequency']
    fs = 48000
    names = [str(n).split('-') for n in g['fr_names']]
    # golden: error = raw - mean(raw) per item, target 0  ->  feed it as the room error of each (speaker, side)
    room = {}
    for (s, e), raw in zip(names, g['fr_raw']):
        room.setdefault(s, {})[e] = FrequencyResponse(name='room', frequency=f, raw=raw, error=raw - np.mean(raw))
    target = FrequencyResponse(name='target', frequency=f, raw=np.zeros(len(f)))
    out = equalization_firs([(s, e) for s, e in names], room, None, None, None, None, target, f, fs)
    assert [(s, e) for s, e, _ in out] == [(s, e) for s, e in names]
    for i, (_, _, fir) in enumerate(out):
        assert fir.shape == (9600,) and rel_peak_err(fir, g['eqw_fir'][i]) < MINPHASE_TOL_F5


def test_magnitude_response_function(T):
    from impulcifer_pip313_b200.audio_io import magnitude_response
    x = np.random.default_rng(8).standard_normal(21600) * np.exp(-np.arange(21600) / 3000.0)
    f, m = magnitude_response(x, 48000)
    fo, mo = O.magnitude_response(x, 48000)
    assert np.array_equal(f, fo) and np.max(np.abs(m - mo)) < 1e-8
    # a length with large prime factors takes the chirp-z plan and gives the same bins as the reference's rfft
    y = np.random.default_rng(9).standard_normal(2 * 3 * 7 * 11 * 13)
    f2, m2 = magnitude_response(y, 48000)
    fo2, mo2 = O.magnitude_response(y, 48000)
    assert np.array_equal(f2, fo2) and np.max(np.abs(10 ** (m2 / 20) - 10 ** (mo2 / 20))) < 1e-11 * np.max(10 ** (mo2 / 20))


def test_ragged_recording_clipped_last_column(T, est48):
    """A file whose last column is clipped (still >= one sweep) keeps that column at its own length
    (reference core/hrir.py:194-216); both columns go through the batched path with different lengths."""
    from impulcifer_pip313_b200 import HRIR
    fs = 48000
    rng = np.random.default_rng(12)
    n_samples = 96000 + 391270 + 350000
    rec = 1e-4 * rng.standard_normal((1, n_samples))
    for start, g in ((96000 + 30000, 0.5), (96000 + 391270 + 20000, 0.3)):
        rec[0, start:start + len(est48)] += g * est48.test_signal
    h = HRIR(est48)
    h.ingest(rec, fs, ['FL', 'FR'], side='left')
    irs = O.open_recording(rec, fs, ['FL', 'FR'], est48.inverse_filter, fs, side='left')
    assert len(h.irs['FL']['left']) == 391270 and len(h.irs['FR']['left']) == 350000
    for s in ('FL', 'FR'):
        assert rel_peak_err(h.irs[s]['left'].data, irs[s]['left']) < 1e-9
        assert h.irs[s]['left'].peak_index() == O.peak_index(irs[s]['left'])
    with pytest.raises(ValueError, match="Sampling rate"):
        h.ingest(rec, 44100, ['FL', 'FR'], side='left')
    with pytest.raises(ValueError, match="No valid columns"):
        h.ingest(rec[:, :96000 + 1000], fs, ['FL'], side='left')


def test_pcm_packing_matches_host_quantiser_and_golden(T, est48):
    """Device track gather + PCM quantise + interleave == host np.rint path, for 16/24/32 bit, incl. silent tracks,
    ties (round half to even) and clipping; 32-bit frames of the demo set == room-responses.wav (see above)."""
    from impulcifer_pip313_b200 import HRIR, ImpulseResponse
    from impulcifer_pip313_b200.audio_io import quantize_pcm
    from impulcifer_pip313_b200.constants import HESUVI_TRACK_ORDER
    rng = np.random.default_rng(31)
    n = 5000
    h = HRIR(est48)
    h.irs = {}
    for spk in ('FL', 'FR', 'FC', 'SL'):
        h.irs[spk] = {}
        for side in ('left', 'right'):
            x = rng.standard_normal(n) * 0.3
            x[:8] = [1.0, -1.0, 1.5, -1.5, 0.5 / 2 ** 31, 1.5 / 2 ** 31, 2.5 / 2 ** 15, -0.5 / 2 ** 23]
            h.irs[spk][side] = ImpulseResponse(x, 48000)
    for order in (None, HESUVI_TRACK_ORDER):
        m = h.track_matrix(order)
        for bits in (16, 24, 32):
            frames = h.pcm_frames(order, bits)
            assert frames.shape == (n, m.shape[0])
            assert np.array_equal(frames, quantize_pcm(m, bits).T), (bits,)
    with pytest.raises(ValueError, match="Invalid bit depth"):
        h.pcm_frames(None, 20)

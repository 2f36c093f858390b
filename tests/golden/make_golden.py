#!/usr/bin/env python
"""Generate the committed golden fixtures from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Imports 115dkk/Impulcifer-pip313 from /root/reference through ``oracle/ref_import.py`` (plot / audio-IO
stubs only, no arithmetic restated) and records what the reference itself computes for the hot path:

  demo_room_inputs.npz   the eight ``data/demo/room-*.wav`` recordings + ``FC.wav`` as int32 PCM
  demo_room_golden.npz   ``data/demo/room-responses.wav`` (the reference's own committed output,
                         32 x 21600 int32) + per-IR peak / knee / crop integers + FC.wav HRIR chain
  synth71_golden.npz     reference results on the seeded synthetic 7.1 set (``synth.recording_set(seed=1234)``):
                         all integers, checksums of every IR, slices, FR / smoothing / min-phase vectors
  room_chain_golden.npz  room / headphone FR chains, channel-balance FIRs, microphone-mismatch estimator on the same set
  decay_chain_golden.npz decay times, reflection levels, decay-adjustment parameters, virtual bass on the same set
  decay_window_golden.npz  ``apply_decay_window`` on fixed parameters + ``process_decay_worker`` on the same set (a11 pin)
  equalize_golden.npz    ``FrequencyResponse.equalize`` on clipped / unclipped error curves (a17 pin)
  small_cases.npz        tiny estimators from the reference's own tests (fs=2000/0.05 s, fs=48000/1 s)
"""
import os
import sys

import numpy as np
from scipy.io import wavfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, ROOT)

import ref_import  # noqa: E402

ref = ref_import.modules()
from impulcifer_pip313_b200 import synth  # noqa: E402

IRE = ref.impulse_response_estimator.ImpulseResponseEstimator
HRIR = ref.hrir.HRIR
ImpulseResponse = ref.impulse_response.ImpulseResponse
FrequencyResponse = ref.frequency_response.FrequencyResponse
DEMO = os.path.join(ref_import.REFERENCE_ROOT, 'data', 'demo')
SWEEP = os.path.join(ref_import.REFERENCE_ROOT, 'data', 'sweep-6.15s-48000Hz-32bit-2.93Hz-24000Hz.wav')


def checksums(x):
    """Order-insensitive-enough fingerprints of a float64 vector (compared at 1e-12 relative)."""
    x = np.asarray(x, dtype=np.float64)
    k = np.arange(len(x), dtype=np.float64)
    probe = np.cos(0.37 * k) + 0.5 * np.sin(0.011 * k * 1.0)
    return np.array([x.sum(), np.dot(x, x), np.dot(x, probe), np.abs(x).max()])


def demo_room():
    est = IRE.from_wav(SWEEP)
    files = [('room-FL,FR-left.wav', ['FL', 'FR'], 'left'), ('room-FL,FR-right.wav', ['FL', 'FR'], 'right'),
             ('room-FC-left.wav', ['FC'], 'left'), ('room-FC-right.wav', ['FC'], 'right'),
             ('room-BL,SL-left.wav', ['BL', 'SL'], 'left'), ('room-BL,SL-right.wav', ['BL', 'SL'], 'right'),
             ('room-SR,BR-left.wav', ['SR', 'BR'], 'left'), ('room-SR,BR-right.wav', ['SR', 'BR'], 'right')]
    inputs = {}
    rir = HRIR(est)
    for name, speakers, side in files:
        fs, d = wavfile.read(os.path.join(DEMO, name))
        assert fs == 48000 and d.dtype == np.int32
        inputs[name] = d
        rir.open_recording(os.path.join(DEMO, name), speakers, side=side)
    fs, fc = wavfile.read(os.path.join(DEMO, 'FC.wav'))
    inputs['FC.wav'] = fc.T.copy()
    fs, sweep = wavfile.read(SWEEP)
    inputs['sweep'] = sweep
    np.savez_compressed(os.path.join(HERE, 'demo_room_inputs.npz'), **inputs)

    gold = {}
    names, peaks, knees, noise, wins, crops, lens = [], [], [], [], [], [], []
    for spk, pair in rir.irs.items():
        for side, ir in pair.items():
            names.append(f'{spk}-{side}')
            gold[f'est_ck_{spk}-{side}'] = checksums(ir.data)
            pk = ir.peak_index()
            peaks.append(pk)
            crops.append(max(pk - 48, 0))
            ir.crop_head()
            p, k, nf, w = ir.decay_params()
            knees.append(k)
            noise.append(nf)
            wins.append(w)
            lens.append(len(ir.data))
    tail = rir.crop_tails()
    _, golden_wav = wavfile.read(os.path.join(DEMO, 'room-responses.wav'))
    mat = np.vstack([rir.irs[s][e].data if s in rir.irs and e in rir.irs[s] else np.zeros(tail)
                     for s, e in (n.split('-') for n in ref.constants.HEXADECAGONAL_TRACK_ORDER)])
    q = np.clip(np.rint(mat * 2 ** 31), -2 ** 31, 2 ** 31 - 1).astype(np.int32)
    assert np.array_equal(q, golden_wav.T), "reference no longer reproduces room-responses.wav"
    gold.update(room_responses=golden_wav.T.copy(), names=np.array(names), peaks=np.array(peaks),
                crops=np.array(crops), knees=np.array(knees), noise=np.array(noise), wins=np.array(wins),
                lens=np.array(lens), tail=np.array(tail))

    # binaural chain on FC.wav (the only complete HRIR recording in the demo set)
    h = HRIR(est)
    h.open_recording(os.path.join(DEMO, 'FC.wav'), ['FC'])
    gold['fc_peaks'] = np.array([h.irs['FC']['left'].peak_index(), h.irs['FC']['right'].peak_index()])
    h.crop_heads()
    gold['fc_len_after_heads'] = np.array([len(h.irs['FC']['left'].data), len(h.irs['FC']['right'].data)])
    h.align_ipsilateral_all()
    gold['fc_tail'] = np.array(h.crop_tails())
    gold['fc_left'] = h.irs['FC']['left'].data.copy()
    gold['fc_right'] = h.irs['FC']['right'].data.copy()
    fr = h.irs['FC']['left'].frequency_response()
    gold['fc_fr_frequency'] = fr.frequency.copy()
    gold['fc_fr_raw'] = fr.raw.copy()
    fr.smoothen_fractional_octave(window_size=1 / 3)
    gold['fc_fr_smoothed'] = fr.smoothed.copy()
    np.savez_compressed(os.path.join(HERE, 'demo_room_golden.npz'), **gold)
    print('demo_room: tail', tail, 'peaks', peaks)


def synth71():
    est = IRE(min_duration=5.0, fs=48000)
    rec = synth.recording_set(est.test_signal, est.fs, synth.LAYOUT_71, seed=1234)
    import tempfile
    gold = {'sweep_len': np.array(len(est)), 'n_octaves': np.array(est.n_octaves),
            'inv_ck': checksums(est.inverse_filter), 'sweep_ck': checksums(est.test_signal),
            'rec_ck': np.array([checksums(rec[0]), checksums(rec[1])])}
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, 'FL,FR,FC,SL,SR,BL,BR.wav')
        wavfile.write(path, est.fs, np.rint(rec.T * 2 ** 31).astype(np.int32))
        h = HRIR(est)
        h.open_recording(path, synth.LAYOUT_71)
    names = [f'{s}-{e}' for s in h.irs for e in ('left', 'right')]
    gold['names'] = np.array(names)
    gold['est_ck'] = np.array([checksums(h.irs[s][e].data) for s in h.irs for e in ('left', 'right')])
    gold['est_peaks'] = np.array([h.irs[s][e].peak_index() for s in h.irs for e in ('left', 'right')])
    # a window of the raw estimate around each peak (pins float values, not just checksums)
    gold['est_slices'] = np.array([h.irs[s][e].data[pk - 64:pk + 1984]
                                   for (s, e), pk in zip(((s, e) for s in h.irs for e in ('left', 'right')),
                                                         gold['est_peaks'])])
    dp = [h.irs[s][e].decay_params() for s in h.irs for e in ('left', 'right')]
    gold['est_knees'] = np.array([d[1] for d in dp])
    gold['est_noise'] = np.array([d[2] for d in dp])
    gold['est_wins'] = np.array([d[3] for d in dp])
    h.crop_heads()
    gold['len_after_heads'] = np.array([len(h.irs[s][e].data) for s in h.irs for e in ('left', 'right')])
    h.align_ipsilateral_all()
    gold['ck_after_ipsi'] = np.array([checksums(h.irs[s][e].data) for s in h.irs for e in ('left', 'right')])
    h.align_onset_groups_peak_leftref()
    gold['peaks_after_align'] = np.array([h.irs[s][e].peak_index() for s in h.irs for e in ('left', 'right')])
    dp = [h.irs[s][e].decay_params() for s in h.irs for e in ('left', 'right')]
    gold['knees_after_align'] = np.array([d[1] for d in dp])
    tail = h.crop_tails()
    gold['tail'] = np.array(tail)
    gold['final_ck'] = np.array([checksums(h.irs[s][e].data) for s in h.irs for e in ('left', 'right')])
    gold['final_FL_left'] = h.irs['FL']['left'].data.copy()
    gold['final_BR_left'] = h.irs['BR']['left'].data.copy()
    # FR post-processing on four IRs
    fr_names = [('FL', 'left'), ('FL', 'right'), ('FC', 'left'), ('BR', 'right')]
    raws, sm13, hl_s, hl_e, mp5, mp10, sm_custom, eqw_eq, eqw_fir = [], [], [], [], [], [], [], [], []
    for s, e in fr_names:
        fr = h.irs[s][e].frequency_response()
        gold['fr_frequency'] = fr.frequency.copy()
        raws.append(fr.raw.copy())
        fr.smoothen_fractional_octave(window_size=1 / 3)
        sm13.append(fr.smoothed.copy())
        fr2 = fr.copy()
        fr2.error = fr2.raw - np.mean(fr2.raw)
        fr2.smoothen_heavy_light()
        hl_s.append(fr2.smoothed.copy())
        hl_e.append(fr2.error_smoothed.copy())
        # the arithmetic of process_equalization_worker (core/parallel_workers.py:98-131)
        fr2.equalize(max_gain=40, treble_f_lower=10000, treble_f_upper=est.fs / 2)
        eqw_eq.append(fr2.equalization.copy())
        eqw_fir.append(fr2.minimum_phase_impulse_response(fs=est.fs, normalize=False, f_res=5))
        fr3 = fr.copy()
        fr3.smoothen_fractional_octave(window_size=2, iterations=2, treble_window_size=1 / 12, treble_iterations=3,
                                       treble_f_lower=2000, treble_f_upper=12000)
        sm_custom.append(fr3.smoothed.copy())
        fr.equalization = -(fr.smoothed - np.mean(fr.smoothed))
        mp5.append(fr.minimum_phase_impulse_response(fs=est.fs, f_res=5, normalize=False))
        mp10.append(fr.minimum_phase_impulse_response(fs=est.fs, f_res=10, normalize=True))
    gold.update(fr_names=np.array([f'{s}-{e}' for s, e in fr_names]), fr_raw=np.array(raws),
                fr_smoothed=np.array(sm13), fr_hl_smoothed=np.array(hl_s), fr_hl_error=np.array(hl_e),
                fr_custom=np.array(sm_custom), minphase_f5=np.array(mp5), minphase_f10=np.array(mp10),
                eqw_equalization=np.array(eqw_eq), eqw_fir=np.array(eqw_fir))
    # EQ-FIR application (ImpulseResponse.equalize, next-row component) on one response
    ir = ImpulseResponse(h.irs['FL']['left'].data.copy(), est.fs)
    ir.equalize(eqw_fir[0])
    gold['equalized_FL_left_ck'] = checksums(ir.data)
    gold['equalized_FL_left_head'] = ir.data[:4096].copy()
    # normalisation gain of the whole set (HRIR.normalize, next-row component)
    hn = h.copy()
    gold['normalize_gain'] = np.array(hn.normalize(peak_target=-0.1, avg_target=None))
    gold['normalized_FC_right_ck'] = checksums(hn.irs['FC']['right'].data)
    # resample 48k -> 44.1k through the reference method (nnresample = restated port: UNPINNED)
    ir = ImpulseResponse(h.irs['FL']['left'].data.copy(), est.fs)
    ir.resample(44100)
    gold['resampled_FL_left_44100'] = ir.data.copy()
    np.savez_compressed(os.path.join(HERE, 'synth71_golden.npz'), **gold)
    print('synth71: tail', tail, 'peaks', gold['est_peaks'])



def room_curves(frequency):
    """Deterministic target / microphone-calibration curves on a frequency grid (shared with the GPU tests)."""
    lf = np.log10(frequency)
    target = -1.2 * (lf - 3.0) + 1.5 * np.exp(-((lf - 2.0) / 0.35) ** 2)
    mic = 0.8 * np.sin(2.3 * lf) + 0.3 * (lf - 3.0) ** 2
    return target, mic


def room_chain():
    """Next-row components on the seeded synthetic 7.1 set: room / headphone frequency-response
    chains, channel-balance FIR design and the interaural microphone-mismatch estimator, all computed by the
    reference's own functions."""
    import copy
    import tempfile
    rc = ref_import.module('core.room_correction')
    mdc = ref_import.module('core.microphone_deviation_correction')
    est = IRE(min_duration=5.0, fs=48000)
    rec = synth.recording_set(est.test_signal, est.fs, synth.LAYOUT_71, seed=1234)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, 'FL,FR,FC,SL,SR,BL,BR.wav')
        wavfile.write(path, est.fs, np.rint(rec.T * 2 ** 31).astype(np.int32))
        h = HRIR(est)
        h.open_recording(path, synth.LAYOUT_71)
    h.crop_heads()
    h.align_ipsilateral_all()
    h.align_onset_groups_peak_leftref()
    h.crop_tails()
    grid = FrequencyResponse.generate_frequencies(f_step=1.01, f_min=10, f_max=est.fs / 2)
    t_raw, m_raw = room_curves(grid)
    target = FrequencyResponse(name='room-target', frequency=grid.copy(), raw=t_raw.copy())
    mic = FrequencyResponse(name='room-mic', frequency=grid.copy(), raw=m_raw.copy())
    keys = [(s, e) for s in h.irs for e in ('left', 'right')]
    gold = {'names': np.array([f'{s}-{e}' for s, e in keys]), 'frequency': grid}

    # speaker-ear specific corrections (core/room_correction.py:185-212)
    frs = rc.calculate_specific_room_corrections(h.copy(), target.copy(), mic_calibration=mic.copy(), limit=400)
    gold['specific_raw'] = np.array([frs[s][e].raw for s, e in keys])
    gold['specific_error'] = np.array([frs[s][e].error for s, e in keys])
    gold['specific_target'] = np.array([frs[s][e].target for s, e in keys])
    frs0 = rc.calculate_specific_room_corrections(h.copy(), target.copy(), mic_calibration=None, limit=0)
    gold['specific_nolimit_error'] = np.array([frs0[s][e].error for s, e in keys])

    # generic (speaker independent) correction from four responses (core/room_correction.py:234-299)
    gen_keys = [('FL', 'left'), ('FR', 'right'), ('FC', 'left'), ('FC', 'right')]
    gold['generic_names'] = np.array([f'{s}-{e}' for s, e in gen_keys])
    for method in ('average', 'conservative'):
        irs = [ImpulseResponse(h.irs[s][e].data.copy(), est.fs) for s, e in gen_keys]
        fr = rc.calculate_generic_room_correction(irs, target.copy(), mic_calibration=mic.copy(), method=method,
                                                  limit=1000)
        for k in ('raw', 'error', 'error_smoothed', 'smoothed', 'target'):
            gold[f'generic_{method}_{k}'] = np.asarray(getattr(fr, k)).copy()
    irs = [ImpulseResponse(h.irs['FC']['left'].data.copy(), est.fs)]
    fr = rc.calculate_generic_room_correction(irs, target.copy(), mic_calibration=None, method='average', limit=0)
    gold['generic_single_error'] = fr.error.copy()
    gold['generic_single_error_smoothed'] = fr.error_smoothed.copy()

    # headphone compensation arithmetic (core/pipeline_stages.py:421-438), FL-left / FR-right as the two ears
    left = h.irs['FL']['left'].frequency_response()
    right = h.irs['FR']['right'].frequency_response()
    gain = left.center([100, 10000])
    right.raw += gain
    zero = FrequencyResponse(name='zero', frequency=left.frequency, raw=np.zeros(len(left.frequency)))
    left.compensate(zero, min_mean_error=False)
    right.compensate(zero, min_mean_error=False)
    gold['hp_gain'] = np.array(gain)
    gold['hp_left_error'] = left.error.copy()
    gold['hp_right_error'] = right.error.copy()

    # channel balance (core/hrir.py:655-799)
    methods = ['trend', 'left', 'right', 'avg', 'min', 'mids', '1.5']
    gold['balance_methods'] = np.array(methods)
    for method in methods:
        hb = h.copy()
        left_fr = ImpulseResponse(np.mean(np.vstack([hb.irs[s]['left'].data for s in ('FL', 'FR')]), axis=0),
                                  est.fs).frequency_response()
        right_fr = ImpulseResponse(np.mean(np.vstack([hb.irs[s]['right'].data for s in ('FL', 'FR')]), axis=0),
                                   est.fs).frequency_response()
        firs = hb.channel_balance_firs(left_fr, right_fr, method)
        gold[f'balance_{method}_left_fir'] = np.asarray(firs[0], dtype=np.float64)
        gold[f'balance_{method}_right_fir'] = np.asarray(firs[1], dtype=np.float64)
        hb.correct_channel_balance(method)
        gold[f'balance_{method}_ck'] = np.array([checksums(hb.irs[s][e].data) for s, e in keys])
        gold[f'balance_{method}_len'] = np.array([len(hb.irs[s][e].data) for s, e in keys])

    # interaural microphone mismatch (core/microphone_deviation_correction.py:106-276)
    for anchor in ('auto', 'diffuse'):
        c = mdc.MicrophoneMatchingCorrector(est.fs, anchor=anchor)
        for s in h.irs:
            li, ri = h.irs[s]['left'], h.irs[s]['right']
            c.collect_speaker(s, li.data, ri.data, li.peak_index(), ri.peak_index())
        if anchor == 'auto':
            gold['mic_frequency'] = c.frequency.copy()
            gold['mic_power_left'] = np.array([c.speaker_power[s]['left'] for s in h.irs])
            gold['mic_power_right'] = np.array([c.speaker_power[s]['right'] for s in h.irs])
        gold[f'mic_{anchor}_mismatch'] = c.estimate_interaural_mismatch().copy()
        gold[f'mic_{anchor}_used'] = np.array(c.anchor_used)
        lf, rf = c.design_correction_filters()
        gold[f'mic_{anchor}_left_fir'] = lf.copy()
        gold[f'mic_{anchor}_right_fir'] = rf.copy()
    hm = h.copy()
    mdc.apply_microphone_deviation_correction_to_hrir(hm, correction_strength=0.7, anchor='auto')
    gold['mic_applied_ck'] = np.array([checksums(hm.irs[s][e].data) for s, e in keys])
    gold['mic_applied_len'] = np.array([len(hm.irs[s][e].data) for s, e in keys])
    np.savez_compressed(os.path.join(HERE, 'room_chain_golden.npz'), **gold)
    print('room_chain: specific error rms', float(np.sqrt(np.mean(gold['specific_error'] ** 2))),
          'mic anchor', gold['mic_auto_used'], 'balance trend fir', gold['balance_trend_right_fir'].shape)



def _cropped_set(stage='final'):
    import tempfile
    est = IRE(min_duration=5.0, fs=48000)
    rec = synth.recording_set(est.test_signal, est.fs, synth.LAYOUT_71, seed=1234)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, 'FL,FR,FC,SL,SR,BL,BR.wav')
        wavfile.write(path, est.fs, np.rint(rec.T * 2 ** 31).astype(np.int32))
        h = HRIR(est)
        h.open_recording(path, synth.LAYOUT_71)
    h.crop_heads()
    if stage == 'heads':
        return est, h
    h.align_ipsilateral_all()
    h.align_onset_groups_peak_leftref()
    h.crop_tails()
    return est, h


def _rt(values):
    return np.array([np.nan if v is None else float(v) for v in values])


def decay_chain():
    """Next-row components: decay times, reflection levels, decay-adjustment parameters and
    virtual bass, computed by the reference on the seeded synthetic 7.1 set."""
    decay = ref.decay
    vb = ref_import.module('core.virtual_bass')
    gold = {}
    for stage in ('final', 'heads'):
        est, h = _cropped_set(stage)
        keys = [(s, e) for s in h.irs for e in ('left', 'right')]
        gold['names'] = np.array([f'{s}-{e}' for s, e in keys])
        params = [h.irs[s][e].decay_params() for s, e in keys]
        gold[f'{stage}_params'] = np.array([[p[0], p[1], p[3]] for p in params])
        gold[f'{stage}_noise'] = np.array([p[2] for p in params])
        gold[f'{stage}_decay_times'] = np.array([_rt(h.irs[s][e].decay_times(*p)) for (s, e), p in zip(keys, params)])
        gold[f'{stage}_len'] = np.array([len(h.irs[s][e].data) for s, e in keys])
    # final set: README metrics and decay adjustment
    refl = h_final_reflections = None
    est, h = _cropped_set('final')
    keys = [(s, e) for s in h.irs for e in ('left', 'right')]
    refl = h.calculate_reflection_levels()
    gold['reflection'] = np.array([[refl[s][e]['early_db'], refl[s][e]['late_db']] for s, e in keys])
    refl2 = h.calculate_reflection_levels(direct_sound_duration_ms=1, early_ref_start_ms=5, early_ref_end_ms=300,
                                          late_ref_start_ms=900, late_ref_end_ms=2000)
    gold['reflection_clipped'] = np.array([[refl2[s][e]['early_db'], refl2[s][e]['late_db']] for s, e in keys])
    adj = []
    for target in (0.05, 0.15, 2.0):
        for s, e in keys[:4]:
            p = decay.decay_adjustment_params(h.irs[s][e].data, est.fs, target)
            adj.append([target] + ([np.nan] * 4 if p is None else [float(v) for v in p]))
    gold['decay_adjustment'] = np.array(adj)
    # virtual bass
    for tag, kw in (('default', {}), ('xo120_inv', dict(crossover_freq=120, invert_polarity=True, head_ms=0.5))):
        hv = h.copy()
        hv.irs['BR']['right'].data = hv.irs['BR']['right'].data[:-1000]  # exercises the zero-padding branch
        vb.apply_virtual_bass_to_hrir(hv, **kw)
        gold[f'vbass_{tag}_ck'] = np.array([checksums(hv.irs[s][e].data) for s, e in keys])
        gold[f'vbass_{tag}_len'] = np.array([len(hv.irs[s][e].data) for s, e in keys])
        gold[f'vbass_{tag}_FL_left'] = hv.irs['FL']['left'].data[:6000].copy()
        gold[f'vbass_{tag}_FC_right'] = hv.irs['FC']['right'].data[:6000].copy()
        gold[f'vbass_{tag}_BR_right_tail'] = hv.irs['BR']['right'].data[-3000:].copy()
    np.savez_compressed(os.path.join(HERE, 'decay_chain_golden.npz'), **gold)
    print('decay_chain: final RT', gold['final_decay_times'][0], 'heads RT', gold['heads_decay_times'][0],
          'reflection', gold['reflection'][0], 'adj', gold['decay_adjustment'][:, 1])


def decay_window():
    """a11 pin: the reference's own ``apply_decay_window`` (core/decay.py:383-403) on fixed parameters, and its
    ``process_decay_worker`` (core/parallel_workers.py:24-39 = decay_adjustment_params + apply_decay_window) on
    responses of the seeded synthetic 7.1 set."""
    decay = ref.decay
    pw = ref.parallel_workers
    gold = {}
    x = np.random.default_rng(0).standard_normal(20000)
    gold['fixed_params'] = np.array([[100, 3000, 3100, -12.0], [0, 500, 500, -3.5], [48, 9000, 9048, -40.0],
                                     [7, 1, 8, -6.0]])
    for i, prm in enumerate(gold['fixed_params']):
        y = decay.apply_decay_window(x.copy(), (int(prm[0]), int(prm[1]), int(prm[2]), float(prm[3])))
        gold[f'fixed_out_{i}'] = y
    est, h = _cropped_set('final')
    keys = [(s, e) for s in h.irs for e in ('left', 'right')]
    gold['names'] = np.array([f'{s}-{e}' for s, e in keys])
    rows = []
    for target in (0.05, 0.15, 2.0):
        for k, (s, e) in enumerate(keys[:6]):
            data = h.irs[s][e].data
            prm = decay.decay_adjustment_params(data, est.fs, target)
            _, _, out = pw.process_decay_worker((s, e, data, est.fs, target))
            rows.append([target, k, -1 if prm is None else 1])
            tag = f't{int(target * 1000)}_{s}_{e}'
            gold[f'worker_ck_{tag}'] = checksums(out)
            gold[f'worker_dec_{tag}'] = out[::29].copy()
            gold[f'worker_changed_{tag}'] = np.array(not np.array_equal(out, data))
    gold['worker_rows'] = np.array(rows)
    np.savez_compressed(os.path.join(HERE, 'decay_window_golden.npz'), **gold)
    print('decay_window: changed', [bool(gold[k]) for k in gold if k.startswith('worker_changed_')])


def equalize_cases():
    """a17 pin: the reference's own ``FrequencyResponse.equalize`` (autoeq/frequency_response.py:1241-1310) on error
    curves that clip against the gain cap in one, several and no regions (kink windows + k=2 FITPACK spline), with the
    default arguments and with the EQ worker's (max_gain=40, treble 10 kHz .. fs/2)."""
    rng = np.random.default_rng(21)
    f = FrequencyResponse.generate_frequencies(f_step=1.01, f_min=10, f_max=24000)
    lf = np.log10(f)
    curves = [
        -3 * np.sin(2.2 * lf) - 2 * np.cos(5 * lf),                                  # never clipped (|gain| < 6)
        -9 * np.sin(2.2 * lf) - 2 * np.cos(5 * lf),                                  # several clipped regions
        -12 * np.exp(-((lf - 2.3) / 0.15) ** 2) + 1.5 * np.sin(9 * lf),              # one clipped hump
        -20 * (lf < 1.4) + 2 * np.sin(7 * lf),                                       # clipped from the first point on
        -15 * (lf > 4.2) - 3 * np.cos(3 * lf),                                       # clipped up to the last point
        6 * rng.standard_normal(len(f)).cumsum() / np.sqrt(np.arange(1, len(f) + 1)),  # rough curve, many kinks
        -45 * np.exp(-((lf - 3.0) / 0.3) ** 2) + 3 * np.sin(4 * lf),                 # clips even against 40 dB
    ]
    gold = {'frequency': f, 'error_smoothed': np.stack(curves)}
    for tag, kw in (('default', {}), ('worker', dict(max_gain=40, treble_f_lower=10000, treble_f_upper=24000.0)),
                    ('nosmooth', dict(smoothen=False)), ('treble', dict(treble_max_gain=0.0, treble_gain_k=0.5))):
        out = []
        for e in curves:
            fr = FrequencyResponse(name='eq', frequency=f.copy(), raw=np.zeros(len(f)), error=e.copy(),
                                   error_smoothed=e.copy(), smoothed=np.ones(len(f)))
            fr.equalize(**kw)
            out.append(np.asarray(fr.equalization, dtype=np.float64))
            if tag == 'default':
                assert np.allclose(fr.equalized_smoothed, 1.0 + fr.equalization)
        gold[f'eq_{tag}'] = np.stack(out)
    np.savez_compressed(os.path.join(HERE, 'equalize_golden.npz'), **gold)
    clipped = [(int(np.sum(-c > 6))) for c in curves]
    print('equalize_cases: points above the 6 dB cap per curve', clipped)


def small_cases():
    gold = {}
    rng = np.random.default_rng(7)
    for tag, (dur, fs) in {'fs2000': (0.05, 2000), 'fs48000_1s': (1.0, 48000), 'fs44100_2s': (2.0, 44100)}.items():
        est = IRE(min_duration=dur, fs=fs)
        full = fs <= 4000  # large cases keep checksums only (recordings are re-synthesised by the test)
        gold[f'{tag}_n_octaves'] = np.array(est.n_octaves)
        gold[f'{tag}_len'] = np.array(len(est))
        gold[f'{tag}_test_signal_ck'] = checksums(est.test_signal)
        gold[f'{tag}_inverse_filter_ck'] = checksums(est.inverse_filter)
        if full:
            gold[f'{tag}_test_signal'] = est.test_signal.copy()
            gold[f'{tag}_inverse_filter'] = est.inverse_filter.copy()
        # delayed-impulse recordings, as in tests/test_estimator_roundtrip.py
        for delay in (0, 1, 127, 1000):
            if delay >= fs:
                continue
            rec = np.concatenate([np.zeros(delay), est.test_signal, np.zeros(int(0.2 * fs))])
            y = est.estimate(rec)
            gold[f'{tag}_d{delay}_est_ck'] = checksums(y)
            gold[f'{tag}_d{delay}_peak'] = np.array(ImpulseResponse(y, fs).peak_index())
            if full:
                gold[f'{tag}_d{delay}_est'] = y
        noise = np.random.default_rng(11).standard_normal(len(est) + fs // 2) * 0.1
        y = est.estimate(noise)
        gold[f'{tag}_noise_est_ck'] = checksums(y)
        if full:
            gold[f'{tag}_noise_est'] = y
    # peak_index edge cases (plateaus, negative first, nothing above threshold, silence)
    cases = {
        'plateau': np.array([0, 0.2, 1.0, 1.0, 1.0, 0.3, 0, -0.5, 0]),
        'neg_first': np.array([0, -0.1, -0.9, -0.2, 0.0, 1.0, 0.1]),
        'monotone': np.linspace(0, 1, 50),
        'silence': np.zeros(32),
        'tiny': np.array([0.0, 1e-21, 0.0]),
        'edge_peak': np.array([1.0, 0.5, 0.2, 0.1]),
        'plateau_to_end': np.array([0, 0.5, 1.0, 1.0]),
        'even_plateau': np.array([0, 1.0, 1.0, 0]),
        'imp1000': np.concatenate([np.zeros(1000), [1.0], np.zeros(500)]),
    }
    for k, v in cases.items():
        gold[f'pk_{k}_x'] = v.astype(np.float64)
        gold[f'pk_{k}_idx'] = np.array(ImpulseResponse(v.astype(np.float64).copy(), 48000).peak_index())
    # decay_params degenerate inputs (tests/test_impulse_response_decay.py)
    dcases = {
        'short': np.array([1.0, 0.5, 0.2]),
        'delta': np.concatenate([[1.0], np.zeros(4799)]),
        'exp': np.exp(-np.arange(48000) / 3000.0) * np.cos(np.arange(48000) * 0.3) + 1e-4 * rng.standard_normal(48000),
        'noise_only': 1e-3 * rng.standard_normal(20000),
        'lt30ms': np.concatenate([[1.0], 0.01 * rng.standard_normal(1000)]),
    }
    for k, v in dcases.items():
        gold[f'dc_{k}_x'] = v.astype(np.float64)
        gold[f'dc_{k}_out'] = np.array([float(t) for t in ref.decay.decay_params(v.astype(np.float64), 48000)])
    # steepest decays decay_params can measure (0 dB window, then a window just above the -200 dB floor + 10 dB):
    # the refined pass then builds the longest window list it ever can (~3840 of K4's 4096)
    for tag, fs, lev_db in (('steep48', 48000, -187.0), ('steep48b', 48000, -150.0), ('steep192', 192000, -187.0)):
        w = int(0.03 * fs)
        x = np.zeros(3 * fs)
        x[:w] = 1.0
        x[w:2 * w] = 10 ** (lev_db / 20)
        gold[f'st_{tag}_x'] = x
        gold[f'st_{tag}_out'] = np.array([float(t) for t in ref.decay.decay_params(x.copy(), fs)])
    np.savez_compressed(os.path.join(HERE, 'small_cases.npz'), **gold)
    print('small_cases written')


if __name__ == '__main__':
    only = sys.argv[1:]
    for fn in (demo_room, synth71, room_chain, decay_chain, decay_window, equalize_cases, small_cases):
        if not only or fn.__name__ in only:
            fn()
    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, 'KiB')

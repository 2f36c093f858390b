"""Next-row components on the GPU against vectors the unmodified reference produced
(tests/golden/make_golden.py::room_chain): room-correction / headphone frequency-response chains, channel-balance
FIR design + application, and the interaural microphone-mismatch estimator."""
import numpy as np
import pytest

from _util import MINPHASE_TOL_F5, MINPHASE_TOL_F10, assert_ck, load, rel_peak_err

pytestmark = pytest.mark.gpu

# dB curves: K8 magnitudes agree with numpy's rfft to ~1e-12 relative, so a dB value moves by < 1e-9
DB_TOL = 2e-9
# FIRs go through the homomorphic minimum-phase chain (tolerances of tests/test_ops_gpu.py)
FIR_TOL = MINPHASE_TOL_F10


def room_curves(frequency):
    """Same closed forms as tests/golden/make_golden.py::room_curves."""
    lf = np.log10(frequency)
    target = -1.2 * (lf - 3.0) + 1.5 * np.exp(-((lf - 2.0) / 0.35) ** 2)
    mic = 0.8 * np.sin(2.3 * lf) + 0.3 * (lf - 3.0) ** 2
    return target, mic


@pytest.fixture(scope="module")
def G():
    return load('room_chain_golden.npz')


@pytest.fixture(scope="module")
def hrir71():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from impulcifer_pip313_b200 import HRIR, ImpulseResponseEstimator, synth
    est = ImpulseResponseEstimator(min_duration=5.0, fs=48000)
    rec = synth.recording_set(est.test_signal, est.fs, synth.LAYOUT_71, seed=1234)
    h = HRIR(est)
    h.ingest(rec, 48000, ['FL', 'FR', 'FC', 'SL', 'SR', 'BL', 'BR'])
    h.crop_heads()
    h.align_ipsilateral_all()
    h.align_onset_groups_peak_leftref()
    h.crop_tails()
    return h


def _curves(grid):
    from impulcifer_pip313_b200 import FrequencyResponse
    t_raw, m_raw = room_curves(grid)
    return (FrequencyResponse(name='room-target', frequency=grid.copy(), raw=t_raw.copy()),
            FrequencyResponse(name='room-mic', frequency=grid.copy(), raw=m_raw.copy()))


def test_specific_room_corrections(hrir71, G):
    from impulcifer_pip313_b200 import room_correction as rc
    target, mic = _curves(G['frequency'])
    frs = rc.calculate_specific_room_corrections(hrir71.copy(), target.copy(), mic_calibration=mic.copy(), limit=400)
    keys = [tuple(n.split('-')) for n in G['names']]
    assert [(s, e) for s in frs for e in frs[s]] == keys
    for k, (s, e) in enumerate(keys):
        fr = frs[s][e]
        np.testing.assert_array_equal(fr.frequency, G['frequency'])
        np.testing.assert_allclose(fr.raw, G['specific_raw'][k], rtol=0, atol=DB_TOL)
        np.testing.assert_allclose(fr.error, G['specific_error'][k], rtol=0, atol=DB_TOL)
        np.testing.assert_allclose(fr.target, G['specific_target'][k], rtol=0, atol=DB_TOL)
        # above the limit the correction is switched off exactly
        assert np.all(fr.error[fr.frequency > 400] == 0)
    frs0 = rc.calculate_specific_room_corrections(hrir71.copy(), target.copy(), mic_calibration=None, limit=0)
    for k, (s, e) in enumerate(keys):
        np.testing.assert_allclose(frs0[s][e].error, G['specific_nolimit_error'][k], rtol=0, atol=DB_TOL)


@pytest.mark.parametrize("method", ["average", "conservative"])
def test_generic_room_correction(hrir71, G, method):
    from impulcifer_pip313_b200 import ImpulseResponse, room_correction as rc
    target, mic = _curves(G['frequency'])
    keys = [tuple(n.split('-')) for n in G['generic_names']]
    irs = [ImpulseResponse(hrir71.irs[s][e].data.copy(), 48000) for s, e in keys]
    fr = rc.calculate_generic_room_correction(irs, target.copy(), mic_calibration=mic.copy(), method=method, limit=1000)
    for k in ('raw', 'error', 'error_smoothed', 'smoothed', 'target'):
        np.testing.assert_allclose(getattr(fr, k), G[f'generic_{method}_{k}'], rtol=0, atol=DB_TOL, err_msg=k)


def test_generic_room_correction_single_and_errors(hrir71, G):
    from impulcifer_pip313_b200 import ImpulseResponse, room_correction as rc
    target, _ = _curves(G['frequency'])
    irs = [ImpulseResponse(hrir71.irs['FC']['left'].data.copy(), 48000)]
    fr = rc.calculate_generic_room_correction(irs, target.copy(), mic_calibration=None, method='average', limit=0)
    np.testing.assert_allclose(fr.error, G['generic_single_error'], rtol=0, atol=DB_TOL)
    np.testing.assert_allclose(fr.error_smoothed, G['generic_single_error_smoothed'], rtol=0, atol=DB_TOL)
    two = [ImpulseResponse(hrir71.irs['FC'][e].data.copy(), 48000) for e in ('left', 'right')]
    with pytest.raises(ValueError, match='Invalid value'):
        rc.calculate_generic_room_correction(two, target.copy(), method='median')


def test_correction_limit_mask_matches_scipy():
    from scipy import signal
    from impulcifer_pip313_b200 import FrequencyResponse, room_correction as rc
    f = FrequencyResponse.generate_frequencies(f_step=1.01, f_min=10, f_max=24000)
    for limit in (400, 1000, 37.5):
        start, end = np.argmax(f > limit / 2), np.argmax(f > limit)
        want = np.concatenate([np.ones(start), signal.windows.hann(end - start), np.zeros(len(f) - end)])
        np.testing.assert_allclose(rc.correction_limit_mask(f, limit), want, rtol=0, atol=1e-15)


def test_headphone_compensation(hrir71, G):
    from impulcifer_pip313_b200 import room_correction as rc
    left, right = rc.headphone_compensation(hrir71.copy())
    np.testing.assert_allclose(left.error, G['hp_left_error'], rtol=0, atol=DB_TOL)
    np.testing.assert_allclose(right.error, G['hp_right_error'], rtol=0, atol=DB_TOL)
    assert len(left.smoothed) == 0 and len(left.equalization) == 0  # compensate() resets the derived curves


@pytest.mark.parametrize("method", ['trend', 'left', 'right', 'avg', 'min', 'mids', '1.5'])
def test_channel_balance(hrir71, G, method):
    from impulcifer_pip313_b200 import ImpulseResponse
    hb = hrir71.copy()
    lm = np.mean(np.vstack([hb.irs[s]['left'].data for s in ('FL', 'FR')]), axis=0)
    rm = np.mean(np.vstack([hb.irs[s]['right'].data for s in ('FL', 'FR')]), axis=0)
    firs = hb.channel_balance_firs(ImpulseResponse(lm, 48000).frequency_response(),
                                   ImpulseResponse(rm, 48000).frequency_response(), method)
    for fir, side in zip(firs, ('left', 'right')):
        want = G[f'balance_{method}_{side}_fir']
        assert len(fir) == len(want)
        assert rel_peak_err(fir, want) < FIR_TOL, (method, side, rel_peak_err(fir, want))
    ret = hb.correct_channel_balance(method)
    assert ret.irs == {}
    keys = [tuple(n.split('-')) for n in G['names']]
    assert [len(hb.irs[s][e].data) for s, e in keys] == list(G[f'balance_{method}_len'])
    for k, (s, e) in enumerate(keys):
        assert_ck(hb.irs[s][e].data, G[f"balance_{method}_ck"][k], rel=20 * FIR_TOL)


def test_channel_balance_rejects_unknown_method(hrir71):
    from impulcifer_pip313_b200 import ImpulseResponse
    fr = ImpulseResponse(hrir71.irs['FC']['left'].data.copy(), 48000).frequency_response()
    with pytest.raises(ValueError, match='not valid value for channel balance'):
        hrir71.copy().channel_balance_firs(fr, fr.copy(), 'loudest')


def test_microphone_mismatch(hrir71, G):
    from impulcifer_pip313_b200.microphone_deviation import (MicrophoneMatchingCorrector,
                                                             apply_microphone_deviation_correction_to_hrir)
    names = list(hrir71.irs)
    for anchor in ('auto', 'diffuse'):
        c = MicrophoneMatchingCorrector(48000, anchor=anchor)
        np.testing.assert_array_equal(c.frequency, G['mic_frequency'])
        c.collect_hrir(hrir71, {s: (hrir71.irs[s]['left'].peak_index(), hrir71.irs[s]['right'].peak_index())
                                for s in names})
        if anchor == 'auto':
            pl = np.array([c.speaker_power[s]['left'] for s in names])
            pr = np.array([c.speaker_power[s]['right'] for s in names])
            np.testing.assert_allclose(pl, G['mic_power_left'], rtol=1e-9, atol=1e-30)
            np.testing.assert_allclose(pr, G['mic_power_right'], rtol=1e-9, atol=1e-30)
        m = c.estimate_interaural_mismatch()
        assert c.anchor_used == str(G[f'mic_{anchor}_used'])
        np.testing.assert_allclose(m, G[f'mic_{anchor}_mismatch'], rtol=0, atol=DB_TOL)
        lf, rf = c.design_correction_filters()
        assert len(lf) == len(G[f'mic_{anchor}_left_fir'])
        assert rel_peak_err(lf, G[f'mic_{anchor}_left_fir']) < FIR_TOL
        assert rel_peak_err(rf, G[f'mic_{anchor}_right_fir']) < FIR_TOL
    # the per-speaker entry point agrees with the batched one
    c1 = MicrophoneMatchingCorrector(48000)
    li, ri = hrir71.irs['SL']['left'], hrir71.irs['SL']['right']
    p = c1.collect_speaker('SL', li.data, ri.data, li.peak_index(), ri.peak_index())
    np.testing.assert_array_equal(p['left'], c.speaker_power['SL']['left'])
    hm = hrir71.copy()
    summary = apply_microphone_deviation_correction_to_hrir(hm, correction_strength=0.7, anchor='auto')
    assert summary['anchor'] == 'frontal' and summary['speakers_processed'] == names
    keys = [tuple(n.split('-')) for n in G['names']]
    assert [len(hm.irs[s][e].data) for s, e in keys] == list(G['mic_applied_len'])
    for k, (s, e) in enumerate(keys):
        assert_ck(hm.irs[s][e].data, G["mic_applied_ck"][k], rel=20 * FIR_TOL)

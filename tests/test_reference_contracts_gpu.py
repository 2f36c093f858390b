"""The contracts the reference's own hot-path tests pin, asserted on the B200 shim classes:
estimator round trip (tests/test_estimator_roundtrip.py), normalize / align / shift (tests/test_dsp_stages.py),
crop_tails (tests/test_dsp_characterization.py:192-224), decay degenerate inputs
(tests/test_impulse_response_decay.py), minimum-phase sanity (tests/test_frequency_response_core.py:101-112).
Same stimuli and thresholds as the reference tests; the code under test is the device path."""
import numpy as np
import pytest
from scipy import signal

pytestmark = pytest.mark.gpu

FS = 48000


@pytest.fixture(scope="module")
def est1s():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from impulcifer_pip313_b200 import ImpulseResponseEstimator
    return ImpulseResponseEstimator(min_duration=1.0, fs=FS)


class _FsOnly:
    fs = FS


def _hrir(channels):
    from impulcifer_pip313_b200 import HRIR
    h = HRIR(_FsOnly())
    h.irs = channels
    return h


def _ir(data):
    from impulcifer_pip313_b200 import ImpulseResponse
    return ImpulseResponse(np.asarray(data, dtype=np.float64), FS)


# ---- estimator round trip ------------------------------------------------------------------------------------
@pytest.mark.parametrize("delay", [0, 1, 127, 1000])
def test_delayed_unit_impulse_moves_the_peak_by_the_delay(est1s, delay):
    base = int(np.argmax(np.abs(est1s.estimate(est1s.test_signal))))
    system = np.zeros(delay + 1)
    system[delay] = 1.0
    rec = signal.convolve(est1s.test_signal, system, mode="full")
    got = est1s.estimate(rec)
    assert isinstance(got, np.ndarray) and got.dtype == np.float64 and len(got) == len(rec)
    assert int(np.argmax(np.abs(got))) == base + delay


def test_unit_impulse_level_and_noise_floor(est1s):
    got = est1s.estimate(est1s.test_signal)
    peak = int(np.argmax(np.abs(got)))
    outside = np.concatenate((got[:peak - 50], got[peak + 50:]))
    assert np.abs(got[peak]) == pytest.approx(1.0, abs=0.05)
    assert 20 * np.log10(np.abs(got[peak]) / np.max(np.abs(outside))) > 40.0


def test_decaying_response_is_recovered(est1s):
    base = int(np.argmax(np.abs(est1s.estimate(est1s.test_signal))))
    n, d = 600, 31
    k = np.arange(n)[d:]
    ir = np.zeros(n)
    ir[d:] = 0.08 * np.exp(-(k - d) / 110.0) * np.cos(2 * np.pi * 1700 * (k - d) / FS)
    ir[d] = 1.0
    got = est1s.estimate(signal.convolve(est1s.test_signal, ir, mode="full"))
    rec = got[base:base + n]
    assert int(np.argmax(np.abs(rec))) == d
    assert np.dot(rec, ir) / (np.linalg.norm(rec) * np.linalg.norm(ir)) > 0.99
    far = np.concatenate((got[:max(base - 1000, 0)], got[base + n + 1000:]))
    assert 20 * np.log10(np.max(np.abs(rec)) / np.max(np.abs(far))) > 40.0


# ---- normalize / align / shift ----------------------------------------------------------------------------------
def _summed(h):
    from impulcifer_pip313_b200.audio_io import magnitude_response
    left = np.sum(np.vstack([p["left"].data for p in h.irs.values()]), axis=0)
    right = np.sum(np.vstack([p["right"].data for p in h.irs.values()]), axis=0)
    return magnitude_response(left, FS), magnitude_response(right, FS)


def test_normalize_peak_and_average_targets(est1s):
    rng = np.random.default_rng(1)
    h = _hrir({"FL": {"left": _ir(rng.standard_normal(4096) * 0.3), "right": _ir(rng.standard_normal(4096) * 0.1)},
               "FR": {"left": _ir(rng.standard_normal(4096) * 0.05), "right": _ir(rng.standard_normal(4096) * 0.2)}})
    gain = h.normalize(peak_target=-0.1)
    assert isinstance(gain, float)
    (_, ml), (_, mr) = _summed(h)
    assert abs(float(np.max(np.vstack([ml, mr]))) + 0.1) < 1e-6
    rng = np.random.default_rng(2)
    h = _hrir({"FL": {"left": _ir(rng.standard_normal(4096) * 0.4), "right": _ir(rng.standard_normal(4096) * 0.25)},
               "FR": {"left": _ir(rng.standard_normal(4096) * 0.15), "right": _ir(rng.standard_normal(4096) * 0.3)}})
    h.normalize(peak_target=None, avg_target=-12.0)
    (fl, ml), (fr, mr) = _summed(h)
    band = np.concatenate([ml[np.logical_and(fl > 80, fl < 6000)], mr[np.logical_and(fr > 80, fr < 6000)]])
    assert abs(float(np.mean(band)) + 12.0) < 1e-6
    with pytest.raises(ValueError):
        _hrir({"FL": {"left": _ir([1.0, 0.0]), "right": _ir([1.0, 0.0])}}).normalize(peak_target=-0.1, avg_target=-12.0)


def test_align_ipsilateral_removes_the_lag_and_keeps_lengths(est1s):
    def impulse_at(i, n=2048):
        d = np.zeros(n)
        d[i] = 1.0
        return _ir(d)

    def lag(a, b, seg=1440):
        corr = signal.correlate(a[:seg], b[:seg], mode="full")
        return int(np.arange(-seg + 1, seg)[np.argmax(corr)])

    h = _hrir({"FL": {"left": impulse_at(60), "right": impulse_at(60)},
               "FR": {"left": impulse_at(60), "right": impulse_at(67)}})
    assert abs(lag(h.irs["FL"]["left"].data, h.irs["FR"]["right"].data)) == 7
    h.align_ipsilateral_all(speaker_pairs=[("FL", "FR")], segment_ms=30)
    assert abs(lag(h.irs["FL"]["left"].data, h.irs["FR"]["right"].data)) <= 1
    assert all(len(ir.data) == 2048 for pair in h.irs.values() for ir in pair.values())


def test_shift_is_exact(est1s):
    ir = _ir([1.0, 2.0, 3.0, 4.0])
    ir.shift(2)
    np.testing.assert_array_equal(ir.data, [0.0, 0.0, 1.0, 2.0])
    ir = _ir([1.0, 2.0, 3.0, 4.0])
    ir.shift(-1)
    np.testing.assert_array_equal(ir.data, [2.0, 3.0, 4.0, 0.0])
    ir.shift(0)
    np.testing.assert_array_equal(ir.data, [2.0, 3.0, 4.0, 0.0])


# ---- crop_tails / decay ----------------------------------------------------------------------------------------
def test_crop_tails_uniform_length_and_faded_end(est1s):
    from impulcifer_pip313_b200 import HRIR
    rng = np.random.default_rng(4)
    n = 2 * FS + len(est1s)

    def decaying(onset, rt):
        x = np.zeros(n)
        m = n - onset
        x[onset:] = rng.standard_normal(m) * 10 ** (-3 * np.arange(m) / (FS * rt))
        x[onset] = 2.0
        return _ir(x + 1e-6 * rng.standard_normal(n))

    h = HRIR(est1s)
    h.irs = {"FL": {"left": decaying(100, 0.2), "right": decaying(110, 0.25)},
             "FR": {"left": decaying(112, 0.3), "right": decaying(101, 0.22)}}
    tail = h.crop_tails()
    assert isinstance(tail, int) and tail > 0
    for pair in h.irs.values():
        for ir in pair.values():
            assert len(ir.data) == tail and abs(ir.data[-1]) < 1e-6
    h2 = HRIR(est1s)
    h2.irs = {"FL": {"left": decaying(100, 0.2), "right": decaying(100, 0.2)}}
    h2.fs = 44100
    with pytest.raises(ValueError):
        h2.crop_tails()


def test_decay_params_degenerate_inputs(est1s):
    for x in (np.zeros(4096), np.ones(5) * 0.1, np.r_[1.0, np.zeros(4095)]):
        pk, knee, noise, win = _ir(x).decay_params()
        assert isinstance(pk, int) and isinstance(knee, int) and np.isfinite(noise) and win >= 1
        assert 0 <= pk <= len(x) and 0 <= knee <= len(x)
    edt, rt20, rt30, rt60 = _ir(np.zeros(4096)).decay_times()
    assert edt is None and rt60 is None


# ---- minimum phase --------------------------------------------------------------------------------------------------
def test_minimum_phase_fir_is_real_finite_and_front_loaded(est1s):
    from impulcifer_pip313_b200 import FrequencyResponse
    f = FrequencyResponse.generate_frequencies(f_min=10, f_max=24000, f_step=1.01)
    fr = FrequencyResponse(name="eq", frequency=f, raw=np.zeros(len(f)))
    fr.equalization = 4.0 * np.sin(np.log10(f) * 2.0)
    fir = fr.minimum_phase_impulse_response(fs=FS, f_res=10, normalize=True)
    assert fir.dtype == np.float64 and np.all(np.isfinite(fir)) and np.isrealobj(fir)
    assert np.sum(fir[:len(fir) // 2] ** 2) > 0.5 * np.sum(fir ** 2)

"""``FrequencyResponse`` — the hot-path subset of the reference's vendored AutoEQ class
(``autoeq/frequency_response.py``), same method names, argument meaning, side effects and error messages,
with smoothing, log-frequency interpolation and minimum-phase FIR design running on the B200.

Host side ("the Python shim owns window-size arithmetic, object construction,
reset() side effects, warnings/exceptions"): ``_window_size``, ``_sigmoid``, ``reset``, ``equalize``'s
kink bookkeeping.  Device side: ``_smoothen_fractional_octave`` (K11), ``interpolate`` /
``center`` (K10), ``minimum_phase_impulse_response`` (K10 + K12 + K13).

The free functions ``smoothen_batch`` / ``minimum_phase_batch`` are what the ``BatchDispatcher`` uses to
run every (speaker, side) curve of many recordings in one launch; the methods call them with batch 1.
"""
from __future__ import annotations

import functools
import math
import warnings

import numpy as np

from . import ops

DEFAULT_F_MIN = 20
DEFAULT_F_MAX = 20000
DEFAULT_STEP = 1.01
DEFAULT_MAX_GAIN = 6.0
DEFAULT_TREBLE_F_LOWER = 6000.0
DEFAULT_TREBLE_F_UPPER = 8000.0
DEFAULT_TREBLE_MAX_GAIN = 6.0
DEFAULT_TREBLE_GAIN_K = 1.0
DEFAULT_SMOOTHING_WINDOW_SIZE = 1 / 3
DEFAULT_SMOOTHING_ITERATIONS = 1
DEFAULT_TREBLE_SMOOTHING_F_LOWER = 100.0
DEFAULT_TREBLE_SMOOTHING_F_UPPER = 10000.0
DEFAULT_TREBLE_SMOOTHING_WINDOW_SIZE = 1 / 3
DEFAULT_TREBLE_SMOOTHING_ITERATIONS = 1
DEFAULT_FS = 44100
DEFAULT_F_RES = 10

_KEYS = 'raw error error_smoothed equalization equalized_raw equalized_smoothed target'.split()


def _device_index():
    from ._lib import default_device
    return default_device()


def _to_dev(a):
    import torch
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=f"cuda:{_device_index()}")


# ---- host-side scalar helpers (reference :1033-1058) ----------------------------------------------------
@functools.lru_cache(maxsize=64)
def _frequency_grid(f_min, f_max, f_step):
    """The geometric grid of ``generate_frequencies`` (running product, as the reference builds it), built once."""
    freq = []
    f = f_min
    while f <= f_max:
        freq.append(f)
        f *= f_step
    return np.array(freq)


@functools.lru_cache(maxsize=256)
def _window_size_cached(grid_bytes, octaves) -> int:
    frequency = np.frombuffer(grid_bytes, dtype=np.float64)
    # the reference adds the ratios one by one (Python sum); cumsum accumulates in the same order, unlike np.sum
    steps = frequency[1:] / frequency[:-1]
    step_size = float(np.cumsum(steps)[-1]) / len(steps)
    w = round(math.log(2 ** octaves) / math.log(step_size))
    return w if w % 2 else w + 1


def window_size(frequency, octaves) -> int:
    """Savitzky-Golay window of ``octaves`` on this grid (reference ``_window_size``, :1033-1049); the 783-iteration
    Python loop of the reference ran twice per smoothing call, on the host, in front of a 10 us kernel."""
    return _window_size_cached(np.ascontiguousarray(frequency, dtype=np.float64).tobytes(), float(octaves))


def sigmoid(frequency, f_lower, f_upper, a_normal=0.0, a_treble=1.0):
    from scipy.special import expit
    f_center = np.sqrt(f_upper / f_lower) * f_lower
    half_range = np.log10(f_upper) - np.log10(f_center)
    a = expit((np.log10(frequency) - np.log10(f_center)) / (half_range / 4))
    return a * -(a_normal - a_treble) + a_normal


_SMOOTH5 = None


def _smooth5_table():
    """Sorted 5-smooth numbers up to 2**26 (host constant, built once)."""
    global _SMOOTH5
    if _SMOOTH5 is None:
        lim = 1 << 26
        vals = []
        p5 = 1
        while p5 <= lim:
            p35 = p5
            while p35 <= lim:
                v = p35
                while v <= lim:
                    vals.append(v)
                    v *= 2
                p35 *= 3
            p5 *= 5
        _SMOOTH5 = np.array(sorted(vals), dtype=np.int64)
    return _SMOOTH5


def next_fast_len(target):
    """scipy.fftpack.next_fast_len: smallest 2^a 3^b 5^c >= target (scalar or array)."""
    tab = _smooth5_table()
    t = np.maximum(np.asarray(target, dtype=np.int64), 1)
    if np.any(t > tab[-1]):
        raise ValueError("length too large")
    out = tab[np.searchsorted(tab, t, side='left')]
    return int(out) if out.ndim == 0 else out


# ---- batched device entry points ------------------------------------------------------------------------
def smoothen_batch(frequency, data, window_size_oct=DEFAULT_SMOOTHING_WINDOW_SIZE, iterations=1,
                   treble_window_size_oct=None, treble_iterations=None, treble_f_lower=100.0, treble_f_upper=10000.0):
    """K11 on rows of ``data`` (numpy [B, F] or CUDA tensor); returns the same kind of container."""
    import torch
    frequency = np.asarray(frequency, dtype=np.float64)
    is_np = isinstance(data, np.ndarray)
    x = _to_dev(np.atleast_2d(data)) if is_np else (data if data.dim() == 2 else data[None, :])
    wn = window_size(frequency, window_size_oct)
    wt = window_size(frequency, treble_window_size_oct)
    k_treble = sigmoid(frequency, treble_f_lower, treble_f_upper)
    y = ops.savgol_blend(x, wn, iterations, wt, treble_iterations, k_treble)
    return y.cpu().numpy() if is_np else y


def equalize_batch(frequency, error, max_gain=DEFAULT_MAX_GAIN, smoothen=True, treble_f_lower=DEFAULT_TREBLE_F_LOWER,
                   treble_f_upper=DEFAULT_TREBLE_F_UPPER, treble_max_gain=DEFAULT_TREBLE_MAX_GAIN,
                   treble_gain_k=DEFAULT_TREBLE_GAIN_K, return_status=False):
    """K21 on rows of (smoothed) error curves that share ``frequency``: the ``equalization`` curves of
    ``FrequencyResponse.equalize`` (reference :1241-1305).  numpy in -> numpy out, CUDA tensor in -> CUDA tensor out.
    The two log-frequency sigmoids (gain cap, gain factor) and the 1/12-octave kink window are host constants of the
    grid, exactly as the reference derives them."""
    frequency = np.asarray(frequency, dtype=np.float64)
    is_np = isinstance(error, np.ndarray)
    if is_np and np.any(np.isnan(error)):
        raise ValueError('NaN values detected during equalization, interpolating data with default parameters.')
    e = _to_dev(np.atleast_2d(error)) if is_np else (error if error.dim() == 2 else error[None, :])
    cap = sigmoid(frequency, treble_f_lower, treble_f_upper, a_normal=max_gain, a_treble=treble_max_gain)
    gain_k = sigmoid(frequency, treble_f_lower, treble_f_upper, a_normal=1.0, a_treble=treble_gain_k)
    half = (window_size(frequency, 1 / 12) - 1) // 2
    eq, status = ops.equalize(e, np.log10(frequency), cap, gain_k, half, smoothen)
    if return_status:  # the caller checks (and synchronises) when it reads its results
        return (eq.cpu().numpy(), status) if is_np else (eq, status)
    if is_np:
        if int(status.max().item()) != 0:  # fewer than 3 points survive the kink windows: dfitpack's error
            raise ValueError("equalize: too few points left for the k=2 interpolating spline (m > k must hold)")
        return eq.cpu().numpy()
    return eq


def minphase_gain_grid(frequency, equalization, fs, f_res, normalize):
    """Host part of minimum_phase_impulse_response (reference :653-676) for rows of ``equalization``:
    log-f interpolation onto linspace(0, fs//2, n) on the device, then the dB -> linear shaping.
    Returns a CUDA tensor [B, n]."""
    import torch
    frequency = np.asarray(frequency, dtype=np.float64)
    f_res = f_res / 2
    f_min = np.max([frequency[0], f_res])
    n = next_fast_len(round(fs // 2 / f_res))
    f = np.linspace(0.0, fs // 2, n)
    fq = f.copy()
    if fq[0] == 0:
        fq[0] = 0.001
    eq = equalization if isinstance(equalization, torch.Tensor) else _to_dev(np.atleast_2d(equalization))
    # f ascends, so "f <= f_min" is a prefix; the interpolation, the prefix fill, [normalize: - max - 0.5], * 2,
    # 10^(./20) and the zero at Nyquist run as one pair of launches for all rows
    n_below = int(np.count_nonzero(f <= f_min))
    return ops.minphase_gain_grid(np.log10(frequency), eq, np.log10(fq), n_below, np.log10(f_min), normalize)


def minimum_phase_batch(frequency, equalization, fs=DEFAULT_FS, f_res=DEFAULT_F_RES, normalize=True):
    """K10 + K12 + K13 on rows of ``equalization``; numpy in -> numpy out, CUDA tensor in -> CUDA tensor out."""
    is_np = isinstance(equalization, np.ndarray)
    g = minphase_gain_grid(frequency, equalization, fs, f_res, normalize)
    fir = ops.minphase_fir(g, fs)
    return fir.cpu().numpy() if is_np else fir


class FrequencyResponse:
    def __init__(self, name=None, frequency=None, raw=None, error=None, smoothed=None, error_smoothed=None,
                 equalization=None, parametric_eq=None, fixed_band_eq=None, equalized_raw=None,
                 equalized_smoothed=None, target=None):
        if not name:
            raise TypeError('Name must not be a non-empty string.')
        self.name = name.strip()
        self.frequency = self._init_data(frequency)
        if not len(self.frequency):
            self.frequency = self.generate_frequencies()
        self.raw = self._init_data(raw)
        self.smoothed = self._init_data(smoothed)
        self.error = self._init_data(error)
        self.error_smoothed = self._init_data(error_smoothed)
        self.equalization = self._init_data(equalization)
        self.parametric_eq = self._init_data(parametric_eq)
        self.fixed_band_eq = self._init_data(fixed_band_eq)
        self.equalized_raw = self._init_data(equalized_raw)
        self.equalized_smoothed = self._init_data(equalized_smoothed)
        self.target = self._init_data(target)
        self._sort()

    def copy(self, name=None):
        return type(self)(  # the merged class after install(), not the bare shim
            name=self.name + '_copy' if name is None else name, frequency=self._init_data(self.frequency),
            raw=self._init_data(self.raw), error=self._init_data(self.error), smoothed=self._init_data(self.smoothed),
            error_smoothed=self._init_data(self.error_smoothed), equalization=self._init_data(self.equalization),
            parametric_eq=self._init_data(self.parametric_eq), fixed_band_eq=self._init_data(self.fixed_band_eq),
            equalized_raw=self._init_data(self.equalized_raw),
            equalized_smoothed=self._init_data(self.equalized_smoothed), target=self._init_data(self.target))

    def _init_data(self, data):
        if data is None:
            return np.array([])
        if isinstance(data, (float, int)) and not isinstance(data, bool):
            data = np.ones(self.frequency.shape) * data
        arr = np.asarray(data)
        if arr.dtype.kind in 'fiub':
            return np.array(arr, dtype=np.float64 if arr.dtype.kind != 'f' else arr.dtype)
        return np.array([np.nan if x is None else x for x in data], dtype=np.float64)

    def _sort(self):
        inds = self.frequency.argsort()
        self.frequency = self.frequency[inds]
        for k in ['raw', 'error', 'smoothed', 'error_smoothed', 'equalization', 'parametric_eq', 'fixed_band_eq',
                  'equalized_raw', 'equalized_smoothed', 'target']:
            v = getattr(self, k)
            if len(v):
                setattr(self, k, v[inds])

    def reset(self, raw=False, smoothed=True, error=True, error_smoothed=True, equalization=True, fixed_band_eq=True,
              parametric_eq=True, equalized_raw=True, equalized_smoothed=True, target=True):
        for key, flag in dict(raw=raw, smoothed=smoothed, error=error, error_smoothed=error_smoothed,
                              equalization=equalization, parametric_eq=parametric_eq, fixed_band_eq=fixed_band_eq,
                              equalized_raw=equalized_raw, equalized_smoothed=equalized_smoothed,
                              target=target).items():
            if flag:
                setattr(self, key, np.array([]))

    @staticmethod
    def generate_frequencies(f_min=DEFAULT_F_MIN, f_max=DEFAULT_F_MAX, f_step=DEFAULT_STEP):
        return _frequency_grid(f_min, f_max, f_step).copy()

    # ---- K10 -----------------------------------------------------------------------------------------
    def interpolate(self, f=None, f_step=DEFAULT_STEP, pol_order=1, f_min=DEFAULT_F_MIN, f_max=DEFAULT_F_MAX):
        """Reference :859-901.  Linear (pol_order=1) interpolation in log-f runs on the device."""
        if len(self.raw):
            raw = np.asarray(self.raw, dtype=float)
            valid = ~np.isnan(raw)
            if not valid.all():
                self.raw = raw[valid]
                self.frequency = self.frequency[valid]
        if pol_order != 1:
            raise NotImplementedError("only pol_order=1 (the order every hot-path caller uses) is implemented")
        present = [k for k in _KEYS if len(getattr(self, k))]
        log_f_in = np.log10(self.frequency)
        stacked = np.stack([getattr(self, k) for k in present]) if present else None
        self.frequency = self.generate_frequencies(f_min=f_min, f_max=f_max, f_step=f_step) if f is None else np.array(f)
        zero_fix = False
        if self.frequency[0] == 0:
            self.frequency[0] = 0.001
            zero_fix = True
        if present:
            out = ops.loglerp(log_f_in, _to_dev(stacked), np.log10(self.frequency)).cpu().numpy()
            for k, row in zip(present, out):
                setattr(self, k, row)
        if zero_fix:
            self.frequency[0] = 0
        self.reset(**{k: False for k in _KEYS})

    def center(self, frequency=1000):
        """Reference :903-940."""
        ee = type(self)(name='equal_energy', frequency=self.frequency.copy(), raw=self.raw.copy())
        ee.interpolate()
        if isinstance(frequency, (list, np.ndarray)) and len(frequency) > 1:
            diff = np.mean(ee.raw[np.logical_and(ee.frequency >= frequency[0], ee.frequency <= frequency[1])])
        else:
            if isinstance(frequency, (list, np.ndarray)):
                frequency = frequency[0]
            diff = ops.loglerp(np.log10(ee.frequency), _to_dev(ee.raw[None, :]),
                               np.log10(np.array([float(frequency)]))).cpu().numpy()[0, 0]
        self.raw -= diff
        if len(self.smoothed):
            self.smoothed -= diff
        if len(self.error):
            self.error += diff
        if len(self.error_smoothed):
            self.error_smoothed += diff
        self.reset(raw=False, smoothed=False, error=False, error_smoothed=False, target=False)
        return -diff

    def compensate(self, compensation, bass_boost_gain=0.0, bass_boost_fc=105.0, bass_boost_q=0.71, tilt=None,
                   sound_signature=None, min_mean_error=False):
        """Reference :984-1031 (flat bass shelf only: the biquad target synthesis is outside the hot path)."""
        if bass_boost_gain:
            raise NotImplementedError("bass-boost target synthesis (autoeq/biquad.py) is out of the hot-path scope")
        comp = FrequencyResponse(name='compensation', frequency=compensation.frequency, raw=compensation.raw)
        comp.center()
        tgt = np.zeros(len(self.frequency))
        if tilt is not None:
            c = DEFAULT_F_MIN * np.sqrt(DEFAULT_F_MAX / DEFAULT_F_MIN)
            tgt = np.log2(self.frequency / c) * tilt
        self.target = comp.raw + tgt
        if sound_signature is not None:
            if not np.all(sound_signature.frequency == self.frequency):
                sound_signature.interpolate(self.frequency)
            self.target += sound_signature.raw
        self.error = self.raw - self.target
        if min_mean_error:
            delta = np.mean(self.error[np.logical_and(self.frequency >= 100, self.frequency <= 10000)])
            self.error -= delta
            self.target += delta
        self.reset(raw=False, smoothed=False, error=False, error_smoothed=True, equalization=True, parametric_eq=True,
                   fixed_band_eq=True, equalized_raw=True, equalized_smoothed=True, target=False)

    # ---- K11 -----------------------------------------------------------------------------------------
    def _window_size(self, octaves):
        return window_size(self.frequency, octaves)

    def _sigmoid(self, f_lower, f_upper, a_normal=0.0, a_treble=1.0):
        return sigmoid(self.frequency, f_lower, f_upper, a_normal, a_treble)

    def _smoothen_fractional_octave(self, data, window_size=DEFAULT_SMOOTHING_WINDOW_SIZE,
                                    iterations=DEFAULT_SMOOTHING_ITERATIONS, treble_window_size=None,
                                    treble_iterations=None, treble_f_lower=DEFAULT_TREBLE_SMOOTHING_F_LOWER,
                                    treble_f_upper=DEFAULT_TREBLE_SMOOTHING_F_UPPER):
        """Reference :1060-1105."""
        if np.any(np.isnan(np.asarray(self.frequency, dtype=float))) or np.any(np.isnan(np.asarray(data, dtype=float))):
            raise ValueError('NaN values present, cannot smoothen!')
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return smoothen_batch(self.frequency, np.asarray(data, dtype=np.float64)[None, :], window_size, iterations,
                                  treble_window_size, treble_iterations, treble_f_lower, treble_f_upper)[0]

    def smoothen_fractional_octave(self, window_size=DEFAULT_SMOOTHING_WINDOW_SIZE,
                                   iterations=DEFAULT_SMOOTHING_ITERATIONS,
                                   treble_window_size=DEFAULT_TREBLE_SMOOTHING_WINDOW_SIZE,
                                   treble_iterations=DEFAULT_TREBLE_SMOOTHING_ITERATIONS,
                                   treble_f_lower=DEFAULT_TREBLE_SMOOTHING_F_LOWER,
                                   treble_f_upper=DEFAULT_TREBLE_SMOOTHING_F_UPPER):
        """Reference :1107-1164: raw -> smoothed, error -> error_smoothed (one launch for both curves)."""
        if treble_f_upper <= treble_f_lower:
            raise ValueError('Upper transition boundary must be greater than lower boundary')
        curves = [self.raw] + ([self.error] if len(self.error) else [])
        if np.any(np.isnan(np.asarray(self.frequency, dtype=float))) or any(np.any(np.isnan(c)) for c in curves):
            raise ValueError('NaN values present, cannot smoothen!')
        out = smoothen_batch(self.frequency, np.stack([np.asarray(c, dtype=np.float64) for c in curves]), window_size,
                             iterations, treble_window_size, treble_iterations, treble_f_lower, treble_f_upper)
        self.smoothed = out[0]
        if len(self.error):
            self.error_smoothed = out[1]
        self.reset(raw=False, smoothed=False, error=False, error_smoothed=False, equalization=True, parametric_eq=True,
                   fixed_band_eq=True, equalized_raw=True, equalized_smoothed=True, target=False)

    def smoothen(self, window_size=DEFAULT_SMOOTHING_WINDOW_SIZE, treble_window_size=DEFAULT_TREBLE_SMOOTHING_WINDOW_SIZE,
                 treble_f_lower=DEFAULT_TREBLE_SMOOTHING_F_LOWER, treble_f_upper=DEFAULT_TREBLE_SMOOTHING_F_UPPER):
        """Reference :1166-1179."""
        self.smoothen_fractional_octave(window_size=window_size, iterations=1, treble_window_size=treble_window_size,
                                        treble_iterations=1, treble_f_lower=treble_f_lower,
                                        treble_f_upper=treble_f_upper)

    def smoothen_heavy_light(self):
        """Reference :1181-1239."""
        light = self._smoothen_fractional_octave(self.error, window_size=1 / 6, iterations=1, treble_f_lower=100,
                                                 treble_f_upper=10000, treble_window_size=1 / 3, treble_iterations=1)
        heavy = self._smoothen_fractional_octave(self.error, window_size=1 / 3, iterations=1, treble_f_lower=1000,
                                                 treble_f_upper=6000, treble_window_size=1.3, treble_iterations=1)
        combination = np.max(np.vstack([light, heavy]), axis=0)
        both = smoothen_batch(self.frequency, np.stack([np.asarray(self.raw, dtype=np.float64), combination]), 1 / 3, 1,
                              1 / 3, 1, 100, 10000)
        self.smoothed, self.error_smoothed = both[0], both[1]
        self.reset(raw=False, smoothed=False, error=False, error_smoothed=False, equalization=True, parametric_eq=True,
                   fixed_band_eq=True, equalized_raw=True, equalized_smoothed=True, target=False)

    # ---- a17 = K21 ----------------------------------------------------------------------------------------
    def equalize(self, max_gain=DEFAULT_MAX_GAIN, smoothen=True, treble_f_lower=DEFAULT_TREBLE_F_LOWER,
                 treble_f_upper=DEFAULT_TREBLE_F_UPPER, treble_max_gain=DEFAULT_TREBLE_MAX_GAIN,
                 treble_gain_k=DEFAULT_TREBLE_GAIN_K):
        """Reference :1241-1310 (clip to the sigmoid gain cap, drop the kink windows, k=2 spline over log f) on the
        device (``impb_equalize``)."""
        self.equalization = []
        self.equalized_raw = []
        if len(self.error_smoothed):
            error = self.error_smoothed
        elif len(self.error):
            error = self.error
        else:
            raise ValueError('Error data is missing. Call FrequencyResponse.compensate().')
        error = np.asarray(error, dtype=float)
        if np.any(np.isnan(error)):
            raise ValueError('NaN values detected during equalization, interpolating data with default parameters.')
        self.equalization = equalize_batch(self.frequency, error[None, :], max_gain, smoothen, treble_f_lower,
                                           treble_f_upper, treble_max_gain, treble_gain_k)[0]
        self.equalized_raw = self.raw + self.equalization
        if len(self.smoothed):
            self.equalized_smoothed = self.smoothed + self.equalization

    # ---- K10 + K12 + K13 --------------------------------------------------------------------------------
    def minimum_phase_impulse_response(self, fs=DEFAULT_FS, f_res=DEFAULT_F_RES, normalize=True):
        """Reference :637-681."""
        return minimum_phase_batch(self.frequency, np.asarray(self.equalization, dtype=np.float64)[None, :], fs=fs,
                                   f_res=f_res, normalize=normalize)[0]

"""``ImpulseResponse`` — drop-in for ``core/impulse_response.py:14-188`` of the reference.

``.data`` stays a real ``numpy.float64`` array at every API boundary (plotters, README writer and tests index
it); device residency is a lazy cache behind it: device methods leave their result on the GPU and the
host array is materialised the next time ``.data`` is read.  Reading ``.data`` hands out a mutable array,
so it also drops the device copy (the caller may edit it in place, as ``HRIR.normalize`` does).
"""
from __future__ import annotations

import functools
import math
from copy import deepcopy

import numpy as np

from . import ops
from .frequency_response import FrequencyResponse

EPSILON = 1e-20


def _device_index():
    from ._lib import default_device
    return default_device()


@functools.lru_cache(maxsize=16)
def resample_taps(fs_new: int, fs_old: int, n_taps: int = 16001, beta: float = 5.0):
    """(up, down, taps*up): the "null-on-Nyquist" Kaiser low-pass of the third-party ``nnresample`` package
    (``nnresample.resample(data, fs_new, fs_old)``, called at reference core/impulse_response.py:123) as
    ``scipy.signal.resample_poly`` applies it.  Host-side constant, computed once per rate pair.
    Parity for this filter is UNPINNED (package absent from the image; algorithm restated)."""
    from scipy import signal
    g = math.gcd(int(fs_new), int(fs_old))
    up, down = int(fs_new) // g, int(fs_old) // g
    max_rate = max(up, down)
    sfact = np.sqrt(1 + (beta / np.pi) ** 2)
    first = signal.firwin(n_taps, 1 / max_rate, window=('kaiser', beta))
    n_fft = 2 ** 19
    nbins = n_fft / 2 + 1
    spec = np.fft.rfft(first, n_fft)
    bot = int(np.floor(nbins / max_rate))
    top = int(np.ceil(nbins * (1 / max_rate + 2 * sfact / n_taps)))
    firstnull = (np.argmin(np.abs(spec[bot:top])) + bot) / nbins
    taps = signal.firwin(n_taps, -firstnull + 2 / max_rate, window=('kaiser', beta))
    return up, down, np.ascontiguousarray(taps * up)


class ImpulseResponse:
    def __init__(self, data, fs, recording=None):
        self.fs = fs
        self._host = None
        self._dev = None
        self.data = data
        self.recording = recording

    # ---- host / device views ---------------------------------------------------------------------------
    @property
    def data(self):
        if self._host is None:
            self._host = self._dev.detach().cpu().numpy()
        self._dev = None  # the caller may mutate the array it gets
        return self._host

    @data.setter
    def data(self, value):
        try:
            import torch
            if isinstance(value, torch.Tensor):
                self._dev = value
                self._host = None
                return
        except ImportError:
            pass
        self._host = value
        self._dev = None

    def device_data(self):
        """1-D float64 CUDA tensor with the current samples (uploads once if the host copy is newer)."""
        if self._dev is None:
            import torch
            self._dev = torch.as_tensor(np.ascontiguousarray(self._host, dtype=np.float64),
                                        device=f"cuda:{_device_index()}")
        return self._dev

    def _set_device(self, tensor):
        self._dev = tensor
        self._host = None

    def copy(self):
        return deepcopy(self)

    def __deepcopy__(self, memo):
        out = type(self).__new__(type(self))  # keeps the merged class install() builds (plot mix-ins, overrides)
        out.fs = self.fs
        out._host = None if self._host is None else self._host.copy()
        out._dev = None if self._dev is None else self._dev.clone()
        out.recording = deepcopy(self.recording, memo)
        return out

    def __len__(self):
        if self._host is not None:
            return len(self._host)
        return int(self._dev.shape[0])

    def duration(self):
        return len(self) / self.fs

    # ---- K2 ------------------------------------------------------------------------------------------
    def peak_index(self, start=0, end=None, peak_height=0.12589):
        """Reference :32-70."""
        n = len(self)
        if n == 0:
            return 0
        if end is None:
            end = n
        s, e, _ = slice(start, end).indices(n)
        if e - s <= 0:
            return start
        x = self.device_data()
        return int(ops.peak_index(x, [s], [e - s], peak_height, max_len=e - s).item()) + s

    # ---- K4 ------------------------------------------------------------------------------------------
    def decay_params(self):
        """Reference :72-73 -> decay.decay_params (core/decay.py:44-260)."""
        n = len(self)
        if n < 10:
            return 0, n, -200.0, n if n > 0 else 1
        pk, knee, noise, win = ops.lundeby(self.device_data(), [0], [n], self.fs, max_len=n)
        return int(pk.item()), int(knee.item()), float(noise.item()), int(win.item())

    # ---- K17 -----------------------------------------------------------------------------------------
    def decay_times(self, peak_ind=None, knee_point_ind=None, noise_floor=None, window_size=None):
        """Reference :75-80 -> decay.decay_times (core/decay.py:263-352): ``(EDT, RT20, RT30, RT60)`` in seconds,
        ``None`` where undefined."""
        if peak_ind is None or knee_point_ind is None or noise_floor is None:
            peak_ind, knee_point_ind, noise_floor, window_size = self.decay_params()
        return decay_times_batch([self], [(peak_ind, knee_point_ind, noise_floor, window_size)])[0]

    def decay_adjustment_params(self, target):
        """``decay.decay_adjustment_params(data, fs, target)`` (core/decay.py:355-380): the
        ``(window_start, half_window, knee_index, level_dB)`` tuple ``apply_decay_window`` takes, or None when the
        response already decays faster than ``target`` seconds of RT60."""
        peak, knee, noise, win = self.decay_params()
        slope = None
        for rt, level in zip(decay_times_batch([self], [(peak, knee, noise, win)])[0], (-10, -20, -30, -60)):
            if not rt:
                break
            slope = level / rt
        target_slope = -60 / target
        if target_slope > slope:  # same TypeError as the reference when no decay time could be measured
            return None
        knee_time = knee / self.fs
        start = peak + 2 * (self.fs // 1000)
        return start, knee - start, knee, target_slope * knee_time - slope * knee_time

    def crop_head(self, head_ms=1):
        """Reference :82-90."""
        if len(self) == 0:
            return
        crop = max(self.peak_index() - int(self.fs * head_ms / 1000), 0)
        if self._dev is not None and self._host is None:
            self._set_device(self._dev[crop:])
        else:
            self.data = self._host[crop:]

    def shift(self, samples):
        """Reference :92-108 (length-preserving integer shift)."""
        n = len(self)
        if samples == 0 or n == 0:
            return
        x = self.device_data()
        lo, hi = max(samples, 0), min(n + samples, n)
        self._set_device(ops.gather_window(x, [0], n, shift=[samples], valid_lo=[lo], valid_hi=[max(hi, lo)])[0])

    def equalize(self, fir):
        """Reference :110-119: ``convolve(data, fir, 'full')`` (next-row component)."""
        from .dispatcher import fft_convolve_full
        fir = np.ascontiguousarray(fir, dtype=np.float64)
        self._set_device(fft_convolve_full(self.device_data()[None, :], fir)[0])

    def convolve(self, x):
        from .dispatcher import fft_convolve_full
        import torch
        xd = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64), device=f"cuda:{_device_index()}")
        return fft_convolve_full(xd[None, :], np.ascontiguousarray(self.data, dtype=np.float64))[0].cpu().numpy()

    # ---- K9 ------------------------------------------------------------------------------------------
    def resample(self, fs):
        """Reference :121-124."""
        up, down, taps = resample_taps(int(fs), int(self.fs))
        if up == down:
            self.fs = fs
            return
        self._set_device(ops.window_resample(self.device_data(), taps, up, down)[0])
        self.fs = fs

    # ---- K15 -----------------------------------------------------------------------------------------
    def apply_decay_window(self, params):
        """``core/decay.py:383-403`` with precomputed ``(window_start, half_window, knee_index, level_dB)``."""
        if params is None:
            return
        self._set_device(ops.window_resample(self.device_data(), None, 1, 1, decay=np.array(params, dtype=np.float64))[0])

    # ---- K8 + K10 --------------------------------------------------------------------------------------
    def magnitude_response(self):
        """Reference :153-155 -> audio_io.magnitude_response (core/audio_io.py:100-113)."""
        n = len(self)
        half = int(np.ceil(n / 2))
        f = np.arange(half) * (self.fs / n)
        return f, ops.magnitude_db(self.device_data()).cpu().numpy()[0]

    def frequency_response(self):
        """Reference :157-188."""
        grid = FrequencyResponse.generate_frequencies(f_step=1.01, f_min=10, f_max=self.fs / 2)
        if len(self) < 2:
            return FrequencyResponse(name="Frequency response (short IR)", frequency=grid, raw=np.zeros_like(grid))
        raw = frequency_response_batch(self.device_data()[None, :], self.fs).cpu().numpy()[0]
        return FrequencyResponse(name="Frequency response", frequency=grid, raw=raw)


def decay_times_batch(irs, params=None):
    """``[ir.decay_times(*p) for ir, p in zip(irs, params)]`` as one K4 (when ``params`` is None) and one K17
    launch.  ``params``: per response ``(peak, knee, noise_floor_dB, window)`` as ``decay_params`` returns them.
    -> list of ``(EDT, RT20, RT30, RT60)`` tuples with ``None`` for undefined entries."""
    import torch
    irs = list(irs)
    if not irs:
        return []
    fs = irs[0].fs
    if any(ir.fs != fs for ir in irs):
        raise ValueError("decay_times_batch: all responses must share one sampling rate")
    tens = [ir.device_data() for ir in irs]
    lens = [int(t.shape[0]) for t in tens]
    flat = torch.cat(tens) if len(tens) > 1 else tens[0]
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    if params is None:
        pk, kn, nf, wn = ops.lundeby(flat, offs, lens, fs, max_len=max(lens))
        pk, kn, nf, wn = (t.cpu().numpy() for t in (pk, kn, nf, wn))
    else:
        pk = np.array([p[0] for p in params], dtype=np.int64)
        kn = np.array([p[1] for p in params], dtype=np.int64)
        nf = np.array([p[2] for p in params], dtype=np.float64)
        wn = np.array([p[3] if p[3] is not None else 0 for p in params], dtype=np.int64)
    out = ops.decay_times(flat, offs, lens, pk, kn, nf, wn, fs, max_len=max(lens))
    return [tuple(None if np.isnan(v) else float(v) for v in row) for row in out]


def frequency_response_batch(irs, fs, row_offsets=None, n=None):
    """rows of equal-length IRs (CUDA [B, n]) -> raw dB on the 1.01-step grid 10 Hz..fs/2 (CUDA [B, F]).
    With ``row_offsets`` (int64 CUDA tensor) and ``n`` the rows are slices ``irs.reshape(-1)[o : o + n]`` of one flat
    buffer (the dispatcher's cropped responses): no gather copy in front of the transform."""
    n = irs.shape[1] if row_offsets is None else int(n)
    half = int(np.ceil(n / 2))
    f = np.arange(half) * (fs / n)
    pts = (fs / 2) / 4.0
    step = 1 if (pts < 2 or len(f) < 2) else (int(round(len(f) / pts)) or 1)
    fd = f[1::step]
    if len(fd) == 0:
        fd, step = f[1:], 1
    if step > 1:  # any response longer than fs/4 samples: only the kept bins are produced, two rows per transform
        md = ops.magnitude_db_decimated(irs, 1, step, row_offsets, n)
    else:
        if row_offsets is not None:
            import torch
            irs = torch.stack([irs.reshape(-1).narrow(0, int(o), n) for o in row_offsets.tolist()])
        md = ops.magnitude_db(irs)[:, 1:].contiguous()
    grid = FrequencyResponse.generate_frequencies(f_step=1.01, f_min=10, f_max=fs / 2)
    return ops.loglerp(np.log10(fd), md, np.log10(grid))

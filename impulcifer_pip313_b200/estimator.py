"""``ImpulseResponseEstimator`` — drop-in for the reference class of the same name
(``core/impulse_response_estimator.py:26-262``), with ``estimate()`` running on the B200.

Host side (unchanged contract): ``fs, high, low, n_octaves, w1, w2, test_signal, duration, inverse_filter``
and ``__len__`` are computed once with numpy exactly as the reference does (they are the constants of
the measurement).  Device side: ``estimate(recording)`` and the batched
``estimate_batch`` call ``impb_deconv_execute`` (hand-written sm_100a FFT convolution) through the C ABI.
No CPU path exists for the deconvolution: without the CUDA library or a device the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

from . import _lib


def _hann_rising(n: int) -> np.ndarray:
    """First half of scipy.signal.windows.hann(n) for even n."""
    from scipy.signal.windows import hann  # the reference's own window (ulp-identical envelope, core/..estimator.py:109-131)
    return hann(n)[:n // 2]


def _next_pow2(n: int) -> int:
    return 1 << (int(n) - 1).bit_length()


class DeconvPlan:
    """Device plan for recordings of up to ``max_len`` samples against one inverse filter."""

    def __init__(self, inverse_filter: np.ndarray, max_len: int, precision: int = 64, device: int = 0):
        import torch
        self.device = int(device)
        self.precision = int(precision)
        self.max_len = int(max_len)
        self.ctx = _lib.context(self.device)
        inv = np.ascontiguousarray(inverse_filter, dtype=np.float64)
        self.M = len(inv)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().impb_deconv_plan_create(self.ctx.handle, inv.ctypes.data, self.M, self.max_len,
                                                          self.precision, C.byref(h)), "impb_deconv_plan_create")
        self.handle = h
        n, n1, n2 = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.lib().impb_deconv_plan_info(h, C.byref(n), C.byref(n1), C.byref(n2))
        self.nfft, self.n1, self.n2 = n.value, n1.value, n2.value
        self._ws = _lib.StreamWorkspace(self.device)
        self._lock = threading.Lock()

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.lib().impb_deconv_plan_destroy(self.handle)
        except Exception:
            pass

    def workspace(self, batch: int):
        return self._ws.get(_lib.lib().impb_deconv_workspace_bytes(self.handle, int(batch)))

    def launches(self, batch: int) -> int:
        return int(_lib.lib().impb_deconv_launches_per_execute(self.handle, int(batch)))

    def blockmax_info(self, out_offset=0):
        """(nblk, blk_len, blk_first) of the block maxima ``execute`` can emit (nblk = 0: not available)."""
        import ctypes as C
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(_lib.lib().impb_deconv_blockmax_info(self.handle, int(out_offset), C.byref(a), C.byref(b), C.byref(c)),
                   "impb_deconv_blockmax_info")
        return int(a.value), int(b.value), int(c.value)

    def execute(self, rec, out=None, length=None, out_offset=0, out_len=None, row_offsets=None, absmax=None,
                argmax=None, blockmax=None):
        """rec: CUDA tensor [batch, ld] (float64 / float32 / int32 PCM), or any CUDA tensor plus
        ``row_offsets`` (int64 CUDA tensor: element offset of every row inside ``rec``) and ``length``.
        ``absmax`` (optional float64 CUDA tensor [batch]) receives max |out[r, :]| at no extra cost; ``argmax``
        (optional int64 CUDA tensor [batch], needs ``absmax``) an index where it is attained; ``blockmax`` (optional
        float64 CUDA tensor [batch, nblk], see ``blockmax_info``) the maxima per time block.
        Returns CUDA tensor [batch, out_len]."""
        import torch
        if row_offsets is None and rec.dim() == 1:
            rec = rec[None, :]
        if not rec.is_cuda or rec.stride(-1) != 1:
            raise ValueError("rec must be a CUDA tensor with contiguous rows")
        fmt = {torch.float64: _lib.FMT_F64, torch.float32: _lib.FMT_F32, torch.int32: _lib.FMT_I32PCM}[rec.dtype]
        if row_offsets is not None:
            if length is None:
                raise ValueError("length is required with row_offsets")
            row_offsets = row_offsets.to(device=rec.device, dtype=torch.int64).contiguous()
            batch = row_offsets.numel()
        else:
            batch = rec.shape[0]
        length = rec.shape[1] if length is None else int(length)
        out_len = length - out_offset if out_len is None else int(out_len)
        out_dtype = torch.float64 if self.precision == 64 else torch.float32
        if out is None:
            out = torch.empty((batch, out_len), dtype=out_dtype, device=rec.device)
        out_f64 = 1 if out.dtype == torch.float64 else 0
        with self._lock:
            ws = self.workspace(batch)
            _lib.check(_lib.lib().impb_deconv_execute(
                self.handle, rec.data_ptr(), None if row_offsets is None else row_offsets.data_ptr(),
                (rec.stride(0) if batch > 1 else rec.shape[-1]) if row_offsets is None else 0, fmt, length,
                out.data_ptr(), out.stride(0) if batch > 1 else out.shape[1], out_f64, int(out_offset), out_len,
                None if absmax is None else absmax.data_ptr(), None if argmax is None else argmax.data_ptr(),
                None if blockmax is None else blockmax.data_ptr(), batch, ws.data_ptr(),
                _lib.current_stream_ptr(rec.device)), "impb_deconv_execute")
        return out


class ImpulseResponseEstimator(object):
    """Farina exponential-sweep measurement: test signal, inverse filter, deconvolution.

    Same constructor, attributes and ``estimate`` contract as the reference class
    (``core/impulse_response_estimator.py:33-50, 149-151``).
    """

    def __init__(self, min_duration=5.0, fs=44100, device=None, precision=None):
        if fs != int(fs):
            raise ValueError('Sampling rate "fs" must be an integer.')
        self.fs = int(fs)
        self.high = self.fs / 2
        self.low = 5
        self.n_octaves = np.ceil(np.log2(self.high / self.low))
        self.low = self.high / 2 ** self.n_octaves
        self.w1 = self.low / self.fs * 2 * np.pi
        self.w2 = self.high / self.fs * 2 * np.pi
        self.test_signal = self.generate_test_signal(min_duration)
        self.duration = len(self.test_signal) / self.fs
        self.inverse_filter = self.generate_inverse_filter()
        self.device = _lib.default_device() if device is None else int(device)
        if precision is None:
            precision = 32 if os.environ.get("IMPULCIFER_B200_PRECISION", "fp64").lower() == "fp32" else 64
        self.precision = int(precision)
        self._plans = {}
        self._plan_key = None

    def __len__(self):
        return len(self.test_signal)

    # ---- host-side constants (reference :73-147) ---------------------------------------------------
    def generate_test_signal(self, min_duration, fade_in=1 / 2, fade_out=None):
        P = self.n_octaves
        M = np.ceil(min_duration * self.fs * (np.pi / 2 ** P) / (np.pi * 2 * np.log(2 ** P)))
        L = M * np.pi * 2 * np.log(2 ** P) / (np.pi / 2 ** P)
        N = np.round(L)
        sweep = np.sin(np.pi / 2 ** P * L / np.log(2 ** P) * np.exp(np.arange(N) / N * np.log(2 ** P)))
        sec_per_octave = N / self.fs / P
        envelope = np.ones(len(sweep))
        if fade_in is not None:
            n = 2 * int(self.fs * sec_per_octave * fade_in)
            n += n % 2
            envelope[:n // 2] = _hann_rising(n)
        if fade_out is not None:
            n = 2 * int(self.fs * sec_per_octave * fade_out)
            n += n % 2
            envelope[len(sweep) - n // 2:] = _hann_rising(n)[::-1]
        return sweep * envelope

    def generate_inverse_filter(self):
        P = self.n_octaves
        N = len(self.test_signal)
        k = np.arange(N)
        inv = self.test_signal[::-1] * (2 ** (P / N)) ** (-k) * P * np.log(2) / (1 - 2 ** -P)
        # level normalisation: |FFT(inv (*) sweep)| at a quarter of the band (reference :80-82)
        n_full = 2 * N - 1
        nfft = _next_pow2(n_full)
        prod = np.fft.irfft(np.fft.rfft(inv, nfft) * np.fft.rfft(self.test_signal, nfft), nfft)[:n_full]
        q = round(n_full / 4)
        ang = -2.0 * np.pi * ((np.arange(n_full) * q) % n_full) / n_full
        level = abs(np.dot(prod, np.cos(ang)) + 1j * np.dot(prod, np.sin(ang)))
        self._plans = {}
        return inv / level

    # ---- device side ---------------------------------------------------------------------------------
    def plan(self, max_len: int, precision=None, device=None) -> DeconvPlan:
        precision = self.precision if precision is None else int(precision)
        device = self.device if device is None else int(device)
        # keyed on a fingerprint of the inverse filter's CONTENT: an in-place edit (est.inverse_filter *= g) or a
        # reassignment that reuses the id must not keep serving the stale device spectrum
        inv = self.inverse_filter
        n_inv = len(inv)
        # (O(1) fingerprint: buffer address, length and four samples; hashing 2.4 MB on every call would cost ~1 ms)
        key = ((inv.ctypes.data if isinstance(inv, np.ndarray) else id(inv), float(inv[0]), float(inv[n_inv // 3]),
                float(inv[n_inv // 2]), float(inv[-1])), n_inv, precision, device)
        if self._plan_key != key[:2]:
            self._plans = {}
            self._plan_key = key[:2]
        best = None
        for (p, d, cap), pl in self._plans.items():
            if p == precision and d == device and cap >= max_len and (best is None or cap < best.max_len):
                best = pl
        if best is None:
            cap = max(int(max_len), 2 * self.fs + len(self.inverse_filter))
            best = DeconvPlan(self.inverse_filter, cap, precision, device)
            self._plans[(precision, device, cap)] = best
        return best

    def estimate(self, recording):
        """Impulse response of one recording: 1-D float64 array of ``len(recording)`` samples."""
        import torch
        rec = np.ascontiguousarray(recording, dtype=np.float64)
        if rec.ndim != 1:
            raise ValueError("recording must be one-dimensional")
        if len(rec) == 0:
            return np.zeros(0)
        dev = torch.device("cuda", self.device)
        with torch.cuda.device(dev):
            x = torch.from_numpy(rec).to(dev, non_blocking=False)
            y = self.plan(len(rec)).execute(x)
            return y[0].to(torch.float64).cpu().numpy()

    def estimate_batch(self, recordings, out_offset=0, out_len=None, precision=None):
        """Batched ``estimate``.  ``recordings``: [batch, len] numpy array or CUDA tensor (f64/f32/int32 PCM).

        Returns the same kind of container ([batch, out_len]); with a CUDA tensor nothing leaves the device.
        """
        import torch
        if isinstance(recordings, np.ndarray):
            dev = torch.device("cuda", self.device)
            with torch.cuda.device(dev):
                x = torch.from_numpy(np.ascontiguousarray(recordings)).to(dev)
                y = self.plan(x.shape[-1], precision).execute(x, out_offset=out_offset, out_len=out_len)
                return y.cpu().numpy()
        with torch.cuda.device(recordings.device):
            return self.plan(recordings.shape[-1], precision, recordings.device.index).execute(
                recordings, out_offset=out_offset, out_len=out_len)

    # ---- construction from a sweep file (reference :234-262) ---------------------------------------------
    @classmethod
    def from_wav(cls, file_path, **kw):
        from .audio_io import read_wav
        fs, data = read_wav(file_path)
        ref = data[0, :] if data.ndim > 1 else data
        ire = cls(min_duration=(len(ref) - 1) / fs, fs=fs, **kw)
        if len(ire.test_signal) != len(ref):
            ire.test_signal = ref
            ire.duration = len(ref) / fs
            ire.inverse_filter = ire.generate_inverse_filter()
        elif np.max(np.abs(ire.test_signal - ref)) > 1e-4:
            print("Warning: Loaded WAV differs slightly from generated signal. "
                  "Re-calculating inverse filter based on WAV.")
            ire.test_signal = ref
            ire.inverse_filter = ire.generate_inverse_filter()
        return ire

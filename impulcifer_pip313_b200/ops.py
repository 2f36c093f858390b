"""Thin torch-tensor wrappers over the C ABI (``include/impb.h``): one function per device entry point.

Everything here takes and returns CUDA tensors and launches on the current torch stream of the tensor's
device; nothing synchronises and nothing runs on the CPU.  Host-side numpy is only used to build the small
constant tables the reference also builds on the host (window shapes, Savitzky-Golay coefficients).
"""
from __future__ import annotations

import ctypes as C
import functools
import math
import threading

import numpy as np

from . import _lib


def _torch():
    import torch
    return torch


def _i64(t):
    torch = _torch()
    assert t.dtype == torch.int64 and t.is_cuda and t.is_contiguous()
    return t.data_ptr()


def _dev(t):
    return t.device.index if t.device.index is not None else 0


def _stream(t):
    return _lib.current_stream_ptr(t.device)


def as_index_tensor(values, device):
    torch = _torch()
    if isinstance(values, torch.Tensor):
        return values.to(device=device, dtype=torch.int64).contiguous()
    return torch.as_tensor(np.asarray(values, dtype=np.int64), device=device)


# ---- K2 -------------------------------------------------------------------------------------------
_win_cache = {}
_win_lock = threading.Lock()


def _dev_window(win, device):
    """Small host-built window -> cached device tensor (keyed by content)."""
    torch = _torch()
    if win is None or isinstance(win, torch.Tensor):
        return win
    arr = np.ascontiguousarray(win, dtype=np.float64)
    if arr.size == 0:
        return None
    key = (arr.tobytes(), str(device))
    with _win_lock:  # dispatcher threads of different GPUs share this cache
        t = _win_cache.get(key)
        if t is None:
            if len(_win_cache) > 256:
                _win_cache.clear()
            t = _win_cache[key] = torch.as_tensor(arr, device=device)
    return t


def _dev_const(values, device):
    """Host-built grid / table (frequency grids, sigmoid weights ...) -> cached device tensor.  A fresh
    ``torch.as_tensor(..., device=...)`` per call is a pageable copy that synchronises the stream: the launches of the
    frequency-response chain would stop queueing behind each other."""
    torch = _torch()
    if isinstance(values, torch.Tensor):
        return values
    arr = np.asarray(values, dtype=np.float64)
    if arr.size == 0:
        return torch.empty(0, dtype=torch.float64, device=device)
    return _dev_window(arr, device)


def _dev_index_const(values, device):
    """int64 twin of ``_dev_const`` (row offsets of a batch geometry that repeats from step to step)."""
    torch = _torch()
    arr = np.ascontiguousarray(values, dtype=np.int64)
    key = (arr.tobytes(), 'i8', str(device))
    with _win_lock:
        t = _win_cache.get(key)
        if t is None:
            if len(_win_cache) > 256:
                _win_cache.clear()
            t = _win_cache[key] = torch.as_tensor(arr, device=device)
    return t


def make_view(device, shift=None, valid_lo=None, valid_hi=None, fade_in=None, hint_max=None, hint_argmax=None,
              blocks=None):
    """(ctypes impb_view or None, keep-alive tensors).  See include/impb.h.  ``hint_max`` / ``hint_argmax``: float64 /
    int64 CUDA tensors (max |x| of the untouched samples and where, relative to the row offset)."""
    if shift is None and valid_lo is None and valid_hi is None and fade_in is None and hint_max is None and blocks is None:
        return None, ()
    sh = None if shift is None else as_index_tensor(shift, device)
    lo = None if valid_lo is None else as_index_tensor(valid_lo, device)
    hi = None if valid_hi is None else as_index_tensor(valid_hi, device)
    fi = _dev_window(fade_in, device)
    hm = ha = None
    if hint_max is not None and hint_argmax is not None:
        hm = hint_max.contiguous()
        ha = as_index_tensor(hint_argmax, device)
    v = _lib.View(None if sh is None else sh.data_ptr(), None if lo is None else lo.data_ptr(),
                  None if hi is None else hi.data_ptr(), None if fi is None else fi.data_ptr(),
                  0 if fi is None else fi.numel(), None if hm is None else hm.data_ptr(),
                  None if ha is None else ha.data_ptr())
    bm = None
    if blocks is not None:  # (block maxima [rows, nblk] CUDA float64, blk_len, blk_first) of untouched rows
        bm, blk_len, blk_first = blocks
        bm = bm.contiguous()
        v.blk_max, v.nblk, v.blk_len, v.blk_first = bm.data_ptr(), bm.shape[1], int(blk_len), int(blk_first)
    return v, (sh, lo, hi, fi, hm, ha, bm)


def peak_index(x, offsets, lengths, height=0.12589, max_len=None, view=None, known_max=None):
    """First-peak index of every segment ``x[offsets[r] : offsets[r]+lengths[r]]`` (relative to the segment).
    ``view`` = dict(shift=, valid_lo=, valid_hi=, fade_in=) evaluates it on the shifted / faded view instead."""
    torch = _torch()
    x = x.contiguous()
    assert x.dtype == torch.float64 and x.is_cuda
    offsets = as_index_tensor(offsets, x.device)
    lengths = as_index_tensor(lengths, x.device)
    batch = offsets.numel()
    out = torch.empty(batch, dtype=torch.int64, device=x.device)
    if max_len is None:
        max_len = int(lengths.max().item()) if batch else 0
    v, keep = make_view(x.device, **(view or {}))
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().impb_peak_index(_lib.context(_dev(x)).handle, x.data_ptr(), _i64(offsets), _i64(lengths),
                                              None if v is None else C.addressof(v),
                                              None if known_max is None else known_max.data_ptr(), float(height),
                                              out.data_ptr(), batch, int(max_len), _stream(x)), "impb_peak_index")
    del keep
    return out


# ---- K4 -------------------------------------------------------------------------------------------
def lundeby(x, offsets, lengths, fs, max_len=None, view=None):
    """decay_params of every segment -> (peak, knee, noise_db, window) tensors (indices relative to the segment)."""
    torch = _torch()
    x = x.contiguous()
    assert x.dtype == torch.float64 and x.is_cuda
    offsets = as_index_tensor(offsets, x.device)
    lengths = as_index_tensor(lengths, x.device)
    batch = offsets.numel()
    peak = torch.empty(batch, dtype=torch.int64, device=x.device)
    knee = torch.empty_like(peak)
    win = torch.empty_like(peak)
    noise = torch.empty(batch, dtype=torch.float64, device=x.device)
    if max_len is None:
        max_len = int(lengths.max().item()) if batch else 0
    v, keep = make_view(x.device, **(view or {}))
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().impb_lundeby(_lib.context(_dev(x)).handle, x.data_ptr(), _i64(offsets), _i64(lengths),
                                           None if v is None else C.addressof(v), int(fs), peak.data_ptr(),
                                           knee.data_ptr(), noise.data_ptr(), win.data_ptr(), batch, int(max_len),
                                           _stream(x)), "impb_lundeby")
    del keep
    return peak, knee, noise, win


# ---- K17 / K18 -------------------------------------------------------------------------------------
def decay_times(x, offsets, lengths, peaks, knees, noise_db, wins, fs, max_len=None):
    """EDT / RT20 / RT30 / RT60 (seconds, NaN = undefined) for segments of the flat CUDA tensor ``x``; the four
    parameter arrays are what ``lundeby`` returned.  -> numpy [batch, 4]."""
    torch = _torch()
    dev = x.device
    offsets, lengths = as_index_tensor(offsets, dev), as_index_tensor(lengths, dev)
    peaks, knees, wins = as_index_tensor(peaks, dev), as_index_tensor(knees, dev), as_index_tensor(wins, dev)
    noise = torch.as_tensor(np.asarray(noise_db, dtype=np.float64), device=dev)
    batch = offsets.numel()
    if max_len is None:
        max_len = int(lengths.max().item()) if batch else 0
    out = torch.empty((batch, 4), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        nbytes = _lib.lib().impb_decay_times_workspace_bytes(batch, int(max_len))
        ws = torch.empty(max(int(nbytes), 8), dtype=torch.uint8, device=dev)
        _lib.check(_lib.lib().impb_decay_times(_lib.context(_dev(x)).handle, x.data_ptr(), _i64(offsets), _i64(lengths),
                                               _i64(peaks), _i64(knees), noise.data_ptr(), _i64(wins), int(fs),
                                               out.data_ptr(), ws.data_ptr(), batch, int(max_len), _stream(x)),
                   "impb_decay_times")
    return out.cpu().numpy()


def segment_energy(x, offsets, seg_lo, seg_hi):
    """mean(x[offsets[r] + lo : offsets[r] + hi] ** 2) for every (row, segment); seg_lo / seg_hi: [batch, nseg]
    integer arrays (row relative).  -> numpy [batch, nseg], NaN for empty ranges."""
    torch = _torch()
    dev = x.device
    seg_lo = np.ascontiguousarray(seg_lo, dtype=np.int64)
    seg_hi = np.ascontiguousarray(seg_hi, dtype=np.int64)
    batch, nseg = seg_lo.shape
    offsets = as_index_tensor(offsets, dev)
    lo, hi = as_index_tensor(seg_lo.reshape(-1), dev), as_index_tensor(seg_hi.reshape(-1), dev)
    out = torch.empty((batch, nseg), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().impb_segment_energy(_lib.context(_dev(x)).handle, x.data_ptr(), _i64(offsets), _i64(lo),
                                                  _i64(hi), nseg, out.data_ptr(), batch, _stream(x)),
                   "impb_segment_energy")
    return out.cpu().numpy()


# ---- K19 / K20 -------------------------------------------------------------------------------------
def dft_bin_abs(x, k):
    """|rfft(row)[k]| for every row of the CUDA matrix ``x`` -> numpy [batch]."""
    torch = _torch()
    x = x.contiguous()
    batch, n = x.shape
    out = torch.empty(batch, dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().impb_dft_bin_abs(_lib.context(_dev(x)).handle, x.data_ptr(), x.stride(0), n, int(k),
                                               out.data_ptr(), batch, _stream(x)), "impb_dft_bin_abs")
    return out.cpu().numpy()


def add_delayed(dst, src, which, delay):
    """dst[r, j] += src[which[r], j - delay[r]] in place (rows of ``src`` are zero outside their length)."""
    torch = _torch()
    assert dst.is_contiguous() and src.is_contiguous() and dst.dtype == torch.float64 and src.dtype == torch.float64
    batch, n = dst.shape
    wh = torch.as_tensor(np.asarray(which, dtype=np.int32), device=dst.device)
    dl = as_index_tensor(delay, dst.device)
    with torch.cuda.device(dst.device):
        _lib.check(_lib.lib().impb_add_delayed(_lib.context(_dev(dst)).handle, dst.data_ptr(), dst.stride(0), n,
                                               src.data_ptr(), src.stride(0), src.shape[1], src.shape[0],
                                               wh.data_ptr(), _i64(dl), batch, _stream(dst)), "impb_add_delayed")
    return dst


# ---- K6 -------------------------------------------------------------------------------------------
def xcorr_lag(x, a_offsets, a_lengths, b_offsets, b_lengths, max_seg_len=None, fade_in=None):
    torch = _torch()
    x = x.contiguous()
    a_offsets, a_lengths = as_index_tensor(a_offsets, x.device), as_index_tensor(a_lengths, x.device)
    b_offsets, b_lengths = as_index_tensor(b_offsets, x.device), as_index_tensor(b_lengths, x.device)
    batch = a_offsets.numel()
    out = torch.empty(batch, dtype=torch.int64, device=x.device)
    if max_seg_len is None:
        max_seg_len = int(max(a_lengths.max().item(), b_lengths.max().item())) if batch else 0
    fi = _dev_window(fade_in, x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().impb_xcorr_lag(_lib.context(_dev(x)).handle, x.data_ptr(), _i64(a_offsets),
                                             _i64(a_lengths), _i64(b_offsets), _i64(b_lengths),
                                             None if fi is None else fi.data_ptr(), 0 if fi is None else fi.numel(),
                                             out.data_ptr(), batch, int(max_seg_len), _stream(x)), "impb_xcorr_lag")
    return out


# ---- K3 + K5 + K7 + K16 ---------------------------------------------------------------------------
def gather_window(src, src_start, out_len, shift=None, valid_lo=None, valid_hi=None, fade_in=None, fade_out=None,
                  gain=1.0, out=None):
    """dst[r, j] = gain * fade_out(j) * fade_in(j - shift) * src[src_start[r] + j - shift[r]] (0 outside valid)."""
    torch = _torch()
    src = src.contiguous()
    assert src.dtype == torch.float64 and src.is_cuda
    dev = src.device
    src_start = as_index_tensor(src_start, dev)
    batch = src_start.numel()
    shift_t = None if shift is None else as_index_tensor(shift, dev)
    lo_t = None if valid_lo is None else as_index_tensor(valid_lo, dev)
    hi_t = None if valid_hi is None else as_index_tensor(valid_hi, dev)
    fi = _dev_window(fade_in, dev)
    fo = _dev_window(fade_out, dev)
    if out is None:
        out = torch.empty((batch, int(out_len)), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().impb_gather_window(
            _lib.context(_dev(src)).handle, src.data_ptr(), _i64(src_start),
            None if shift_t is None else _i64(shift_t), None if lo_t is None else _i64(lo_t),
            None if hi_t is None else _i64(hi_t), int(out_len),
            None if fi is None else fi.data_ptr(), 0 if fi is None else fi.numel(),
            None if fo is None else fo.data_ptr(), 0 if fo is None else fo.numel(),
            float(gain), out.data_ptr(), out.stride(0) if batch > 1 else max(int(out_len), 1), batch, _stream(src)),
            "impb_gather_window")
    return out


def gather_rows(src, src_start, out_lens, dst_offsets, total, max_out_len, shift=None, valid_lo=None, valid_hi=None,
                fade_in=None, fade_out=None, gain=1.0):
    """Ragged ``gather_window``: row r -> ``dst[dst_offsets[r] : dst_offsets[r] + out_lens[r]]`` of one flat CUDA
    tensor of ``total`` samples (returned).  All index arguments are CUDA int64 tensors (or array-likes)."""
    torch = _torch()
    src = src.contiguous()
    dev = src.device
    src_start = as_index_tensor(src_start, dev)
    out_lens, dst_offsets = as_index_tensor(out_lens, dev), as_index_tensor(dst_offsets, dev)
    batch = src_start.numel()
    shift_t = None if shift is None else as_index_tensor(shift, dev)
    lo_t = None if valid_lo is None else as_index_tensor(valid_lo, dev)
    hi_t = None if valid_hi is None else as_index_tensor(valid_hi, dev)
    fi, fo = _dev_window(fade_in, dev), _dev_window(fade_out, dev)
    out = torch.empty(int(total), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().impb_gather_rows(
            _lib.context(_dev(src)).handle, src.data_ptr(), _i64(src_start),
            None if shift_t is None else _i64(shift_t), None if lo_t is None else _i64(lo_t),
            None if hi_t is None else _i64(hi_t), _i64(out_lens), _i64(dst_offsets), int(max_out_len),
            None if fi is None else fi.data_ptr(), 0 if fi is None else fi.numel(),
            None if fo is None else fo.data_ptr(), 0 if fo is None else fo.numel(), float(gain), out.data_ptr(), batch,
            _stream(src)), "impb_gather_rows")
    return out


# ---- K8 -------------------------------------------------------------------------------------------
class FftPlan:
    """Exact-length transform plan used by ``magnitude_db``."""

    def __init__(self, n: int, device: int = 0):
        torch = _torch()
        self.n, self.device = int(n), int(device)
        self.ctx = _lib.context(self.device)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().impb_fft_create(self.ctx.handle, self.n, 64, C.byref(h)), "impb_fft_create")
        self.handle = h
        self._ws = _lib.StreamWorkspace(self.device)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.lib().impb_fft_destroy(self.handle)
        except Exception:
            pass

    def workspace(self, batch):
        return self._ws.get(_lib.lib().impb_fft_workspace_bytes(self.handle, int(batch)))


_plan_lock = threading.Lock()
_fft_plans = {}
_minphase_plans = {}


def fft_plan(n, device=0) -> FftPlan:
    with _plan_lock:
        p = _fft_plans.get((n, device))
        if p is None:
            p = _fft_plans[(n, device)] = FftPlan(n, device)
        return p


def magnitude_db(x):
    """20*log10(|rfft(x)|)[: ceil(n/2)] for every row of ``x`` ([batch, n] float64 CUDA)."""
    torch = _torch()
    if x.dim() == 1:
        x = x[None, :]
    x = x.contiguous()
    batch, n = x.shape
    plan = fft_plan(n, _dev(x))
    half = (n + 1) // 2
    out = torch.empty((batch, half), dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        ws = plan.workspace(batch)
        _lib.check(_lib.lib().impb_magnitude_db(plan.handle, x.data_ptr(), n, out.data_ptr(), half, batch, ws.data_ptr(),
                                                _stream(x)), "impb_magnitude_db")
    return out


def magnitude_db_decimated(x, first, step, row_offsets=None, n=None):
    """``magnitude_db(x)[:, first::step]`` without the bins in between (the decimation of
    ``ImpulseResponse.frequency_response``, reference core/impulse_response.py:170-178).  When ``step`` divides the row
    length the kept bins are one (n / step)-point transform of the folded row; otherwise rows ride in pairs through one
    complex transform and only the kept bins are read back.  Lengths that need the chirp-z plan take the full transform.
    ``row_offsets`` (int64 CUDA tensor) + ``n``: the rows are ``x.reshape(-1)[o : o + n]`` for every offset ``o``
    (slices of one flat buffer, no gather copy)."""
    torch = _torch()
    if row_offsets is None:
        if x.dim() == 1:
            x = x[None, :]
        x = x.contiguous()
        batch, n = x.shape
    else:
        x = x.reshape(-1)
        batch, n = int(row_offsets.numel()), int(n)
    first, step = int(first), int(step)
    half = (n + 1) // 2
    n_out = max(0, -(-(half - first) // step))
    out = torch.empty((batch, n_out), dtype=torch.float64, device=x.device)
    if n_out == 0 or batch == 0:
        return out
    offs_ptr = None if row_offsets is None else row_offsets.data_ptr()
    if 2 <= step <= 64 and first < step and n % step == 0 and n // step >= 4:
        # step divides the row: the kept bins are the (n / step)-point DFT of the folded, modulated row
        small = fft_plan(n // step, _dev(x))
        with torch.cuda.device(x.device):
            ws = small.workspace(batch)
            rc = _lib.lib().impb_magnitude_db_folded(small.handle, x.data_ptr(), n, offs_ptr, first, step,
                                                     out.data_ptr(), n_out, batch, ws.data_ptr(), _stream(x))
        if rc != -4:  # IMPB_EUNSUPPORTED: no specialised plan for n / step -> the paired transform of the full row
            _lib.check(rc, "impb_magnitude_db_folded")
            return out
    plan = fft_plan(n, _dev(x))
    with torch.cuda.device(x.device):
        ws = plan.workspace(batch)
        rc = _lib.lib().impb_magnitude_db_decimated(plan.handle, x.data_ptr(), n, offs_ptr, first, step,
                                                    out.data_ptr(), n_out, batch, ws.data_ptr(), _stream(x))
    if rc == -4:  # IMPB_EUNSUPPORTED: chirp-z plan
        if row_offsets is not None:
            x = torch.stack([x.narrow(0, int(o), n) for o in row_offsets.tolist()])
        return magnitude_db(x)[:, first::step].contiguous()
    _lib.check(rc, "impb_magnitude_db_decimated")
    return out


def magnitude_lin(x, nfft, n_bins=None):
    """|rfft(x, n=nfft)|[:n_bins] for every row of ``x`` ([batch, n_valid <= nfft] float64 CUDA)."""
    torch = _torch()
    if x.dim() == 1:
        x = x[None, :]
    x = x.contiguous()
    batch, n_valid = x.shape
    n_bins = nfft // 2 + 1 if n_bins is None else int(n_bins)
    plan = fft_plan(int(nfft), _dev(x))
    out = torch.empty((batch, n_bins), dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        ws = plan.workspace(batch)
        _lib.check(_lib.lib().impb_magnitude_lin(plan.handle, x.data_ptr(), n_valid, n_valid, out.data_ptr(), n_bins,
                                                 n_bins, batch, ws.data_ptr(), _stream(x)), "impb_magnitude_lin")
    return out


# ---- K10 ------------------------------------------------------------------------------------------
def loglerp(logf_in, y, logf_out):
    """Piecewise-linear interpolation of rows of ``y`` from grid ``logf_in`` to ``logf_out`` (log10 Hz)."""
    torch = _torch()
    if y.dim() == 1:
        y = y[None, :]
    y = y.contiguous()
    dev = y.device
    lx, lo = _dev_const(logf_in, dev), _dev_const(logf_out, dev)
    batch, n_in = y.shape
    out = torch.empty((batch, lo.numel()), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().impb_loglerp(_lib.context(_dev(y)).handle, lx.data_ptr(), y.data_ptr(), n_in, n_in,
                                           lo.data_ptr(), lo.numel(), out.data_ptr(), lo.numel(), batch, _stream(y)),
                   "impb_loglerp")
    return out


def minphase_gain_grid(logf_in, eq_db, logf_out, n_below, logf_min, normalize):
    """Rows of equalization curves (dB, CUDA [B, F] on the log10 grid ``logf_in``) -> the linear gain grid that
    ``minphase_fir`` takes: interpolation, the f <= f_min prefix, [normalize], * 2, 10^(./20), last point 0
    (reference autoeq/frequency_response.py:653-676) in two launches."""
    torch = _torch()
    y = (eq_db if eq_db.dim() == 2 else eq_db[None, :]).contiguous()
    dev = y.device
    lx, lo = _dev_const(logf_in, dev), _dev_const(logf_out, dev)
    batch, n_in = y.shape
    n = lo.numel()
    out = torch.empty((batch, n), dtype=torch.float64, device=dev)
    scratch = torch.empty(max(batch, 1), dtype=torch.int64, device=dev) if normalize else None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().impb_minphase_gain_grid(
            _lib.context(_dev(y)).handle, lx.data_ptr(), y.data_ptr(), n_in, n_in, lo.data_ptr(), n, int(n_below),
            float(logf_min), 1 if normalize else 0, out.data_ptr(), n, batch,
            scratch.data_ptr() if normalize else None, _stream(y)), "impb_minphase_gain_grid")
    return out


# ---- K11 ------------------------------------------------------------------------------------------
@functools.lru_cache(maxsize=64)
def savgol_tables(window: int, order: int = 2):
    """(conv coefficients [w], edge-fit matrix [w//2, w]) exactly as scipy builds them
    (``savgol_coeffs`` least squares; ``np.polyfit`` on the first ``window`` samples)."""
    half = window // 2
    x = np.arange(-half, window - half, dtype=np.float64)[::-1]
    A = x ** np.arange(order + 1, dtype=np.float64).reshape(-1, 1)
    y = np.zeros(order + 1)
    y[0] = 1.0
    coef = np.linalg.lstsq(A, y, rcond=None)[0]
    t = np.arange(window, dtype=np.float64)
    V = np.vander(t, order + 1)
    edge = np.vander(np.arange(half, dtype=np.float64), order + 1) @ np.linalg.pinv(V)
    return np.ascontiguousarray(coef), np.ascontiguousarray(edge)


_dev_tables = {}


def _savgol_dev(window, device):
    torch = _torch()
    key = (window, str(device))
    with _plan_lock:
        t = _dev_tables.get(key)
        if t is None:
            c, e = savgol_tables(window)
            t = _dev_tables[key] = (torch.as_tensor(c, device=device), torch.as_tensor(e, device=device))
        return t


def savgol_blend(y, w_normal, it_normal, w_treble, it_treble, k_treble):
    """rows of y -> savgol^it_normal(y) * (1 - k) + savgol^it_treble(y) * k   (order 2, mode='interp')."""
    torch = _torch()
    if y.dim() == 1:
        y = y[None, :]
    y = y.contiguous()
    dev = y.device
    batch, n = y.shape
    if w_normal > n or w_treble > n:
        raise ValueError("If mode is 'interp', window_length must be less than or equal to the size of x.")
    cn, en = _savgol_dev(int(w_normal), dev)
    ct, et = _savgol_dev(int(w_treble), dev)
    kt = _dev_const(k_treble, dev)
    out = torch.empty_like(y)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().impb_savgol_blend(_lib.context(_dev(y)).handle, y.data_ptr(), n, n, int(w_normal),
                                                int(it_normal), cn.data_ptr(), en.data_ptr(), int(w_treble),
                                                int(it_treble), ct.data_ptr(), et.data_ptr(), kt.data_ptr(),
                                                out.data_ptr(), n, batch, _stream(y)), "impb_savgol_blend")
    return out


# ---- K22: per-set integer bookkeeping ---------------------------------------------------------------------
class SetLayoutDev:
    """Device copy of a recording-set layout (``impb_set_layout``): keeps the int32 / int64 index tensors alive."""

    def __init__(self, device, n, left, right, delay, pa, pb, pos_rows, pos_k, neg_rows, neg_k, gl, g_ref, g_rows,
                 g_of_row):
        torch = _torch()
        i32 = lambda v: torch.as_tensor(np.asarray(v, dtype=np.int32).reshape(-1), device=device)
        spk_of_row = np.zeros(n, dtype=np.int32)
        for k, (a, b) in enumerate(zip(left, right)):
            spk_of_row[a] = spk_of_row[b] = k
        self.t = dict(left=i32(left), right=i32(right),
                      delay=torch.as_tensor(np.asarray(delay, dtype=np.int64), device=device), spk_of_row=i32(spk_of_row),
                      pa=i32(pa), pb=i32(pb), pos_rows=i32(pos_rows), pos_k=i32(pos_k), neg_rows=i32(neg_rows),
                      neg_k=i32(neg_k), gl=i32(gl), g_rows=i32(g_rows), g_of_row=i32(g_of_row))
        t = self.t
        ptr = lambda k: t[k].data_ptr() if t[k].numel() else None
        self.n, self.n_pairs, self.n_groups = int(n), len(pa), len(gl)
        self.c = _lib.SetLayout(int(n), len(left), ptr('left'), ptr('right'), ptr('delay'), ptr('spk_of_row'),
                                len(pa), ptr('pa'), ptr('pb'), len(pos_rows), ptr('pos_rows'), ptr('pos_k'),
                                len(neg_rows), ptr('neg_rows'), ptr('neg_k'), len(gl), ptr('gl'), int(g_ref),
                                len(g_rows), ptr('g_rows'), ptr('g_of_row'))
        self.device = device


def _p(t):
    return None if t is None else t.data_ptr()


def set_heads(lay: SetLayoutDev, peaks, e_arg, L, seg, S):
    """-> crop, n1, off1, hint_arg [S*n] and (xa_off, xa_len, xb_off, xb_len) [S*n_pairs] (None without pairs)."""
    torch = _torch()
    dev = peaks.device
    B = S * lay.n
    out = torch.empty((4, B), dtype=torch.int64, device=dev)
    x = torch.empty((4, S * lay.n_pairs), dtype=torch.int64, device=dev) if lay.n_pairs else None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().impb_set_heads(
            _lib.context(_dev(peaks)).handle, C.byref(lay.c), peaks.data_ptr(), e_arg.data_ptr(), int(L), int(seg), int(S),
            out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), out[3].data_ptr(),
            *([x[i].data_ptr() for i in range(4)] if x is not None else [None] * 4), _stream(peaks)), "impb_set_heads")
    return out[0], out[1], out[2], out[3], x


def set_shift(lay: SetLayoutDev, lags, n1, off1, hint_arg, e_max, S, want_groups):
    """-> shift, lo, hi [S*n]; group-leader views (off, len, shift, lo, hi [int64], max [f64], arg) or None."""
    torch = _torch()
    dev = n1.device
    out = torch.empty((3, S * lay.n), dtype=torch.int64, device=dev)
    gi = gm = None
    if want_groups:
        gi = torch.empty((6, S * lay.n_groups), dtype=torch.int64, device=dev)
        gm = torch.empty(S * lay.n_groups, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().impb_set_shift(
            _lib.context(_dev(n1)).handle, C.byref(lay.c), _p(lags), n1.data_ptr(), off1.data_ptr(), hint_arg.data_ptr(),
            e_max.data_ptr(), int(S), out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
            *([gi[i].data_ptr() for i in range(5)] if gi is not None else [None] * 5), _p(gm),
            gi[5].data_ptr() if gi is not None else None, _stream(n1)), "impb_set_shift")
    return out[0], out[1], out[2], gi, gm


def set_shift2(lay: SetLayoutDev, group_peaks, n1, S, shift, lo, hi):
    """Composes the group-onset shifts into ``shift / lo / hi`` in place."""
    torch = _torch()
    with torch.cuda.device(n1.device):
        _lib.check(_lib.lib().impb_set_shift2(_lib.context(_dev(n1)).handle, C.byref(lay.c), _p(group_peaks), n1.data_ptr(),
                                              int(S), shift.data_ptr(), lo.data_ptr(), hi.data_ptr(), _stream(n1)),
                   "impb_set_shift2")


@functools.lru_cache(maxsize=8)
def _smooth5_host(limit: int):
    vals = []
    p2 = 1
    while p2 <= 2 * limit:
        p3 = p2
        while p3 <= 2 * limit:
            p5 = p3
            while p5 <= 2 * limit:
                vals.append(p5)
                p5 *= 5
            p3 *= 3
        p2 *= 2
    return np.array(sorted(vals), dtype=np.int64)


_smooth_cache = {}


def set_tails(lay: SetLayoutDev, knee, n1, S, head, fade_out_len, tail_bound, table_limit):
    """-> tails [S], set_off [S+1], out_len [S*n], dst_off [S*n], flags [1] (all int64, device)."""
    torch = _torch()
    dev = n1.device
    key = (int(table_limit), str(dev))
    sm = _smooth_cache.get(key)
    if sm is None:
        sm = _smooth_cache[key] = torch.as_tensor(_smooth5_host(int(table_limit)), device=dev)
    B = S * lay.n
    buf = torch.empty(S + (S + 1) + 2 * B + 1, dtype=torch.int64, device=dev)
    tails, set_off = buf[:S], buf[S:2 * S + 1]
    out_len, dst_off, flags = buf[2 * S + 1:2 * S + 1 + B], buf[2 * S + 1 + B:2 * S + 1 + 2 * B], buf[-1:]
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().impb_set_tails(
            _lib.context(_dev(n1)).handle, C.byref(lay.c), knee.data_ptr(), n1.data_ptr(), int(S), sm.data_ptr(),
            int(sm.numel()), int(head), int(fade_out_len), int(tail_bound), tails.data_ptr(), set_off.data_ptr(),
            out_len.data_ptr(), dst_off.data_ptr(), flags.data_ptr(), _stream(n1)), "impb_set_tails")
    return tails, set_off, out_len, dst_off, flags


# ---- K21 --------------------------------------------------------------------------------------------
def equalize(error, logf, cap, gain_k, half_window, smoothen=True):
    """rows of error curves -> (equalization rows, int32 status per row) — ``impb_equalize`` (K21)."""
    torch = _torch()
    if error.dim() == 1:
        error = error[None, :]
    error = error.contiguous()
    dev = error.device
    batch, n = error.shape
    tabs = [_dev_const(t, dev) for t in (logf, cap, gain_k)]
    out = torch.empty_like(error)
    status = torch.zeros(batch, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().impb_equalize(_lib.context(_dev(error)).handle, error.data_ptr(), n, tabs[0].data_ptr(),
                                            tabs[1].data_ptr(), tabs[2].data_ptr(), n, int(half_window),
                                            1 if smoothen else 0, out.data_ptr(), n, status.data_ptr(), batch,
                                            _stream(error)), "impb_equalize")
    return out, status


# ---- K12 + K13 ------------------------------------------------------------------------------------
class MinPhasePlan:
    def __init__(self, n: int, device: int = 0):
        torch = _torch()
        from scipy.signal import get_window
        self.n, self.device = int(n), int(device)
        self.ctx = _lib.context(self.device)
        ham = np.ascontiguousarray(get_window('hamming', 2 * self.n, fftbins=False), dtype=np.float64)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().impb_minphase_create(self.ctx.handle, self.n, ham.ctypes.data, C.byref(h)),
                       "impb_minphase_create")
        self.handle = h
        self._ws = _lib.StreamWorkspace(self.device)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.lib().impb_minphase_destroy(self.handle)
        except Exception:
            pass

    def workspace(self, batch):
        return self._ws.get(_lib.lib().impb_minphase_workspace_bytes(self.handle, int(batch)))


def minphase_plan(n, device=0) -> MinPhasePlan:
    with _plan_lock:
        p = _minphase_plans.get((n, device))
        if p is None:
            p = _minphase_plans[(n, device)] = MinPhasePlan(n, device)
        return p


def minphase_fir(gain, fs):
    """rows of linear gain on linspace(0, fs//2, n) (last = 0) -> n-tap minimum-phase FIRs."""
    torch = _torch()
    if gain.dim() == 1:
        gain = gain[None, :]
    gain = gain.contiguous()
    batch, n = gain.shape
    plan = minphase_plan(n, _dev(gain))
    out = torch.empty_like(gain)
    with torch.cuda.device(gain.device):
        ws = plan.workspace(batch)
        _lib.check(_lib.lib().impb_minphase_fir(plan.handle, gain.data_ptr(), n, int(fs), out.data_ptr(), n, batch, ws.data_ptr(),
                                                _stream(gain)), "impb_minphase_fir")
    return out


# ---- K15 + K9 -------------------------------------------------------------------------------------
def window_resample(x, taps, up, down, decay=None):
    """resample_poly(x * decay_window, up, down, window=taps/up) for every row; taps already scaled by up."""
    torch = _torch()
    if x.dim() == 1:
        x = x[None, :]
    x = x.contiguous()
    dev = x.device
    batch, n = x.shape
    g = math.gcd(int(up), int(down))
    n_out = _lib.lib().impb_resample_out_len(n, int(up), int(down))
    g = math.gcd(int(up), int(down))
    planned = isinstance(taps, np.ndarray) and int(up) // g > 1  # host taps, true rational ratio: ResamplePlan below
    tp = None
    if taps is not None and not planned:
        tp = taps if isinstance(taps, torch.Tensor) else torch.as_tensor(np.asarray(taps, np.float64), device=dev)
    dc = None
    if decay is not None:
        dc = decay if isinstance(decay, torch.Tensor) else torch.as_tensor(np.asarray(decay, np.float64), device=dev)
        dc = dc.reshape(batch, 4).contiguous()
    u, d = int(up) // g, int(down) // g
    if u == 1 and d > 1 and tp is not None and tp.numel() >= 256 and n >= 4096:
        # pure decimation: every output needs ALL taps (16 001 for nnresample), i.e. out[j] is sample j*d of the
        # 'same' convolution of the row with the taps -> one K1 launch (FFT domain) and a strided pick, instead of
        # n_out * ntaps multiply-adds per row (14 rows of 230 400 samples, 192 -> 48 kHz: 15.7 ms direct)
        if dc is not None:
            xw = torch.empty_like(x)
            with torch.cuda.device(dev):
                _lib.check(_lib.lib().impb_window_resample(_lib.context(_dev(x)).handle, x.data_ptr(), n, n, dc.data_ptr(),
                                                           None, 0, 1, 1, xw.data_ptr(), n, batch, _stream(x)),
                           "impb_window_resample")
            x = xw
        full = _decimation_plan(taps if isinstance(taps, np.ndarray) else tp, n, _dev(x)).execute(x)
        return full[:, ::d][:, :n_out].contiguous()
    out = torch.empty((batch, n_out), dtype=torch.float64, device=dev)
    if planned:
        # host taps (what nnresample hands to resample_poly): planned path, per-run tap tables in the kernel's
        # parameter space (csrc/resample.cu: k_resample_runs_c); bit-identical to the generic kernels
        plan = _resample_plan(taps, u, d, _dev(x))
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().impb_resample_execute(plan.handle, x.data_ptr(), n, n,
                                                        None if dc is None else dc.data_ptr(), out.data_ptr(), n_out,
                                                        batch, _stream(x)), "impb_resample_execute")
        return out
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().impb_window_resample(
            _lib.context(_dev(x)).handle, x.data_ptr(), n, n, None if dc is None else dc.data_ptr(),
            None if tp is None else tp.data_ptr(), 0 if tp is None else tp.numel(), u, d,
            out.data_ptr(), n_out, batch, _stream(x)), "impb_window_resample")
    return out


class ResamplePlan:
    """``impb_resample``: host tap tables of one (taps, up, down) on one device."""

    def __init__(self, taps: np.ndarray, up: int, down: int, device: int = 0):
        taps = np.ascontiguousarray(taps, dtype=np.float64)
        self.ctx = _lib.context(device)
        h = C.c_void_p()
        _lib.check(_lib.lib().impb_resample_plan_create(self.ctx.handle, taps.ctypes.data, taps.size, int(up), int(down),
                                                        C.byref(h)), "impb_resample_plan_create")
        self.handle = h

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.lib().impb_resample_plan_destroy(self.handle)
        except Exception:
            pass


_resample_plans = {}
_resample_lock = threading.Lock()


def _resample_plan(taps: np.ndarray, up: int, down: int, device: int) -> ResamplePlan:
    key = (hash(taps.tobytes()), taps.size, int(up), int(down), int(device))
    with _resample_lock:
        plan = _resample_plans.get(key)
        if plan is None:
            if len(_resample_plans) > 16:
                _resample_plans.clear()
            plan = _resample_plans[key] = ResamplePlan(taps, up, down, device)
    return plan


_decim_plans = {}


def _decimation_plan(taps_dev, n, device):
    """Cached K1 plan whose "inverse filter" is the anti-alias FIR (keyed by content, length and device)."""
    from .estimator import DeconvPlan
    h = taps_dev if isinstance(taps_dev, np.ndarray) else taps_dev.detach().cpu().numpy()
    h = np.ascontiguousarray(h, dtype=np.float64)
    key = (h.tobytes(), int(n), int(device))
    plan = _decim_plans.get(key)
    if plan is None:
        if len(_decim_plans) > 16:
            _decim_plans.clear()
        plan = _decim_plans[key] = DeconvPlan(np.ascontiguousarray(h), int(n), 64, int(device))
    return plan


# ---- output packing -----------------------------------------------------------------------------------
def pack_tracks(src, row_offsets, n, bit_depth=32):
    """Interleaved PCM frames [n, n_tracks] (int32, or int16 for 16 bit) from track rows inside flat ``src``.
    ``row_offsets[t]`` = element offset of track t, -1 = silent track."""
    torch = _torch()
    src = src.contiguous()
    offs = as_index_tensor(row_offsets, src.device)
    n_tracks = offs.numel()
    if bit_depth not in (16, 24, 32):
        raise ValueError('Invalid bit depth. Accepted values are 16, 24 and 32.')
    out = torch.empty((int(n), n_tracks), dtype=torch.int16 if bit_depth == 16 else torch.int32, device=src.device)
    with torch.cuda.device(src.device):
        _lib.check(_lib.lib().impb_pack_tracks(_lib.context(_dev(src)).handle, src.data_ptr(), _i64(offs), int(n),
                                               n_tracks, int(bit_depth), out.data_ptr(), _stream(src)),
                   "impb_pack_tracks")
    return out

"""ctypes binding of ``libimpulcifer_b200.so`` (C ABI declared in ``include/impb.h``).

There is no CPU implementation behind this module: if the shared library is missing, or no CUDA device
is present when a context is requested, the call raises.  Build the library with
``python -c "import __graft_entry__ as g; g.build()"`` (or ``make -C impulcifer_pip313_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# IMPB_LIB: developer override used to A/B differently compiled builds of the same sources (tools/)
LIB_PATH = os.environ.get("IMPB_LIB") or os.path.join(_HERE, "libimpulcifer_b200.so")

FMT_F64, FMT_F32, FMT_I32PCM = 0, 1, 2

_lib = None
_lock = threading.Lock()
_contexts = {}

c_i32, c_i64, c_vp, c_dbl = C.c_int32, C.c_int64, C.c_void_p, C.c_double

# name -> (restype, argtypes); mirrors include/impb.h one to one (tests/test_abi.py checks the export list)
SIGNATURES = {
    "impb_version": (C.c_uint32, []),
    "impb_ctx_create": (c_i32, [c_i32, C.POINTER(c_vp)]),
    "impb_ctx_destroy": (None, [c_vp]),
    "impb_last_error": (C.c_char_p, []),
    "impb_deconv_plan_create": (c_i32, [c_vp, c_vp, c_i64, c_i64, c_i32, C.POINTER(c_vp)]),
    "impb_deconv_plan_destroy": (None, [c_vp]),
    "impb_deconv_workspace_bytes": (c_i64, [c_vp, c_i64]),
    "impb_deconv_plan_info": (c_i32, [c_vp, C.POINTER(c_i64), C.POINTER(c_i64), C.POINTER(c_i64)]),
    "impb_deconv_execute": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_i32, c_i64, c_vp, c_i64, c_i32, c_i64, c_i64, c_vp,
                                    c_vp, c_vp, c_i64, c_vp, c_vp]),
    "impb_deconv_blockmax_info": (c_i32, [c_vp, c_i64, C.POINTER(c_i64), C.POINTER(c_i64), C.POINTER(c_i64)]),
    "impb_deconv_launches_per_execute": (c_i32, [c_vp, c_i64]),
    "impb_peak_index": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_dbl, c_vp, c_i64, c_i64, c_vp]),
    "impb_lundeby": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp]),
    "impb_xcorr_lag": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_i64, c_vp]),
    "impb_gather_window": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_dbl, c_vp,
                                   c_i64, c_i64, c_vp]),
    "impb_gather_rows": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_dbl,
                                 c_vp, c_i64, c_vp]),
    "impb_fft_create": (c_i32, [c_vp, c_i64, c_i32, C.POINTER(c_vp)]),
    "impb_fft_destroy": (None, [c_vp]),
    "impb_fft_workspace_bytes": (c_i64, [c_vp, c_i64]),
    "impb_magnitude_db": (c_i32, [c_vp, c_vp, c_i64, c_vp, c_i64, c_i64, c_vp, c_vp]),
    "impb_magnitude_db_decimated": (c_i32, [c_vp, c_vp, c_i64, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_vp, c_vp]),
    "impb_magnitude_db_folded": (c_i32, [c_vp, c_vp, c_i64, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_vp, c_vp]),
    "impb_decay_times_workspace_bytes": (c_i64, [c_i64, c_i64]),
    "impb_decay_times": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_vp, c_i64, c_i64, c_vp]),
    "impb_segment_energy": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_i64, c_vp]),
    "impb_dft_bin_abs": (c_i32, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp]),
    "impb_add_delayed": (c_i32, [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_i32, c_vp, c_vp, c_i64, c_vp]),
    "impb_magnitude_lin": (c_i32, [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "impb_loglerp": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64, c_i64, c_vp]),
    "impb_minphase_gain_grid": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_dbl, c_i32, c_vp, c_i64,
                                        c_i64, c_vp, c_vp]),
    "impb_savgol_blend": (c_i32, [c_vp, c_vp, c_i64, c_i64, c_i32, c_i32, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp,
                                  c_vp, c_vp, c_i64, c_i64, c_vp]),
    "impb_minphase_create": (c_i32, [c_vp, c_i64, c_vp, C.POINTER(c_vp)]),
    "impb_minphase_destroy": (None, [c_vp]),
    "impb_minphase_workspace_bytes": (c_i64, [c_vp, c_i64]),
    "impb_minphase_fir": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_vp, c_i64, c_i64, c_vp, c_vp]),
    "impb_equalize": (c_i32, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_vp, c_i64, c_vp, c_i64, c_vp]),
    "impb_set_heads": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64] + [c_vp] * 8 + [c_vp]),
    "impb_set_shift": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64] + [c_vp] * 10 + [c_vp]),
    "impb_set_shift2": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "impb_set_tails": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i32, c_i64, c_i64, c_i64] + [c_vp] * 5 + [c_vp]),
    "impb_pack_tracks": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_vp, c_vp]),
    "impb_resample_out_len": (c_i64, [c_i64, c_i32, c_i32]),
    "impb_resample_plan_create": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_i32, C.POINTER(c_vp)]),
    "impb_resample_plan_destroy": (None, [c_vp]),
    "impb_resample_execute": (c_i32, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_i64, c_i64, c_vp]),
    "impb_plan_probe": (c_i32, [c_i64, c_i32, c_i32, C.POINTER(c_i64), C.POINTER(c_i64), C.POINTER(c_i32 * 6),
                                C.POINTER(c_i32 * 6)]),
    "impb_window_resample": (c_i32, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_i64, c_i32, c_i32, c_vp, c_i64,
                                     c_i64, c_vp]),
}


class View(C.Structure):
    """``impb_view`` of include/impb.h (all members optional)."""
    _fields_ = [("shift", c_vp), ("valid_lo", c_vp), ("valid_hi", c_vp), ("fade_in", c_vp), ("fade_in_len", c_i64),
                ("hint_max", c_vp), ("hint_argmax", c_vp), ("blk_max", c_vp), ("nblk", c_i64), ("blk_len", c_i64),
                ("blk_first", c_i64)]


class SetLayout(C.Structure):
    """``impb_set_layout`` of include/impb.h."""
    _fields_ = [("n", c_i32), ("n_spk", c_i32), ("left", c_vp), ("right", c_vp), ("delay", c_vp), ("spk_of_row", c_vp),
                ("n_pairs", c_i32), ("pa", c_vp), ("pb", c_vp), ("n_pos", c_i32), ("pos_rows", c_vp), ("pos_k", c_vp),
                ("n_neg", c_i32), ("neg_rows", c_vp), ("neg_k", c_vp), ("n_groups", c_i32), ("gl", c_vp),
                ("g_ref", c_i32), ("n_grows", c_i32), ("g_rows", c_vp), ("g_of_row", c_vp)]


class ImpbError(RuntimeError):
    pass


def lib():
    """The loaded shared library.  Raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise ImpbError(
                        f"{LIB_PATH} is missing: the CUDA extension has not been built and this package has "
                        f"no CPU path.  Run `python -c 'import __graft_entry__ as g; g.build()'`.")
                handle = C.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(handle, name, None)
                    if fn is None:  # header / library mismatch: calling it raises AttributeError
                        continue
                    fn.restype = res
                    fn.argtypes = args
                _lib = handle
    return _lib


def missing_symbols():
    """Entry points declared in include/impb.h that the built library does not export (must be empty)."""
    handle = lib()
    return [n for n in SIGNATURES if not hasattr(handle, n)]


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().impb_last_error()
        raise ImpbError(f"{what or 'impb call'} failed with status {rc}: {msg.decode() if msg else ''}")


class Context:
    """One ``impb_ctx`` per CUDA device, created lazily and shared by every object on that device."""

    def __init__(self, device: int):
        self.device = int(device)
        h = c_vp()
        check(lib().impb_ctx_create(self.device, C.byref(h)), "impb_ctx_create")
        self.handle = h

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                lib().impb_ctx_destroy(self.handle)
        except Exception:
            pass


def context(device: int = 0) -> Context:
    with _lock:
        ctx = _contexts.get(device)
    if ctx is None:
        ctx = Context(device)
        with _lock:
            ctx = _contexts.setdefault(device, ctx)
    return ctx


def default_device() -> int:
    """Device index for objects created without an explicit device: ``IMPULCIFER_B200_DEVICE`` when set (a process
    pinned to one GPU), else the calling THREAD's current torch device, so one dispatcher thread per GPU
    (``torch.cuda.set_device(i)`` in each) uploads to its own GPU."""
    env = os.environ.get("IMPULCIFER_B200_DEVICE")
    if env is not None:
        return int(env)
    import torch
    return int(torch.cuda.current_device()) if torch.cuda.is_available() else 0


def current_stream_ptr(device) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream


class StreamWorkspace:
    """Scratch memory of a plan, one buffer per CUDA stream that launches it: executes issued from two streams must not
    share scratch (nothing orders them), and on one stream the caching allocator already orders a grown buffer after
    the launches that used the old one."""

    def __init__(self, device: int):
        self.device = int(device)
        self._by_stream = {}

    def get(self, need: int):
        import torch
        key = torch.cuda.current_stream(self.device).cuda_stream
        ws = self._by_stream.get(key)
        if ws is None or ws.numel() < need:
            ws = self._by_stream[key] = torch.empty(int(need), dtype=torch.uint8, device=f"cuda:{self.device}")
        return ws

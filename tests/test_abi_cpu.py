"""No-GPU checks: the C-ABI library loads, exports every symbol include/impb.h declares, and the host-only
entry points (planner introspection, resample length) behave."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'impb.h')).read()
    return sorted(set(re.findall(r'IMPB_API\s+[\w\s\*]+?\b(impb_\w+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from impulcifer_pip313_b200 import _lib
    declared = _declared()
    assert len(declared) >= 25
    handle = _lib.lib()
    missing = [n for n in declared if not hasattr(handle, n)]
    assert not missing, missing
    assert sorted(_lib.SIGNATURES) == declared, set(declared) ^ set(_lib.SIGNATURES)
    assert _lib.missing_symbols() == []
    assert handle.impb_version() >= 0x00010000


def test_no_device_fails_loudly():
    """The product path has no CPU fallback: without a CUDA device the context cannot be created."""
    import torch
    from impulcifer_pip313_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.ImpbError):
        _lib.Context(0)
    from impulcifer_pip313_b200 import ImpulseResponseEstimator
    est = ImpulseResponseEstimator(min_duration=0.05, fs=2000)
    with pytest.raises(Exception):
        est.estimate(est.test_signal)


def _probe(n, exact, elem=8):
    from impulcifer_pip313_b200 import _lib
    n1, n2 = C.c_int64(), C.c_int64()
    rc_, rr = (C.c_int32 * 6)(), (C.c_int32 * 6)()
    rc = _lib.lib().impb_plan_probe(n, exact, elem, C.byref(n1), C.byref(n2), C.byref(rc_), C.byref(rr))
    return rc, n1.value, n2.value, [r for r in rc_ if r], [r for r in rr if r]


def test_planner_headline_and_exact_sizes():
    # 6.15 s @ 48 kHz: L = 391270, M = 295270 -> circular length >= L + ceil((M-1)/2) = 538905
    rc, n1, n2, rcol, rrow = _probe(391270 + (295270 - 1) - (295270 - 1) // 2, 0)
    assert rc == 0 and n1 * n2 >= 538905 and (n1, n2) == (768, 768)
    assert rcol == [8, 12, 8] and rrow == [8, 12, 8]
    for n in (19200, 65536, 9600, 32768, 21600, 22500, 57600, 96000, 84375):
        rc, n1, n2, rcol, rrow = _probe(n, 1)
        assert rc == 0 and n1 * n2 == n, n
        assert int(__import__('numpy').prod(rcol)) == n1 and int(__import__('numpy').prod(rrow)) == n2
    # a length with a prime factor > 5 has no exact plan: the call reports it instead of approximating
    rc, *_ = _probe(2 * 3 * 7 * 11 * 13, 1)
    assert rc != 0
    # 88.2 / 96 kHz and 192 kHz sweeps plan the lengths that have specialised half-width-tile kernels
    for n_min, want in ((1077811, (768, 1536)), (990000, (768, 1536)), (2428174, (1536, 1600))):
        rc, n1, n2, rcol, rrow = _probe(n_min, 0)
        assert rc == 0 and n1 * n2 >= n_min and (n1, n2) == want, (n_min, n1, n2)
    assert _probe(2428174, 0)[3:] == ([12, 8, 16], [16, 10, 10])
    # 44.1 kHz shares the 48 kHz plan
    assert _probe(495118, 0)[1:3] == (768, 768)
    # exact-length transforms of the frequency-response chain factor into menu lengths (plan.cu:fixed_radices)
    menu = {768, 675, 432, 400, 300, 256, 243, 162, 160, 135, 128, 125, 120, 108, 81, 75}
    for n, both in ((65536, True), (19200, True), (96000, True), (84375, True), (57600, True), (21600, True)):
        _, n1, n2, _, _ = _probe(n, 1)
        assert (n1 in menu and n2 in menu) == both, (n, n1, n2)


def test_resample_out_len():
    from impulcifer_pip313_b200 import _lib
    f = _lib.lib().impb_resample_out_len
    assert f(96000, 147, 160) == 88200
    assert f(96000, 44100, 48000) == 88200
    assert f(21601, 1, 4) == 5401
    assert f(10, 3, 3) == 10

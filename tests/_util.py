"""Shared helpers for the parity tests."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def checksums(x):
    """Same fingerprints as tests/golden/make_golden.py:checksums."""
    x = np.asarray(x, dtype=np.float64)
    k = np.arange(len(x), dtype=np.float64)
    probe = np.cos(0.37 * k) + 0.5 * np.sin(0.011 * k * 1.0)
    return np.array([x.sum(), np.dot(x, x), np.dot(x, probe), np.abs(x).max()])


def assert_ck(x, ck, rel=1e-9, what=''):
    """Checksums are sums of up to ~4e5 terms: compare against the vector's own scale."""
    got = checksums(x)
    scale = max(np.sqrt(abs(ck[1])) * np.sqrt(len(x)), abs(ck[3]), 1e-300)
    assert abs(got[0] - ck[0]) <= rel * scale, (what, 'sum', got[0], ck[0])
    assert abs(got[1] - ck[1]) <= rel * max(abs(ck[1]), 1e-300), (what, 'energy', got[1], ck[1])
    assert abs(got[2] - ck[2]) <= rel * scale, (what, 'probe', got[2], ck[2])
    assert abs(got[3] - ck[3]) <= rel * max(abs(ck[3]), 1e-300), (what, 'max', got[3], ck[3])


def rel_peak_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


# Minimum-phase FIR gates (relative to the FIR's peak).  The homomorphic transform takes log|H| at the Nyquist bin,
# where the type-II firwin2 filter is analytically zero and the reference sees only the imprint of its own rounding
# (1.2e-12 .. 3.5e-12 at 9 600 taps, 2e-14 .. 8e-14 at 4 800 taps); a relative error e of that ONE bin moves the FIR by
# ~1e-4 e of its peak.  pocketfft itself only knows the bin to ~1e-16 absolute, so the reference is reproducible to
# ~1e-8 (9 600 taps) / ~2e-7 (4 800 taps) of peak between two machines: measured, scipy's formula on the B200 box's
# CPU against the golden recorded in the build container (profiles/r02_minphase_localise.md).  K12 evaluates the bin
# analytically (fr.cu: k_mp_nyq_imprint) and sits at <= 1.5e-8 / <= 2.0e-7 on the same cases; the gates below are
# that floor with headroom, not a tolerance on the transform (residual after removing the one-bin component: 1e-11).
MINPHASE_TOL_F5 = 1e-7    # f_res = 5  -> 9 600 taps @ 48 kHz (EQ worker)
MINPHASE_TOL_F10 = 5e-7   # f_res = 10 -> 4 800 taps @ 48 kHz (channel balance, microphone deviation)

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

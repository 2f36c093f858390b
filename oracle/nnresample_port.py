"""TEST INFRASTRUCTURE ONLY — CPU restatement of the third-party ``nnresample`` package.

PARITY UNPINNED: ``nnresample>=0.2.4`` (reference ``pyproject.toml:36``; call site
``core/impulse_response.py:121-124``) is neither vendored in the reference nor installed in this
image, and the reference's own tests pin only output *length* and ``fs``
(``tests/test_dsp_characterization.py:229-241``).  This file restates the package's published algorithm
("null-on-Nyquist" Kaiser resampler):

  1. reduce ``up/down`` by their gcd, ``max_rate = max(up, down)``
  2. first guess ``firwin(N, 1/max_rate, window=('kaiser', beta))``, N = 16001 taps, beta = 5.0
  3. zero-pad to 2**19, ``rfft``; find the first spectral null between ``1/max_rate`` and
     ``1/max_rate + 2*sqrt(1+(beta/pi)**2)/N``
  4. redesign with the cutoff shifted so that the first null lands on the new Nyquist:
     ``firwin(N, 2/max_rate - firstnull, window=('kaiser', beta))``
  5. ``scipy.signal.resample_poly(x, up, down, window=filt)``

It is installed as ``sys.modules['nnresample']`` by ``oracle/ref_import.py`` when the reference is
imported in the build container, so the reference's ``ImpulseResponse.resample`` runs against the same
restatement the CUDA path is checked against.
"""
from math import gcd

import numpy as np
from scipy import signal

N_FFT = 2 ** 19


def compute_filt(up, down, fc='nn', beta=5.0, N=32001, return_fc=False):
    g = gcd(int(up), int(down))
    up = int(up) // g
    down = int(down) // g
    max_rate = max(up, down)
    sfact = np.sqrt(1 + (beta / np.pi) ** 2)

    init_filt = signal.firwin(N, 1 / max_rate, window=('kaiser', beta))

    nbins = N_FFT / 2 + 1
    padded = np.zeros(N_FFT)
    padded[:N] = init_filt
    ffilt = np.fft.rfft(padded)

    bot = int(np.floor(nbins / max_rate))
    top = int(np.ceil(nbins * (1 / max_rate + 2 * sfact / N)))
    firstnull = (np.argmin(np.abs(ffilt[bot:top])) + bot) / nbins

    if fc == 'nn':
        fc_used = -firstnull + 2 / max_rate
    elif fc == 'kaiser':
        fc_used = 1 / max_rate
    else:
        fc_used = float(fc)
    filt = signal.firwin(N, fc_used, window=('kaiser', beta))
    if return_fc:
        return filt, fc_used
    return filt


def resample(s, up, down, axis=0, fc='nn', **kwargs):
    """``nnresample.resample(data, fs_new, fs_old)`` as called by core/impulse_response.py:123."""
    up = int(up)
    down = int(down)
    kwargs.setdefault('N', 16001)
    kwargs.setdefault('beta', 5.0)
    filt = compute_filt(up, down, fc, **kwargs)
    return signal.resample_poly(s, up, down, axis=axis, window=filt)

"""Synthetic sweep recordings of the shapes BASELINE.json names.

Host-side numpy only; used by tests, ``bench.py`` and ``__graft_entry__.smoke()`` to make inputs.  Not part
of the compute path.  The same seeds give the same recordings on every machine (``np.random.default_rng``).

Layouts (track/column structure of ``core/hrir.py:146-355`` of the reference):
  * 7.1     : speakers FL,FR,FC,SL,SR,BL,BR -> one stereo file, 2 tracks x (2 s + 7 x (2 s + sweep))
  * 11 / 13 : ``TRUEHD_11CH_ORDER`` / ``TRUEHD_13CH_ORDER`` (``core/constants.py:28-29``)
"""
from __future__ import annotations

import numpy as np

LAYOUT_71 = ['FL', 'FR', 'FC', 'SL', 'SR', 'BL', 'BR']
LAYOUT_TRUEHD_11 = ['FL', 'FR', 'FC', 'BL', 'BR', 'SL', 'SR', 'TFL', 'TFR', 'TBL', 'TBR']
LAYOUT_TRUEHD_13 = ['FL', 'FR', 'FC', 'BL', 'BR', 'SL', 'SR', 'TFL', 'TFR', 'TSL', 'TSR', 'TBL', 'TBR']
_LEFT = {'FL', 'SL', 'BL', 'WL', 'TFL', 'TSL', 'TBL'}
_RIGHT = {'FR', 'SR', 'BR', 'WR', 'TFR', 'TSR', 'TBR'}


def _fftconv_head(a: np.ndarray, b: np.ndarray, n_keep: int) -> np.ndarray:
    n = len(a) + len(b) - 1
    nfft = 1 << (n - 1).bit_length()
    return np.fft.irfft(np.fft.rfft(a, nfft) * np.fft.rfft(b, nfft), nfft)[:n_keep]


def room_ir(rng: np.random.Generator, fs: int, delay: int, gain: float, length_s: float = 0.8) -> np.ndarray:
    """g*delta[n-d] + 0.05*g*N(0,1)*10^(-3 n/(fs*RT60)) for n > d."""
    n = int(length_s * fs)
    rt60 = rng.uniform(0.25, 0.6)
    h = np.zeros(n)
    k = np.maximum(np.arange(n) - delay, 0)
    tail = 0.05 * gain * rng.standard_normal(n) * 10.0 ** (-3.0 * k / (fs * rt60))
    h[delay + 1:] = tail[delay + 1:]
    h[delay] = gain
    return h


def pcm32_roundtrip(x: np.ndarray) -> np.ndarray:
    return np.clip(np.rint(x * 2.0 ** 31), -2 ** 31, 2 ** 31 - 1) / 2.0 ** 31


def recording_set(test_signal: np.ndarray, fs: int, speakers=None, seed: int = 1234,
                  silence_s: float = 2.0, noise_sigma: float = 1e-5, level: float = 0.25) -> np.ndarray:
    """One binaural recording file: float64 [2, silence + n_speakers*(silence + len(sweep))].

    Track 0 = left ear, track 1 = right ear; speaker ``i`` plays in column ``i``.
    """
    if speakers is None:
        speakers = LAYOUT_71
    rng = np.random.default_rng(seed)
    sil = int(silence_s * fs)
    col = sil + len(test_signal)
    total = sil + col * len(speakers)
    rec = np.zeros((2, total))
    for i, spk in enumerate(speakers):
        latency = int(rng.uniform(0, 0.5) * fs)
        for ear in (0, 1):
            if spk in _LEFT:
                ipsi = ear == 0
            elif spk in _RIGHT:
                ipsi = ear == 1
            else:
                ipsi = None
            if ipsi is None:
                d, g = 127, 0.8
            elif ipsi:
                d, g = 120, 1.0
            else:
                d, g = 134, 0.6
            h = room_ir(rng, fs, d + latency, g * level, 0.8 + 0.5)
            seg = _fftconv_head(test_signal, h, col)
            a = sil + i * col
            rec[ear, a:a + col] += seg
    rec += noise_sigma * rng.standard_normal(rec.shape)
    return pcm32_roundtrip(rec)


def columns_of(rec: np.ndarray, fs: int, sweep_len: int, n_speakers: int, silence_s: float = 2.0) -> np.ndarray:
    """[2*n_speakers, silence+sweep_len] view-copy: the columns ``estimate()`` is called on, in the
    reference's order (speaker-major, left then right)."""
    sil = int(silence_s * fs)
    col = sil + sweep_len
    out = np.empty((2 * n_speakers, col))
    for i in range(n_speakers):
        a = sil + i * col
        out[2 * i] = rec[0, a:a + col]
        out[2 * i + 1] = rec[1, a:a + col]
    return out

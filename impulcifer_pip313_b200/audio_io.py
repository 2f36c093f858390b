"""WAV container I/O for the shim (reference ``core/audio_io.py:67-97``).

Container parsing is not part of the accelerated path (libsndfile I/O is out of scope); this
module only has to hand float64 ``[tracks, samples]`` arrays to the device code and write the reference's
track layouts back.  ``soundfile`` is used when importable (as in the reference); otherwise PCM WAVs go
through ``scipy.io.wavfile`` with libsndfile's scaling rules (int32 / 2**31, ``lrint(x * 2**31)`` + clip).
"""
from __future__ import annotations

import os

import numpy as np


def read_wav(file_path, expand=False):
    """-> (fs, float64 array with one row per track, samples in -1..1)  (reference :67-79)."""
    try:
        import soundfile as sf
        data, fs = sf.read(file_path, dtype='float64', always_2d=False)
    except ImportError:
        from scipy.io import wavfile
        fs, raw = wavfile.read(file_path)
        if raw.dtype == np.int32:
            data = raw.astype(np.float64) / 2 ** 31
        elif raw.dtype == np.int16:
            data = raw.astype(np.float64) / 2 ** 15
        elif raw.dtype == np.uint8:
            data = (raw.astype(np.float64) - 128) / 2 ** 7
        else:
            data = raw.astype(np.float64)
    if data.ndim > 1:
        data = np.transpose(data)
    elif expand:
        data = np.expand_dims(data, axis=0)
    return int(fs), np.ascontiguousarray(data)


def quantize_pcm(data: np.ndarray, bit_depth: int) -> np.ndarray:
    """libsndfile double -> PCM conversion with clipping on (what python-soundfile selects): every depth is
    ``lrint(x * 2**31)`` clipped to int32, PCM_16 / PCM_24 keep the top 16 / 24 bits of it (``>> 16`` / ``>> 8``,
    libsndfile pcm.c ``d2les_clip_array`` / ``d2let_clip_array``) — a floor, not a round at the target depth."""
    if bit_depth not in (16, 24, 32):
        raise ValueError('Invalid bit depth. Accepted values are 16, 24 and 32.')
    q = np.clip(np.rint(np.asarray(data, dtype=np.float64) * 2.0 ** 31), -2.0 ** 31, 2.0 ** 31 - 1).astype(np.int64)
    q >>= 32 - bit_depth
    return q.astype(np.int16 if bit_depth == 16 else np.int32)


def write_wav(file_path, fs, data, bit_depth=32):
    """Writes tracks-on-rows data as an interleaved PCM WAV (reference :82-97, same transpose rule)."""
    d = os.path.dirname(file_path)
    if d:
        os.makedirs(d, exist_ok=True)
    if bit_depth not in (16, 24, 32):
        raise ValueError('Invalid bit depth. Accepted values are 16, 24 and 32.')
    data = np.asarray(data)
    if len(data.shape) > 1 and data.shape[1] > data.shape[0]:
        data = np.transpose(data)
    try:
        import soundfile as sf
        sf.write(file_path, data, samplerate=fs, subtype={16: "PCM_16", 24: "PCM_24", 32: "PCM_32"}[bit_depth])
        return
    except ImportError:
        pass
    from scipy.io import wavfile
    if bit_depth == 24:
        q = quantize_pcm(data, 24).astype(np.int32) << 8  # stored in a 32-bit container
    else:
        q = quantize_pcm(data, bit_depth)
    wavfile.write(file_path, int(fs), q)


def magnitude_response(x, fs):
    """Reference ``core/audio_io.py:100-113``: ``(f, 20*log10|rfft(x)|)`` for the first ``ceil(n/2)`` bins,
    computed on the device (K8).  No CPU path: 5-smooth lengths use the two-pass FFT, every other length the chirp-z
    (Bluestein) plan built on the same passes."""
    import torch
    from . import ops
    x = np.ascontiguousarray(x, dtype=np.float64)
    nfft = len(x)
    half = int(np.ceil(nfft / 2))
    from ._lib import default_device
    dev = torch.device("cuda", default_device())
    mag = ops.magnitude_db(torch.from_numpy(x).to(dev)).cpu().numpy()[0]
    return np.arange(half) * (fs / nfft), mag

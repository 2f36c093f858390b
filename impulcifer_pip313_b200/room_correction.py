"""Room / headphone frequency-response chains on the device (next row).

Mirrors the arithmetic of the reference's ``core/room_correction.py:185-306`` and of the headphone
compensation block ``core/pipeline_stages.py:421-438``; file discovery, CSV parsing and plotting stay in the
reference (they are outside the hot path).  The expensive part of every function here is the magnitude
response of each impulse response: it is ONE K8 + K10 launch pair over all responses of equal length
(``impulse_response.frequency_response_batch``) instead of one rfft per response; smoothing runs as one K11
launch over all curves.  What remains per curve is O(783) arithmetic on the host side of the shim.
"""
from __future__ import annotations

import numpy as np

from . import ops
from .constants import IR_ROOM_SPL
from .frequency_response import FrequencyResponse, smoothen_batch
from .impulse_response import frequency_response_batch


def _hann(m: int) -> np.ndarray:
    """Symmetric Hann window of ``m`` points (``scipy.signal.windows.hann(m)``)."""
    if m <= 0:
        return np.zeros(0)
    if m == 1:
        return np.ones(1)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(m) / (m - 1))


def correction_limit_mask(frequency, limit):
    """1 below ``limit/2``, falling Hann half... (reference ``_correction_limit_mask``, :293-300): ones up to the
    first frequency above ``limit/2``, a full Hann window of the points up to the first frequency above ``limit``,
    zeros after it."""
    frequency = np.asarray(frequency)
    start = int(np.argmax(frequency > limit / 2))
    end = int(np.argmax(frequency > limit))
    return np.concatenate([np.ones(max(start, 0)), _hann(end - start), np.zeros(len(frequency) - end)])


def apply_correction_limit(fr, limit):
    fr.error *= correction_limit_mask(fr.frequency, limit)


def frequency_responses(irs):
    """``[ImpulseResponse.frequency_response() for ir in irs]`` with one launch pair per distinct length."""
    irs = list(irs)
    out = [None] * len(irs)
    by_len = {}
    for i, ir in enumerate(irs):
        by_len.setdefault((len(ir), int(ir.fs)), []).append(i)
    import torch
    for (n, fs), idx in by_len.items():
        grid = FrequencyResponse.generate_frequencies(f_step=1.01, f_min=10, f_max=fs / 2)
        if n < 2:
            for i in idx:
                out[i] = FrequencyResponse(name="Frequency response (short IR)", frequency=grid.copy(),
                                           raw=np.zeros_like(grid))
            continue
        rows = torch.stack([irs[i].device_data() for i in idx])
        raw = frequency_response_batch(rows, fs).cpu().numpy()
        for k, i in enumerate(idx):
            out[i] = FrequencyResponse(name="Frequency response", frequency=grid.copy(), raw=raw[k].copy())
    return out


def calculate_specific_room_corrections(rir, target, mic_calibration=None, limit=400):
    """Speaker-ear correction responses of an already loaded (cropped) room HRIR.  Reference
    ``core/room_correction.py:185-212``: the first response fixes the reference gain of the whole set."""
    items = [(speaker, side, ir) for speaker, pair in rir.irs.items() for side, ir in pair.items()]
    curves = frequency_responses(ir for _, _, ir in items)
    frs = {}
    reference_gain = None
    for (speaker, side, _), fr in zip(items, curves):
        if mic_calibration is not None:
            fr.raw -= mic_calibration.raw
        if reference_gain is None:
            reference_gain = fr.center([100, 10000])
        else:
            fr.raw += reference_gain
        wanted = target.copy()
        wanted.raw += IR_ROOM_SPL[speaker][side]
        fr.compensate(wanted, min_mean_error=False)
        if limit > 0:
            apply_correction_limit(fr, limit)
        frs.setdefault(speaker, {})[side] = fr
    return frs


def calculate_generic_room_correction(irs, target, mic_calibration=None, method='average', limit=1000,
                                      collect_raws=False):
    """Combine generic room responses into one correction curve.  Reference ``:234-290``.  Returns the
    ``FrequencyResponse`` (and the centred per-response curves when ``collect_raws``)."""
    irs = list(irs)
    if method not in ('average', 'conservative') and len(irs) > 1:
        raise ValueError(f'Invalid value "{method}" for method. Supported values are "conservative" and "average"')
    grid = FrequencyResponse.generate_frequencies(f_min=10, f_max=irs[0].fs / 2, f_step=1.01)
    room_fr = FrequencyResponse(name='generic_room', frequency=grid, raw=0, error=0, target=target.raw)
    curves = frequency_responses(irs)
    raws = []
    for fr in curves:
        if mic_calibration is not None:
            fr.raw -= mic_calibration.raw
        fr.center([100, 10000])
        room_fr.raw += fr.raw
        if collect_raws:
            raws.append(fr.copy())
        fr.compensate(target, min_mean_error=True)
    errors = np.vstack([fr.error for fr in curves])
    if method == 'conservative' and len(irs) > 1:
        # fr.smoothen(1/3, 1/3) of every error curve: one K11 launch
        errors = smoothen_batch(grid, errors, 1 / 3, 1, 1 / 3, 1, 100.0, 10000.0)
    room_fr.raw /= len(irs)
    if errors.shape[0] > 1:
        if method == 'conservative':
            share = np.mean(errors > 0, axis=0)
            positive, negative = share == 1, share == 0
            room_fr.error[positive] = np.min(errors[:, positive], axis=0)
            room_fr.error[negative] = np.max(errors[:, negative], axis=0)
            room_fr.smoothen(window_size=1 / 6, treble_window_size=1 / 6)
            room_fr.error = room_fr.error_smoothed.copy()
        else:
            room_fr.error = np.mean(errors, axis=0)
            room_fr.smoothen(window_size=1 / 3, treble_window_size=1 / 3)
    else:
        room_fr.error = errors[0, :]
        room_fr.smoothen(window_size=1 / 3, treble_window_size=1 / 3)
    if limit > 0:
        apply_correction_limit(room_fr, limit)
        room_fr.error_smoothed *= correction_limit_mask(room_fr.frequency, limit)
    return (room_fr, raws) if collect_raws else room_fr


def headphone_compensation(hp_irs):
    """Left / right headphone error curves against a flat target (reference ``core/pipeline_stages.py:421-438``):
    FL-left and FR-right responses, the right one shifted by the left one's centring gain."""
    left, right = frequency_responses([hp_irs.irs["FL"]["left"], hp_irs.irs["FR"]["right"]])
    gain = left.center([100, 10000])
    right.raw += gain
    zero = FrequencyResponse(name="zero", frequency=left.frequency, raw=np.zeros(len(left.frequency)))
    left.compensate(zero, min_mean_error=False)
    right.compensate(zero, min_mean_error=False)
    return left, right

"""Virtual-bass synthesis on the device (next row).  Mirrors ``core/virtual_bass.py:82-197`` of
the reference: a shared band-limited bass impulse, gain-matched at the crossover to the mean of all high-passed
responses, is added (with the head delay and the interaural delay) to every response after an 8th-order high-pass.

Device mapping:
  * the reference's ``sosfilt(sos_hp8_xo, ir)`` is a causal LTI filter started from rest, i.e. the first n samples
    of ``ir (*) h`` with ``h = sosfilt(sos_hp8_xo, unit impulse)``: ONE K1 launch over all responses (the filter's
    own impulse response is a plan constant computed once on the host, like the inverse sweep filter);
  * ``_mag_at`` = one DFT bin per response: ``impb_dft_bin_abs`` (K19);
  * onset peaks: one ``impb_peak_index`` (K2) launch;
  * ``_delay_signal`` + the final add: ``impb_add_delayed`` (K20), one launch.
The filter designs (Butterworth sections, RBJ shelves) and the two synthetic bass vectors are host constants.
"""
from __future__ import annotations

import numpy as np

from . import ops
from .constants import speaker_side

SHELVES = ((150.0, -1.5, 0.760), (400.0, -3.0, 0.660), (800.0, -3.5, 0.610))  # core/virtual_bass.py:113-117


def rbj_high_shelf(fc, fs, gain_db, q):
    """RBJ cookbook high-shelf biquad as one SOS row (reference :66-79)."""
    from scipy import signal
    a = 10 ** (gain_db / 40.0)
    w0 = 2 * np.pi * fc / fs
    alpha = np.sin(w0) / (2 * q)
    cw = np.cos(w0)
    ra = 2 * np.sqrt(a) * alpha
    b = [a * ((a + 1) + (a - 1) * cw + ra), -2 * a * ((a - 1) + (a + 1) * cw), a * ((a + 1) + (a - 1) * cw - ra)]
    den = [(a + 1) - (a - 1) * cw + ra, 2 * ((a - 1) - (a + 1) * cw), (a + 1) - (a - 1) * cw - ra]
    return signal.tf2sos(b, den)


def bass_constants(n_ir, fs, crossover_freq, hp_freq):
    """Host constants of one synthesis: impulse response of the 8th-order crossover high-pass, the shared bass impulse
    ``mpbass`` and the ILD shelf chain (reference :104-121)."""
    from scipy import signal
    imp = np.zeros(n_ir)
    imp[0] = 1.0
    nyq = fs / 2
    mpbass = signal.sosfilt(signal.butter(4, hp_freq / nyq, btype="high", output="sos"), imp)
    lp4 = signal.butter(4, crossover_freq / nyq, btype="low", output="sos")
    mpbass = signal.sosfilt(np.vstack([lp4, lp4]), mpbass)
    hp4 = signal.butter(4, crossover_freq / nyq, btype="high", output="sos")
    h_hp8 = signal.sosfilt(np.vstack([hp4, hp4]), imp)
    sos_ild = np.vstack([rbj_high_shelf(fc, fs, gain, q) for fc, gain, q in SHELVES])
    return h_hp8, mpbass, sos_ild


def synthesize_virtual_bass(irs, fs, crossover_freq=250, head_ms=1.0, hp_freq=15.0, invert_polarity=None):
    """In-place virtual bass for ``irs = {speaker: {'left': ImpulseResponse, 'right': ImpulseResponse}}``."""
    import torch
    from scipy import signal
    from .dispatcher import fft_convolve_full
    if crossover_freq >= fs / 2:
        return
    every = [ir for pair in irs.values() for ir in pair.values()]
    if not every:
        return
    n_ir = max(len(ir) for ir in every)
    pairs = [(spk, pair) for spk, pair in irs.items() if "left" in pair and "right" in pair]
    dev = every[0].device_data().device
    # zero-pad every response to the common length (pairs and unpaired alike, as the reference does)
    for ir in every:
        if len(ir) < n_ir:
            x = torch.zeros(n_ir, dtype=torch.float64, device=dev)
            x[:len(ir)] = ir.device_data()
            ir._set_device(x)
    if not pairs:
        return
    h_hp8, mpbass, sos_ild = bass_constants(n_ir, fs, crossover_freq, hp_freq)
    rows = [pair[side] for _, pair in pairs for side in ("left", "right")]
    hi = fft_convolve_full(torch.stack([ir.device_data() for ir in rows]), h_hp8, out_len=n_ir)
    # gain that matches the bass impulse to the mean high-passed level at the crossover
    freqs = np.fft.rfftfreq(n_ir, 1.0 / fs)
    k = int(np.argmin(np.abs(freqs - crossover_freq)))
    mean_xo_mag = float(np.mean(ops.dft_bin_abs(hi, k)))
    mpbass_mag = float(np.abs(np.fft.rfft(mpbass))[k]) + 1e-20
    gain = mean_xo_mag / mpbass_mag
    polarity = -1.0 if invert_polarity else 1.0
    direct = mpbass * gain * polarity
    cross = signal.sosfilt(sos_ild, direct)
    synth = torch.as_tensor(np.stack([direct, cross]), device=dev)
    # onset peaks of the (padded, unfiltered) responses
    flat = torch.stack([ir.device_data() for ir in rows]).reshape(-1)
    offs = np.arange(len(rows), dtype=np.int64) * n_ir
    peaks = ops.peak_index(flat, offs, np.full(len(rows), n_ir, dtype=np.int64), max_len=n_ir).cpu().numpy()
    head = int(round(head_ms * 1e-3 * fs))
    which, delay = [], []
    for i, (spk, _) in enumerate(pairs):
        on_left = speaker_side(spk) == "left"
        itd = int(peaks[2 * i + 1]) - int(peaks[2 * i])
        cross_delay = head + (itd if on_left else -itd)
        # left ear, right ear: 0 = direct, 1 = cross
        which += [0, 1] if on_left else [1, 0]
        delay += [head, cross_delay] if on_left else [cross_delay, head]
    ops.add_delayed(hi, synth, which, delay)
    for i, ir in enumerate(rows):
        ir._set_device(hi[i])


def apply_virtual_bass_to_hrir(hrir, crossover_freq=250, head_ms=1.0, hp_freq=15.0, invert_polarity=None):
    """Reference :178-197."""
    synthesize_virtual_bass(irs=hrir.irs, fs=hrir.fs, crossover_freq=crossover_freq, head_ms=head_ms, hp_freq=hp_freq,
                            invert_polarity=invert_polarity)

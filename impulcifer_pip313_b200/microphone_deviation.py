"""Interaural microphone-mismatch correction (next row): device version of the reference's
``core/microphone_deviation_correction.py:56-276, 386-455`` (``MicrophoneMatchingCorrector`` v4.0).

The per-speaker work — window the direct sound around the peak, |rfft| at 8192+ points, resample the magnitude onto
the 1.01-step grid — is batched over every (speaker, ear) response: one ``impb_gather_window`` (crop + both fades),
one ``impb_magnitude_lin`` and one ``impb_loglerp`` launch.  The curve arithmetic on the 700-point grid (power
averages, band taper, clamps) is host numpy, the smoothing and the two minimum-phase FIR designs go through the
``FrequencyResponse`` shim (K11, K12+K13), and the filters are applied with one K1 launch per ear.
"""
from __future__ import annotations

import warnings

import numpy as np

from . import ops
from .frequency_response import FrequencyResponse, next_fast_len

_CENTER_SPEAKERS = ("FC", "TFC", "BC")
_SINGLE_SPEAKER = "SINGLE"


def _device():
    import os
    import torch
    from ._lib import default_device
    return torch.device("cuda", default_device())


class MicrophoneMatchingCorrector:
    def __init__(self, sample_rate, correction_strength=0.7, max_correction_db=6.0, smoothing_octave=1.0 / 6.0,
                 f_min=200.0, f_max=16000.0, window_ms=5.0, pre_ms=0.5, anchor="auto"):
        self.fs = int(sample_rate)
        self.correction_strength = float(np.clip(correction_strength, 0.0, 1.0))
        self.max_correction_db = float(max_correction_db)
        self.smoothing_octave = float(smoothing_octave)
        nyq = self.fs / 2.0
        self.f_min = float(np.clip(f_min, 1.0, nyq * 0.5))
        self.f_max = float(np.clip(f_max, self.f_min * 2.0, nyq * 0.98))
        self.window_ms = float(window_ms)
        self.pre_ms = float(pre_ms)
        self.anchor = anchor
        self.win_samples = max(int(round(self.window_ms * self.fs / 1000.0)), 32)
        self.pre_samples = max(int(round(self.pre_ms * self.fs / 1000.0)), 0)
        self.frequency = FrequencyResponse.generate_frequencies(f_step=1.01, f_min=20.0, f_max=nyq)
        self.speaker_power = {}
        self.mismatch_db = None
        self.anchor_used = None

    # ---- direct-sound power spectra (reference :106-140), batched ----------------------------------------------
    def windowed_power_batch(self, rows, peaks):
        """``rows``: list of 1-D float64 CUDA tensors (or numpy arrays); ``peaks``: peak index per row (None =
        argmax|x|).  Returns numpy [len(rows), F] power on ``self.frequency``."""
        import torch
        dev = _device()
        rows = [r if isinstance(r, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(r, dtype=np.float64), device=dev)
                for r in rows]
        out = np.zeros((len(rows), len(self.frequency)))
        plan = {}
        for i, (x, pk) in enumerate(zip(rows, peaks)):
            n = int(x.shape[0])
            if n == 0:
                continue
            if pk is None:
                pk = int(torch.argmax(torch.abs(x)).item())
            pk = int(np.clip(int(pk), 0, n - 1))
            start, end = max(pk - self.pre_samples, 0), min(pk + self.win_samples, n)
            if end - start < 8:
                continue
            plan.setdefault(end - start, []).append((i, start))
        for seg_len, members in plan.items():
            fade_in = min(self.pre_samples, seg_len // 4)
            fade_out = max(seg_len // 4, 1)
            fi = np.hanning(2 * fade_in)[:fade_in] if fade_in > 1 else None
            fo = np.hanning(2 * fade_out)[fade_out:] if fade_out > 1 else None
            flat = torch.cat([rows[i] for i, _ in members])
            base = np.concatenate([[0], np.cumsum([int(rows[i].shape[0]) for i, _ in members])[:-1]])
            seg = ops.gather_window(flat, [int(b) + s for b, (_, s) in zip(base, members)], seg_len, fade_in=fi,
                                    fade_out=fo)
            nfft = next_fast_len(max(seg_len, 8192))
            mag = ops.magnitude_lin(seg, nfft)
            fft_freq = np.fft.rfftfreq(nfft, 1.0 / self.fs)
            # np.interp(self.frequency, fft_freq, mag): the grid lies inside [0, fs/2], no clamping needed
            grid_mag = ops.loglerp(fft_freq, mag, self.frequency).cpu().numpy()
            for k, (i, _) in enumerate(members):
                out[i] = grid_mag[k] ** 2
        return out

    def _windowed_power(self, ir, peak_index):
        return self.windowed_power_batch([np.asarray(ir, dtype=float)], [peak_index])[0]

    def collect_speaker(self, speaker_name, left_ir, right_ir, left_peak_index=None, right_peak_index=None):
        p = self.windowed_power_batch([left_ir, right_ir], [left_peak_index, right_peak_index])
        self.speaker_power[speaker_name] = {"left": p[0], "right": p[1]}
        return self.speaker_power[speaker_name]

    def collect_hrir(self, hrir, peaks=None):
        """All speakers of an HRIR in one batch; ``peaks``: optional {speaker: (left_peak, right_peak)}."""
        names = list(hrir.irs)
        rows, pk = [], []
        for s in names:
            for k, side in enumerate(("left", "right")):
                rows.append(hrir.irs[s][side].device_data())
                pk.append(None if peaks is None else peaks[s][k])
        p = self.windowed_power_batch(rows, pk)
        for k, s in enumerate(names):
            self.speaker_power[s] = {"left": p[2 * k], "right": p[2 * k + 1]}

    def collect_speaker_deviation(self, speaker_name, left_ir, right_ir, left_peak_index=None, right_peak_index=None):
        self.collect_speaker(speaker_name, left_ir, right_ir, left_peak_index, right_peak_index)
        p = self.speaker_power[speaker_name]
        delta = 10.0 * np.log10((p["left"] + 1e-20) / (p["right"] + 1e-20))
        return {b: float(np.interp(b, self.frequency, delta)) for b in (250, 500, 1000, 2000, 4000, 8000)
                if b < self.fs / 2}

    # ---- mismatch curve (reference :165-242) --------------------------------------------------------------------
    def _band_weight(self):
        logf = np.log10(self.frequency)
        w = np.ones_like(self.frequency)
        lo_edge, lo_zero = np.log10(self.f_min), np.log10(max(self.f_min / 2.0, 1.0))
        hi_edge = np.log10(self.f_max)
        hi_zero = np.log10(min(self.f_max * 2.0, self.fs / 2.0 * 0.999))
        rising = (logf >= lo_zero) & (logf < lo_edge)
        w[logf < lo_zero] = 0.0
        if np.any(rising):
            w[rising] = 0.5 - 0.5 * np.cos(np.pi * (logf[rising] - lo_zero) / max(lo_edge - lo_zero, 1e-9))
        falling = (logf > hi_edge) & (logf <= hi_zero)
        w[logf > hi_zero] = 0.0
        if np.any(falling):
            w[falling] = 0.5 + 0.5 * np.cos(np.pi * (logf[falling] - hi_edge) / max(hi_zero - hi_edge, 1e-9))
        return w

    def estimate_interaural_mismatch(self):
        if not self.speaker_power:
            warnings.warn("no speaker data collected; call collect_speaker first")
            self.mismatch_db = np.zeros_like(self.frequency)
            self.anchor_used = "none"
            return self.mismatch_db
        center = [s for s in self.speaker_power if s in _CENTER_SPEAKERS]
        if self.anchor in ("frontal", "auto") and center:
            pool, self.anchor_used = center, "frontal"
        else:
            pool, self.anchor_used = list(self.speaker_power), "diffuse"
        left = np.mean([self.speaker_power[s]["left"] for s in pool], axis=0)
        right = np.mean([self.speaker_power[s]["right"] for s in pool], axis=0)
        raw_delta = 10.0 * np.log10((left + 1e-20) / (right + 1e-20))
        fr = FrequencyResponse(name="interaural_mismatch", frequency=self.frequency.copy(), raw=raw_delta)
        try:
            fr.smoothen(window_size=self.smoothing_octave, treble_window_size=self.smoothing_octave)
            smoothed = fr.smoothed if len(fr.smoothed) else raw_delta
        except ValueError:
            smoothed = raw_delta
        smoothed = smoothed * self._band_weight()
        self.mismatch_db = np.clip(smoothed, -2.0 * self.max_correction_db, 2.0 * self.max_correction_db)
        return self.mismatch_db

    def separate_microphone_error(self):
        delta = self.estimate_interaural_mismatch()
        return {b: float(np.interp(b, self.frequency, delta)) for b in (250, 500, 1000, 2000, 4000, 8000)
                if b < self.fs / 2}

    # ---- correction filters (reference :244-276) ----------------------------------------------------------------
    def design_correction_filters(self):
        if self.mismatch_db is None:
            self.estimate_interaural_mismatch()
        half = np.clip(self.mismatch_db * self.correction_strength / 2.0, -self.max_correction_db,
                       self.max_correction_db)
        return self._fir_from_curve(-half, "left_mic_correction"), self._fir_from_curve(half, "right_mic_correction")

    def _fir_from_curve(self, curve_db, name):
        fr = FrequencyResponse(name=name, frequency=self.frequency.copy(), raw=curve_db.copy())
        fr.equalization = curve_db.copy()
        fir = fr.minimum_phase_impulse_response(fs=self.fs, normalize=False)
        return fir[:min(2048, self.fs // 10)]

    def get_analysis_summary(self):
        if self.mismatch_db is None:
            return {"error": "analysis not run"}
        applied = np.clip(self.mismatch_db * self.correction_strength / 2.0, -self.max_correction_db,
                          self.max_correction_db)
        nz = np.abs(applied[self._band_weight() > 0])
        return {"method": "interaural_v4", "anchor": self.anchor_used,
                "avg_error_db": float(np.mean(nz)) if len(nz) else 0.0,
                "max_error_db": float(np.max(nz)) if len(nz) else 0.0,
                "speakers_analyzed": list(self.speaker_power.keys()),
                "correction_strength": self.correction_strength}


def apply_microphone_deviation_correction_to_hrir(hrir, correction_strength=0.7, anchor="auto", plot_analysis=False,
                                                  plot_dir=None):
    """Reference :386-455 (plots are the reference's business).  Returns the analysis summary dict."""
    corrector = MicrophoneMatchingCorrector(sample_rate=hrir.fs, correction_strength=correction_strength, anchor=anchor)
    if not hrir.irs:
        return {"error": "no speaker data"}
    peaks = {s: (pair["left"].peak_index(), pair["right"].peak_index()) for s, pair in hrir.irs.items()}
    corrector.collect_hrir(hrir, peaks)
    corrector.estimate_interaural_mismatch()
    summary = corrector.get_analysis_summary()
    if summary["max_error_db"] < 0.05:
        summary["speakers_processed"] = []
        return summary
    left_fir, right_fir = corrector.design_correction_filters()
    hrir._equalize_rows([(pair[side], left_fir if side == "left" else right_fir)
                         for pair in hrir.irs.values() for side in ("left", "right")])
    summary["speakers_processed"] = list(hrir.irs.keys())
    return summary

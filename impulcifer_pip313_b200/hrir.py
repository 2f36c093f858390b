"""``HRIR`` — drop-in for ``core/hrir.py`` of the reference: recording ingest (track x time-column split),
head/tail crop, onset alignment, normalisation, resampling and WAV track layout.

Every method keeps the reference's signature, in-place mutation semantics, exceptions and warnings; the
arithmetic runs as batched launches over all (speaker, ear) responses of the set:
  open_recording  one ``impb_deconv_execute`` for all columns of the file      (ref :76-363, 396-424)
  crop_heads      one ``impb_peak_index`` + one ``impb_gather_window``           (ref :548-612)
  align_*         ``impb_xcorr_lag`` / ``impb_peak_index`` + gather shifts       (ref :921-1001)
  crop_tails      one ``impb_lundeby`` + one ``impb_gather_window``              (ref :614-653)
  normalize       ``impb_magnitude_db`` of the summed ears + gain in the gather  (ref :457-546)
  resample        one ``impb_window_resample``                                   (ref :890-919)
The integer bookkeeping (column map, crop indices, lags, tail index) is host Python, exactly as §8(b) of
SURVEY.md assigns it, and is bit-exact with the reference.
"""
from __future__ import annotations

import warnings

import numpy as np

from . import ops
from .audio_io import read_wav, write_wav
from .constants import (HEXADECAGONAL_TRACK_ORDER, IPSILATERAL_PAIRS, ONSET_GROUPS, SPEAKER_DELAYS, SPEAKER_NAMES,
                        speaker_side, track_name)
from .frequency_response import FrequencyResponse, next_fast_len
from .impulse_response import ImpulseResponse, resample_taps


def ingest_layout(n_tracks, n_samples, speakers, est_len, fs, side=None, silence_length=2.0):
    """Integer logic of ``_ingest_recording`` (ref :146-355): which (track, sample range) becomes which
    (speaker, side).  Returns ``(columns, plan)``: ``columns[j] = (start, end)`` in recording samples,
    ``plan = [(speaker, side, track, column)]`` in the order the reference estimates them."""
    if silence_length * fs != int(silence_length * fs):
        raise ValueError("Silence length must produce full samples with given sampling rate.")
    sil = int(silence_length * fs)
    tracks_k = 2 if side is None else 1
    n_columns = round(len(speakers) / (n_tracks // tracks_k))
    avail = n_samples - sil
    col = sil + est_len
    if col > avail:
        if n_columns <= 1:
            col, n_columns = avail, 1
        else:
            col = avail // n_columns
    base = sil

    def cut(origin, width, limit, need):
        out = []
        for i in range(n_columns):
            a, b = i * width, min((i + 1) * width, limit)
            if b > a and (b - a) >= need:
                out.append((origin + a, origin + b))
        return out

    columns = cut(base, col, avail, est_len)
    if not columns:
        if sil > 0:  # fallback 1: shorter leading silence, back-to-back sweeps (ref :221-262)
            min_sil = int(0.5 * fs)
            spare = avail - est_len
            if spare >= min_sil:
                adj = max(min_sil, spare)
                columns = cut(sil + adj, est_len, avail - adj, est_len)
                if columns:
                    base, avail = sil + adj, avail - adj
        if not columns and avail > est_len * 0.8:  # fallback 2: accept >= 80 % of a sweep (ref :264-287)
            columns = [(base, base + avail)] if n_columns == 1 else cut(base, avail // n_columns, avail, 1)
        if not columns:
            raise ValueError(
                f"No valid columns could be extracted even with fallback methods.\n"
                f"Recording length ({avail} samples, {avail / fs:.2f}s) is too short "
                f"for the required estimator length ({est_len} samples, {est_len / fs:.2f}s).\n"
                f"Solutions:\n"
                f"  1. Re-record with longer duration (minimum {(est_len + sil) / fs:.1f}s)\n"
                f"  2. Use a shorter test signal\n"
                f"  3. Check if the correct test signal file was used for recording")
    plan = []
    i = 0
    while i < n_tracks:
        for j in range(len(columns)):
            n = int(i // 2 * len(columns) + j)
            if n >= len(speakers):
                continue
            speaker = speakers[n]
            if speaker not in SPEAKER_NAMES:
                continue
            if side is None:
                if i + 1 < n_tracks:
                    plan.append((speaker, 'left', i, j))
                    plan.append((speaker, 'right', i + 1, j))
            else:
                plan.append((speaker, side, i, j))
        i += tracks_k
    return columns, plan


def _pack(irs):
    """list of ImpulseResponse -> (flat CUDA tensor, offsets, lengths) without copying when they already
    are views of one storage (the usual case: rows of one deconvolution output)."""
    import torch
    tens = [ir.device_data() for ir in irs]
    lens = [int(t.shape[0]) for t in tens]
    base = tens[0].untyped_storage().data_ptr() if tens else 0
    shared = all(t.untyped_storage().data_ptr() == base and t.stride(0) == 1 for t in tens)
    if shared and tens:
        st = tens[0].untyped_storage()
        flat = torch.empty(0, dtype=torch.float64, device=tens[0].device).set_(st, 0, (st.nbytes() // 8,))
        return flat, [int(t.storage_offset()) for t in tens], lens
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64).tolist() if tens else []
    return (torch.cat(tens) if tens else None), offs, lens


class HRIR:
    def __init__(self, estimator):
        self.estimator = estimator
        self.fs = self.estimator.fs
        self.irs = dict()

    def copy(self):
        hrir = HRIR(self.estimator)
        hrir.irs = {spk: {"left": pair["left"].copy(), "right": pair["right"].copy()} for spk, pair in self.irs.items()}
        return hrir

    def subset(self, speakers, copy_irs=False):
        hrir = HRIR(self.estimator)
        for speaker in speakers:
            pair = self.irs.get(speaker)
            if pair is not None:
                hrir.irs[speaker] = {side: ir.copy() if copy_irs else ir for side, ir in pair.items()}
        return hrir

    def _all(self):
        return [(spk, side, ir) for spk, pair in self.irs.items() for side, ir in pair.items()]

    # ---- ingest ----------------------------------------------------------------------------------------
    def open_recording(self, file_path, speakers, side=None, silence_length=2.0, debug=False):
        if self.fs != self.estimator.fs:
            raise ValueError("Refusing to open recording because HRIR's sampling rate doesn't match impulse response "
                             "estimator's sampling rate.")
        fs, recording = read_wav(file_path, expand=True)
        self.ingest(recording, fs, speakers, side=side, silence_length=silence_length)

    def ingest(self, recording, fs, speakers, side=None, silence_length=2.0):
        """``_ingest_recording`` on an already decoded ``[tracks, samples]`` float64 array (numpy or CUDA)."""
        import torch
        if fs != self.fs:
            raise ValueError("Sampling rate of recording must match sampling rate of test signal.")
        is_np = isinstance(recording, np.ndarray)
        if is_np:
            recording = np.atleast_2d(np.ascontiguousarray(recording, dtype=np.float64))
        n_tracks, n_samples = recording.shape
        columns, plan = ingest_layout(n_tracks, n_samples, speakers, len(self.estimator), fs, side, silence_length)
        dev = torch.device("cuda", self.estimator.device)
        rec_d = torch.from_numpy(recording).to(dev) if is_np else recording.contiguous()
        # group by column length (only a clipped last column differs) -> one launch per length
        by_len = {}
        for idx, (spk, sd, trk, col) in enumerate(plan):
            a, b = columns[col]
            by_len.setdefault(b - a, []).append((idx, trk * n_samples + a))
        results = [None] * len(plan)
        with torch.cuda.device(dev):
            for length, rows in by_len.items():
                offs = torch.as_tensor([r[1] for r in rows], dtype=torch.int64, device=dev)
                out = self.estimator.plan(length).execute(rec_d.reshape(-1), length=length, row_offsets=offs)
                for k, (idx, _) in enumerate(rows):
                    results[idx] = out[k]
        for (spk, sd, trk, col), est in zip(plan, results):
            a, b = columns[col]
            view = recording[trk, a:b] if is_np else recording[trk, a:b]
            ir = ImpulseResponse(est.to(torch.float64), fs, view)
            self.irs.setdefault(spk, {})[sd] = ir

    # ---- output layout ---------------------------------------------------------------------------------
    def track_matrix(self, track_order=None):
        if track_order is None:
            track_order = HEXADECAGONAL_TRACK_ORDER
        by_name = {track_name(spk, side): ir.data for spk, side, ir in self._all()}
        if not by_name:
            raise ValueError("No impulse responses available for WAV output.")
        ref_len = len(next(iter(by_name.values())))
        return np.vstack([by_name.get(ch, np.zeros(ref_len)) for ch in track_order])

    def pcm_frames(self, track_order=None, bit_depth=32):
        """Interleaved PCM frames ``[n, n_tracks]`` exactly as ``write_wav`` stores them (gather in track order,
        quantise, interleave on the device); host numpy int array."""
        if track_order is None:
            track_order = HEXADECAGONAL_TRACK_ORDER
        if bit_depth not in (16, 24, 32):
            raise ValueError('Invalid bit depth. Accepted values are 16, 24 and 32.')
        named = {track_name(spk, side): ir for spk, side, ir in self._all()}
        if not named:
            raise ValueError("No impulse responses available for WAV output.")
        ref_len = len(next(iter(named.values())))
        present = [named[ch] for ch in track_order if ch in named]
        if any(len(ir) != ref_len for ir in present):
            raise ValueError("all the input array dimensions except for the concatenation axis must match exactly")
        flat, offs, _ = _pack(present) if present else (None, [], [])
        it = iter(offs)
        row_offsets = [next(it) if ch in named else -1 for ch in track_order]
        if flat is None:
            return np.zeros((ref_len, len(track_order)), dtype=np.int16 if bit_depth == 16 else np.int32)
        return ops.pack_tracks(flat, row_offsets, ref_len, bit_depth).cpu().numpy()

    def write_wav(self, file_path, track_order=None, bit_depth=32):
        """Reference :426-455 + core/audio_io.py:82-97.  Frames are packed on the device; the RIFF container is
        written on the host (soundfile when importable, else scipy.io.wavfile)."""
        import os
        frames = self.pcm_frames(track_order, bit_depth)
        d = os.path.dirname(file_path)
        if d:
            os.makedirs(d, exist_ok=True)
        if frames.shape[0] < frames.shape[1]:
            # the reference transposes only when tracks outnumber samples' worth of rows (shape[1] > shape[0]);
            # with fewer samples than tracks it would write the matrix untransposed: keep that quirk
            write_wav(file_path, self.fs, self.track_matrix(track_order), bit_depth=bit_depth)
            return
        try:
            import soundfile as sf
            sub = {16: "PCM_16", 24: "PCM_24", 32: "PCM_32"}[bit_depth]
            sf.write(file_path, frames if bit_depth != 24 else frames << 8, samplerate=self.fs, subtype=sub)
        except ImportError:
            from scipy.io import wavfile
            wavfile.write(file_path, int(self.fs), frames if bit_depth != 24 else frames << 8)

    # ---- normalisation ------------------------------------------------------------------------------------
    def normalize(self, peak_target=-0.1, avg_target=None):
        import torch
        left = [pair["left"] for pair in self.irs.values() if len(pair["left"]) > 0]
        right = [pair["right"] for pair in self.irs.values() if len(pair["right"]) > 0]
        if not left or not right:
            raise ValueError("No valid impulse response data found for normalization. All channels appear to be empty.")
        if (peak_target is None) == (avg_target is None):
            raise ValueError('One and only one of the parameters "peak_target" and "avg_target" must be given!')
        sums = []
        for group in (left, right):
            n = max(len(ir) for ir in group)
            acc = torch.zeros(n, dtype=torch.float64, device=group[0].device_data().device)
            for ir in group:  # np.sum(np.vstack(...), axis=0): sequential row adds
                d = ir.device_data()
                acc[:len(d)] += d
            sums.append(acc)
        if len(sums[0]) == len(sums[1]):
            mags = ops.magnitude_db(torch.stack(sums))
            mr_l, mr_r = mags[0], mags[1]
        else:
            mr_l, mr_r = ops.magnitude_db(sums[0])[0], ops.magnitude_db(sums[1])[0]
        if avg_target is None:
            gain = float(torch.maximum(mr_l.max(), mr_r.max()).item()) * -1 + peak_target
        else:
            parts = []
            for mr, acc in ((mr_l, sums[0]), (mr_r, sums[1])):
                f = np.arange(len(mr)) * (self.fs / len(acc))
                sel = torch.as_tensor(np.logical_and(f > 80, f < 6000), device=mr.device)
                parts.append(mr[sel])
            gain = float(torch.cat(parts).mean().item()) * -1 + avg_target
        gain_scalar = 10 ** (gain / 20)
        items = [ir for _, _, ir in self._all()]
        flat, offs, lens = _pack(items)
        for n in sorted(set(lens)):
            sel = [i for i, ln in enumerate(lens) if ln == n]
            if n == 0:
                continue
            out = ops.gather_window(flat, [offs[i] for i in sel], n, gain=gain_scalar)
            if gain_scalar == 1.0:
                out = out.clone()
            for k, i in enumerate(sel):
                items[i]._set_device(out[k])
        return gain

    # ---- crops ------------------------------------------------------------------------------------------------
    def crop_heads(self, head_ms=1):
        from scipy.signal.windows import hann
        if self.fs != self.estimator.fs:
            raise ValueError("Refusing to crop heads because HRIR sampling rate doesn't match impulse response "
                             "estimator's sampling rate.")
        speakers = list(self.irs.keys())
        if not speakers:
            return
        items = [self.irs[s][e] for s in speakers for e in ("left", "right")]
        flat, offs, lens = _pack(items)
        peaks = ops.peak_index(flat, offs, lens, max_len=max(lens)).cpu().numpy()
        head = int(head_ms * self.fs / 1000)
        window = hann(head * 2)[:head]
        jobs = {}  # (new_len, faded) -> [(item index, source start)]
        for k, speaker in enumerate(speakers):
            peak_left, peak_right = int(peaks[2 * k]), int(peaks[2 * k + 1])
            itd = np.abs(peak_left - peak_right) / self.fs
            delay = int(np.round(SPEAKER_DELAYS[speaker] * self.fs)) + head
            if peak_left < peak_right:
                if speaker_side(speaker) == "right":
                    warnings.warn(
                        f"Warning: {speaker} measurement has lower delay to left ear than to right ear. "
                        f"{speaker} should be at the right side of the head so the sound should arrive first "
                        "in the right ear. This is usually a problem with the measurement process or the "
                        "speaker order given is not correct. Detected delay difference is "
                        f"{itd * 1000:.4f} milliseconds.")
                crop_index = max(0, peak_left - delay)
            else:
                if speaker_side(speaker) == "left":
                    warnings.warn(
                        f"Warning: {speaker} measurement has lower delay to right ear than to left ear. "
                        f"{speaker} should be at the left side of the head so the sound should arrive first "
                        "in the left ear. This is usually a problem with the measurement process or the "
                        "speaker order given is not correct. Detected delay difference is "
                        f"{itd * 1000:.4f} milliseconds.")
                crop_index = max(0, peak_right - delay)
            new_l, new_r = max(lens[2 * k] - crop_index, 0), max(lens[2 * k + 1] - crop_index, 0)
            faded = new_l >= head and new_r >= head
            for i, n_new in ((2 * k, new_l), (2 * k + 1, new_r)):
                jobs.setdefault((n_new, faded), []).append((i, offs[i] + min(crop_index, lens[i])))
        for (n_new, faded), rows in jobs.items():
            if n_new == 0:
                for i, _ in rows:
                    items[i].data = np.zeros(0)
                continue
            out = ops.gather_window(flat, [r[1] for r in rows], n_new, fade_in=window if faded else None)
            if not faded:
                out = out.clone()
            for k, (i, _) in enumerate(rows):
                items[i]._set_device(out[k])

    def crop_tails(self):
        from scipy.signal.windows import hann
        if self.fs != self.estimator.fs:
            raise ValueError("Refusing to crop tails because HRIR sampling rate doesn't match estimator sampling rate.")
        items = [ir for _, _, ir in self._all()]
        if not items:
            return 0
        flat, offs, lens = _pack(items)
        _, knee, _, _ = ops.lundeby(flat, offs, lens, self.fs, max_len=max(lens))
        tail_indices = knee.cpu().numpy()
        seconds_per_octave = len(self.estimator) / self.estimator.fs / self.estimator.n_octaves
        fade_out = 2 * int(self.fs * seconds_per_octave * (1 / 24))
        window = hann(fade_out)[fade_out // 2:]
        tail_ind = int(min(np.min(lens), next_fast_len(int(max(tail_indices)))))
        if tail_ind < len(window):  # numpy would raise on np.ones(negative)
            raise ValueError("negative dimensions are not allowed")
        out = ops.gather_window(flat, offs, tail_ind, fade_out=window)
        for k, ir in enumerate(items):
            ir._set_device(out[k])
        return tail_ind

    # ---- alignment ------------------------------------------------------------------------------------------
    def _shift_many(self, shifts):
        """{ImpulseResponse: samples} -> length-preserving shifts, one launch per distinct length."""
        todo = [(ir, s) for ir, s in shifts if s != 0 and len(ir) > 0]
        if not todo:
            return
        flat, offs, lens = _pack([ir for ir, _ in todo])
        for n in sorted(set(lens)):
            sel = [i for i, ln in enumerate(lens) if ln == n]
            sh = [todo[i][1] for i in sel]
            lo = [max(s, 0) for s in sh]
            hi = [max(min(n + s, n), max(s, 0)) for s in sh]
            out = ops.gather_window(flat, [offs[i] for i in sel], n, shift=sh, valid_lo=lo, valid_hi=hi)
            for k, i in enumerate(sel):
                todo[i][0]._set_device(out[k])

    def align_ipsilateral_all(self, speaker_pairs=None, segment_ms=30):
        if speaker_pairs is None:
            speaker_pairs = list(IPSILATERAL_PAIRS)
        segment_len = int(self.fs * segment_ms / 1000)
        # the reference walks the pairs in order and shifts as it goes; pairs never share a speaker, so all
        # lags can be measured first in one launch and applied afterwards
        work = []
        for sp1, sp2 in speaker_pairs:
            if sp1 not in self.irs or sp2 not in self.irs:
                continue
            a = self.irs[sp1]["left"]
            b = self.irs[sp1]["right"] if sp1 == sp2 else self.irs[sp2]["right"]
            work.append((sp1, sp2, a, b))
        seen = [s for w in work for s in {w[0], w[1]}]
        if len(seen) != len(set(seen)):  # overlapping custom pairs: keep the sequential semantics
            for w in work:
                self._align_pairs([w], segment_len)
        else:
            self._align_pairs(work, segment_len)

    def _align_pairs(self, work, segment_len):
        if not work:
            return
        flat, offs, lens = _pack([x for w in work for x in (w[2], w[3])])
        k = len(work)
        seg = [min(n, segment_len) for n in lens]
        lags = ops.xcorr_lag(flat, offs[0::2], seg[0::2], offs[1::2], seg[1::2], max_seg_len=max(seg + [1])).cpu().numpy()
        shifts = []
        for (sp1, sp2, _, _), lag in zip(work, lags):
            lag = int(lag)
            if sp1 == sp2:
                if lag > 0:
                    shifts.append((self.irs[sp1]["right"], lag))
                elif lag < 0:
                    shifts.append((self.irs[sp1]["left"], -lag))
            elif lag > 0:
                shifts += [(self.irs[sp2][sd], lag) for sd in ("left", "right")]
            elif lag < 0:
                shifts += [(self.irs[sp1][sd], -lag) for sd in ("left", "right")]
        assert k == len(lags)
        self._shift_many(shifts)

    def align_onset_groups_peak_leftref(self, groups=None):
        if groups is None:
            groups = list(ONSET_GROUPS)
        ref_group = ("FL", "FR")
        leaders = [g for g in groups if g[0] in self.irs and "left" in self.irs[g[0]]]
        if "FL" not in self.irs or "left" not in self.irs["FL"]:
            raise RuntimeError("Cannot find FL left channel reference for onset alignment.")
        lead_irs = [self.irs["FL"]["left"]] + [self.irs[g[0]]["left"] for g in leaders]
        flat, offs, lens = _pack(lead_irs)
        peaks = ops.peak_index(flat, offs, lens, max_len=max(lens)).cpu().numpy()
        ref_peak = int(peaks[0])
        shifts = []
        for g, pk in zip(leaders, peaks[1:]):
            if g == ref_group:
                continue
            d = int(pk) - ref_peak
            for speaker in g:
                if speaker in self.irs:
                    shifts += [(self.irs[speaker][sd], -d) for sd in ("left", "right")]
        self._shift_many(shifts)

    # ---- EQ / resample ------------------------------------------------------------------------------------------
    def equalize(self, fir):
        """Reference :858-888.  Responses of equal length that take the same filter share one K1 launch."""
        if isinstance(fir, list):
            if isinstance(fir[0], np.ndarray):
                fir = np.vstack(fir)
            elif isinstance(fir[0], list):
                fir = np.array(fir)
            elif isinstance(fir[0], ImpulseResponse):
                fir = np.vstack([fir[0].data, fir[1].data]) if len(fir) > 1 else fir[0].data.copy()
        if len(fir.shape) == 1 or fir.shape[0] == 1:
            fir = np.tile(fir, (2, 1))
        self._equalize_rows([(ir, fir[0] if side == "left" else fir[1]) for _, side, ir in self._all()])

    @staticmethod
    def _equalize_rows(work):
        """``ir.equalize(fir)`` for every (ir, fir) of ``work``, batched by (filter, length)."""
        import torch
        from .dispatcher import fft_convolve_full
        groups = {}
        for ir, fir in work:
            fir = np.ascontiguousarray(fir, dtype=np.float64)
            groups.setdefault((fir.tobytes(), len(ir)), (fir, []))[1].append(ir)
        for fir, irs in groups.values():
            out = fft_convolve_full(torch.stack([ir.device_data() for ir in irs]), fir)
            for k, ir in enumerate(irs):
                ir._set_device(out[k])

    # ---- README metrics (next row) ---------------------------------------------------------
    def calculate_reflection_levels(self, direct_sound_duration_ms=2, early_ref_start_ms=20, early_ref_end_ms=50,
                                    late_ref_start_ms=50, late_ref_end_ms=150, epsilon=1e-12):
        """Early / late reflection levels relative to the direct sound for every response.  Reference
        ``core/hrir.py:1003-1090``; one K2 and one K18 launch for the whole set."""
        items = self._all()
        if not items:
            return {}
        flat, offs, lens = _pack([ir for _, _, ir in items])
        peaks = ops.peak_index(flat, offs, lens, max_len=max(lens)).cpu().numpy()
        n = np.asarray(lens, dtype=np.int64)

        def at(ms):
            return np.minimum(peaks + int(ms * self.fs / 1000), n)

        lo = np.stack([peaks, at(early_ref_start_ms), at(late_ref_start_ms)], axis=1)
        hi = np.stack([at(direct_sound_duration_ms), at(early_ref_end_ms), at(late_ref_end_ms)], axis=1)
        ms = ops.segment_energy(flat, offs, lo, np.maximum(hi, lo))
        out = {}
        for (speaker, side, _), (direct, early, late) in zip(items, ms):
            rms_direct = epsilon if np.isnan(direct) else np.sqrt(direct)
            rms_direct = rms_direct if rms_direct > epsilon else epsilon
            rms_early = 0 if np.isnan(early) else np.sqrt(early)
            rms_late = 0 if np.isnan(late) else np.sqrt(late)
            out.setdefault(speaker, {})[side] = {"early_db": 20 * np.log10(rms_early / rms_direct + epsilon),
                                                 "late_db": 20 * np.log10(rms_late / rms_direct + epsilon)}
        return out

    def decay_times(self):
        """``{speaker: {side: (EDT, RT20, RT30, RT60)}}`` for the whole set: one K4 + one K17 launch (the per-response
        loop of the README writer, reference ``core/pipeline_stages.py:590-625``)."""
        from .impulse_response import decay_times_batch
        items = self._all()
        out = {}
        for (speaker, side, _), rt in zip(items, decay_times_batch([ir for _, _, ir in items])):
            out.setdefault(speaker, {})[side] = rt
        return out

    # ---- channel balance (next row) -------------------------------------------------------
    def channel_balance_firs(self, left_fr, right_fr, method):
        """Left / right FIR pair that evens out the two ears.  Reference ``core/hrir.py:655-762``; every smoothing
        and minimum-phase design below runs on the device through the ``FrequencyResponse`` shim."""
        fs = self.fs
        nyq_int = int(round(fs / 2))

        def impulse(n):
            u = np.zeros(n)
            u[0] = 1.0
            return u

        def band_level(fr, band):  # reference get_center_value (:43-76) for a two-frequency band
            sel = np.logical_and(fr.frequency >= band[0], fr.frequency <= band[1])
            return -np.mean(fr.raw[sel])

        def fixed_gain(gain_db):
            n = int(round(fs * 0.1))
            return [impulse(n), impulse(n) * 10 ** (gain_db / 20)]

        if method == "mids":
            return fixed_gain(band_level(right_fr, [100, 3000]) - band_level(left_fr, [100, 3000]))
        if method == "trend":
            trend = FrequencyResponse(name="trend", frequency=left_fr.frequency, raw=left_fr.raw - right_fr.raw)
            trend.smoothen_fractional_octave(window_size=2, treble_f_lower=20000, treble_f_upper=nyq_int)
            right_fr.equalization = trend.smoothed
            fir = right_fr.minimum_phase_impulse_response(fs=fs, normalize=False)
            return [impulse(len(fir)), fir]
        if method in ("left", "right"):
            ref, subj = (left_fr, right_fr) if method == "left" else (right_fr, left_fr)
            ref.smoothen_fractional_octave(window_size=1 / 3, treble_f_lower=20000, treble_f_upper=nyq_int)
            subj.raw += ref.center([100, 10000])
            subj.target = ref.smoothed
            subj.error = subj.raw - subj.target
            subj.smoothen_heavy_light()
            subj.equalize(max_gain=15, treble_f_lower=20000, treble_f_upper=fs / 2)
            fir = subj.minimum_phase_impulse_response(fs=fs, normalize=False)
            return [impulse(len(fir)), fir] if method == "left" else [fir, impulse(len(fir))]
        if method in ("avg", "min"):
            gain = (band_level(left_fr, [100, 10000]) + band_level(right_fr, [100, 10000])) / 2
            for fr in (left_fr, right_fr):
                fr.raw += gain
                fr.smoothen_fractional_octave(window_size=1 / 3, treble_f_lower=20000, treble_f_upper=23999)
            goal = (left_fr.raw + right_fr.raw) / 2 if method == "avg" else np.min([left_fr.raw, right_fr.raw], axis=0)
            firs = []
            for fr in (left_fr, right_fr):
                fr.target = goal
                fr.error = fr.raw - fr.target
                fr.smoothen_fractional_octave(window_size=1 / 3, treble_f_lower=20000, treble_f_upper=23999)
                fr.equalize(max_gain=15, treble_f_lower=2000, treble_f_upper=fs / 2)
                firs.append(fr.minimum_phase_impulse_response(fs=fs, normalize=False))
            return firs
        try:
            return fixed_gain(float(method))
        except ValueError:
            raise ValueError(f'"{method}" is not valid value for channel balance method.')

    def correct_channel_balance(self, method):
        """Reference ``core/hrir.py:764-799``: per ipsilateral group, the mean left / mean right responses give the
        FIR pair which is then applied to every response of the group.  The means, the magnitude responses of all
        groups (one K8 launch) and the FIR application (one K1 launch per filter) run on the device."""
        import torch
        from .constants import IPSILATERAL_PAIRS
        from .room_correction import frequency_responses
        eqir = HRIR(self.estimator)
        groups = [[a] if a == b else [a, b] for a, b in IPSILATERAL_PAIRS]
        groups = [g for g in groups if all(ch in self.irs for ch in g)]
        means = []
        for g in groups:
            for side in ("left", "right"):
                rows = torch.stack([self.irs[ch][side].device_data() for ch in g])
                # np.mean(np.vstack(rows), axis=0): pairwise sum over at most two rows, then one division
                acc = rows[0].clone()
                for r in rows[1:]:
                    acc += r
                means.append(ImpulseResponse(acc / float(len(g)), self.fs))
        curves = frequency_responses(means)
        work = []
        for k, g in enumerate(groups):
            firs = self.channel_balance_firs(curves[2 * k], curves[2 * k + 1], method)
            for ch in g:
                work += [(self.irs[ch]["left"], firs[0]), (self.irs[ch]["right"], firs[1])]
        self._equalize_rows(work)
        return eqir

    def resample(self, fs):
        items = [ir for _, _, ir in self._all()]
        if items:
            up, down, taps = resample_taps(int(fs), int(self.fs))
            if up != down:
                flat, offs, lens = _pack(items)
                for n in sorted(set(lens)):
                    sel = [i for i, ln in enumerate(lens) if ln == n]
                    x = ops.gather_window(flat, [offs[i] for i in sel], n).clone() if n else None
                    if x is None:
                        continue
                    out = ops.window_resample(x, taps, up, down)
                    for k, i in enumerate(sel):
                        items[i]._set_device(out[k])
            for ir in items:
                ir.fs = fs
        self.fs = fs

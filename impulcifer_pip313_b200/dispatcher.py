"""GPU batch dispatcher — the replacement for the reference's process / thread pools
(``core/parallel_utils.py:97-152``, ``core/parallel_processing.py:84-190``, ``core/parallel_workers.py:24-131``).

The reference maps a Python worker over (speaker, side) items.  Here every item of MANY recording sets is a
row of one launch:

  ``BatchDispatcher.process``       deconvolution -> crop_heads -> align_ipsilateral_all ->
                                    align_onset_groups_peak_leftref -> crop_tails  for S recording sets at once
  ``BatchDispatcher.postprocess``   frequency_response -> smoothen_fractional_octave -> minimum-phase FIR
                                    (the arithmetic of ``process_equalization_worker``) for every row
  ``DeviceShard``                   which recording sets a rank owns (sets never split across GPUs) and the
                                    final result gather; no collective on the compute path.

Integer decisions (crop indices, lags, shifts) are taken with torch integer ops on the device, so a batch
needs ONE host synchronisation: reading the knee indices that fix each set's output length.
"""
from __future__ import annotations

import os
import threading
import time
import warnings

import numpy as np

from . import ops
from .constants import IPSILATERAL_PAIRS, ONSET_GROUPS, SPEAKER_DELAYS, speaker_side
from .frequency_response import equalize_batch, minimum_phase_batch, next_fast_len, smoothen_batch
from .hrir import ingest_layout
from .impulse_response import frequency_response_batch

_conv_lock = threading.Lock()
_conv_plans = {}


def fft_convolve_full(x, fir, out_len=None):
    """``scipy.signal.convolve(x_row, fir, 'full')`` for every row of ``x`` (CUDA [B, n]) -> CUDA [B, n+m-1]
    (``ImpulseResponse.equalize``, reference core/impulse_response.py:110-119; next-row component).
    ``out_len``: keep only the first ``out_len`` samples of every row (a causal filter applied in place)."""
    import torch
    from .estimator import DeconvPlan
    fir = np.ascontiguousarray(fir, dtype=np.float64)
    n, m = x.shape[1], len(fir)
    total = n + m - 1
    s = (m - 1) // 2
    key = (fir.tobytes(), total, x.device.index)
    with _conv_lock:
        plan = _conv_plans.get(key)
        if plan is None:
            if len(_conv_plans) > 64:
                _conv_plans.clear()
            plan = _conv_plans[key] = DeconvPlan(fir, total, 64, x.device.index or 0)
    padded = torch.zeros((x.shape[0], total), dtype=torch.float64, device=x.device)
    padded[:, s:s + n] = x
    return plan.execute(padded, out_len=out_len)


class SetResult:
    """Cropped responses of one recording set, in the reference's estimation order.

    ``data`` is the CUDA tensor ``[n_irs, tail_ind]``.  The dispatcher hands over the set's slice of one flat output
    buffer and the 2-D view is made on first access (64 sets x a few microseconds of tensor bookkeeping stay off the
    critical path between two steps)."""
    __slots__ = ("names", "tail_ind", "peaks", "crop", "knees", "warnings", "_data", "_flat", "_off")

    def __init__(self, names, data, tail_ind, peaks, crop, knees, warnings=None, flat=None, offset=0):
        self.names = names            # ["FL-left", "FL-right", ...]
        self.tail_ind = tail_ind
        self.peaks = peaks            # raw-estimate peak index per IR
        self.crop = crop              # head crop index per IR
        self.knees = knees            # Lundeby knee index per IR (after alignment)
        self.warnings = [] if warnings is None else warnings
        self._data = data
        self._flat = flat             # the dispatcher's flat output buffer; this set starts at element `offset`
        self._off = offset

    @property
    def data(self):
        if self._data is None:
            n = len(self.names)
            self._data = self._flat.narrow(0, self._off, n * self.tail_ind).view(n, self.tail_ind)
        return self._data

    @data.setter
    def data(self, value):
        self._data = value
        self._flat = None

    # pickling (DeviceShard.gather): ship this set's own samples as a host tensor, never the shared flat buffer
    def __getstate__(self):
        data = self.data
        if hasattr(data, "detach"):
            data = data.detach().cpu().contiguous().clone()
        return {"names": self.names, "tail_ind": self.tail_ind, "peaks": self.peaks, "crop": self.crop,
                "knees": self.knees, "warnings": self.warnings, "data": data}

    def __setstate__(self, state):
        self.names, self.tail_ind, self.peaks = state["names"], state["tail_ind"], state["peaks"]
        self.crop, self.knees, self.warnings = state["crop"], state["knees"], state["warnings"]
        self._data, self._flat, self._off = state["data"], None, 0

    def as_dict(self):
        out = {}
        host = self.data.cpu().numpy()
        for name, row in zip(self.names, host):
            spk, side = name.split('-')
            out.setdefault(spk, {})[side] = row
        return out


class BatchDispatcher:
    def __init__(self, estimator, speakers, side=None, silence_length=2.0, head_ms=1, segment_ms=30):
        self.estimator = estimator
        self.speakers = list(speakers)
        self.side = side
        self.silence_length = silence_length
        self.head_ms = head_ms
        self.segment_ms = segment_ms
        self.fs = estimator.fs
        self._layout = None
        self.launches = 0
        self.phase_timing = False   # developer hook: record device time per phase of process() into phase_ms
        self._read_state = {}       # device index -> (side stream, pinned int64 buffer) of the step's host read
        self.phase_ms = []

    # ---- static layout of one set ---------------------------------------------------------------------
    def _prepare(self, n_tracks, n_samples, device):
        import torch
        key = (n_tracks, n_samples, str(device))
        if self._layout is not None and self._layout['key'] == key:
            return self._layout
        fs = self.fs
        columns, plan = ingest_layout(n_tracks, n_samples, self.speakers, len(self.estimator), fs, self.side,
                                      self.silence_length)
        lengths = {columns[c][1] - columns[c][0] for _, _, _, c in plan}
        if len(lengths) != 1:
            raise ValueError("BatchDispatcher needs full-length columns; use HRIR.open_recording for clipped files")
        L = lengths.pop()
        names = [f"{spk}-{sd}" for spk, sd, _, _ in plan]
        row_of = {nm: i for i, nm in enumerate(names)}
        offs = np.array([trk * n_samples + columns[c][0] for _, _, trk, c in plan], dtype=np.int64)
        spk_list = []
        for spk, _, _, _ in plan:
            if spk not in spk_list:
                spk_list.append(spk)
        paired = self.side is None
        lay = dict(key=key, L=L, names=names, row_of=row_of, n=len(plan), speakers=spk_list, paired=paired,
                   offs=torch.as_tensor(offs, device=device))
        if paired:
            left = np.array([row_of[f"{s}-left"] for s in spk_list])
            right = np.array([row_of[f"{s}-right"] for s in spk_list])
            lay['left'] = torch.as_tensor(left, device=device)
            lay['right'] = torch.as_tensor(right, device=device)
            lay['left_h'], lay['right_h'] = left.tolist(), right.tolist()  # host copies: no device round trip later
            head = int(self.head_ms * fs / 1000)
            lay['delay'] = torch.as_tensor([int(np.round(SPEAKER_DELAYS[s] * fs)) + head for s in spk_list],
                                           device=device)
            lay['side_code'] = [speaker_side(s) for s in spk_list]
            # ipsilateral pairs present in this layout, in the reference's order
            pairs = [(a, b) for a, b in IPSILATERAL_PAIRS if a in spk_list and b in spk_list]
            lay['pairs'] = pairs
            lay['pa'] = torch.as_tensor([row_of[f"{a}-left"] for a, _ in pairs], device=device)
            lay['pb'] = torch.as_tensor([row_of[f"{b}-right"] for _, b in pairs], device=device)
            groups = [g for g in ONSET_GROUPS if g[0] in spk_list]
            lay['groups'] = groups
            lay['gl'] = torch.as_tensor([row_of[f"{g[0]}-left"] for g in groups], device=device)
            # which rows a positive / negative ipsilateral lag of pair k shifts (ref core/hrir.py:936-955)
            pos_rows, pos_k, neg_rows, neg_k = [], [], [], []
            for k, (a, b) in enumerate(pairs):
                if a == b:
                    pos_rows.append(row_of[f"{a}-right"]); pos_k.append(k)
                    neg_rows.append(row_of[f"{a}-left"]); neg_k.append(k)
                else:
                    for sd in ('left', 'right'):
                        pos_rows.append(row_of[f"{b}-{sd}"]); pos_k.append(k)
                        neg_rows.append(row_of[f"{a}-{sd}"]); neg_k.append(k)
            for name, v in (('pos_rows', pos_rows), ('pos_k', pos_k), ('neg_rows', neg_rows), ('neg_k', neg_k)):
                lay[name] = torch.as_tensor(v, device=device, dtype=torch.int64)
            # rows moved by the onset alignment of group g (every group but FL/FR) (ref core/hrir.py:981-1001)
            g_rows, g_of_row = [], []
            for gi, g in enumerate(groups):
                if g == ("FL", "FR"):
                    continue
                for spk in g:
                    if spk in spk_list:
                        for sd in ('left', 'right'):
                            g_rows.append(row_of[f"{spk}-{sd}"]); g_of_row.append(gi)
            lay['g_rows'] = torch.as_tensor(g_rows, device=device, dtype=torch.int64)
            lay['g_of_row'] = torch.as_tensor(g_of_row, device=device, dtype=torch.int64)
            lay['g_ref'] = [i for i, g in enumerate(groups) if g[0] == 'FL'][0] if 'FL' in spk_list else -1
            lay['set_rows'] = torch.arange(len(plan), device=device, dtype=torch.int64)
            # the same tables for the device-side bookkeeping kernels (K22, csrc/setplan.cu)
            lay['dev'] = ops.SetLayoutDev(
                device, len(plan), left, right, [int(np.round(SPEAKER_DELAYS[s] * fs)) + head for s in spk_list],
                [row_of[f"{a}-left"] for a, _ in pairs], [row_of[f"{b}-right"] for _, b in pairs], pos_rows, pos_k,
                neg_rows, neg_k, [row_of[f"{g[0]}-left"] for g in groups], lay['g_ref'], g_rows, g_of_row)
        self._layout = lay
        return lay

    # ---- the batched crop / align pipeline ------------------------------------------------------------------
    def process(self, recordings, precision=None):
        """recordings: CUDA tensor [S, n_tracks, n_samples] (float64, float32 or int32 PCM).
        Returns a list of ``SetResult`` (one per recording set)."""
        import torch
        from scipy.signal.windows import hann
        if recordings.dim() == 2:
            recordings = recordings[None]
        S, n_tracks, n_samples = recordings.shape
        dev = recordings.device
        lay = self._prepare(n_tracks, n_samples, dev)
        if not lay['paired']:
            raise ValueError("BatchDispatcher.process handles binaural (two tracks per speaker) sets")
        est, fs, L, n = self.estimator, self.fs, lay['L'], lay['n']
        B = S * n
        head = int(self.head_ms * fs / 1000)
        launches = 0
        marks = [] if self.phase_timing else None

        def mark(name):  # developer hook (tools/profile_dispatch.py): device-time stamps between the phases
            if marks is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append((name, ev, time.perf_counter()))

        with torch.cuda.device(dev):
            mark('start')
            set_base = torch.arange(S, device=dev, dtype=torch.int64)[:, None]
            # 1. deconvolution of every column of every set (K1)
            row_offs = (set_base * (n_tracks * n_samples) + lay['offs'][None, :]).reshape(-1)
            plan = est.plan(L, precision)
            if plan.precision == 32 and recordings.dtype == torch.float64:
                # the FP32 path rounds every sample to float on load anyway: do it once up front so that the column
                # pass can stage raw float tiles with cp.async (12.1 -> 8.0 ms of K1 per 1792 rows for a 1.3 ms cast)
                recordings = recordings.to(torch.float32)
            e_max = torch.empty(B, dtype=torch.float64, device=dev)
            e_arg = torch.empty(B, dtype=torch.int64, device=dev)
            nblk, blk_len, blk_first = plan.blockmax_info()
            e_blk = torch.empty((B, nblk), dtype=torch.float64, device=dev) if nblk else None
            # the FP32 fast path widens on store (exact), so the crop kernels always see float64 rows
            E = torch.empty((B, L), dtype=torch.float64, device=dev)
            plan.execute(recordings.reshape(-1), out=E, length=L, row_offsets=row_offs, absmax=e_max, argmax=e_arg,
                         blockmax=e_blk)
            launches += plan.launches(B) + 1  # + argmax pick
            mark('K1 deconvolution')
            flatE = E.reshape(-1)
            rows = torch.arange(B, device=dev, dtype=torch.int64)
            full = torch.full((B,), L, device=dev, dtype=torch.int64)
            # 2. first peaks (K2) and the pairwise head crop (ref core/hrir.py:573-606), integer ops on device
            peaks = ops.peak_index(flatE, rows * L, full, max_len=L, known_max=e_max,
                                   view=None if e_blk is None else dict(blocks=(e_blk, blk_len, blk_first)))
            launches += 3
            mark('first peaks')
            # From here on nothing is materialised and nothing is read back: the cropped, faded and shifted responses
            # are VIEWS over E (impb_view: start offset, composite shift, surviving index range, fade-in table) and
            # the integer bookkeeping between the kernels runs on the device (K22, csrc/setplan.cu).
            ld = lay['dev']
            seg = int(fs * self.segment_ms / 1000)
            window_in = hann(head * 2)[:head] if head > 0 else None
            if 'FL' not in lay['speakers']:
                raise RuntimeError("Cannot find FL left channel reference for onset alignment.")
            # pairwise head crop (ref core/hrir.py:573-606) + the segments of the ipsilateral pairs
            crop, n1, off1, hint_arg, xs = ops.set_heads(ld, peaks, e_arg, L, seg, S)
            launches += 1
            # 3. ipsilateral onset alignment (ref :921-955): lag of the first 30 ms, K6
            lags = None
            if xs is not None:
                lags = ops.xcorr_lag(flatE, xs[0], xs[1], xs[2], xs[3], max_seg_len=seg, fade_in=window_in)
                launches += 1
            with_groups = len(lay['groups']) > 1
            shift, lo, hi, gi, gm = ops.set_shift(ld, lags, n1, off1, hint_arg, e_max, S, with_groups)
            launches += 1
            mark('head crop + ipsilateral lag')
            # 4. group onset alignment against FL-left (ref :957-1001): K2 on the shifted views of the leaders
            gp = None
            if with_groups:
                gp = ops.peak_index(flatE, gi[0], gi[1], max_len=L,
                                    view=dict(shift=gi[2], valid_lo=gi[3], valid_hi=gi[4], fade_in=window_in, hint_max=gm,
                                              hint_argmax=gi[5]))
                launches += 4
            ops.set_shift2(ld, gp, n1, S, shift, lo, hi)  # shift is now the composite shift, [lo, hi) what survives
            launches += 1
            mark('group onset peaks')
            view = dict(shift=shift, valid_lo=lo, valid_hi=hi, fade_in=window_in, hint_max=e_max, hint_argmax=hint_arg)
            # 5. Lundeby knee per response (K4) -> per-set tail index and ragged output geometry (ref :614-653), on
            #    the device: the final gather is launched without the host having seen a single integer
            _, knee, _, _ = ops.lundeby(flatE, off1, n1, fs, max_len=L, view=view)
            launches += 7  # peak_index (init, max, first peak, final) + window max + block sums + Lundeby
            mark('Lundeby')
            seconds_per_octave = len(est) / est.fs / est.n_octaves
            fade_out = 2 * int(fs * seconds_per_octave * (1 / 24))
            window_out = hann(fade_out)[fade_out // 2:]
            # knee <= first peak of the cropped view + 2 s; the views start within head + delay + |shifts| of it
            tail_bound = min(int(L), next_fast_len(int(2 * fs) + 8192))
            tails_d, set_off_d, out_len_d, dst_off_d, flags_d = ops.set_tails(ld, knee, n1, S, head, len(window_out),
                                                                              tail_bound, L)
            # ---- the ONE host read of the step: integers the results report.  They are final before the gather is
            # launched, so they travel on a side stream while the gather (the step's last and longest crop kernel) runs:
            # the host's bookkeeping and the caller's next launches no longer wait for it.
            ints_d = torch.cat([knee, n1, peaks, crop, tails_d, set_off_d, flags_d])
            ready = torch.cuda.Event()
            ready.record()
            flat = ops.gather_rows(flatE, off1, out_len_d, dst_off_d, S * n * tail_bound, tail_bound, shift=shift,
                                   valid_lo=lo, valid_hi=hi, fade_in=window_in, fade_out=window_out)
            launches += 2
            mark('tail gather launch')
            ints = self._read_ints(ints_d, ready)
            mark('host sync')
            B4 = 4 * B
            knees_h, n1_h, peaks_h, crop_h = (ints[i * B:(i + 1) * B].reshape(S, n) for i in range(4))
            tails_np, set_off, flags = ints[B4:B4 + S], ints[B4 + S:B4 + 2 * S + 1], int(ints[-1])
            if flags & 1:
                raise ValueError("an impulse response is shorter than the fade-in; use the HRIR class for such files")
            if flags & 2:
                raise ValueError("negative dimensions are not allowed")
            if flags & 4:
                # a tail above the preallocated bound (onsets more than 8192 samples into the cropped views): redo the
                # geometry exactly on the host and gather again, as round 1 did for every batch
                tails_np = np.minimum(n1_h.min(axis=1), next_fast_len(knees_h.max(axis=1))).astype(np.int64)
                if tails_np.min() < len(window_out):
                    raise ValueError("negative dimensions are not allowed")
                set_off = np.concatenate([[0], np.cumsum(tails_np * n)])
                geom = np.stack([np.repeat(tails_np, n),
                                 (set_off[:-1, None] + np.arange(n)[None, :] * tails_np[:, None]).reshape(-1)])
                geom_d = torch.from_numpy(geom).to(dev)
                flat = ops.gather_rows(flatE, off1, geom_d[0], geom_d[1], int(set_off[-1]), int(tails_np.max()),
                                       shift=shift, valid_lo=lo, valid_hi=hi, fade_in=window_in, fade_out=window_out)
                launches += 1
            tails = tails_np.tolist()
            # host-only bookkeeping after the last launch: it overlaps the gather
            side_msgs = self._side_warnings(lay, peaks_h)
            offs_l = set_off.tolist()
            names = lay['names']
            results = [SetResult(names, None, tails[s], peaks_h[s], crop_h[s], knees_h[s], side_msgs[s], flat, offs_l[s])
                       for s in range(S)]
            mark('tail gather')
        if marks is not None:
            torch.cuda.synchronize(dev)
            self.phase_ms = [(b[0], a[1].elapsed_time(b[1])) for a, b in zip(marks[:-1], marks[1:])]
            self.phase_host_ms = [(b[0], (b[2] - a[2]) * 1e3) for a, b in zip(marks[:-1], marks[1:])]
        self.launches = launches
        return results

    def _read_ints(self, ints_d, ready):
        """Device int64 vector -> numpy, through a pinned buffer on a side stream that only waits for ``ready`` (not for
        what the calling stream was given afterwards)."""
        import torch
        if os.environ.get("IMPULCIFER_B200_SYNC_READ"):  # A/B: the read waits for everything launched so far
            return ints_d.cpu().numpy()
        dev = ints_d.device
        st = self._read_state.get(dev.index)
        if st is None or st[1].numel() < ints_d.numel():
            st = self._read_state[dev.index] = (st[0] if st else torch.cuda.Stream(dev),
                                                torch.empty(max(ints_d.numel(), 1 << 14), dtype=torch.int64).pin_memory())
        side, pinned = st
        side.wait_event(ready)
        with torch.cuda.stream(side):
            host = pinned[:ints_d.numel()]
            host.copy_(ints_d, non_blocking=True)
        ints_d.record_stream(side)
        side.synchronize()
        return host.numpy().copy()

    # ---- host buffers in, host buffers out ------------------------------------------------------------------
    def process_host(self, recordings, out=None, chunk_sets=None, precision=None, out_sets=None):
        """``process`` for recordings that live in (pinned) HOST memory: ``[S, n_tracks, n_samples]`` float64,
        float32 or int32 PCM (the samples of a PCM WAV exactly as stored; scaled by 2^-31 inside K1).

        The sets are cut into chunks; the H2D copy of chunk k+1 (copy stream), the kernels of chunk k (compute
        stream) and the D2H copy of the cropped responses of chunk k-1 (third stream) overlap.  Returns the list of
        ``SetResult`` with ``data`` in host memory (views of ``out`` if a pinned float64 buffer is passed)."""
        import torch
        S = recordings.shape[0]
        if chunk_sets is None:
            chunk_sets = int(__import__('os').environ.get("IMPULCIFER_B200_CHUNK_SETS", "8"))
        dev = torch.device("cuda", self.estimator.device)
        key = (tuple(recordings.shape), recordings.dtype, chunk_sets)
        st = getattr(self, "_host_state", None)
        if st is None or st['key'] != key:
            st = dict(key=key, h2d=torch.cuda.Stream(dev), d2h=torch.cuda.Stream(dev),
                      stage=torch.empty(recordings.shape, dtype=recordings.dtype, device=dev))
            self._host_state = st
        stage, h2d, d2h = st['stage'], st['h2d'], st['d2h']
        compute = torch.cuda.current_stream(dev)
        bounds = [(a, min(a + chunk_sets, S)) for a in range(0, S, chunk_sets)]
        h2d.wait_stream(compute)  # the staging buffer may still be read by the previous call
        ready = []
        with torch.cuda.stream(h2d):
            for a, b in bounds:
                src = recordings[a:b]
                if src.is_contiguous():
                    stage[a:b].copy_(src, non_blocking=True)
                else:  # sets of a strided shard view (DeviceShard: every G-th set of the shared arena): one DMA per
                    for i in range(a, b):  # set, straight from the page-locked arena (a strided copy_ would first
                        stage[i].copy_(recordings[i], non_blocking=True)  # gather on the host)
                ev = torch.cuda.Event()
                ev.record(h2d)
                ready.append(ev)
        results, off, launches = [], 0, 0
        done = []
        for (a, b), ev in zip(bounds, ready):
            compute.wait_event(ev)
            part = self.process(stage[a:b], precision)
            launches += self.launches
            fin = torch.cuda.Event()
            fin.record(compute)
            d2h.wait_event(fin)
            with torch.cuda.stream(d2h):
                for j, r in enumerate(part):
                    n = r.data.numel()
                    if out_sets is not None:  # fixed slot per set (ResultArena.data rows, page-locked)
                        if n > out_sets.shape[1]:
                            raise ValueError(f"result of {n} samples does not fit the arena slot of {out_sets.shape[1]}")
                        host = out_sets[len(results) + j, :n].view(r.data.shape)
                    elif out is not None:
                        host = out[off:off + n].view(r.data.shape)
                    else:
                        host = torch.empty(r.data.shape, dtype=r.data.dtype).pin_memory()
                    host.copy_(r.data, non_blocking=True)
                    r.data.record_stream(d2h)
                    r.data = host
                    off += n
            results += part
        d2h.synchronize()
        self.launches = launches
        self.host_bytes_out = off * 8
        return results

    def _side_warnings(self, lay, peaks):
        """Wrong-side ITD warnings of crop_heads (ref core/hrir.py:578-606) for all sets: [S][...] messages."""
        li, ri = lay['left_h'], lay['right_h']
        pl, pr = peaks[:, li], peaks[:, ri]
        is_right = np.array([c == 'right' for c in lay['side_code']])
        is_left = np.array([c == 'left' for c in lay['side_code']])
        bad = ((pl < pr) & is_right[None, :]) | ((pl >= pr) & is_left[None, :])
        out = [[] for _ in range(peaks.shape[0])]
        for s, k in zip(*np.nonzero(bad)):
            spk = lay['speakers'][k]
            first, second = ('left', 'right') if pl[s, k] < pr[s, k] else ('right', 'left')
            msg = f"Warning: {spk} measurement has lower delay to {first} ear than to {second} ear."
            warnings.warn(msg)
            out[s].append(msg)
        return out

    # ---- per-IR frequency-response post-processing -----------------------------------------------------------
    def postprocess(self, results, window_size=1 / 3, f_res=5, normalize=False, worker=False):
        """frequency_response -> smoothen_fractional_octave(window_size) -> minimum-phase FIR of the inverse
        smoothed response, for every row of every set (sets of equal tail share launches).
        Returns per set ``(frequency[F], raw[n,F], smoothed[n,F], fir[n,taps])`` as CUDA tensors."""
        import torch
        from .frequency_response import FrequencyResponse
        grid = FrequencyResponse.generate_frequencies(f_step=1.01, f_min=10, f_max=self.fs / 2)
        out = [None] * len(results)
        by_tail = {}
        for i, r in enumerate(results):
            by_tail.setdefault(r.tail_ind, []).append(i)
        for tail, idxs in by_tail.items():
            flat = results[idxs[0]]._flat
            if flat is not None and all(results[i]._flat is flat for i in idxs):
                # the sets of one process() call are slices of one flat buffer: hand K8 the row offsets, no gather copy
                n_rows = len(results[idxs[0]].names)
                offs = np.concatenate([results[i]._off + np.arange(n_rows, dtype=np.int64) * tail for i in idxs])
                raw = frequency_response_batch(flat, self.fs, row_offsets=ops._dev_index_const(offs, flat.device), n=tail)
            else:
                data = torch.cat([results[i].data for i in idxs]) if len(idxs) > 1 else results[idxs[0]].data
                raw = frequency_response_batch(data, self.fs)
            if worker:  # the whole EQ-worker arithmetic, equalize() included (core/parallel_workers.py:98-131)
                sm, _eq, fir = equalization_worker_batch(grid, raw, self.fs, f_res=f_res)
            else:
                sm = smoothen_batch(grid, raw, window_size, 1, window_size, 1, 100.0, 10000.0)
                eq = -(sm - sm.mean(dim=1, keepdim=True))
                fir = minimum_phase_batch(grid, eq, fs=self.fs, f_res=f_res, normalize=normalize)
            n = results[idxs[0]].data.shape[0]
            for k, i in enumerate(idxs):
                sl = slice(k * n, (k + 1) * n)
                out[i] = (grid, raw[sl], sm[sl], fir[sl])
        return out


def equalization_worker_batch(frequency, raw, fs, max_gain=40, f_res=5):
    """The arithmetic of ``process_equalization_worker`` (reference core/parallel_workers.py:98-131) for rows of raw
    responses already on the device: error = raw - mean(raw) (flat target), ``smoothen_heavy_light`` (four
    Savitzky-Golay blends), ``equalize(max_gain, treble 10 kHz .. fs/2)`` (K21) and the ``f_res``-Hz minimum-phase
    FIR.  CUDA tensors in and out: ``(error_smoothed, equalization, fir)``."""
    import torch
    err = raw - raw.mean(dim=1, keepdim=True)
    light = smoothen_batch(frequency, err, 1 / 6, 1, 1 / 3, 1, 100, 10000)
    heavy = smoothen_batch(frequency, err, 1 / 3, 1, 1.3, 1, 1000, 6000)
    err_s = smoothen_batch(frequency, torch.maximum(light, heavy), 1 / 3, 1, 1 / 3, 1, 100, 10000)
    eq = equalize_batch(frequency, err_s, max_gain=max_gain, treble_f_lower=10000, treble_f_upper=fs / 2)
    fir = minimum_phase_batch(frequency, eq, fs=fs, f_res=f_res, normalize=False)
    return err_s, eq, fir


def equalization_firs(items, room_frs, hp_left, hp_right, eq_left, eq_right, target, common_freq, estimator_fs,
                      max_gain=40, f_res=5):
    """Batched ``process_equalization_worker`` (reference ``core/parallel_workers.py:69-131``) — the work the
    reference maps over a ProcessPoolExecutor at ``core/pipeline.py:674-688``.

    ``items``: [(speaker, side), ...].  Returns ``[(speaker, side, fir), ...]`` in input order, like
    ``parallel_map`` does.  The error curves are summed on the host exactly as the worker does; the four
    Savitzky-Golay blends of ``smoothen_heavy_light`` and the minimum-phase design run as ONE launch each over all
    items, and so does ``FrequencyResponse.equalize`` (K21: gain clip, kink windows, k=2 spline): no per-item loop."""
    from .frequency_response import FrequencyResponse
    items = list(items)
    if not items:
        return []
    freq = np.asarray(common_freq, dtype=np.float64)
    frs = []
    for speaker, side in items:
        fr = FrequencyResponse(name=f'{speaker}-{side} eq', frequency=freq.copy(), raw=0, error=0)
        if room_frs is not None and speaker in room_frs and side in room_frs[speaker]:
            fr.error += room_frs[speaker][side].error
        hp_eq = hp_left if side == 'left' else hp_right
        if hp_eq is not None:
            fr.error += hp_eq.error
        eq = eq_left if side == 'left' else eq_right
        if eq is not None and hasattr(eq, 'error'):
            fr.error += eq.error
        fr.error -= target.raw
        frs.append(fr)
    err = np.stack([fr.error for fr in frs])
    raw = np.stack([fr.raw for fr in frs])
    if np.any(np.isnan(err)) or np.any(np.isnan(raw)):
        raise ValueError('NaN values present, cannot smoothen!')
    # smoothen_heavy_light for every item at once (reference autoeq/frequency_response.py:1181-1239), device resident
    # from here to the FIR taps: one upload of the summed error rows, one download of the filters
    import torch
    from .frequency_response import _to_dev
    err_d, raw_d = _to_dev(err), _to_dev(raw)
    light = smoothen_batch(freq, err_d, 1 / 6, 1, 1 / 3, 1, 100, 10000)
    heavy = smoothen_batch(freq, err_d, 1 / 3, 1, 1.3, 1, 1000, 6000)
    both = smoothen_batch(freq, torch.cat([raw_d, torch.maximum(light, heavy)]), 1 / 3, 1, 1 / 3, 1, 100, 10000)
    n = len(frs)
    # FrequencyResponse.equalize for every item in ONE launch (K21), straight from the smoothed error rows
    eq, status = equalize_batch(freq, both[n:], max_gain=max_gain, treble_f_lower=10000, treble_f_upper=estimator_fs / 2,
                                return_status=True)
    firs = minimum_phase_batch(freq, eq, fs=estimator_fs, f_res=f_res, normalize=False).cpu().numpy()
    if int(status.max().item()) != 0:
        raise ValueError("equalize: too few points left for the k=2 interpolating spline (m > k must hold)")
    return [(spk, side, firs[i]) for i, (spk, side) in enumerate(items)]


def decay_adjust(tasks):
    """Batched ``process_decay_worker`` (reference ``core/parallel_workers.py:24-39``) — the work the reference
    maps over a process pool at ``core/pipeline.py:694-716``.

    ``tasks``: ``[(speaker, side, data, fs, target_seconds), ...]``.  Returns ``[(speaker, side, adjusted), ...]`` in
    input order; ``adjusted`` is a new float64 array (the worker copies before it multiplies).  One K4 launch
    (Lundeby parameters) and one K17 launch (Schroeder decay times) cover every response of a sampling rate, the
    four-number window parameters are host scalars exactly as ``decay_adjustment_params`` derives them
    (core/decay.py:355-380), and K15 applies the windows of all equal-length rows in one launch."""
    import torch
    from . import ops
    from .impulse_response import _device_index
    tasks = list(tasks)
    if not tasks:
        return []
    dev = torch.device("cuda", _device_index())
    out = [None] * len(tasks)
    by_fs = {}
    for i, t in enumerate(tasks):
        by_fs.setdefault(int(t[3]), []).append(i)
    for fs, idxs in by_fs.items():
        rows = [np.ascontiguousarray(tasks[i][2], dtype=np.float64) for i in idxs]
        lens = [len(r) for r in rows]
        for i, n in zip(idxs, lens):
            if n < 10:  # decay_params' degenerate branch has no decay time -> the reference raises TypeError
                raise TypeError("'>' not supported between instances of 'float' and 'NoneType'")
        flat = torch.from_numpy(np.concatenate(rows)).to(dev)
        offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
        pk, kn, nf, wn = (t.cpu().numpy() for t in ops.lundeby(flat, offs, lens, fs, max_len=max(lens)))
        rts = ops.decay_times(flat, offs, lens, pk, kn, nf, wn, fs, max_len=max(lens))
        params = np.full((len(idxs), 4), np.nan)
        for r, i in enumerate(idxs):
            target = tasks[i][4]
            slope = None
            for rt, level in zip(rts[r], (-10, -20, -30, -60)):
                if np.isnan(rt) or not rt:
                    break
                slope = level / float(rt)
            target_slope = -60 / target
            if target_slope > slope:  # TypeError when nothing could be measured, as in the reference
                continue
            knee_time = int(kn[r]) / fs
            start = int(pk[r]) + 2 * (fs // 1000)
            params[r] = (start, int(kn[r]) - start, int(kn[r]), target_slope * knee_time - slope * knee_time)
        by_len = {}
        for r, n in enumerate(lens):
            by_len.setdefault(n, []).append(r)
        for n, rs in by_len.items():
            x = torch.stack([flat[int(offs[r]):int(offs[r]) + n] for r in rs])
            y = ops.window_resample(x, None, 1, 1, decay=params[rs]).cpu().numpy()
            for k, r in enumerate(rs):
                i = idxs[r]
                out[i] = (tasks[i][0], tasks[i][1], y[k])
    return out


class DeviceShard:
    """Partition of recording sets over ranks (one process per GPU).  Sets are independent and never split
    (crop_heads / align / crop_tails / normalize couple all responses of a set), so the
    compute path has no collective; ``gather`` is the final result gather."""

    def __init__(self, world_size: int = 1, rank: int = 0):
        if not (0 <= rank < world_size):
            raise ValueError("rank out of range")
        self.world_size, self.rank = int(world_size), int(rank)

    def owned(self, n_sets: int):
        """Indices of the sets this rank processes (round-robin: set s -> rank s mod world_size)."""
        return list(range(self.rank, n_sets, self.world_size))

    def owner(self, set_index: int) -> int:
        return set_index % self.world_size

    def count(self, n_sets: int) -> int:
        return len(range(self.rank, n_sets, self.world_size))

    def gather(self, local_results, n_sets: int, dst: int = 0):
        """Collect ``{set_index: picklable result}`` from every rank on ``dst`` in set order
        (``torch.distributed.gather_object``; identity when world_size == 1).  Kept for small metadata; sample data
        goes through ``run`` / ``ResultArena`` (shared page-locked memory, no pickling)."""
        if self.world_size == 1:
            return [local_results[i] for i in range(n_sets)]
        import torch.distributed as dist
        bucket = [None] * self.world_size if self.rank == dst else None
        dist.gather_object(local_results, bucket, dst=dst)
        if self.rank != dst:
            return None
        merged = {}
        for part in bucket:
            merged.update(part)
        return [merged[i] for i in range(n_sets)]

    def run(self, dispatcher, recordings_host, results, chunk_sets=None, precision=None):
        """The designed multi-GPU step: this rank takes the recording sets it owns (s mod G) out of the
        shared input arena ``recordings_host`` [n_sets, tracks, samples] (page-locked), runs the hot path on its GPU and
        lands every set's cropped responses in its slot of ``results`` (``ResultArena``) — the final result gather IS
        that D2H copy.  No collective; the caller closes the step with a barrier, after which every rank sees all sets
        in set order.  Returns the number of sets processed here."""
        n_sets = recordings_host.shape[0]
        mine = recordings_host[self.rank::self.world_size]      # strided view: sets rank, rank + G, ...
        slots = results.data[self.rank::self.world_size]
        if mine.shape[0] == 0:
            return 0
        part = dispatcher.process_host(mine, chunk_sets=chunk_sets, precision=precision, out_sets=slots)
        for k, r in enumerate(part):
            results.put_meta(self.rank + k * self.world_size, r)
        assert self.rank + (len(part) - 1) * self.world_size < n_sets
        return len(part)

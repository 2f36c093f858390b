"""TEST INFRASTRUCTURE ONLY — CPU oracle for the Impulcifer hot path.

A numpy (float64) restatement of the reference's algorithm for the path scoped in SURVEY.md §8:
sweep deconvolution -> peak / crop / align / tail-crop -> per-IR frequency-response post-processing.
Every function cites the reference ``file:line`` it follows (paths relative to the reference root;
``scipy:`` = SciPy 1.18.1 in site-packages, the library that performs the arithmetic for the reference).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module, and only as the checker / baseline — never on the product path.
The product (``impulcifer_pip313_b200``) never imports it and raises if its CUDA library is missing.

Pinning: ``tests/test_oracle_golden.py`` checks these functions against
  * vectors produced by the *unmodified* reference imported in the build container
    (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``), and
  * the reference's own committed golden ``data/demo/room-responses.wav`` (bit-exact PCM_32).
The one unpinned piece is ``resample`` (third-party ``nnresample`` absent; see ``nnresample_port.py``).
"""
from __future__ import annotations

import math

import numpy as np

EPSILON = 1e-20  # core/impulse_response.py:12, core/decay.py:9

SPEAKER_NAMES = ['FL', 'FR', 'FC', 'BL', 'BR', 'SL', 'SR', 'WL', 'WR',
                 'TFL', 'TFR', 'TSL', 'TSR', 'TBL', 'TBR']  # core/constants.py:5
LEFT_SIDE = {'FL', 'SL', 'BL', 'WL', 'TFL', 'TSL', 'TBL'}  # core/constants.py:14-16
RIGHT_SIDE = {'FR', 'SR', 'BR', 'WR', 'TFR', 'TSR', 'TBR'}
IPSILATERAL_PAIRS = (("FL", "FR"), ("SL", "SR"), ("BL", "BR"), ("TFL", "TFR"), ("TSL", "TSR"),
                     ("TBL", "TBR"), ("FC", "FC"), ("WL", "WR"))  # core/constants.py:63-72
ONSET_GROUPS = [("FL", "FR"), ("SL", "SR"), ("BL", "BR"), ("WL", "WR"), ("TFL", "TFR"),
                ("TSL", "TSR"), ("TBL", "TBR"), ("FC",)]  # core/hrir.py:964-974
HEXADECAGONAL_TRACK_ORDER = [f'{s}-{e}' for s in ['FL', 'FR', 'FC', 'LFE', 'BL', 'BR', 'SL', 'SR', 'WL', 'WR',
                                                  'TFL', 'TFR', 'TSL', 'TSR', 'TBL', 'TBR']
                             for e in ('left', 'right')]  # core/constants.py:100-104
HESUVI_TRACK_ORDER = ['FL-left', 'FL-right', 'SL-left', 'SL-right', 'BL-left', 'BL-right', 'FC-left', 'FR-right',
                      'FR-left', 'SR-right', 'SR-left', 'BR-right', 'BR-left', 'FC-right', 'WL-left', 'WL-right',
                      'WR-left', 'WR-right', 'TFL-left', 'TFL-right', 'TFR-left', 'TFR-right', 'TSL-left',
                      'TSL-right', 'TSR-left', 'TSR-right', 'TBL-left', 'TBL-right', 'TBR-left',
                      'TBR-right']  # core/constants.py:95-98


# ----------------------------------------------------------------------------------------------
# small library restatements
# ----------------------------------------------------------------------------------------------
def next_fast_len_5smooth(target: int) -> int:
    """scipy.fftpack.next_fast_len / scipy.fft.next_fast_len(real=True): smallest 2^a 3^b 5^c >= target."""
    target = int(target)
    if target <= 6:
        return max(target, 1) if target > 0 else 1
    best = 1 << (target - 1).bit_length()
    p5 = 1
    while p5 < best:
        p35 = p5
        while p35 < best:
            q = -(-target // p35)
            p2 = 1 << max(q - 1, 0).bit_length()
            cand = p2 * p35
            if cand >= target and cand < best:
                best = cand
            p35 *= 3
            if p35 == target:
                return target
        p5 *= 5
        if p5 == target:
            return target
    return best


def hann(n: int) -> np.ndarray:
    """scipy.signal.windows.hann(n) (symmetric).  SciPy is the reference's own dependency (pyproject.toml:31-33) and
    evaluates the window as general_cosine over linspace(-pi, pi, n), which differs from 0.5 - 0.5 cos(2 pi k/(n-1))
    in the last ulp; calling it keeps the sweep envelope (core/impulse_response_estimator.py:109-131), the crop fades
    and the decay windows bit-identical to the reference's."""
    from scipy.signal.windows import hann as _hann
    if n <= 0:
        return np.zeros(0)
    return _hann(int(n))


def hamming(n: int) -> np.ndarray:
    k = np.arange(n)
    return 0.54 - 0.46 * np.cos(2.0 * np.pi * k / (n - 1))


def linregress(x, y):
    """scipy:stats/_stats_py.py linregress -> slope, intercept (cov with bias=1)."""
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    xm, ym = np.mean(x), np.mean(y)
    ssxm, ssxym, _, _ = np.cov(x, y, bias=1).flat
    with np.errstate(divide='ignore', invalid='ignore'):
        slope = ssxym / ssxm
    return slope, ym - slope * xm


# ----------------------------------------------------------------------------------------------
# a1/a2: estimator  (core/impulse_response_estimator.py)
# ----------------------------------------------------------------------------------------------
def sweep_parameters(fs: int):
    """core/impulse_response_estimator.py:33-46 -> (n_octaves P, low, high)."""
    high = fs / 2
    P = np.ceil(np.log2(high / 5))
    return P, high / 2 ** P, high


def generate_test_signal(min_duration: float, fs: int, fade_in=0.5) -> np.ndarray:
    """core/impulse_response_estimator.py:86-147 (fade_out=None branch, the only one the reference uses)."""
    P, _, _ = sweep_parameters(fs)
    M = np.ceil(min_duration * fs * (np.pi / 2 ** P) / (np.pi * 2 * np.log(2 ** P)))
    L = M * np.pi * 2 * np.log(2 ** P) / (np.pi / 2 ** P)
    N = np.round(L)
    phase = np.pi / 2 ** P * L / np.log(2 ** P) * np.exp(np.arange(N) / N * np.log(2 ** P))
    sig = np.sin(phase)
    sec_per_oct = N / fs / P
    n_in = 2 * int(fs * sec_per_oct * fade_in)
    if n_in % 2:
        n_in += 1
    win = np.ones(len(sig))
    win[:n_in // 2] = hann(n_in)[:n_in // 2]
    return sig * win


def generate_inverse_filter(test_signal: np.ndarray, n_octaves: float) -> np.ndarray:
    """core/impulse_response_estimator.py:73-84."""
    P = n_octaves
    N = len(test_signal)
    inv = np.flip(test_signal) * (2 ** (P / N)) ** (np.arange(N) * -1) * P * np.log(2) / (1 - 2 ** -P)
    full = fft_convolve_full(inv, test_signal)
    spec = np.fft.fft(full)
    inv = inv / np.abs(spec[round(spec.shape[0] / 4)])
    return inv


def fft_convolve_full(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """scipy:signal/_signaltools.py fftconvolve(mode='full') for real 1-D input (rfft at next_fast_len)."""
    n = len(a) + len(b) - 1
    nfft = next_fast_len_5smooth(n)
    return np.fft.irfft(np.fft.rfft(a, nfft) * np.fft.rfft(b, nfft), nfft)[:n]


def estimate(recording: np.ndarray, inverse_filter: np.ndarray) -> np.ndarray:
    """core/impulse_response_estimator.py:149-151: convolve(recording, inverse_filter, 'same', 'auto').

    SciPy picks the FFT method for every size on this path; 'same' = the centred len(recording)
    samples of the full convolution, start (M-1)//2 (scipy:signal/_signaltools.py:414-421,485-551).
    """
    recording = np.asarray(recording, dtype=np.float64)
    L, M = len(recording), len(inverse_filter)
    start = (M - 1) // 2
    return fft_convolve_full(recording, inverse_filter)[start:start + L].copy()


# ----------------------------------------------------------------------------------------------
# a4: peak index  (core/impulse_response.py:32-70, twin core/decay.py:12-41)
# ----------------------------------------------------------------------------------------------
def local_maxima(x: np.ndarray) -> np.ndarray:
    """scipy:signal/_peak_finding_utils.pyx _local_maxima_1d: midpoints of strict (plateau) maxima.

    Run-length formulation: a run of equal samples is a peak iff it is neither the first nor the last
    run and both neighbouring runs are strictly lower; the reported index is (first+last)//2.
    """
    n = len(x)
    if n < 3:
        return np.zeros(0, dtype=np.intp)
    change = np.flatnonzero(x[1:] != x[:-1]) + 1
    starts = np.concatenate(([0], change))
    ends = np.concatenate((change - 1, [n - 1]))
    vals = x[starts]
    if len(vals) < 3:
        return np.zeros(0, dtype=np.intp)
    mid = (vals[1:-1] > vals[:-2]) & (vals[1:-1] > vals[2:])
    idx = np.flatnonzero(mid) + 1
    return ((starts[idx] + ends[idx]) // 2).astype(np.intp)


def peak_index(data: np.ndarray, start=0, end=None, peak_height=0.12589) -> int:
    """core/impulse_response.py:32-70."""
    if len(data) == 0:
        return 0
    if end is None:
        end = len(data)
    seg = np.array(data[start:end], dtype=np.float64)
    if len(seg) == 0:
        return start
    m = np.max(np.abs(seg))
    if m < EPSILON:
        return start
    seg /= m
    cands = []
    for s in (seg, seg * -1.0):
        p = local_maxima(s)
        p = p[s[p] >= peak_height]
        if len(p):
            cands.append(p[0])
    if not cands:
        return int(np.argmax(np.abs(seg)) + start)
    return int(min(cands) + start)


# ----------------------------------------------------------------------------------------------
# a8: Lundeby decay parameters  (core/decay.py:44-260)
# ----------------------------------------------------------------------------------------------
def _db10(v):
    return 10 * np.log10(np.maximum(v, EPSILON))


def _first_true(mask):
    nz = np.flatnonzero(mask)
    return int(nz[0]) if len(nz) else None


def decay_params(ir: np.ndarray, fs: int):
    """core/decay.py:44-260 -> (peak_index, knee_point_index, noise_floor_dB, window_size)."""
    ir = np.asarray(ir, dtype=np.float64)
    n_ir = len(ir)
    if n_ir < 10:
        return 0, n_ir, -200.0, n_ir if n_ir > 0 else 1

    pk = peak_index(ir)
    stop = min(pk + int(2 * fs), n_ir)
    seg = ir[pk:stop].copy() if pk < stop else ir[min(max(pk, 0), n_ir - 1):][:1].copy()
    mx = np.max(np.abs(seg))
    if mx >= EPSILON:
        seg = seg / mx
    sq = seg ** 2
    ns = len(sq)
    t_sq = np.linspace(0, ns / fs, ns)

    def nearest_t(t):
        return int(np.argmin(np.abs(t_sq - t)))

    def mean_db_all():
        return 10 * np.log10(max(np.mean(sq), EPSILON))

    # --- first pass: 30 ms blocks ----------------------------------------------------------
    wd = 0.03
    n = int(ns / fs / wd)
    if n == 0:
        return pk, pk + ns, mean_db_all(), max(1, ns)
    w = max(int(ns / n), 1)
    w_first = w
    t_win = np.arange(n) * wd + wd / 2
    win = _db10(np.mean(sq[:n * w].reshape(n, w), axis=1))

    tail = sq[int(ns * 0.9):]
    if len(tail) == 0:
        tail = sq
    noise = 10 * np.log10(np.maximum(np.mean(tail), EPSILON))

    slope_end = len(win)
    c0 = _first_true(win <= noise + 10.0)
    if c0 is not None and c0 > 0:
        slope_end = c0
    if slope_end < 2:
        if len(win) >= 2:
            slope_end = len(win)
        else:
            return pk, pk + ns, noise, w_first

    slope, icpt = linregress(t_win[:slope_end], win[:slope_end])
    if np.isnan(slope) or abs(slope) < EPSILON:
        return pk, pk + ns, noise, w_first
    knee_t = np.clip((noise - icpt) / slope, t_sq[0], t_sq[-1])

    # --- second pass: 3 blocks per 10 dB of decay ------------------------------------------
    den = abs(slope) * 3
    wd = (t_sq[-1] / 3.0) if den < EPSILON else 10 / den
    n = 1 if wd <= EPSILON else int(ns / fs / wd)
    if n == 0:
        n = 1
    w = max(int(ns / n), 1)
    t_win = np.arange(n) * wd + wd / 2
    if n * w > ns:  # only reachable with w forced to 1 (core/decay.py:150-155)
        n = ns // w
    win = _db10(np.mean(sq[:n * w].reshape(n, w), axis=1))

    k_idx = _first_true(t_win >= knee_t)
    if k_idx is None or k_idx >= len(win):  # the reference's ``except IndexError`` (core/decay.py:166-175)
        knee_t = t_win[-1]
        k_idx = len(t_win) - 1
        knee_val = win[-1]
    else:
        knee_val = win[k_idx]

    noise_it, knee_t_it, knee_val_it, k_idx_it = noise, knee_t, knee_val, k_idx
    total = t_sq[-1]
    for _ in range(5):
        nf0 = _first_true(win <= knee_val_it - 5)
        if nf0 is None:
            break
        t0 = max(t_win[nf0], 0.1 * total)
        if t0 > t_win[-1]:
            break
        t1 = min(t0 + knee_t_it, total)
        a, b = nearest_t(t0), nearest_t(t1)
        if a >= b:
            break
        noise_it = 10 * np.log10(np.maximum(np.mean(sq[a:b]), EPSILON))

        e = _first_true(win <= noise_it + 8)
        s = _first_true(win <= noise_it + 8 + 20)
        if e is None or s is None:
            break
        e -= 1
        s = max(s - 1, 0)
        if e <= s + 1:
            break
        l_slope, l_icpt = linregress(t_win[s:e], win[s:e])
        if np.isnan(l_slope) or abs(l_slope) < EPSILON:
            break
        new_t = np.clip((noise_it - l_icpt) / l_slope, t_win[0], t_win[-1])
        new_idx = _first_true(t_win >= new_t)
        if new_idx is None:
            new_idx = len(t_win) - 1
        if new_idx == k_idx_it:
            knee_t_it = t_win[k_idx_it]
            break
        k_idx_it = new_idx
        knee_t_it = t_win[k_idx_it]
        knee_val_it = win[k_idx_it]

    return pk, pk + nearest_t(knee_t_it), noise_it, w


# ----------------------------------------------------------------------------------------------
# a3: recording ingest — integer column/track logic  (core/hrir.py:76-363)
# ----------------------------------------------------------------------------------------------
def ingest_layout(n_tracks: int, n_samples: int, speakers, est_len: int, fs: int, side=None,
                  silence_length=2.0):
    """Index arithmetic of core/hrir.py:146-355.

    Returns (silence, [(col_start, col_end)...] relative to the *uncropped* recording,
             [(speaker, side, track, column)...] in the order the reference estimates them).
    """
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
    columns = []
    for i in range(n_columns):
        a, b = i * col, min((i + 1) * col, avail)
        if b > a and (b - a) >= est_len:
            columns.append((base + a, base + b))
    if not columns:
        # fallback 1: shorter leading silence (core/hrir.py:221-262)
        if sil > 0:
            min_sil = int(0.5 * fs)
            spare = avail - est_len
            if spare >= min_sil:
                adj = max(min_sil, spare)
                avail2 = avail - adj
                for i in range(n_columns):
                    a, b = i * est_len, min((i + 1) * est_len, avail2)
                    if b > a and (b - a) >= est_len:
                        columns.append((sil + adj + a, sil + adj + b))
                if columns:
                    base, avail = sil + adj, avail2
        # fallback 2: accept >= 80 % of the sweep length (core/hrir.py:264-287)
        if not columns and avail > est_len * 0.8:
            if n_columns == 1:
                columns.append((base, base + avail))
            else:
                c = avail // n_columns
                for i in range(n_columns):
                    a, b = i * c, min((i + 1) * c, avail)
                    if b > a:
                        columns.append((base + a, base + b))
        if not columns:
            raise ValueError("No valid columns could be extracted even with fallback methods.")
    plan = []
    i = 0
    while i < n_tracks:
        for j in range(len(columns)):
            n = int(i // 2 * len(columns) + j)
            if n >= len(speakers):
                continue
            spk = speakers[n]
            if spk not in SPEAKER_NAMES:
                continue
            if side is None:
                if i + 1 < n_tracks:
                    plan.append((spk, 'left', i, j))
                    plan.append((spk, 'right', i + 1, j))
            else:
                plan.append((spk, side, i, j))
        i += tracks_k
    return sil, columns, plan


def open_recording(recording: np.ndarray, fs: int, speakers, inverse_filter: np.ndarray, est_fs: int,
                   side=None, silence_length=2.0):
    """core/hrir.py:76-363,396-424 with the WAV already decoded to float64 [tracks, samples]."""
    if fs != est_fs:
        raise ValueError("Sampling rate of recording must match sampling rate of test signal.")
    recording = np.atleast_2d(recording)
    _, columns, plan = ingest_layout(recording.shape[0], recording.shape[1], speakers, len(inverse_filter),
                                     fs, side, silence_length)
    irs = {}
    for spk, sd, trk, col in plan:
        a, b = columns[col]
        irs.setdefault(spk, {})[sd] = estimate(recording[trk, a:b], inverse_filter)
    return irs


# ----------------------------------------------------------------------------------------------
# a5/a6/a7/a9: crop / align  (core/impulse_response.py:82-108, core/hrir.py:548-653, 921-1001)
# ----------------------------------------------------------------------------------------------
def crop_head(data: np.ndarray, fs: int, head_ms=1) -> np.ndarray:
    """core/impulse_response.py:82-90 (room-measurement path; no fade-in)."""
    if len(data) == 0:
        return data
    return data[max(peak_index(data) - int(fs * head_ms / 1000), 0):]


def crop_heads(irs: dict, fs: int, head_ms=1) -> dict:
    """core/hrir.py:548-612. ``irs`` = {speaker: {'left': arr, 'right': arr}}; returns crop indices."""
    crops = {}
    head = int(head_ms * fs / 1000)
    delay = int(np.round(0 * fs)) + head  # SPEAKER_DELAYS are all 0 (core/constants.py:52-54)
    fade = hann(head * 2)[:head]
    for spk, pair in irs.items():
        pl, pr = peak_index(pair['left']), peak_index(pair['right'])
        c = max(0, (pl if pl < pr else pr) - delay)
        crops[spk] = c
        pair['left'] = pair['left'][c:].copy()
        pair['right'] = pair['right'][c:].copy()
        if len(pair['left']) >= head and len(pair['right']) >= head:
            pair['left'][:head] *= fade
            pair['right'][:head] *= fade
    return crops


def shift(data: np.ndarray, samples: int) -> np.ndarray:
    """core/impulse_response.py:92-108: length-preserving integer shift."""
    n = len(data)
    out = np.zeros(n)
    if samples >= 0:
        if samples < n:
            out[samples:] = data[:n - samples]
    else:
        k = -samples
        if k < n:
            out[:n - k] = data[k:]
    return out


def xcorr_lag(a: np.ndarray, b: np.ndarray) -> int:
    """core/hrir.py:932-935: argmax of scipy.signal.correlate(a, b, 'full'), as a lag."""
    from scipy import signal
    corr = signal.correlate(a, b, mode="full")
    return int(np.arange(-len(a) + 1, len(a))[np.argmax(corr)])


def align_ipsilateral_all(irs: dict, fs: int, segment_ms=30) -> None:
    """core/hrir.py:921-955."""
    seg = int(fs * segment_ms / 1000)
    for s1, s2 in IPSILATERAL_PAIRS:
        if s1 not in irs or s2 not in irs:
            continue
        if s1 == s2:
            lag = xcorr_lag(irs[s1]['left'][:seg], irs[s1]['right'][:seg])
            if lag > 0:
                irs[s1]['right'] = shift(irs[s1]['right'], lag)
            elif lag < 0:
                irs[s1]['left'] = shift(irs[s1]['left'], -lag)
            continue
        lag = xcorr_lag(irs[s1]['left'][:seg], irs[s2]['right'][:seg])
        if lag > 0:
            for sd in ('left', 'right'):
                irs[s2][sd] = shift(irs[s2][sd], lag)
        elif lag < 0:
            for sd in ('left', 'right'):
                irs[s1][sd] = shift(irs[s1][sd], -lag)


def align_onset_groups_peak_leftref(irs: dict) -> None:
    """core/hrir.py:957-1001."""
    if 'FL' not in irs or 'left' not in irs['FL']:
        raise RuntimeError("Cannot find FL left channel reference for onset alignment.")
    ref = peak_index(irs['FL']['left'])
    for grp in ONSET_GROUPS:
        if grp == ("FL", "FR"):
            continue
        if grp[0] not in irs or 'left' not in irs[grp[0]]:
            continue
        d = peak_index(irs[grp[0]]['left']) - ref
        for spk in grp:
            if spk in irs:
                for sd in ('left', 'right'):
                    irs[spk][sd] = shift(irs[spk][sd], -d)


def crop_tails(irs: dict, fs: int, est_len: int, n_octaves: float, est_fs: int) -> int:
    """core/hrir.py:614-653."""
    knees, lens = [], []
    for pair in irs.values():
        for d in pair.values():
            knees.append(decay_params(d, fs)[1])
            lens.append(len(d))
    if not knees:
        return 0
    fade = 2 * int(fs * (est_len / est_fs / n_octaves) * (1 / 24))
    win = hann(fade)[fade // 2:]
    tail = min(np.min(lens), next_fast_len_5smooth(max(knees)))
    for pair in irs.values():
        for sd in pair:
            d = pair[sd][:tail].copy()
            d *= np.concatenate([np.ones(len(d) - len(win)), win])
            pair[sd] = d
    return int(tail)


def track_matrix(irs: dict, track_order=None) -> np.ndarray:
    """core/hrir.py:426-455: rows in track order, missing tracks are zeros of the first IR's length."""
    if track_order is None:
        track_order = HEXADECAGONAL_TRACK_ORDER
    by_name = {f'{s}-{e}': d for s, pair in irs.items() for e, d in pair.items()}
    if not by_name:
        raise ValueError("No impulse responses available for WAV output.")
    n = len(next(iter(by_name.values())))
    return np.vstack([by_name.get(ch, np.zeros(n)) for ch in track_order])


def pcm32(x: np.ndarray) -> np.ndarray:
    """libsndfile PCM_32 quantisation of float64 in [-1, 1): lrint(x * 2^31) with clipping."""
    return np.clip(np.rint(np.asarray(x, dtype=np.float64) * 2.0 ** 31), -2 ** 31, 2 ** 31 - 1).astype(np.int32)


# ----------------------------------------------------------------------------------------------
# a11: decay window  (core/decay.py:383-403)
# ----------------------------------------------------------------------------------------------
def apply_decay_window(data: np.ndarray, params) -> np.ndarray:
    if params is None:
        return data
    start, half, knee, level = params
    w = np.concatenate([np.ones(start), hann(half * 2)[half:], np.zeros(len(data) - knee)]) - 1.0
    return data * 10 ** (w * -level / 20)


# ----------------------------------------------------------------------------------------------
# a12/a13/a19: frequency response  (core/audio_io.py:100-113, core/impulse_response.py:157-188,
#                                   autoeq/frequency_response.py:850-901)
# ----------------------------------------------------------------------------------------------
def magnitude_response(x: np.ndarray, fs: int):
    n = len(x)
    half = int(np.ceil(n / 2))
    with np.errstate(divide='ignore'):
        mag = 20 * np.log10(np.abs(np.fft.rfft(x)[:half]))
    return np.arange(half) * (fs / n), mag


def generate_frequencies(f_min=20, f_max=20000, f_step=1.01) -> np.ndarray:
    out, f = [], f_min
    while f <= f_max:
        out.append(f)
        f *= f_step
    return np.array(out, dtype=np.float64)


def loglerp(f_in: np.ndarray, y_in: np.ndarray, f_out: np.ndarray) -> np.ndarray:
    """InterpolatedUnivariateSpline(log10 f, y, k=1)(log10 f_out): piecewise-linear in log-f with linear
    extrapolation from the end segments (autoeq/frequency_response.py:874-894)."""
    x = np.log10(np.asarray(f_in, dtype=np.float64))
    fo = np.array(f_out, dtype=np.float64)
    if len(fo) and fo[0] == 0:
        fo[0] = 0.001
    xo = np.log10(fo)
    j = np.clip(np.searchsorted(x, xo, side='right') - 1, 0, len(x) - 2)
    x0, x1 = x[j], x[j + 1]
    return (y_in[j] * (x1 - xo) + y_in[j + 1] * (xo - x0)) / (x1 - x0)


def frequency_response(ir: np.ndarray, fs: int):
    """core/impulse_response.py:157-188 -> (frequency[F], raw[F]) on the 1.01-step grid 10 Hz..fs/2."""
    grid = generate_frequencies(f_min=10, f_max=fs / 2, f_step=1.01)
    if len(ir) < 2:
        return grid, np.zeros_like(grid)
    f, m = magnitude_response(ir, fs)
    pts = (fs / 2) / 4.0
    step = 1 if (pts < 2 or len(f) < 2) else (int(round(len(f) / pts)) or 1)
    fd, md = f[1::step], m[1::step]
    if len(fd) == 0:
        fd, md = f[1:], m[1:]
        if len(fd) == 0:
            return grid, np.zeros_like(grid)
    return grid, loglerp(fd, md, grid)


# ----------------------------------------------------------------------------------------------
# a14-a16: fractional-octave smoothing  (autoeq/frequency_response.py:1033-1239)
# ----------------------------------------------------------------------------------------------
def window_size(frequency: np.ndarray, octaves: float) -> int:
    """autoeq/frequency_response.py:1033-1050 (sequential Python sum and banker's ``round`` preserved)."""
    steps = [frequency[i] / frequency[i - 1] for i in range(1, len(frequency))]
    step = sum(steps) / len(steps)
    w = round(math.log(2 ** octaves) / math.log(step))
    return w if w % 2 else w + 1


def sigmoid(frequency: np.ndarray, f_lower: float, f_upper: float) -> np.ndarray:
    """autoeq/frequency_response.py:1052-1058 with a_normal=0, a_treble=1."""
    fc = np.sqrt(f_upper / f_lower) * f_lower
    half = np.log10(f_upper) - np.log10(fc)
    z = (np.log10(frequency) - np.log10(fc)) / (half / 4)
    a = 1.0 / (1.0 + np.exp(-z))
    return a * -(0.0 - 1.0) + 0.0


def savgol_coeffs(window: int, order: int = 2) -> np.ndarray:
    """scipy:signal/_savitzky_golay.py:100-160 (use='conv', deriv=0, pos=centre)."""
    half = window // 2
    x = np.arange(-half, window - half, dtype=np.float64)[::-1]
    A = x ** np.arange(order + 1, dtype=np.float64).reshape(-1, 1)
    y = np.zeros(order + 1)
    y[0] = 1.0
    return np.linalg.lstsq(A, y, rcond=None)[0]


def savgol_edge_matrix(window: int, order: int = 2) -> np.ndarray:
    """Rows p=0..half-1: weights w_p with  fitted(p) = w_p . y[0:window]  for the least-squares
    polynomial of degree ``order`` through the first ``window`` samples
    (scipy:signal/_savitzky_golay.py:190-258, mode='interp').  The trailing edge uses the mirrored rows."""
    half = window // 2
    t = np.arange(window, dtype=np.float64)
    V = np.vander(t, order + 1)  # columns t^order .. t^0  (np.polyfit convention)
    pinv = np.linalg.pinv(V)
    P = np.vander(np.arange(half, dtype=np.float64), order + 1)
    return P @ pinv


def savgol_filter(y: np.ndarray, window: int, order: int = 2) -> np.ndarray:
    """scipy.signal.savgol_filter(y, window, order) with default mode='interp'."""
    y = np.asarray(y, dtype=np.float64)
    n = len(y)
    if window > n:
        raise ValueError("If mode is 'interp', window_length must be less than or equal to the size of x.")
    half = window // 2
    c = savgol_coeffs(window, order)
    out = np.convolve(y, c, mode='same') if n >= window else np.zeros(n)
    E = savgol_edge_matrix(window, order)
    out[:half] = E @ y[:window]
    out[n - half:] = (E @ y[::-1][:window])[::-1]
    return out


def smoothen_fractional_octave(frequency, data, window=1 / 3, iterations=1, treble_window=1 / 3,
                               treble_iterations=1, treble_f_lower=100.0, treble_f_upper=10000.0):
    """autoeq/frequency_response.py:1060-1105."""
    frequency = np.asarray(frequency, dtype=np.float64)
    data = np.asarray(data, dtype=np.float64)
    if np.any(np.isnan(frequency)) or np.any(np.isnan(data)):
        raise ValueError('NaN values present, cannot smoothen!')
    yn = data
    wn = window_size(frequency, window)
    for _ in range(iterations):
        yn = savgol_filter(yn, wn, 2)
    yt = data
    wt = window_size(frequency, treble_window)
    for _ in range(treble_iterations):
        yt = savgol_filter(yt, wt, 2)
    kt = sigmoid(frequency, treble_f_lower, treble_f_upper)
    kn = kt * -1 + 1
    return yn * kn + yt * kt


def smoothen_heavy_light(frequency, raw, error):
    """autoeq/frequency_response.py:1181-1239 -> (smoothed, error_smoothed)."""
    light = smoothen_fractional_octave(frequency, error, 1 / 6, 1, 1 / 3, 1, 100, 10000)
    heavy = smoothen_fractional_octave(frequency, error, 1 / 3, 1, 1.3, 1, 1000, 6000)
    comb = np.max(np.vstack([light, heavy]), axis=0)
    smoothed = smoothen_fractional_octave(frequency, raw, 1 / 3, 1, 1 / 3, 1, 100, 10000)
    err_s = smoothen_fractional_octave(frequency, comb, 1 / 3, 1, 1 / 3, 1, 100, 10000)
    return smoothed, err_s


# ----------------------------------------------------------------------------------------------
# a18: minimum-phase FIR  (autoeq/frequency_response.py:637-681, scipy firwin2 + minimum_phase)
# ----------------------------------------------------------------------------------------------
def minphase_grid(frequency, equalization, fs, f_res=10, normalize=True):
    """First half of autoeq/frequency_response.py:637-676: the linear-amplitude gain on
    ``linspace(0, fs//2, n)`` that is handed to firwin2 -> (f[n], gain[n])."""
    frequency = np.asarray(frequency, dtype=np.float64)
    eq = np.asarray(equalization, dtype=np.float64)
    f_res = f_res / 2
    f_min = np.max([frequency[0], f_res])
    g_min = loglerp(frequency, eq, np.array([f_min]))[0]
    n = next_fast_len_5smooth(round(fs // 2 / f_res))
    f = np.linspace(0.0, fs // 2, n)
    g = loglerp(frequency, eq, f)
    g[f <= f_min] = g_min
    if normalize:
        g -= np.max(g)
        g -= 0.5
    g *= 2
    g = 10 ** (g / 20)
    g[-1] = 0.0
    return f, g


def firwin2_even(numtaps: int, freq: np.ndarray, gain: np.ndarray, fs: float) -> np.ndarray:
    """scipy:signal/_fir_filter_design.py:577-775 for even numtaps (type II), Hamming window."""
    nyq = 0.5 * fs
    nfreqs = 1 + 2 ** int(math.ceil(math.log(numtaps, 2)))
    x = np.linspace(0.0, nyq, nfreqs)
    fx = np.interp(x, freq, gain)
    shift_ = np.exp(-(numtaps - 1) / 2. * 1j * np.pi * x / nyq)
    full = np.fft.irfft(fx * shift_)
    return full[:numtaps] * hamming(numtaps)


def minimum_phase_half(h: np.ndarray, n_fft: int) -> np.ndarray:
    """scipy:signal/_fir_filter_design.py:1383-1406, method='homomorphic', half=True."""
    n_half = len(h) // 2
    t = np.abs(np.fft.fft(h, n_fft))
    t += 1e-7 * np.min(t[t > 0])
    t = np.log(t)
    t *= 0.5
    t = np.fft.ifft(t).real
    win = np.zeros(n_fft)
    win[0] = 1
    stop = n_fft // 2
    win[1:stop] = 2
    win[stop] = 1 + (n_fft % 2)
    t *= win
    t = np.fft.ifft(np.exp(np.fft.fft(t)))
    return t.real[:n_half + len(h) % 2]


def minimum_phase_impulse_response(frequency, equalization, fs=44100, f_res=10, normalize=True):
    """autoeq/frequency_response.py:637-681."""
    f, g = minphase_grid(frequency, equalization, fs, f_res, normalize)
    h = firwin2_even(len(f) * 2, f, g, fs)
    return minimum_phase_half(h, len(h))


# ----------------------------------------------------------------------------------------------
# a10: resample  (core/impulse_response.py:121-124) — PARITY UNPINNED, see nnresample_port.py
# ----------------------------------------------------------------------------------------------
def resample_filter(fs_new: int, fs_old: int):
    """Returns (up, down, taps) exactly as resample_poly will use them (taps NOT yet scaled by up)."""
    import nnresample_port
    g = math.gcd(int(fs_new), int(fs_old))
    return int(fs_new) // g, int(fs_old) // g, nnresample_port.compute_filt(fs_new, fs_old, N=16001, beta=5.0)


def resample(data: np.ndarray, fs_new: int, fs_old: int) -> np.ndarray:
    import nnresample_port
    return nnresample_port.resample(data, fs_new, fs_old)

"""K1 parity: CUDA sweep deconvolution vs the CPU oracle and the reference golden vectors."""
import numpy as np
import pytest

import impulcifer_oracle as O
from _util import assert_ck, load, rel_peak_err

pytestmark = pytest.mark.gpu

FP64_TOL = 1e-9          # north_star: FP64 kernel within 1e-9 of the peak IR magnitude
FP32_TOL_DB = -90.0      # north_star: FP32 fast path within -90 dB re peak


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _est(dur, fs, **kw):
    from impulcifer_pip313_b200.estimator import ImpulseResponseEstimator
    return ImpulseResponseEstimator(min_duration=dur, fs=fs, **kw)


@pytest.mark.parametrize("dur,fs", [(0.05, 2000), (1.0, 48000), (2.0, 44100)])
def test_small_estimators_match_reference_golden(torch_cuda, dur, fs):
    g = load('small_cases.npz')
    tag = {2000: 'fs2000', 48000: 'fs48000_1s', 44100: 'fs44100_2s'}[fs]
    est = _est(dur, fs)
    assert len(est) == int(g[f'{tag}_len'])
    assert_ck(est.test_signal, g[f'{tag}_test_signal_ck'], rel=1e-12)
    assert_ck(est.inverse_filter, g[f'{tag}_inverse_filter_ck'], rel=1e-10)
    for delay in (0, 1, 127, 1000):
        if delay >= fs:
            continue
        rec = np.concatenate([np.zeros(delay), est.test_signal, np.zeros(int(0.2 * fs))])
        y = est.estimate(rec)
        assert y.dtype == np.float64 and y.shape == rec.shape
        ref = O.estimate(rec, est.inverse_filter)
        assert rel_peak_err(y, ref) < FP64_TOL
        assert_ck(y, g[f'{tag}_d{delay}_est_ck'], rel=1e-8, what=f'{tag} d{delay}')
        assert O.peak_index(y) == int(g[f'{tag}_d{delay}_peak'])
        if f'{tag}_d{delay}_est' in g.files:
            assert rel_peak_err(y, g[f'{tag}_d{delay}_est']) < FP64_TOL


def test_batch_odd_count_and_formats(torch_cuda):
    torch = torch_cuda
    est = _est(1.0, 48000)
    rng = np.random.default_rng(5)
    L = 2 * est.fs + len(est)
    rec = rng.standard_normal((5, L)) * 0.05
    ref = np.stack([O.estimate(r, est.inverse_filter) for r in rec])
    got = est.estimate_batch(rec)
    assert got.shape == ref.shape
    assert rel_peak_err(got, ref) < FP64_TOL
    # int32 PCM input is scaled by 2^-31 on load, exactly like libsndfile
    pcm = np.rint(rec * 2 ** 31).astype(np.int32)
    ref_pcm = np.stack([O.estimate(r.astype(np.float64) / 2 ** 31, est.inverse_filter) for r in pcm])
    got_pcm = est.estimate_batch(torch.from_numpy(pcm).cuda()).cpu().numpy()
    assert rel_peak_err(got_pcm, ref_pcm) < FP64_TOL
    # cropped output window
    got_win = est.estimate_batch(torch.from_numpy(rec).cuda(), out_offset=1000, out_len=4096).cpu().numpy()
    assert rel_peak_err(got_win, ref[:, 1000:1000 + 4096]) < FP64_TOL
    # shorter (clipped) recording through the same plan
    short = rec[:3, :L - 12345]
    ref_s = np.stack([O.estimate(r, est.inverse_filter) for r in short])
    assert rel_peak_err(est.estimate_batch(np.ascontiguousarray(short)), ref_s) < FP64_TOL


def test_fp32_fast_path(torch_cuda):
    est = _est(1.0, 48000)
    rec = np.zeros((2, 2 * est.fs + len(est)))
    rec[0, 5000:5000 + len(est)] = est.test_signal
    rec[1, 9000:9000 + len(est)] = 0.3 * est.test_signal
    rec += 1e-4 * np.random.default_rng(1).standard_normal(rec.shape)
    ref = np.stack([O.estimate(r, est.inverse_filter) for r in rec])
    got = est.estimate_batch(rec, precision=32).astype(np.float64)
    for i in range(2):
        err_db = 20 * np.log10(np.max(np.abs(got[i] - ref[i])) / np.max(np.abs(ref[i])))
        assert err_db < FP32_TOL_DB, err_db


def test_headline_shape_synthetic_71(torch_cuda):
    """6.15 s @ 48 kHz sweep, 7.1 layout: 14 columns of 391 270 samples (BASELINE config 1)."""
    from impulcifer_pip313_b200 import synth
    g = load('synth71_golden.npz')
    est = _est(5.0, 48000)
    assert len(est) == 295270
    rec = synth.recording_set(est.test_signal, est.fs, synth.LAYOUT_71, seed=1234)
    cols = synth.columns_of(rec, est.fs, len(est), 7)
    assert cols.shape == (14, 391270)
    plan = est.plan(cols.shape[1])
    assert plan.nfft >= 538905
    got = est.estimate_batch(cols)
    for i in range(14):
        assert_ck(got[i], g['est_ck'][i], rel=1e-8, what=str(g['names'][i]))
        pk = int(g['est_peaks'][i])
        assert O.peak_index(got[i]) == pk
        assert rel_peak_err(got[i][pk - 64:pk + 1984], g['est_slices'][i]) < FP64_TOL
    ref0 = O.estimate(cols[0], est.inverse_filter)
    assert rel_peak_err(got[0], ref0) < FP64_TOL
    # linearity (size-independent property): estimate(a x + b y) == a estimate(x) + b estimate(y)
    mix = 0.7 * cols[2] - 1.3 * cols[5]
    lin = est.estimate(mix)
    assert rel_peak_err(lin, 0.7 * got[2] - 1.3 * got[5]) < 1e-11


def test_demo_recording_golden(torch_cuda):
    inp, gold = load('demo_room_inputs.npz'), load('demo_room_golden.npz')
    from impulcifer_pip313_b200.estimator import ImpulseResponseEstimator
    est = ImpulseResponseEstimator(min_duration=(295270 - 1) / 48000, fs=48000)
    fc = inp['FC.wav'].astype(np.float64) / 2 ** 31
    cols = np.ascontiguousarray(fc[:, 96000:96000 + 391270])
    got = est.estimate_batch(cols)
    assert [O.peak_index(got[0]), O.peak_index(got[1])] == list(gold['fc_peaks'])
    rec = inp['room-FC-left.wav'].astype(np.float64) / 2 ** 31
    y = est.estimate(rec[96000:])
    assert_ck(y, gold['est_ck_FC-left'], rel=1e-8)


def test_full_size_properties(torch_cuda):
    """Size-independent properties at the headline size (L = 391 270): time-shift equivariance, homogeneity,
    empty input, and agreement of the three input formats."""
    torch = torch_cuda
    est = _est(5.0, 48000)
    L = 2 * est.fs + len(est)
    rng = np.random.default_rng(21)
    x = np.zeros(L)
    x[40000:40000 + len(est)] = est.test_signal * 0.4
    x += 1e-4 * rng.standard_normal(L)
    y = est.estimate(x)
    assert y.shape == (L,) and y.dtype == np.float64
    # delaying the recording by d samples delays the response by d (samples pushed past the end are lost)
    d = 777
    xd = np.concatenate([np.zeros(d), x[:-d]])
    yd = est.estimate(xd)
    pk = int(np.argmax(np.abs(y)))
    assert int(np.argmax(np.abs(yd))) == pk + d
    assert rel_peak_err(yd[pk + d - 2000:pk + d + 20000], y[pk - 2000:pk + 20000]) < 1e-9
    # homogeneity
    assert rel_peak_err(est.estimate(-2.5 * x), -2.5 * y) < 1e-12
    # empty recording
    assert est.estimate(np.zeros(0)).shape == (0,)
    # f32 and int32-PCM containers of exactly representable samples give the f64 result
    q = np.rint(x * 2 ** 23) / 2 ** 23            # 24-bit samples: exact in f32 and in PCM_32
    ref = est.estimate_batch(q[None, :])[0]
    got32 = est.estimate_batch(torch.from_numpy(q.astype(np.float32)[None, :]).cuda()).cpu().numpy()[0]
    gotpcm = est.estimate_batch(torch.from_numpy(np.rint(q * 2 ** 31).astype(np.int32)[None, :]).cuda()).cpu().numpy()[0]
    assert np.array_equal(got32, ref) and np.array_equal(gotpcm, ref)


def test_absmax_and_argmax_outputs(torch_cuda):
    T = torch_cuda
    """absmax_out / argmax_out of impb_deconv_execute: exact max |row| and an index attaining it, for the
    specialised 768x768 plan (tracked in the final pass) and a generic plan (extra pass), odd batch."""
    from impulcifer_pip313_b200 import ImpulseResponseEstimator
    for dur, fs, batch in ((5.0, 48000, 5), (0.3, 44100, 3)):
        est = ImpulseResponseEstimator(min_duration=dur, fs=fs)
        L = 2 * fs + len(est)
        g = T.Generator(device='cuda').manual_seed(11)
        x = T.randn(batch, L, dtype=T.float64, device='cuda', generator=g) * 0.05
        x[1] = 0.0  # an all-zero row
        plan = est.plan(L)
        amax = T.empty(batch, dtype=T.float64, device='cuda')
        arg = T.empty(batch, dtype=T.int64, device='cuda')
        out = plan.execute(x, absmax=amax, argmax=arg)
        ref = out.abs().max(dim=1).values
        assert T.equal(amax, ref)
        a = arg.cpu().numpy()
        o = out.cpu().numpy()
        for r in range(batch):
            if ref[r] > 0:
                assert 0 <= a[r] < L and abs(o[r, a[r]]) == float(ref[r]), r
        # block maxima (specialised plans only) and the peak search that starts from them
        nblk, blk_len, blk_first = plan.blockmax_info()
        if nblk:
            from impulcifer_pip313_b200 import ops
            bm = T.empty((batch, nblk), dtype=T.float64, device='cuda')
            out_b = plan.execute(x, absmax=amax, argmax=arg, blockmax=bm)
            assert T.equal(out_b, out)
            ob = out.abs().cpu().numpy()
            want = np.zeros((batch, nblk))
            for q in range(nblk):
                lo, hi = max(blk_first + q * blk_len, 0), min(blk_first + (q + 1) * blk_len, L)
                if hi > lo:
                    want[:, q] = ob[:, lo:hi].max(axis=1)
            assert np.array_equal(bm.cpu().numpy(), want)
            offs = T.arange(batch, device='cuda', dtype=T.int64) * L
            lens = T.full((batch,), L, device='cuda', dtype=T.int64)
            p0 = ops.peak_index(out.reshape(-1), offs, lens, max_len=L, known_max=amax)
            p1 = ops.peak_index(out.reshape(-1), offs, lens, max_len=L, known_max=amax,
                                view=dict(blocks=(bm, blk_len, blk_first)))
            assert T.equal(p0, p1)
        else:
            assert blk_len == 0
        # a cropped output window shifts the index origin
        out2 = plan.execute(x, out_offset=1000, out_len=50000, absmax=amax, argmax=arg)
        assert T.equal(amax, out2.abs().max(dim=1).values)
        a2, o2 = arg.cpu().numpy(), out2.cpu().numpy()
        assert abs(o2[0, a2[0]]) == float(amax[0])


@pytest.mark.parametrize("fs,plan", [(96000, (768, 1536)), (192000, (1536, 1600))])
def test_high_rate_plans_use_half_width_tiles_and_match_oracle(torch_cuda, fs, plan):
    """88.2 ... 192 kHz sweeps: specialised kernels at half tile width (768 x 1536, 1536 x 1600), FP64 and FP32."""
    T = torch_cuda
    est = _est(5.0, fs)
    L = 2 * fs + len(est)
    p64 = est.plan(L)
    assert (p64.n1, p64.n2) == plan
    rng = np.random.default_rng(fs)
    rec = np.zeros((3, L))
    for r in range(3):
        d = int(rng.integers(1000, 50000))
        h = np.zeros(4000)
        h[0] = 1.0
        h[1:] = 0.1 * rng.standard_normal(3999) * np.exp(-np.arange(3999) / 600.0)
        seg = np.convolve(est.test_signal, h)[:L - d]
        rec[r, d:d + len(seg)] = seg
    want = np.stack([O.estimate(rec[r], est.inverse_filter) for r in range(3)])
    amax = T.empty(3, dtype=T.float64, device='cuda')
    arg = T.empty(3, dtype=T.int64, device='cuda')
    nblk, blk_len, blk_first = p64.blockmax_info()
    assert nblk == p64_radix0(plan)
    bm = T.empty((3, nblk), dtype=T.float64, device='cuda')
    got = p64.execute(T.from_numpy(rec).cuda(), absmax=amax, argmax=arg, blockmax=bm).cpu().numpy()
    for r in range(3):
        assert rel_peak_err(got[r], want[r]) < 1e-9
        assert float(amax[r]) == np.abs(got[r]).max() and abs(got[r, int(arg[r])]) == float(amax[r])
    ob = np.abs(got)
    for q in range(nblk):
        lo, hi = max(blk_first + q * blk_len, 0), min(blk_first + (q + 1) * blk_len, L)
        if hi > lo:
            assert np.array_equal(bm[:, q].cpu().numpy(), ob[:, lo:hi].max(axis=1))
    g32 = est.plan(L, precision=32).execute(T.from_numpy(rec).cuda().float()).cpu().numpy().astype(np.float64)
    for r in range(3):
        assert 20 * np.log10(np.max(np.abs(g32[r] - want[r])) / np.max(np.abs(want[r]))) < -90.0


def p64_radix0(plan):
    return {768: 8, 1536: 12}[plan[0]]

"""Next-row components on the GPU against vectors from the unmodified reference
(tests/golden/make_golden.py::decay_chain): Schroeder decay times, reflection levels, decay-adjustment parameters
and virtual-bass synthesis."""
import numpy as np
import pytest

from _util import assert_ck, load, rel_peak_err

pytestmark = pytest.mark.gpu

RT_RTOL = 1e-9   # seconds, relative: regression over a curve that agrees to ~1e-13 dB
DB_TOL = 1e-9


@pytest.fixture(scope="module")
def G():
    return load('decay_chain_golden.npz')


def _build(stage):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from impulcifer_pip313_b200 import HRIR, ImpulseResponseEstimator, synth
    est = ImpulseResponseEstimator(min_duration=5.0, fs=48000)
    rec = synth.recording_set(est.test_signal, est.fs, synth.LAYOUT_71, seed=1234)
    h = HRIR(est)
    h.ingest(rec, 48000, ['FL', 'FR', 'FC', 'SL', 'SR', 'BL', 'BR'])
    h.crop_heads()
    if stage == 'heads':
        return h
    h.align_ipsilateral_all()
    h.align_onset_groups_peak_leftref()
    h.crop_tails()
    return h


@pytest.fixture(scope="module")
def final_set():
    return _build('final')


def _same_rt(got, want):
    got = np.array([np.nan if v is None else v for v in got], dtype=float)
    assert np.array_equal(np.isnan(got), np.isnan(want)), (got, want)
    ok = ~np.isnan(want)
    np.testing.assert_allclose(got[ok], want[ok], rtol=RT_RTOL, atol=0)


@pytest.mark.parametrize("stage", ["final", "heads"])
def test_decay_times_match_reference(G, final_set, stage):
    h = final_set if stage == 'final' else _build('heads')
    keys = [tuple(n.split('-')) for n in G['names']]
    assert [len(h.irs[s][e]) for s, e in keys] == list(G[f'{stage}_len'])
    # batched over the set (K4 + K17) ...
    rts = h.decay_times()
    for k, (s, e) in enumerate(keys):
        _same_rt(rts[s][e], G[f'{stage}_decay_times'][k])
    # ... and per response with the reference's own parameters handed in (the README writer's call)
    for k, (s, e) in enumerate(keys[:4]):
        pk, kn, wn = (int(v) for v in G[f'{stage}_params'][k])
        ir = h.irs[s][e]
        assert ir.decay_params() == (pk, kn, pytest.approx(float(G[f'{stage}_noise'][k]), abs=1e-9), wn)
        _same_rt(ir.decay_times(pk, kn, float(G[f'{stage}_noise'][k]), wn), G[f'{stage}_decay_times'][k])


def test_decay_times_degenerate_rows(final_set):
    from impulcifer_pip313_b200 import ImpulseResponse
    from impulcifer_pip313_b200.impulse_response import decay_times_batch
    x = final_set.irs['FL']['left'].data.copy()
    n = len(x)
    rows = [ImpulseResponse(x.copy(), 48000), ImpulseResponse(np.zeros(4096), 48000), ImpulseResponse(x[:64].copy(), 48000)]
    params = [(n - 1, n, -60.0, 1440),      # knee at the peak: empty Schroeder curve -> all undefined
              (0, 2048, -200.0, 1440),     # all-zero response
              (10, 40, 0.0, 1440)]         # noise floor too high for every target
    got = decay_times_batch(rows, params)
    assert got == [(None,) * 4] * 3


def test_reflection_levels(G, final_set):
    keys = [tuple(n.split('-')) for n in G['names']]
    r = final_set.calculate_reflection_levels()
    got = np.array([[r[s][e]['early_db'], r[s][e]['late_db']] for s, e in keys])
    np.testing.assert_allclose(got, G['reflection'], rtol=0, atol=DB_TOL)
    r2 = final_set.calculate_reflection_levels(direct_sound_duration_ms=1, early_ref_start_ms=5, early_ref_end_ms=300,
                                               late_ref_start_ms=900, late_ref_end_ms=2000)
    got2 = np.array([[r2[s][e]['early_db'], r2[s][e]['late_db']] for s, e in keys])
    np.testing.assert_allclose(got2, G['reflection_clipped'], rtol=0, atol=DB_TOL)


def test_decay_adjustment_params(G, final_set):
    keys = [tuple(n.split('-')) for n in G['names']]
    rows = iter(G['decay_adjustment'])
    for target in (0.05, 0.15, 2.0):
        for s, e in keys[:4]:
            want = next(rows)
            assert want[0] == target
            got = final_set.irs[s][e].decay_adjustment_params(target)
            if np.isnan(want[1]):
                assert got is None
            else:
                assert got[:3] == tuple(int(v) for v in want[1:4])
                assert got[3] == pytest.approx(want[4], rel=1e-9)


@pytest.mark.parametrize("tag,kw", [('default', {}),
                                    ('xo120_inv', dict(crossover_freq=120, invert_polarity=True, head_ms=0.5))])
def test_virtual_bass(G, final_set, tag, kw):
    from impulcifer_pip313_b200.virtual_bass import apply_virtual_bass_to_hrir
    hv = final_set.copy()
    hv.irs['BR']['right'].data = hv.irs['BR']['right'].data[:-1000]
    apply_virtual_bass_to_hrir(hv, **kw)
    keys = [tuple(n.split('-')) for n in G['names']]
    assert [len(hv.irs[s][e].data) for s, e in keys] == list(G[f'vbass_{tag}_len'])
    for k, (s, e) in enumerate(keys):
        assert_ck(hv.irs[s][e].data, G[f'vbass_{tag}_ck'][k], rel=1e-9, what=f'{s}-{e}')
    assert rel_peak_err(hv.irs['FL']['left'].data[:6000], G[f'vbass_{tag}_FL_left']) < 1e-9
    assert rel_peak_err(hv.irs['FC']['right'].data[:6000], G[f'vbass_{tag}_FC_right']) < 1e-9
    assert rel_peak_err(hv.irs['BR']['right'].data[-3000:], G[f'vbass_{tag}_BR_right_tail']) < 1e-9


def test_virtual_bass_rejects_crossover_above_nyquist(final_set):
    from impulcifer_pip313_b200.virtual_bass import synthesize_virtual_bass
    hv = final_set.copy()
    before = hv.irs['FL']['left'].data.copy()
    synthesize_virtual_bass(hv.irs, 48000, crossover_freq=24000)
    np.testing.assert_array_equal(hv.irs['FL']['left'].data, before)


def test_decay_window_kernel_matches_reference():
    """K15 against the reference's own ``apply_decay_window`` output (core/decay.py:383-403), fixed parameters."""
    import torch
    from impulcifer_pip313_b200 import ops
    g = load('decay_window_golden.npz')
    x = np.random.default_rng(0).standard_normal(20000)
    prm = g['fixed_params']
    xs = torch.from_numpy(np.tile(x, (len(prm), 1))).cuda()
    got = ops.window_resample(xs, None, 1, 1, decay=prm).cpu().numpy()
    for i in range(len(prm)):
        assert np.max(np.abs(got[i] - g[f'fixed_out_{i}'])) <= 1e-13 * np.max(np.abs(x)), i


def test_batched_decay_worker_matches_reference(final_set):
    """dispatcher.decay_adjust (what install() routes ``parallel_map(process_decay_worker, ...)`` to) against the
    reference's own worker (core/parallel_workers.py:24-39) on the synthetic 7.1 set: one batched call over all
    (response, target) tasks, results in input order, untouched rows returned as copies."""
    from impulcifer_pip313_b200.dispatcher import decay_adjust
    g = load('decay_window_golden.npz')
    keys = [tuple(n.split('-')) for n in g['names']]
    tasks, tags = [], []
    for target, k, _has in g['worker_rows']:
        s, e = keys[int(k)]
        tasks.append((s, e, final_set.irs[s][e].data, 48000, float(target)))
        tags.append(f't{int(target * 1000)}_{s}_{e}')
    out = decay_adjust(tasks)
    assert [(o[0], o[1]) for o in out] == [(t[0], t[1]) for t in tasks]
    for (s, e, y), task, tag in zip(out, tasks, tags):
        changed = not np.array_equal(y, task[2])
        assert changed == bool(g[f'worker_changed_{tag}']), tag
        assert y is not task[2]
        ref = g[f'worker_dec_{tag}']
        assert np.max(np.abs(y[::29] - ref)) <= 1e-9 * np.max(np.abs(ref)), tag
        assert_ck(y, g[f'worker_ck_{tag}'], what=tag)

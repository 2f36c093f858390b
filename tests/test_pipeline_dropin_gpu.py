"""Pipeline-level drop-in: the reference's own ``BRIRPipeline.run()`` (core/pipeline.py:416-503) with
``impulcifer_pip313_b200.install()`` active on the B200 against the same pipeline un-shimmed on the CPU.

Both arms are fresh interpreters over the staged, unmodified reference (``oracle/_ref``, ``oracle/make_ref.py``)
on the same synthetic measurement folder; the scenarios are the flag sets of the reference's own integrity
test (``tests/test_brir_integrity.py:39-61``) plus a room-correction and a microphone-deviation folder.
Compared: the set of files written, every WAV's sampling rate / frame count / channel layout (must be
identical) and the int32 PCM samples (byte-identical is the aim; the bound asserted is stated per scenario in
LSB of the 32-bit full scale, i.e. units of 4.7e-10, because the FP64 device FFTs round differently from
pocketfft at the 1e-16 level and a final ``lrint`` can flip on a tie).

When ``oracle/_ref`` did not travel, the 'default' scenario is still checked against the committed
fingerprint ``tests/golden/pipeline_default_golden.npz`` (made by ``tests/golden/make_pipeline_golden.py``).
"""
import os
import shutil

import numpy as np
import pytest

import _pipeline_case as pc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'oracle', '_ref')
HAVE_REF = os.path.isfile(os.path.join(REF, 'core', 'pipeline.py'))

# max |PCM_32 difference| allowed per scenario, in LSB (1 LSB = 2^-31 of full scale = 4.66e-10).
# Measured on the B200 (profiles/r02_dropin_report.md): every file written BEFORE a minimum-phase FIR is applied
# (responses.wav, headphone-responses.wav, room-responses.wav) is byte-identical (0 LSB; one 1-LSB tie in
# room-responses.wav).  The offending stage for everything after it is the minimum-phase FIR DESIGN
# (autoeq/frequency_response.py:672-680 -> scipy minimum_phase): the reference takes the log of its own rounding
# noise at the Nyquist bin, so its FIR is reproducible only to ~1e-8 (9 600 taps) / ~2e-7 (4 800 taps) of peak even
# between two machines running the reference itself (tests/_util.py MINPHASE_TOL_*, profiles/r02_minphase_localise.md);
# convolving with such a FIR moves samples by a few LSB: measured max 5 (EQ FIRs, 9 600 taps), 29 (channel-balance
# FIRs, 4 800 taps), 144 (microphone-deviation FIRs, 4 800 taps, applied before responses.wav).
LSB_BOUND = {
    'default': 16,
    'virtual_bass': 16,
    'no_headphone_compensation': 16,
    'room_correction': 16,
    'dsp_shaping': 96,
    'mic_deviation': 400,
    # --fs=44100 goes through nnresample, whose source is absent (parity unpinned): both arms run a
    # restatement (numpy port vs K9), agreement is to rounding of a 16 001-tap FP64 dot product (measured: 2 LSB)
    'resample_and_extra_outputs': 16,
}
# files that must be byte-identical (no FIR design upstream of them) in the scenarios that write them
BYTE_IDENTICAL = {'responses.wav': {'default', 'virtual_bass', 'dsp_shaping', 'no_headphone_compensation',
                                    'room_correction', 'resample_and_extra_outputs'},
                  'headphone-responses.wav': set(LSB_BOUND)}


def _sweep(fs_expected=48000):
    from scipy.io import wavfile
    fs, sw = wavfile.read(os.path.join(REF, 'data', pc.SWEEP_NAME))
    assert fs == fs_expected
    return fs, sw.astype(np.float64) / 2 ** 31


def _compare(ref_dir, got_dir, bound):
    ref = pc.read_outputs(ref_dir)
    got = pc.read_outputs(got_dir)
    assert sorted(ref) == sorted(got), (sorted(ref), sorted(got))
    assert ref, "the pipeline wrote no WAV"
    report = {}
    for name in sorted(ref):
        (fs_r, a), (fs_g, b) = ref[name], got[name]
        assert fs_r == fs_g, name
        assert a.shape == b.shape and a.dtype == b.dtype, (name, a.shape, b.shape, a.dtype, b.dtype)
        # identical channel layout: the same tracks are silent in both
        assert np.array_equal(np.any(a != 0, axis=0), np.any(b != 0, axis=0)), name
        d = np.abs(a.astype(np.int64) - b.astype(np.int64))
        report[name] = (int(d.max()) if d.size else 0, int(np.count_nonzero(d)), d.size)
    out_dir = os.path.join(ROOT, 'gpurun_out')
    if os.path.isdir(out_dir):  # evidence for profiles/: per-file LSB statistics of this run
        import json
        with open(os.path.join(out_dir, 'dropin_report.jsonl'), 'a') as fh:
            fh.write(json.dumps({'scenario': os.path.basename(os.path.dirname(ref_dir)), 'bound_lsb': bound,
                                 'files': {k: {'max_lsb': v[0], 'n_differ': v[1], 'n': v[2]} for k, v in report.items()}}) + '\n')
    scenario = os.path.basename(os.path.dirname(ref_dir))
    for name, v in report.items():
        assert v[0] <= bound, (name, v)
        if scenario in BYTE_IDENTICAL.get(name, ()):
            assert v[0] == 0, (name, v, "expected byte-identical PCM")
    return report


@pytest.mark.gpu
@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not staged (python oracle/make_ref.py)")
@pytest.mark.parametrize('scenario', sorted(pc.SCENARIOS))
def test_reference_pipeline_with_install_matches_unshimmed(tmp_path, scenario):
    fs, sweep = _sweep()
    kind, _ = pc.SCENARIOS[scenario]
    ref_dir, got_dir = str(tmp_path / scenario / 'ref'), str(tmp_path / scenario / 'b200')
    pc.write_folder(ref_dir, kind, sweep, fs)
    shutil.copytree(ref_dir, got_dir)
    pc.run_subprocess(ref_dir, scenario, shim=False)
    info = pc.run_subprocess(got_dir, scenario, shim=True)
    assert info['rebound'] >= 4
    assert info['hrir_cls'].startswith('impulcifer_pip313_b200') and info['est_cls'].startswith('impulcifer_pip313_b200')
    report = _compare(ref_dir, got_dir, LSB_BOUND[scenario])
    print(scenario, {k: f'max {v[0]} LSB, {v[1]}/{v[2]} differ' for k, v in report.items()})
    # README metrics (PNR / ITD / length / RT60 table, reflection levels) come from device-side decay analysis
    with open(os.path.join(ref_dir, 'README.md')) as f1, open(os.path.join(got_dir, 'README.md')) as f2:
        r1 = [ln for ln in f1.read().splitlines() if not ln.startswith('Processed on')]
        r2 = [ln for ln in f2.read().splitlines() if not ln.startswith('Processed on')]
    assert r1 == r2


@pytest.mark.gpu
def test_default_scenario_against_committed_fingerprint(tmp_path):
    path = os.path.join(os.path.dirname(__file__), 'golden', 'pipeline_default_golden.npz')
    if not os.path.isfile(path):
        pytest.skip("fingerprint not generated")
    if not HAVE_REF:
        pytest.skip("the shimmed arm still needs the reference's pipeline code (oracle/_ref)")
    g = np.load(path)
    fs, sweep = _sweep()
    got_dir = str(tmp_path / 'b200')
    pc.write_folder(got_dir, 'hp', sweep, fs)
    pc.run_subprocess(got_dir, 'default', shim=True)
    got = pc.read_outputs(got_dir)
    for name in ('hesuvi.wav', 'hrir.wav', 'responses.wav', 'headphone-responses.wav'):
        key = name.replace('.', '_').replace('-', '_')
        fs_g, b = got[name]
        assert fs_g == int(g[key + '_fs']) and list(b.shape) == list(g[key + '_shape'])
        step = int(g['decimation'])
        d = np.abs(b[::step].astype(np.int64) - g[key + '_dec'].astype(np.int64))
        assert d.max() <= LSB_BOUND['default'], (name, int(d.max()))

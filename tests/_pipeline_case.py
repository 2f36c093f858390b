"""Pipeline-level drop-in cases: the reference's OWN ``BRIRPipeline.run()`` (core/pipeline.py:416-503) on a
synthetic demo-shaped folder, as is (CPU) or with ``impulcifer_pip313_b200.install()`` active (B200).

Mirrors how the reference tests itself: the folder is synthesised like ``tests/test_pipeline_direct.py:175-204``
(sweep ⊛ synthetic room response per ear, written as PCM_32 WAV) and the scenarios are the flag sets of
``tests/test_brir_integrity.py:39-61`` (default, virtual bass, decay + channel balance + bass boost,
``--fs=44100`` with the extra output layouts, no headphone compensation) plus a room-correction folder.

Runs as a script so each arm gets a fresh interpreter (``install()`` rebinds module globals for good):

    python tests/_pipeline_case.py --dir D --scenario default [--shim]

TEST INFRASTRUCTURE: the un-shimmed arm imports the reference through ``oracle/ref_import.py``.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

SCENARIOS = {
    # name -> (folder kind, ProcessingConfig overrides)
    'default': ('hp', {}),
    'virtual_bass': ('hp', {'vbass': True, 'vbass_freq': 250}),
    'dsp_shaping': ('hp', {'decay': 100.0, 'channel_balance': 'trend', 'bass_boost_gain': 4.0}),
    'resample_and_extra_outputs': ('hp', {'fs': 44100, 'output_truehd_layouts': True, 'jamesdsp': True,
                                          'hangloose': True}),
    'no_headphone_compensation': ('hp', {'do_headphone_compensation': False}),
    'room_correction': ('room', {}),
    'mic_deviation': ('hp', {'do_headphone_compensation': False, 'microphone_deviation_correction': True}),
}

SPEAKERS_71 = ['FL', 'FR', 'FC', 'SL', 'SR', 'BL', 'BR']
SWEEP_NAME = 'sweep-6.15s-48000Hz-32bit-2.93Hz-24000Hz.wav'


def _write_pcm32(path, fs, tracks):
    """tracks [n_tracks, n] float in -1..1 -> interleaved PCM_32 WAV (container only, no reference code)."""
    from scipy.io import wavfile
    q = np.clip(np.rint(np.asarray(tracks, dtype=np.float64) * 2.0 ** 31), -2.0 ** 31, 2.0 ** 31 - 1).astype(np.int32)
    wavfile.write(path, int(fs), np.ascontiguousarray(q.T))


def write_folder(path, kind, test_signal, fs, speakers=None, seed=4242):
    """Synthesise the measurement folder: ``<speakers>.wav`` (2 ear tracks, one column per speaker),
    ``headphones.wav`` (FL,FR layout) and, for kind 'room', single-track ``room-<spk>-<side>.wav`` + target."""
    sys.path.insert(0, ROOT)
    from impulcifer_pip313_b200 import synth
    speakers = speakers or SPEAKERS_71
    os.makedirs(path, exist_ok=True)
    rec = synth.recording_set(test_signal, fs, speakers, seed=seed)
    _write_pcm32(os.path.join(path, ','.join(speakers) + '.wav'), fs, rec)
    hp = synth.recording_set(test_signal, fs, ['FL', 'FR'], seed=seed + 1, level=0.2)
    # headphone measurement: left driver -> left ear only, right driver -> right ear only
    sil = int(2.0 * fs)
    col = sil + len(test_signal)
    hp[1, sil:sil + col] *= 1e-3
    hp[0, sil + col:sil + 2 * col] *= 1e-3
    _write_pcm32(os.path.join(path, 'headphones.wav'), fs, hp)
    if kind == 'room':
        for k, spk in enumerate(speakers[:3]):
            for side_i, side in enumerate(('left', 'right')):
                r = synth.recording_set(test_signal, fs, [spk], seed=seed + 10 + 2 * k + side_i, level=0.2)
                _write_pcm32(os.path.join(path, f'room-{spk}-{side}.wav'), fs, r[side_i:side_i + 1])
        f = np.array([20.0, 50, 100, 200, 500, 1000, 2000, 5000, 10000, 20000])
        raw = np.array([4.0, 3.5, 2.0, 1.0, 0.3, 0.0, -0.5, -1.5, -3.0, -6.0])
        with open(os.path.join(path, 'room-target.csv'), 'w') as fh:
            fh.write('frequency,raw\n')
            for a, b in zip(f, raw):
                fh.write(f'{a:.2f},{b:.2f}\n')


def run_inprocess(dir_path, scenario, shim):
    """Load the reference (stubs for plotting / WAV container only), optionally install() the B200 path, run."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    sys.path.insert(0, ROOT)
    import ref_import
    ref_import.load()
    info = {}
    if shim:
        import impulcifer_pip313_b200 as b200
        got = b200.install()
        info['rebound'] = got['rebound']
    from core.pipeline import BRIRPipeline, ProcessingConfig
    kind, overrides = SCENARIOS[scenario]
    cfg = dict(dir_path=dir_path, test_signal='default', plot=False)
    cfg.update(overrides)
    if 'decay' in cfg and not isinstance(cfg['decay'], dict):
        from core.constants import SPEAKER_NAMES
        cfg['decay'] = {ch: float(cfg['decay']) / 1000 for ch in SPEAKER_NAMES}
    pipeline = BRIRPipeline(ProcessingConfig(**cfg))
    if shim:
        # evidence that the B200 classes are the ones the reference pipeline instantiates
        import core.hrir
        import core.impulse_response_estimator as e
        info['hrir_cls'] = core.hrir.HRIR.__mro__[1].__module__
        info['est_cls'] = e.ImpulseResponseEstimator.__mro__[1].__module__
    pipeline.run()
    return info


def run_subprocess(dir_path, scenario, shim, timeout=900):
    cmd = [sys.executable, os.path.abspath(__file__), '--dir', dir_path, '--scenario', scenario]
    if shim:
        cmd.append('--shim')
    env = dict(os.environ)
    env.setdefault('OMP_NUM_THREADS', '4')
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    if p.returncode != 0:
        raise RuntimeError(f"pipeline arm failed (shim={shim}, scenario={scenario}):\n{p.stdout[-3000:]}\n{p.stderr[-6000:]}")
    for line in reversed(p.stdout.splitlines()):
        if line.startswith('PIPELINE_CASE_INFO '):
            return json.loads(line[len('PIPELINE_CASE_INFO '):])
    return {}


def read_outputs(dir_path):
    """{relative file name: (fs, int array [frames, tracks])} for every WAV the pipeline wrote."""
    from scipy.io import wavfile
    inputs = {'headphones.wav'}
    out = {}
    for base, _dirs, files in os.walk(dir_path):
        for name in sorted(files):
            rel = os.path.relpath(os.path.join(base, name), dir_path)
            if not name.endswith('.wav') or rel in inputs or rel.startswith('room-') and 'responses' not in rel:
                continue
            if ',' in name or name.split('.')[0] in SPEAKERS_71 and os.sep not in rel:
                continue  # the measurement files themselves
            fs, d = wavfile.read(os.path.join(base, name))
            out[rel] = (fs, d)
    return out


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--dir', required=True)
    ap.add_argument('--scenario', required=True, choices=sorted(SCENARIOS))
    ap.add_argument('--shim', action='store_true')
    a = ap.parse_args()
    info = run_inprocess(a.dir, a.scenario, a.shim)
    print('PIPELINE_CASE_INFO ' + json.dumps(info))

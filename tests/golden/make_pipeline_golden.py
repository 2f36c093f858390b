#!/usr/bin/env python
"""Fingerprint of what the UNMODIFIED reference pipeline writes for the 'default' scenario of
``tests/_pipeline_case.py`` (run in the build container; needs ``oracle/_ref`` or /root/reference).

    python tests/golden/make_pipeline_golden.py

Stores, per output WAV, the sampling rate, the [frames, tracks] shape and every 16th frame of the int32 PCM
(the full files are 5-13 MB each; the live two-arm comparison in ``tests/test_pipeline_dropin_gpu.py`` covers
every sample when ``oracle/_ref`` travels to the GPU box).
"""
import os
import sys
import tempfile

import numpy as np
from scipy.io import wavfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import _pipeline_case as pc  # noqa: E402
import ref_import  # noqa: E402

DECIMATION = 16


def main():
    fs, sw = wavfile.read(os.path.join(ref_import.REFERENCE_ROOT, 'data', pc.SWEEP_NAME))
    sw = sw.astype(np.float64) / 2 ** 31
    out = {'decimation': np.array(DECIMATION)}
    with tempfile.TemporaryDirectory() as td:
        pc.write_folder(td, 'hp', sw, fs)
        pc.run_subprocess(td, 'default', shim=False)
        for name, (fs_o, d) in pc.read_outputs(td).items():
            key = name.replace('.', '_').replace('-', '_')
            out[key + '_fs'] = np.array(fs_o)
            out[key + '_shape'] = np.array(d.shape)
            out[key + '_dec'] = d[::DECIMATION].copy()
            print(name, fs_o, d.shape)
    np.savez_compressed(os.path.join(HERE, 'pipeline_default_golden.npz'), **out)


if __name__ == '__main__':
    main()

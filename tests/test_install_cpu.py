"""install(): the merged classes sit in front of the reference's in the MRO and replace its bindings.
Runs only where the reference checkout exists (the build container); needs no GPU."""
import sys

import pytest

import ref_import


@pytest.mark.skipif(not ref_import.available(), reason="reference checkout not present")
def test_install_rebinds_reference_classes():
    ref_import.load()
    from core import hrir as ref_hrir
    from core import room_correction  # noqa: F401  (an importer that must be re-bound)
    old_hrir = ref_hrir.HRIR
    import impulcifer_pip313_b200 as b200
    from impulcifer_pip313_b200 import hrir as shim_hrir
    got = b200.install()
    assert got['rebound'] >= 4
    from core import hrir as h2, impulse_response as ir2, impulse_response_estimator as e2
    from autoeq import frequency_response as fr2
    assert h2.HRIR is got['HRIR'] and issubclass(h2.HRIR, old_hrir) and issubclass(h2.HRIR, shim_hrir.HRIR)
    # hot methods come from the shim, everything else from the reference
    assert h2.HRIR.crop_heads is shim_hrir.HRIR.crop_heads
    assert h2.HRIR.correct_channel_balance is shim_hrir.HRIR.correct_channel_balance
    assert h2.HRIR.plot is old_hrir.plot
    from core import room_correction as rc2
    assert rc2.calculate_specific_room_corrections.__module__.startswith('impulcifer_pip313_b200')
    assert rc2._correction_limit_mask.__module__.startswith('impulcifer_pip313_b200')
    assert ir2.ImpulseResponse.peak_index.__module__.startswith('impulcifer_pip313_b200')
    assert hasattr(ir2.ImpulseResponse, 'plot_fr')  # plotting mix-in of the reference still there
    assert fr2.FrequencyResponse.minimum_phase_impulse_response.__module__.startswith('impulcifer_pip313_b200')
    assert hasattr(fr2.FrequencyResponse, 'optimize_parametric_eq') or hasattr(fr2.FrequencyResponse, 'write_to_csv')
    est = e2.ImpulseResponseEstimator(min_duration=0.05, fs=2000)
    assert est.estimate.__func__.__module__.startswith('impulcifer_pip313_b200')
    assert hasattr(est, 'sweep_sequence')  # reference-only method reachable through the merged class
    assert sys.modules['core.room_correction'].__dict__.get('HRIR', got['HRIR']) is got['HRIR']
    # the EQ process pool is routed to the batched device path, everything else keeps the reference pool
    from core import parallel_utils
    assert parallel_utils.parallel_map is got['parallel_map']
    assert parallel_utils.parallel_map(abs, [-1, -2, 3], use_threads=True) == [1, 2, 3]


@pytest.mark.skipif(not ref_import.available(), reason="reference checkout not present")
def test_copies_keep_the_merged_classes():
    """ir.copy() / fr.copy() / HRIR.copy() after install() must return the MERGED classes (reference mix-ins such as
    plot_fr / write_to_csv and the overrides install() adds), not the bare shim classes."""
    import numpy as np
    ref_import.load()
    import impulcifer_pip313_b200 as b200
    got = b200.install()
    IR, FR = got['ImpulseResponse'], got['FrequencyResponse']
    ir = IR(np.linspace(0.0, 1.0, 32), 48000)
    c = ir.copy()
    assert type(c) is IR and hasattr(c, 'plot_fr') and np.array_equal(c.data, ir.data) and c.data is not ir.data
    f = np.geomspace(20.0, 20000.0, 64)
    fr = FR(name='t', frequency=f, raw=np.zeros(64))
    fc = fr.copy()
    assert type(fc) is FR and fc.name == 't_copy' and np.array_equal(fc.frequency, fr.frequency)


def test_reference_staging_recipe_is_a_plain_copy(tmp_path):
    """oracle/make_ref.py stages the reference unmodified: every staged Python file is byte-identical to its source."""
    import filecmp
    import os
    src = os.environ.get("IMPULCIFER_REFERENCE_SRC", "/root/reference")
    staged = os.path.join(os.path.dirname(os.path.abspath(ref_import.__file__)), "_ref")
    if not (os.path.isdir(os.path.join(src, "core")) and os.path.isdir(os.path.join(staged, "core"))):
        pytest.skip("needs both the reference checkout and the staged copy (build container only)")
    checked = 0
    for pkg in ("core", "autoeq", "infra"):
        for base, _dirs, files in os.walk(os.path.join(staged, pkg)):
            for name in files:
                if name.endswith(".py"):
                    rel = os.path.relpath(os.path.join(base, name), staged)
                    assert filecmp.cmp(os.path.join(staged, rel), os.path.join(src, rel), shallow=False), rel
                    checked += 1
    assert checked > 40

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

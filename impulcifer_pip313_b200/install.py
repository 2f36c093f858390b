"""Drop the B200 hot path under an unmodified checkout of the reference (115dkk/Impulcifer-pip313).

    import impulcifer_pip313_b200 as b200
    b200.install()                # before the pipeline modules are imported (or after: importers are re-bound)
    import impulcifer; impulcifer.main(dir_path=..., test_signal='default')

``install()`` builds three classes that put this package's device-backed methods IN FRONT of the reference's
own classes in the MRO, so everything outside the hot path (plotting mix-ins, README metrics, CSV I/O ...)
keeps running the reference's code on top of them:

    ImpulseResponseEstimator  <- estimator.ImpulseResponseEstimator , core.impulse_response_estimator.*
    ImpulseResponse           <- impulse_response.ImpulseResponse   , core.impulse_response.ImpulseResponse
    HRIR                      <- hrir.HRIR                          , core.hrir.HRIR
    FrequencyResponse         <- frequency_response.FrequencyResponse, autoeq.frequency_response.FrequencyResponse

and rebinds every already-imported module attribute that still points at the reference classes.  The EQ stage's
``parallel_map(process_equalization_worker, ...)`` call (core/pipeline.py:674-688) is routed to
``dispatcher.equalization_firs`` and the decay stage's ``parallel_map(process_decay_worker, ...)``
(core/pipeline.py:694-716) to ``dispatcher.decay_adjust``; every other ``parallel_map`` use keeps the reference's pools.  The module-level
room-correction, microphone-mismatch, decay-time and virtual-bass functions are replaced by their batched twins
(``room_correction.py``, ``microphone_deviation.py``, ``impulse_response.decay_times_batch``, ``virtual_bass.py``).
Nothing in the reference tree is edited.  See INTEGRATION.md.
"""
from __future__ import annotations

import sys

_installed = {}


def install(verbose: bool = False):
    if _installed:
        return _installed
    from core import hrir as ref_hrir
    from core import impulse_response as ref_ir
    from core import impulse_response_estimator as ref_est
    from autoeq import frequency_response as ref_fr

    from . import estimator, frequency_response, hrir, impulse_response

    RefEst, RefIR, RefHRIR, RefFR = (ref_est.ImpulseResponseEstimator, ref_ir.ImpulseResponse, ref_hrir.HRIR,
                                     ref_fr.FrequencyResponse)

    class FrequencyResponse(frequency_response.FrequencyResponse, RefFR):
        __doc__ = RefFR.__doc__

        # argument variants outside the hot path (bass-shelf target synthesis, cubic interpolation) stay the reference's
        def compensate(self, compensation, bass_boost_gain=0.0, *args, **kwargs):
            if bass_boost_gain:
                return RefFR.compensate(self, compensation, bass_boost_gain, *args, **kwargs)
            return frequency_response.FrequencyResponse.compensate(self, compensation, bass_boost_gain, *args, **kwargs)

        def interpolate(self, f=None, f_step=frequency_response.DEFAULT_STEP, pol_order=1, **kwargs):
            if pol_order != 1:
                return RefFR.interpolate(self, f=f, f_step=f_step, pol_order=pol_order, **kwargs)
            return frequency_response.FrequencyResponse.interpolate(self, f=f, f_step=f_step, pol_order=1, **kwargs)

    class ImpulseResponse(impulse_response.ImpulseResponse, RefIR):
        __doc__ = RefIR.__doc__

        def frequency_response(self):
            fr = impulse_response.ImpulseResponse.frequency_response(self)
            return FrequencyResponse(name=fr.name, frequency=fr.frequency, raw=fr.raw)

    class HRIR(hrir.HRIR, RefHRIR):
        __doc__ = RefHRIR.__doc__

    class ImpulseResponseEstimator(estimator.ImpulseResponseEstimator, RefEst):
        __doc__ = RefEst.__doc__

    # the shim modules construct these classes themselves: point them at the merged ones
    from . import microphone_deviation, room_correction
    hrir.ImpulseResponse = ImpulseResponse
    impulse_response.FrequencyResponse = FrequencyResponse
    for shim_mod in (hrir, room_correction, microphone_deviation):
        shim_mod.FrequencyResponse = FrequencyResponse

    mapping = {RefEst: ImpulseResponseEstimator, RefIR: ImpulseResponse, RefHRIR: HRIR, RefFR: FrequencyResponse}

    # next-row components: the batched room / microphone chains replace the module-level
    # functions of core/room_correction.py:185-306 and core/microphone_deviation_correction.py
    from core import microphone_deviation_correction as ref_md
    from core import room_correction as ref_rc

    def _calculate_generic_room_correction(irs, target, mic_calibration=None, method='average', limit=1000,
                                           collect_raws=False):
        got = room_correction.calculate_generic_room_correction(irs, target, mic_calibration=mic_calibration,
                                                                method=method, limit=limit, collect_raws=collect_raws)
        return got if collect_raws else (got, [])

    class MicrophoneMatchingCorrector(microphone_deviation.MicrophoneMatchingCorrector,
                                      ref_md.MicrophoneMatchingCorrector):
        __doc__ = ref_md.MicrophoneMatchingCorrector.__doc__

    # rank 4: decay analysis and virtual bass
    from core import decay as ref_decay
    from core import virtual_bass as ref_vb
    from . import virtual_bass

    def decay_times(data, fs, peak_ind=None, knee_point_ind=None, noise_floor=None, window_size=None):
        return ImpulseResponse(data, fs).decay_times(peak_ind, knee_point_ind, noise_floor, window_size)

    def decay_adjustment_params(data, fs, target):
        return ImpulseResponse(data, fs).decay_adjustment_params(target)

    decay_times.__doc__ = ref_decay.decay_times.__doc__
    decay_adjustment_params.__doc__ = ref_decay.decay_adjustment_params.__doc__
    mapping.update({
        ref_decay.decay_times: decay_times,
        ref_decay.decay_adjustment_params: decay_adjustment_params,
        ref_vb.synthesize_virtual_bass: virtual_bass.synthesize_virtual_bass,
        ref_vb.apply_virtual_bass_to_hrir: virtual_bass.apply_virtual_bass_to_hrir,
    })
    mapping.update({
        ref_rc.calculate_specific_room_corrections: room_correction.calculate_specific_room_corrections,
        ref_rc._calculate_generic_room_correction: _calculate_generic_room_correction,
        ref_rc._correction_limit_mask: room_correction.correction_limit_mask,
        ref_md.MicrophoneMatchingCorrector: MicrophoneMatchingCorrector,
        ref_md.apply_microphone_deviation_correction_to_hrir:
            microphone_deviation.apply_microphone_deviation_correction_to_hrir,
    })
    rebound = 0
    for mod in list(sys.modules.values()):
        d = getattr(mod, '__dict__', None)
        if not d or getattr(mod, '__name__', '').startswith('impulcifer_pip313_b200'):
            continue
        for k, v in list(d.items()):
            if not (isinstance(v, type) or callable(v)):
                continue
            try:
                new = mapping.get(v)
            except TypeError:  # unhashable callable
                continue
            if new is not None:
                d[k] = new
                rebound += 1
    # the process pools of the EQ and decay stages (core/pipeline.py:674-688, 694-716) -> batched launches
    from core import parallel_utils as ref_pu
    pool_map = ref_pu.parallel_map

    def parallel_map(func, items, max_workers=None, initializer=None, initargs=(), **kw):
        name = getattr(func, '__name__', '')
        if name == 'process_equalization_worker' and len(initargs) == 8:
            from .dispatcher import equalization_firs
            return equalization_firs(list(items), *initargs)
        if name == 'process_decay_worker':  # core/pipeline.py:694-716 -> K4 + K17 + K15, one launch each
            from .dispatcher import decay_adjust
            return decay_adjust(list(items))
        return pool_map(func, items, max_workers, initializer, initargs, **kw)

    parallel_map.__doc__ = pool_map.__doc__
    parallel_map.__wrapped__ = pool_map
    for mod in list(sys.modules.values()):
        d = getattr(mod, '__dict__', None)
        if d and d.get('parallel_map') is pool_map:
            d['parallel_map'] = parallel_map
            rebound += 1
    _installed.update(parallel_map=parallel_map)
    _installed.update(ImpulseResponseEstimator=ImpulseResponseEstimator, ImpulseResponse=ImpulseResponse, HRIR=HRIR,
                      FrequencyResponse=FrequencyResponse, rebound=rebound)
    if verbose:
        print(f"impulcifer_pip313_b200: installed over the reference ({rebound} bindings replaced)")
    return _installed


# ``import impulcifer_pip313_b200 as b200; b200.install()`` must keep working after this submodule has been imported:
# the import machinery binds the submodule over the package attribute of the same name, so the module itself is callable.
import types as _types


class _CallableModule(_types.ModuleType):
    def __call__(self, *args, **kwargs):
        return install(*args, **kwargs)


sys.modules[__name__].__class__ = _CallableModule

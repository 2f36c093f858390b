"""impulcifer_pip313_b200 — B200-native hot path of 115dkk/Impulcifer-pip313.

Sweep deconvolution, crop / align / resample and per-IR frequency-response post-processing as hand-written
sm_100a CUDA kernels behind the reference's own Python class API.  See DESIGN.md and INTEGRATION.md.
Importing this package does not touch CUDA; the kernels load on first use and there is no CPU fallback.
"""
__version__ = "0.1.0"

__all__ = ["ImpulseResponseEstimator", "ImpulseResponse", "HRIR", "FrequencyResponse", "BatchDispatcher",
           "DeviceShard", "install"]


def __getattr__(name):
    if name == "ImpulseResponseEstimator":
        from .estimator import ImpulseResponseEstimator
        return ImpulseResponseEstimator
    if name == "ImpulseResponse":
        from .impulse_response import ImpulseResponse
        return ImpulseResponse
    if name == "HRIR":
        from .hrir import HRIR
        return HRIR
    if name == "FrequencyResponse":
        from .frequency_response import FrequencyResponse
        return FrequencyResponse
    if name in ("BatchDispatcher", "DeviceShard"):
        from . import dispatcher
        return getattr(dispatcher, name)
    if name == "install":
        from .install import install
        return install
    raise AttributeError(name)

"""TEST INFRASTRUCTURE ONLY — loader for the *unmodified* reference (115dkk/Impulcifer-pip313).

Imports the reference's own hot-path modules from the staged copy ``oracle/_ref`` (``oracle/make_ref.py``;
git-ignored, travels to the GPU box) or, failing that, from ``/root/reference`` (build container only), so
that golden vectors can be generated from the real implementation (``tests/golden/make_golden.py``), the
oracle restatement in ``oracle/impulcifer_oracle.py`` can be pinned against it, the pipeline-level drop-in
test can run the reference's own ``BRIRPipeline`` and ``bench.py --impl reference`` can time it.

The product package never imports this module (``tests/test_abi_cpu.py`` greps for it).

Only plotting / audio-device / WAV-container modules are stubbed.  No arithmetic is restated here:
  * ``matplotlib*``, ``seaborn``, ``bokeh*``, ``sounddevice``  -> MagicMock (imported at module top by
    core/impulse_response_estimator.py:10, core/plotting/*.py)
  * ``soundfile`` -> 20-line shim over ``scipy.io.wavfile`` (PCM_32 -> float64 via /2**31, exactly what
    libsndfile does for ``dtype='float64'``); write() quantises with ``np.round(x * 2**31)`` + clip
    (PCM_16: the top 16 bits of that value, as libsndfile's clipping converters do).
  * ``nnresample`` -> the restatement in ``oracle/nnresample_port.py`` (package absent from the image;
    parity for that one step is therefore UNPINNED).
"""
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_STAGED = os.path.join(_HERE, "_ref")          # oracle/make_ref.py: unmodified copy that travels to the GPU box


def _default_root():
    env = os.environ.get("IMPULCIFER_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile(os.path.join(_STAGED, "core", "pipeline.py")):
        return _STAGED
    return "/root/reference"


REFERENCE_ROOT = _default_root()

_STUBS = [
    'matplotlib', 'matplotlib.pyplot', 'matplotlib.ticker', 'matplotlib.font_manager', 'matplotlib.figure',
    'matplotlib.axes', 'matplotlib.colors', 'matplotlib.cm', 'matplotlib.mlab', 'matplotlib.gridspec',
    'mpl_toolkits', 'mpl_toolkits.mplot3d', 'mpl_toolkits.axes_grid1', 'seaborn', 'bokeh', 'bokeh.plotting',
    'bokeh.models', 'bokeh.layouts', 'bokeh.palettes', 'bokeh.io', 'bokeh.embed', 'bokeh.resources',
    'sounddevice',
]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "core"))


def _soundfile_shim():
    sf = types.ModuleType('soundfile')

    def _read(path, dtype='float64', always_2d=False, **kw):
        from scipy.io import wavfile
        fs, d = wavfile.read(path)
        if d.dtype == np.int32:
            d = d.astype(np.float64) / 2 ** 31
        elif d.dtype == np.int16:
            d = d.astype(np.float64) / 2 ** 15
        else:
            d = d.astype(np.float64)
        if always_2d and d.ndim == 1:
            d = d[:, None]
        return d, fs

    def _write(path, data, samplerate=48000, subtype="PCM_32", **kw):
        from scipy.io import wavfile
        data = np.asarray(data)
        if data.dtype.kind == 'i':
            # integer frames are already PCM (libsndfile i2lei / s2les / i2les: int32 -> PCM_16 keeps the top 16 bits)
            if subtype == "PCM_32":
                q = data.astype(np.int32) if data.dtype == np.int32 else (data.astype(np.int32) << 16)
            elif subtype == "PCM_16":
                q = data.astype(np.int16) if data.dtype == np.int16 else (data.astype(np.int64) >> 16).astype(np.int16)
            else:
                raise ValueError("shim supports PCM_16/PCM_32 only")
            wavfile.write(path, int(samplerate), q)
            return
        data = data.astype(np.float64)
        if subtype == "PCM_32":
            q = np.clip(np.round(data * 2 ** 31), -2 ** 31, 2 ** 31 - 1).astype(np.int32)
        elif subtype == "PCM_16":
            # libsndfile d2les_clip_array (python-soundfile turns SFC_SET_CLIPPING on): scale to 2^31, clip,
            # lrint, then keep the top 16 bits (arithmetic shift = floor), NOT a round at 16-bit depth
            q32 = np.clip(np.round(data * 2 ** 31), -2 ** 31, 2 ** 31 - 1).astype(np.int64)
            q = (q32 >> 16).astype(np.int16)
        else:
            raise ValueError("shim supports PCM_16/PCM_32 only")
        wavfile.write(path, int(samplerate), q)

    class _Info:
        def __init__(self, path):
            from scipy.io import wavfile
            fs, d = wavfile.read(path, mmap=True)
            self.samplerate = fs
            self.frames = d.shape[0]
            self.channels = 1 if d.ndim == 1 else d.shape[1]
            self.subtype = {np.dtype('int32'): 'PCM_32', np.dtype('int16'): 'PCM_16'}.get(d.dtype, 'FLOAT')

    sf.read, sf.write, sf.info = _read, _write, _Info
    return sf


def _tiny_png(path, *a, **k):
    from PIL import Image
    Image.new('RGB', (2, 2), (255, 255, 255)).save(path)


def _plot_stubs():
    """Plotting only: enough of pyplot for the unconditional ``_stage_plot_results`` (core/pipeline.py:783-790,
    core/plotting/hrir_plotter.py:324-398: ``fig, ax = plt.subplots()``, ``fig.savefig`` then PIL re-opens the PNG)."""
    plt = sys.modules['matplotlib.pyplot']
    if not isinstance(plt, MagicMock) or plt.__dict__.get('_impb_stub'):
        return

    def subplots(*a, **k):
        fig, ax = MagicMock(), MagicMock()
        fig.savefig = _tiny_png
        return fig, ax

    plt.subplots = subplots
    plt._impb_stub = True
    sys.modules['matplotlib'].pyplot = plt


def _patch_plotting_utils():
    """``sync_axes`` reduces axis limits read back from real Axes; ``save_fig_as_png`` rasterises. No arithmetic
    of the path is in either (core/plotting_utils.py:19-46, 122-128)."""
    if not isinstance(sys.modules.get('matplotlib'), MagicMock):
        return
    import core.plotting_utils as pu
    pu.sync_axes = lambda *a, **k: None
    pu.save_fig_as_png = lambda file_path, fig, *a, **k: _tiny_png(file_path)


_loaded = False


def load():
    """Make ``core.*`` and ``autoeq.*`` of the reference importable. Idempotent."""
    global _loaded
    if _loaded:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for m in _STUBS:
        sys.modules.setdefault(m, MagicMock())
    _plot_stubs()
    sys.modules.setdefault('soundfile', _soundfile_shim())
    if 'nnresample' not in sys.modules:
        here = os.path.dirname(os.path.abspath(__file__))
        if here not in sys.path:
            sys.path.insert(0, here)
        import nnresample_port
        sys.modules['nnresample'] = nnresample_port
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _patch_plotting_utils()
    _loaded = True


def module(name):
    """Import one more module of the reference by dotted name (e.g. ``core.room_correction``)."""
    import importlib
    load()
    return importlib.import_module(name)


def modules():
    """Return the reference's hot-path modules as a namespace."""
    load()
    ns = types.SimpleNamespace()
    from core import impulse_response_estimator, impulse_response, hrir, decay, constants, audio_io
    from core import parallel_utils, parallel_processing, parallel_workers
    from autoeq import frequency_response
    ns.impulse_response_estimator = impulse_response_estimator
    ns.impulse_response = impulse_response
    ns.hrir = hrir
    ns.decay = decay
    ns.constants = constants
    ns.audio_io = audio_io
    ns.parallel_utils = parallel_utils
    ns.parallel_processing = parallel_processing
    ns.parallel_workers = parallel_workers
    ns.frequency_response = frequency_response
    return ns

"""One pass over every device kernel of the hot path for an ncu capture (developer tool, not the bench):
BatchDispatcher.process + postprocess + HRIR-level extras on synthetic 7.1 sets, inside a cudaProfilerStart/Stop
range.  Run as

  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none --csv --log-file gpurun_out/all_kernels.csv python tools/all_kernels.py [sets]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from impulcifer_pip313_b200 import BatchDispatcher, HRIR, ImpulseResponseEstimator, synth
from impulcifer_pip313_b200.impulse_response import decay_times_batch

S = int(sys.argv[1]) if len(sys.argv) > 1 else 32
est = ImpulseResponseEstimator(5.0, 48000)
base = synth.recording_set(est.test_signal, 48000, synth.LAYOUT_71, seed=1234)
batch = torch.from_numpy(base).cuda()[None].repeat(S, 1, 1).contiguous()
disp = BatchDispatcher(est, synth.LAYOUT_71)


def next_rows(h):
    """room correction, headphone compensation on an uncropped (non-5-smooth) length,
    channel balance, virtual bass."""
    from impulcifer_pip313_b200 import FrequencyResponse
    from impulcifer_pip313_b200 import room_correction as rc
    from impulcifer_pip313_b200.virtual_bass import apply_virtual_bass_to_hrir
    grid = FrequencyResponse.generate_frequencies(f_step=1.01, f_min=10, f_max=24000)
    lf = np.log10(grid)
    target = FrequencyResponse(name='t', frequency=grid.copy(), raw=-1.2 * (lf - 3.0))
    rc.calculate_specific_room_corrections(h.copy(), target, mic_calibration=None, limit=400)
    raw = HRIR(est)
    raw.ingest(base, 48000, synth.LAYOUT_71)
    rc.headphone_compensation(raw)
    hb = h.copy()
    hb.correct_channel_balance('avg')
    apply_virtual_bass_to_hrir(h.copy())


def one_pass():
    res = disp.process(batch)
    post = disp.postprocess(res)
    h = HRIR(est)
    h.ingest(base, 48000, synth.LAYOUT_71)
    h.crop_heads()
    h.align_ipsilateral_all()
    h.align_onset_groups_peak_leftref()
    h.crop_tails()
    h.decay_times()
    h.calculate_reflection_levels()
    next_rows(h)
    h.normalize()
    h.resample(44100)
    h.pcm_frames()
    return res, post


one_pass()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
one_pass()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")

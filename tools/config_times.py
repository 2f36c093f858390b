"""Wall time of the named BASELINE configs 3 and 4 on one GPU (developer tool; synthetic inputs of the named shapes).

config 3: 192 kHz 7.1 recording set -> deconvolution + crop chain (HRIR) -> resample to 48 kHz -> 1/3-octave smoothing
          -> 9 600-tap minimum-phase FIR per response
config 4: TrueHD 11-channel set (22 responses) -> dispatcher -> specific room corrections + headphone compensation
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from impulcifer_pip313_b200 import BatchDispatcher, FrequencyResponse, HRIR, ImpulseResponseEstimator, synth
from impulcifer_pip313_b200 import room_correction as rc
from impulcifer_pip313_b200.frequency_response import minimum_phase_batch, smoothen_batch
from impulcifer_pip313_b200.impulse_response import frequency_response_batch


def best(fn, reps=4):
    fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts)


# ---- config 3 ---------------------------------------------------------------------------------------------
est = ImpulseResponseEstimator(5.0, 192000)
rec = synth.recording_set(est.test_signal, 192000, synth.LAYOUT_71, seed=7)
grid = FrequencyResponse.generate_frequencies(f_step=1.01, f_min=10, f_max=24000)


def config3():
    h = HRIR(est)
    h.ingest(rec, 192000, synth.LAYOUT_71)
    h.crop_heads()
    h.align_ipsilateral_all()
    h.align_onset_groups_peak_leftref()
    h.crop_tails()
    h.resample(48000)
    rows = torch.stack([ir.device_data() for _, _, ir in h._all()])
    raw = frequency_response_batch(rows, 48000)
    sm = smoothen_batch(grid, raw, 1 / 3, 1, 1 / 3, 1, 100.0, 10000.0)
    fir = minimum_phase_batch(grid, -(sm - sm.mean(dim=1, keepdim=True)), fs=48000, f_res=5, normalize=False)
    return fir.shape, rows.shape


print('config 3 (192 kHz 7.1 set, 14 responses):', config3(), f'{best(config3):.1f} ms')

# ---- config 4 ---------------------------------------------------------------------------------------------
est48 = ImpulseResponseEstimator(5.0, 48000)
rec11 = torch.from_numpy(synth.recording_set(est48.test_signal, 48000, synth.LAYOUT_TRUEHD_11, seed=5)[None]).cuda()
hp = synth.recording_set(est48.test_signal, 48000, ['FL', 'FR'], seed=6)
disp = BatchDispatcher(est48, synth.LAYOUT_TRUEHD_11)
g48 = FrequencyResponse.generate_frequencies(f_step=1.01, f_min=10, f_max=24000)
target = FrequencyResponse(name='t', frequency=g48.copy(), raw=-1.2 * (np.log10(g48) - 3.0))


def config4():
    res = disp.process(rec11)[0]
    d = res.as_dict()
    from impulcifer_pip313_b200 import ImpulseResponse
    h = HRIR(est48)
    h.irs = {s: {e: ImpulseResponse(v, 48000) for e, v in p.items()} for s, p in d.items()}
    frs = rc.calculate_specific_room_corrections(h, target, limit=400)
    hh = HRIR(est48)
    hh.ingest(hp, 48000, ['FL', 'FR'])
    left, right = rc.headphone_compensation(hh)
    return len(frs), left.error.shape


print('config 4 (TrueHD 11ch set, 22 responses + room corrections + headphone compensation):', config4(),
      f'{best(config4):.1f} ms')

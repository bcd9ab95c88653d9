"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from astroz_b200 import Constellation, synth
from astroz_b200.api import Satrec, WGS72
from tests.golden import tles as G

tles = synth.mixed_catalog(203, n_geo=20, n_molniya=12, n_gps=8)
jd, fr = synth.time_grid(97)
c = Constellation(tles)
for layout in (0, 1):
    for mode in (0, 1, 2):
        for vel in (True, False):
            p, v = c.propagate(jd, fr, outputMode=mode, layout=layout, velocities=vel)
            assert np.isfinite(p).all()
ne = Constellation(synth.near_earth_catalog(75))
t = np.arange(0.0, 200.0, 1.0)
off = (2460437.5 - ne.epochs) * 1440.0
ne.propagate_into(t, epoch_offsets=off)
# time-major with the pair transpose (even and odd row strides), masks, wider blocks
mask = np.ones(75, dtype=np.uint8)
mask[[0, 7, 8, 74]] = 0
for stride in (75, 76, 81):
    for tm in (True, False):
        shape = (len(t), stride, 3) if tm else (stride, len(t), 3)
        p, v = np.zeros(shape), np.zeros(shape)
        ne.propagate_into(t, p, v, epoch_offsets=off, time_major=tm, output_stride=stride)
        ne.propagate_into(t, p, v, epoch_offsets=off, satellite_mask=mask, time_major=tm, output_stride=stride)
# deep-space members only, strided
pd, vd = np.zeros((len(jd), 210, 3)), np.zeros((len(jd), 210, 3))
c.propagate_sdp4_into(jd, fr, pd, vd, output_stride=210, sat_offset=5)
# device-side element init
import torch
el = synth.elements_from_tles(tles)
dv = Constellation.from_device_elements(torch.from_numpy(el).cuda())
dv.propagate(jd, fr)
ne.screen_conjunction(t, 3, 100.0, epoch_offsets=off)
ne.screen_all(t, 50.0, epoch_offsets=off)
for tle in (G.ISS, G.GEO28626, G.HEO09880):
    s = Satrec.twoline2rv(*tle, WGS72)
    s.sgp4(s.jdsatepoch, s.jdsatepochF + 0.1)
    s.sgp4_array(np.full(300, s.jdsatepoch), s.jdsatepochF + np.arange(300) / 1440.0)
print("sanitize_small: all kernels exercised")

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
ne.screen_conjunction(t, 3, 100.0, epoch_offsets=off)
ne.screen_all(t, 50.0, epoch_offsets=off)
for tle in (G.ISS, G.GEO28626, G.HEO09880):
    s = Satrec.twoline2rv(*tle, WGS72)
    s.sgp4(s.jdsatepoch, s.jdsatepochF + 0.1)
    s.sgp4_array(np.full(300, s.jdsatepoch), s.jdsatepochF + np.arange(300) / 1440.0)
print("sanitize_small: all kernels exercised")

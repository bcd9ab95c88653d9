"""End-to-end timing of the fused single-target screen on the headline grid (host call, 12 B/sat back)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from astroz_b200 import Constellation, synth
tles = synth.near_earth_catalog(); c = Constellation(tles); c.set_timing(True)
times = np.arange(1440.0); ref = 2460437.5; off = (ref - c.epochs) * 1440.0
for _ in range(3): d, ti = c.screen_conjunction(times, 0, 10.0, epoch_offsets=off, reference_jd=ref)
t0 = time.perf_counter(); K = 20
for _ in range(K): d, ti = c.screen_conjunction(times, 0, 10.0, epoch_offsets=off, reference_jd=ref)
dt = (time.perf_counter() - t0) / K
print(json.dumps({"screen_e2e_ms": dt * 1e3, "cells": len(tles) * len(times), "Gprops_e2e": len(tles) * len(times) / dt / 1e9,
                  "kernel_ms": c.last_kernel_ms()[0], "below_threshold": int((d < 10.0).sum())}))

for _ in range(2): pairs, ti = c.screen_all(times, 10.0, epoch_offsets=off)
t0 = time.perf_counter(); K = 5
for _ in range(K): pairs, ti = c.screen_all(times, 10.0, epoch_offsets=off)
dt = (time.perf_counter() - t0) / K
print(json.dumps({"screen_all_e2e_ms": dt * 1e3, "Gprops_e2e": len(tles) * len(times) / dt / 1e9, "hits": int(len(ti)),
                  "what": "all-vs-all: propagate (positions, time-major) + device cell-list screen, hits only come back"}))

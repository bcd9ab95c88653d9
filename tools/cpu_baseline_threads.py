"""CPU SIMD port (the timed reference arm): pass-time distribution against the thread count on this host."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from astroz_b200 import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

orc.build()
tles = synth.near_earth_catalog()
jd, fr = synth.time_grid()
sim = orc.SimdConstellation(tles)
n, nt = len(tles), len(jd)
pos = np.zeros((nt, n, 3))
vel = np.zeros((nt, n, 3))
try:
    quota = open("/sys/fs/cgroup/cpu.max").read().strip()
except OSError:
    quota = "n/a"
print(json.dumps({"cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0)), "cgroup_cpu_max": quota}))
for threads in (8, 16, 32, 64, 96, 128):
    sim.propagate(jd, fr, layout=1, threads=threads, out=(pos, vel))
    ts = []
    for _ in range(40):
        t0 = time.perf_counter()
        sim.propagate(jd, fr, layout=1, threads=threads, out=(pos, vel))
        ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e3
    print(json.dumps({"threads": threads, "ms_min": round(ts.min(), 2), "ms_p50": round(float(np.median(ts)), 2),
                      "ms_mean": round(ts.mean(), 2), "ms_max": round(ts.max(), 2),
                      "Gprops_mean": round(n * nt / ts.mean() / 1e6, 3)}))

"""Host-API call time (pinned host buffers in and out) against the number of kernel/D2H pipeline chunks
(ASTROZ_D2H_CHUNKS) on the headline grid."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import astroz_b200 as az  # noqa: E402
from astroz_b200 import synth  # noqa: E402

tles = synth.near_earth_catalog()
jd, fr = synth.time_grid()
n, nt = len(tles), len(jd)
hp = az.pinned_empty((n, nt, 3))
hv = az.pinned_empty((n, nt, 3))
for chunks in (1, 2, 4, 8, 12, 16, 24, 32, 48, 64):
    os.environ["ASTROZ_D2H_CHUNKS"] = str(chunks)
    c = az.Constellation(tles)
    for _ in range(3):
        c.propagate(jd, fr, hp, hv, layout=az.Layout.satelliteMajor)
    t0 = time.perf_counter()
    K = 15
    for _ in range(K):
        c.propagate(jd, fr, hp, hv, layout=az.Layout.satelliteMajor)
    ms = (time.perf_counter() - t0) / K * 1e3
    print(json.dumps({"chunks": chunks, "ms": round(ms, 3), "Gprops": round(n * nt / ms / 1e6, 4),
                      "d2h_GBs": round(2 * hp.nbytes / ms / 1e6, 2)}))
    c.deinit()

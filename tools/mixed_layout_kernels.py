#!/usr/bin/env python
"""Per-kernel times of a mixed (config 3) call in both layouts: K1 (near-earth grid), K2 (deep-space grid), call span."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from astroz_b200 import Constellation, synth  # noqa: E402

tles = synth.mixed_catalog()
jd, fr = synth.time_grid()
dev = torch.device("cuda", 0)
n, nt = len(tles), len(jd)
c = Constellation(tles)
c.set_timing(True)
pos = torch.empty((n * nt * 3,), dtype=torch.float64, device=dev)
vel = torch.empty_like(pos)
for layout in (0, 1):
    for v in (True, False):
        ks = []
        for _ in range(6):
            c.propagate_device(jd, fr, pos, vel if v else None, None, 0, layout)
            c.synchronize()
            ks.append(c.last_kernel_ms())
        k = ks[-1]
        print(json.dumps({"layout": layout, "vel": v, "k1_ms": k[0], "span_ms": k[1], "k2_ms": k[2]}), flush=True)

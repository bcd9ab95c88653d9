#!/usr/bin/env python
"""Time-major near-earth grid with the block's rows shifted by 0 / 1 (pairs 16-byte aligned / not) and with a mask."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from astroz_b200 import Constellation, synth  # noqa: E402

tles = synth.near_earth_catalog()
jd, fr = synth.time_grid()
dev = torch.device("cuda", 0)
n, nt = len(tles), len(jd)
c = Constellation(tles)
c.set_timing(True)
pos = torch.empty(((n + 2) * nt * 3,), dtype=torch.float64, device=dev)
vel = torch.empty_like(pos)
for off in (0, 1):
    for v in (True, False):
        for _ in range(5):
            c.propagate_device(jd, fr, pos, vel if v else None, None, 0, 1, out_num_sats=n + 2, out_sat_offset=off)
            c.synchronize()
        print(json.dumps({"row_offset": off, "vel": v, "k1_ms": c.last_kernel_ms()[0]}), flush=True)

#!/usr/bin/env python
"""Times the grid kernels of whichever library build ASTROZ_B200_LIB names (tools/variant_sweep.sh): config 2 (K1),
config 3 mixed (K1 + K2 side by side) and the deep-space grid alone."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from astroz_b200 import Constellation, _lib, synth  # noqa: E402

dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
jd, fr = synth.time_grid()
nt = len(jd)
out = {"lib": os.environ.get("ASTROZ_B200_LIB", "default"), "tag": os.environ.get("AZ_TAG", "")}


def timed(fn, K=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K


for name, tles in (("config2", synth.near_earth_catalog()), ("config3", synth.mixed_catalog())):
    c = Constellation(tles)
    n = len(tles)
    block = torch.empty((2, n, nt, 3), dtype=torch.float64, device=dev)
    ms = timed(lambda: c.propagate_device(jd, fr, block[0], block[1], stream=stream.cuda_stream))
    out[name + "_ms"] = ms
    out[name + "_Gprops"] = n * nt / ms / 1e6
    if c.numSdp4:
        nd = c.numSdp4
        d = torch.empty((2, nd, nt, 3), dtype=torch.float64, device=dev)
        call = lambda: _lib.check(_lib.lib().astroz_cuda_sdp4_propagate_into_device(  # noqa: E731
            c._h, _lib.dptr(jd), _lib.dptr(fr), nt, C.c_void_p(d[0].data_ptr()), C.c_void_p(d[1].data_ptr()), 0, 0, nd, 0,
            C.c_void_p(stream.cuda_stream)))
        out["k2_alone_ms"] = timed(call)
        out["k2_alone_Gprops"] = nd * nt / out["k2_alone_ms"] / 1e6
    del c, block
print(json.dumps(out), flush=True)

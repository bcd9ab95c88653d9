#!/usr/bin/env python
"""astroz_cuda_sgp4_array on a 31.5 M-epoch axis, called directly: pageable vs pinned epochs, by chunk count."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import astroz_b200
from astroz_b200 import _lib
from astroz_b200.api import Satrec, WGS72
from tests.golden import tles as G
sat = Satrec.twoline2rv(*G.ISS, WGS72)
L = _lib.lib()
for n in (2_000_000, 8_000_000, 31_536_000):
    out = astroz_b200.pinned_empty((n, 6))
    for kind in ("pageable", "pinned"):
        if kind == "pinned":
            jd = astroz_b200.pinned_empty((n,)); fr = astroz_b200.pinned_empty((n,))
        else:
            jd = np.empty(n); fr = np.empty(n)
        jd[:] = sat.jdsatepoch; fr[:] = sat.jdsatepochF + np.arange(n) / 86400.0
        ep = sat.jdsatepoch + sat.jdsatepochF
        for _ in range(2):
            _lib.check(L.astroz_cuda_sgp4_array(sat._h, _lib.dptr(jd), _lib.dptr(fr), ep, _lib.dptr(out), n))
        t0 = time.perf_counter(); K = 4
        for _ in range(K):
            _lib.check(L.astroz_cuda_sgp4_array(sat._h, _lib.dptr(jd), _lib.dptr(fr), ep, _lib.dptr(out), n))
        dt = (time.perf_counter() - t0) / K
        print(json.dumps({"n": n, "epochs": kind, "ms": dt * 1e3, "Mprops": n / dt / 1e6, "GBs_total": n * 64 / dt / 1e9}), flush=True)

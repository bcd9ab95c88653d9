"""1 satellite x N epochs through Satrec.sgp4_array (host in, host out) -- the reference's single-satellite benchmark family."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from astroz_b200.api import Satrec, WGS72
from tests.golden import tles as G
sat = Satrec.twoline2rv(*G.ISS, WGS72)
for n in (1440, 1_209_600, 31_536_000 // 4, 31_536_000):   # ... "1 year (second)", benchmarks/zig_sgp4_bench.zig:46-52
    jd = np.full(n, sat.jdsatepoch); fr = sat.jdsatepochF + np.arange(n) / 86400.0
    for _ in range(3): e, r, v = sat.sgp4_array(jd, fr)   # results held like in the timed loop: both pinned blocks of the pool exist before the clock starts
    t0 = time.perf_counter(); K = 5
    for _ in range(K): e, r, v = sat.sgp4_array(jd, fr)
    dt = (time.perf_counter() - t0) / K
    print(json.dumps({"n_times": n, "ms": dt * 1e3, "Mprops_e2e": n / dt / 1e6}), flush=True)

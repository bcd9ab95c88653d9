"""Ingest of a large Monte-Carlo element cloud (SURVEY 8f-3 / BASELINE config 5 scaled to 10^6 draws):
host init (threaded az_elements.hpp builders + table upload) vs device init (K5, az_ingest.cu).

    python tools/ingest_timing.py [n_draws]
"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import astroz_b200 as az  # noqa: E402
from astroz_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
base = synth.elements_from_tles(synth.monte_carlo_catalog(1))[:, 0]
g = torch.Generator(device="cuda").manual_seed(12345)
sig = torch.tensor([0.0, 1e-5, 1e-5, 0.01, 0.01, 0.01, 0.01, 0.0], dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
el = torch.from_numpy(base).cuda()[:, None] + sig[:, None] * torch.randn((8, n), generator=g, dtype=torch.float64, device="cuda")
el[2].clamp_(min=1e-7)
el[7] = base[7] * (1.0 + 0.1 * torch.randn(n, generator=g, dtype=torch.float64, device="cuda"))
torch.cuda.synchronize()
t_draw = time.perf_counter() - t0

res = {"n_draws": n, "draw_ms": t_draw * 1e3}
for rep in range(3):
    t0 = time.perf_counter()
    dev = az.Constellation.from_device_elements(el)
    dev.set_timing(True)
    t_dev = time.perf_counter() - t0
    res.setdefault("device_init_ms", []).append(round(t_dev * 1e3, 3))
    if rep < 2:
        dev.deinit()
host_el = el.cpu().numpy()
for rep in range(2):
    t0 = time.perf_counter()
    host = az.Constellation.from_elements(*host_el)
    t_host = time.perf_counter() - t0
    res.setdefault("host_init_ms", []).append(round(t_host * 1e3, 3))
    if rep < 1:
        host.deinit()
# a short propagation of the whole cloud, results left in HBM
nt = 16
jd = np.full(nt, float(base[0]))
fr = np.arange(nt) * (90.0 / 1440.0 / nt)
out_p = torch.empty((n, nt, 3), dtype=torch.float64, device="cuda")
out_v = torch.empty_like(out_p)
ref_p = torch.empty_like(out_p)
ref_v = torch.empty_like(out_p)
for c, (p, v) in ((dev, (out_p, out_v)), (host, (ref_p, ref_v))):
    c.propagate_device(jd, fr, p, v)
    c.synchronize()
t0 = time.perf_counter()
dev.propagate_device(jd, fr, out_p, out_v)
dev.synchronize()
res["propagate_ms"] = round((time.perf_counter() - t0) * 1e3, 3)
res["kernel_ms"] = dev.last_kernel_ms()
res["max_dr_km_device_vs_host_init"] = float((out_p - ref_p).abs().max())
res["max_dv_kms_device_vs_host_init"] = float((out_v - ref_v).abs().max())
print(json.dumps(res))

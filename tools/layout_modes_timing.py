"""Kernel-only timing of every output specialisation on the headline grid (device-resident)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from astroz_b200 import Constellation, synth
tles = synth.mixed_catalog() if os.environ.get("AZ_CATALOG") == "mixed" else synth.near_earth_catalog()  # AZ_CATALOG=mixed: config 3
jd, fr = synth.time_grid()
dev = torch.device("cuda", 0); n, nt = len(tles), len(jd)
c = Constellation(tles)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
pos = torch.empty((n * nt * 3,), dtype=torch.float64, device=dev); vel = torch.empty_like(pos)
for layout in (0, 1):
    for mode in (0, 1, 2):
        for v in (True, False):
            f = lambda: c.propagate_device(jd, fr, pos, vel if v else None, None, mode, layout, stream=stream.cuda_stream)
            for _ in range(3): f()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(10): f()
            e1.record(stream); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(json.dumps({"layout": layout, "mode": mode, "vel": v, "ms": round(ms, 4), "Gprops": round(n * nt / ms / 1e6, 2)}), flush=True)

"""K2 tuning sweep on the deep-space part of config 3 (1,536 deep-space satellites x 1,440 epochs)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from astroz_b200 import Constellation, synth
tles = [t for t in synth.mixed_catalog() if float(t[1][52:63]) < 6.0]
jd, fr = synth.time_grid()
dev = torch.device("cuda", 0); n, nt = len(tles), len(jd)
pos = torch.empty((n, nt, 3), dtype=torch.float64, device=dev); vel = torch.empty_like(pos)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
for v in (-1, 0, 1, 2):
    os.environ["ASTROZ_SDP4_VARIANT"] = str(v)
    c = Constellation(tles)
    for _ in range(3): c.propagate_device(jd, fr, pos, vel, stream=stream.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(20): c.propagate_device(jd, fr, pos, vel, stream=stream.cuda_stream)
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(json.dumps({"variant": v, "n_sdp4": c.numSdp4, "ms": round(ms, 4), "Gprops": round(n * nt / ms / 1e6, 2)}), flush=True)

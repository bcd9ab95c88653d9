"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): satellite-sharded propagate with
(a) one ncclAllGather and (b) the fused NVLink kernels must reproduce the single-GPU block bit for bit."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("case", [("2003", "777", "mixed"), ("4096", "1440")])
def test_sharded_gather_bitwise(case):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "multi_gpu_check.py"), *case]
    env = dict(os.environ, NCCL_DEBUG="WARN")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["nccl_equal"] and res["fused_peer_equal"]
    if res.get("multicast"):
        assert res["fused_multicast_equal"]


def test_one_handle_over_all_gpus_matches_one_gpu():
    """device = -1 on a multi-GPU box: one propagate call, every GPU, each device copying its rows over its own PCIe
    link; and the replicated (all-gather) propagate over NVLink peer mappings.  Bit for bit the one-GPU result."""
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    import numpy as np

    import astroz_b200 as az
    from astroz_b200 import synth
    from tests.test_gpu_parity import _device_block

    tles = synth.mixed_catalog(4001, n_geo=200, n_molniya=60, n_gps=60)
    jd, fr = synth.time_grid(480)
    single = az.Constellation(tles, device=0)
    multi = az.Constellation(tles, device=-1)
    ids, rows = multi.devices
    assert ids == list(range(min(n, 8))) and rows[-1] == 4001
    for layout in (0, 1):
        ps, vs = single.propagate(jd, fr, layout=layout)
        pm, vm = multi.propagate(jd, fr, layout=layout)
        assert np.array_equal(ps, pm) and np.array_equal(vs, vm)
    ps, vs = single.propagate(jd, fr, layout=0)
    blk_p, blk_v = multi.host_block(len(jd), 0), multi.host_block(len(jd), 0)   # each GPU's rows on its own NUMA node
    multi.propagate(jd, fr, blk_p, blk_v, layout=0)
    assert np.array_equal(blk_p, ps) and np.array_equal(blk_v, vs)
    ids, ppos, pvel = multi.propagate_replicated(jd, fr)
    from cuda import cudart

    for k, d in enumerate(ids):
        cudart.cudaSetDevice(d)
        assert np.array_equal(_device_block(ppos[k], ps.shape), ps)
        assert np.array_equal(_device_block(pvel[k], vs.shape), vs)
    cudart.cudaSetDevice(0)

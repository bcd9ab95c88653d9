"""Multi-GPU check (torchrun, one rank per GPU): satellite-sharded config-2-style catalog.
Verifies that (a) sharded propagate + NCCL all-gather, (b) the fused multimem.st kernel and (c) the fused
peer-store kernel all give the single-GPU block bit for bit, and times them.
    python -m torch.distributed.run --nproc-per-node N tools/multi_gpu_check.py [n_sats] [n_times]
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from astroz_b200 import Constellation, synth  # noqa: E402
from astroz_b200.parallel import ShardedPropagator, SymmetricBlock, propagate_gather  # noqa: E402

rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n_sats = int(sys.argv[1]) if len(sys.argv) > 1 else 13478
n_times = int(sys.argv[2]) if len(sys.argv) > 2 else 1440
mixed = len(sys.argv) > 3 and sys.argv[3] == "mixed"
tles = synth.mixed_catalog(n_sats) if mixed else synth.near_earth_catalog(n_sats)
jd, fr = synth.time_grid(n_times)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)

sp = ShardedPropagator(tles, rank, world, device=lr, dist=dist)
rows, pr = sp.rows, sp.padded_rows
sym = SymmetricBlock(pr, n_times, dev)

# single-GPU truth (every rank computes the whole catalog once)
whole = Constellation(tles, device=lr)
whole.referenceEpochJd = sp.reference_epoch
truth = torch.zeros((2, pr, n_times, 3), dtype=torch.float64, device=dev)
whole.propagate_device(jd, fr, truth[0], truth[1], out_num_sats=pr, stream=stream.cuda_stream)
torch.cuda.synchronize()

def timed(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

res = {"world": world, "n_sats": n_sats, "n_times": n_times, "multicast": sym.has_multicast}

# (a) shard-local propagate then one NCCL all-gather of the [pos|vel] rows
nccl = torch.zeros((world, 2, rows, n_times, 3), dtype=torch.float64, device=dev)
mine = torch.zeros((2, rows, n_times, 3), dtype=torch.float64, device=dev)
def step_nccl():
    if sp.local is not None:
        sp.local.propagate_device(jd, fr, mine[0], mine[1], out_num_sats=rows, stream=stream.cuda_stream)
    dist.all_gather_into_tensor(nccl, mine)
res["nccl_ms"] = timed(step_nccl)
got = nccl.permute(1, 0, 2, 3, 4).reshape(2, pr, n_times, 3)
res["nccl_equal"] = bool(torch.equal(got[:, :n_sats], truth[:, :n_sats]))

def step_local():
    if sp.local is not None:
        sp.local.propagate_device(jd, fr, mine[0], mine[1], out_num_sats=rows, stream=stream.cuda_stream)
res["local_only_ms"] = timed(step_local)

# (b) fused multicast
if sym.has_multicast:
    sym.block.zero_(); torch.cuda.synchronize(); sym.barrier()
    def step_mc():
        propagate_gather(sp, sym, jd, fr, stream=stream.cuda_stream, use_multicast=True)
        sym.barrier()
    res["fused_multicast_ms"] = timed(step_mc)
    res["fused_multicast_equal"] = bool(torch.equal(sym.block[:, :n_sats], truth[:, :n_sats]))

# (c) fused peer stores
sym.block.zero_(); torch.cuda.synchronize(); sym.barrier()
def step_p2p():
    propagate_gather(sp, sym, jd, fr, stream=stream.cuda_stream, use_multicast=False)
    sym.barrier()
res["fused_peer_ms"] = timed(step_p2p)
res["fused_peer_equal"] = bool(torch.equal(sym.block[:, :n_sats], truth[:, :n_sats]))

cells = n_sats * n_times
for k in ("nccl_ms", "local_only_ms", "fused_multicast_ms", "fused_peer_ms"):
    if k in res:
        res[k.replace("_ms", "_Gprops")] = cells / res[k] / 1e6
res["recv_GB_per_gpu"] = (world - 1) * 2 * rows * n_times * 24 / 1e9
if rank == 0:
    print(json.dumps(res), flush=True)
dist.barrier()
dist.destroy_process_group()

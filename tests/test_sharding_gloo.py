"""N > 1 host logic on CPU: world_size-2 (and 3) gloo runs of the satellite sharding + all-gather assembly.
The GPU propagate is replaced by an oracle-backed stand-in with the same interface, so what is tested is the
sharding arithmetic, the reference-epoch agreement across ranks, row offsets and the collective."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds():
    from astroz_b200.parallel import shard_bounds, shard_rows

    assert shard_rows(13478, 8) == 1688 and shard_rows(13478, 1) == 13480
    b = shard_bounds(13478, 8)
    assert b[0] == (0, 1688) and b[-1] == (11816, 13478)
    assert all(x[0] % 8 == 0 for x in b)
    assert shard_bounds(5, 4) == [(0, 5), (5, 5), (5, 5), (5, 5)]
    for n in (1, 7, 8, 9, 1000, 13478):
        for w in (1, 2, 3, 4, 8):
            bb = shard_bounds(n, w)
            assert bb[0][0] == 0 and bb[-1][1] == n and all(bb[i][1] == bb[i + 1][0] for i in range(w - 1))


class _OracleBackend:
    """Stand-in for astroz_b200.Constellation on a CPU-only box (test infrastructure)."""

    def __init__(self, tles, grav=1, device=0):
        from oracle import oracle as orc

        self.orc, self.tles, self.grav = orc, tles, grav
        self.numSatellites = len(tles)
        self._k = orc.constellation_propagate(tles, np.array([2460437.5]), np.array([0.0]), grav=grav)[3]
        self.numSgp4 = int((self._k == 0).sum())
        self.numSdp4 = self.numSatellites - self.numSgp4
        self.classes = self._k
        self.epochs = np.array([orc.parse_tle(*t)["epochJd"] for t in tles])
        self.referenceEpochJd = None

    def propagate_device(self, jd, fr, pos, vel, status, mode, layout, out_num_sats, out_sat_offset, stream):
        # the oracle's own reference epoch is its first near-earth satellite; emulate an externally set one
        orc = self.orc
        n, nt = self.numSatellites, len(jd)
        for i, t in enumerate(self.tles):
            if self._k[i] == 0:
                s = orc.Sgp4(*t, grav=self.grav)
                ref = self.referenceEpochJd
                ts = ((jd + fr) - ref) * 1440.0 + (ref - s.epochJd) * 1440.0
                rv = np.array([s.propagate(x) for x in ts])
            else:
                s = orc.Sdp4(*t, grav=self.grav)
                ts = ((jd + fr) - s.epochJd) * 1440.0
                rv = np.array([s.propagate(x)[1:] for x in ts])
            pos[out_sat_offset + i] = torch.from_numpy(rv[:, 0])
            if vel is not None:
                vel[out_sat_offset + i] = torch.from_numpy(rv[:, 1])


def _worker(rank, world, port, n_sats, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from astroz_b200 import synth
    from astroz_b200.parallel import ShardedPropagator

    # deep-space objects first so rank 0's shard starts with non-SGP4 rows: the reference epoch must still
    # be the whole catalog's first near-earth satellite
    tles = synth.mixed_catalog(n_sats, n_geo=3, n_molniya=2, n_gps=2)
    jd, fr = synth.time_grid(5)
    sp = ShardedPropagator(tles, rank, world, backend=_OracleBackend, dist=dist)
    full = torch.zeros((sp.padded_rows, len(jd), 6), dtype=torch.float64)
    pos, vel = full[..., :3], full[..., 3:]
    sp.propagate_into_full(jd, fr, _View(full, 0), _View(full, 3))
    sp.all_gather(full)
    if rank == world - 1:
        np.save(out_path, np.concatenate([[sp.reference_epoch], full[:n_sats].numpy().ravel()]))
    dist.barrier()
    dist.destroy_process_group()


class _View:
    """Row-assignable view of columns [c0, c0+3) of the interleaved [rows, nt, 6] block."""

    def __init__(self, full, c0):
        self.full, self.c0 = full, c0

    def __setitem__(self, row, value):
        self.full[row, :, self.c0:self.c0 + 3] = value


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_all_gather_matches_single_process(tmp_path, world, oracle):
    from astroz_b200 import synth

    n_sats = 29
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(world, port, n_sats, out), nprocs=world, join=True)
    got = np.load(out)
    tles = synth.mixed_catalog(n_sats, n_geo=3, n_molniya=2, n_gps=2)
    jd, fr = synth.time_grid(5)
    po, vo, err, klass = oracle.constellation_propagate(tles, jd, fr)
    first_sgp4 = int(np.flatnonzero(klass == 0)[0])
    assert got[0] == oracle.parse_tle(*tles[first_sgp4])["epochJd"]
    block = got[1:].reshape(n_sats, len(jd), 6)
    assert np.array_equal(block[..., :3], po) and np.array_equal(block[..., 3:], vo)

"""The CPU SIMD baseline (oracle/simd_baseline.c, the port bench.py times as the reference arm) against
the scalar oracle at the reference's own SIMD-vs-scalar tolerance (src/Sgp4Batch.zig:180-189)."""
import numpy as np

from tests.golden import tles as G


def test_simd_port_matches_scalar(oracle):
    from astroz_b200 import synth

    tles = synth.near_earth_catalog(203)
    jd, fr = synth.time_grid(1440)
    jd, fr = jd[::45].copy(), fr[::45].copy()
    sim = oracle.SimdConstellation(tles)
    for layout in (0, 1):
        ps, vs = sim.propagate(jd, fr, layout=layout, threads=3)
        po, vo, _, _ = oracle.constellation_propagate(tles, jd, fr, layout=layout)
        assert np.max(np.abs(ps - po)) < 1e-3 and np.max(np.abs(vs - vo)) < 1e-6
    p1, _ = sim.propagate(jd, fr, layout=1, threads=1, velocities=False)
    assert np.array_equal(p1, sim.propagate(jd, fr, layout=1, threads=5)[0])


def test_simd_port_vallado(oracle):
    # src/Sgp4Batch.zig:235-296: 0.01 km / 1e-6 km/s
    for tle, t, pos, vel in G.VALLADO:
        sim = oracle.SimdConstellation([tle, tle, tle])
        ep = oracle.parse_tle(*tle)["epochJd"]
        p, v = sim.propagate(np.array([ep]), np.array([t / 1440.0]), layout=0)
        assert np.max(np.abs(p[0, 0] - np.array(pos))) < 0.01
        assert np.max(np.abs(v[0, 0] - np.array(vel))) < 1e-6

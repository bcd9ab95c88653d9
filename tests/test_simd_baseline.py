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


def test_simd_port_mixed_catalog_matches_scalar(oracle):
    """The Sdp4Batch port (all three resonance classes, the carry walking the time axis, classification and the
    original-index scatter of Constellation.init) at the reference's batch-vs-scalar tolerance
    (src/Sdp4Batch.zig:543-548: 1e-3 km, 1e-6 km/s), both layouts, both deep-space thread policies."""
    from astroz_b200 import synth

    tles = synth.mixed_catalog(330, n_geo=40, n_molniya=21, n_gps=19)
    jd, fr = synth.time_grid(1440)
    jd, fr = jd[::20].copy(), fr[::20].copy()      # 72 epochs over a day: the 720-minute lattice is crossed
    sim = oracle.SimdConstellation(tles)
    assert (sim.numSgp4, sim.numSdp4) == (250, 80)
    po0, vo0, err, klass = oracle.constellation_propagate(tles, jd, fr, layout=0)
    assert not err.any() and set(klass.tolist()) == {0, 1, 2, 3}
    for layout in (0, 1):
        for sdp4_threads in (0, 3):
            ps, vs = sim.propagate(jd, fr, layout=layout, threads=4, sdp4_threads=sdp4_threads)
            if layout == 1:
                ps, vs = ps.transpose(1, 0, 2), vs.transpose(1, 0, 2)
            assert np.max(np.abs(ps - po0)) < 1e-3 and np.max(np.abs(vs - vo0)) < 1e-6
    # a week out and backwards in time: the integrator restarts, the port still follows the scalar path
    jd2 = np.full(9, jd[0])
    fr2 = np.array([7.0, 6.5, 3.0, 0.2, -0.4, -2.0, -5.5, 1.0, 4.0])
    ps, vs = sim.propagate(jd2, fr2, layout=0, threads=2)
    po, vo, err, _ = oracle.constellation_propagate(tles, jd2, fr2, layout=0)
    assert not err.any()
    assert np.max(np.abs(ps - po)) < 1e-3 and np.max(np.abs(vs - vo)) < 1e-6


def test_simd_port_sdp4_vectors(oracle):
    # src/Sdp4.zig:1481-1559 (python-sgp4 vectors for GPS / GEO / HEO); the batch path is held to 1e-3 km / 1e-6 km/s
    # of the scalar path (src/Sdp4Batch.zig:543-548), which itself reproduces these to print precision
    for tle, t, pos, vel in G.SDP4_VECTORS:
        sim = oracle.SimdConstellation([tle])
        assert sim.numSdp4 == 1
        ep = oracle.parse_tle(*tle)["epochJd"]
        p, v = sim.propagate(np.array([ep]), np.array([t / 1440.0]), layout=0, threads=1)
        assert np.max(np.abs(p[0, 0] - np.array(pos))) < 2e-3
        if vel is not None:
            assert np.max(np.abs(v[0, 0] - np.array(vel))) < 2e-6

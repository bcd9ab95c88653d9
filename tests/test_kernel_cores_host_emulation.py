"""The product's __host__ __device__ per-cell cores (astroz_b200/csrc/az_device.cuh) and its host-side
element builder, run on the CPU by a test-only harness (tests/host_emul/emul.cu) against the oracle.
This checks the kernel arithmetic in the container that has no GPU; the real device run is in
tests/test_gpu_parity.py."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from tests.golden import tles as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL_DIR = os.path.join(ROOT, "tests", "host_emul")


@pytest.fixture(scope="module")
def emul():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc unavailable")
    so = os.path.join(EMUL_DIR, "libemul.so")
    src = os.path.join(EMUL_DIR, "emul.cu")
    csrc = os.path.join(ROOT, "astroz_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".hpp"))]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "--expt-relaxed-constexpr",
                        "-Xcompiler", "-fPIC", "-shared", "-I" + csrc, "-o", so, src], check=True, capture_output=True)
    L = C.CDLL(so)

    def run(tles, jd, fr, grav=1):
        n, nt = len(tles), len(jd)
        a1 = (C.c_char_p * n)(*[t[0].encode() for t in tles])
        a2 = (C.c_char_p * n)(*[t[1].encode() for t in tles])
        p = np.zeros((n, nt, 3))
        v = np.zeros((n, nt, 3))
        st = np.zeros((n, nt), dtype=np.uint8)
        dp = C.POINTER(C.c_double)
        rc = L.emul_constellation_propagate(a1, a2, n, grav, jd.ctypes.data_as(dp), fr.ctypes.data_as(dp), nt,
                                            p.ctypes.data_as(dp), v.ctypes.data_as(dp),
                                            st.ctypes.data_as(C.POINTER(C.c_uint8)))
        assert rc == 0
        return p, v, st

    run.lib = L
    return run


def test_sincos_kernel_accuracy(emul):
    """sincos_full (az_math.cuh): reduction to the 1024-point lattice of the circle, table entry, two-term remainder
    series, angle addition.  Dense sweep, lattice edges (the remainder at its maximum), large arguments; against numpy
    everywhere and against 40-digit values on a subset.  The reference's own bound is 1e-12 (src/simdMath.zig:214-232)."""
    rng = np.random.default_rng(5)
    h = 2.0 * np.pi / 1024.0
    k = np.arange(-1100, 1100)
    x = np.concatenate([np.linspace(-20.0, 20.0, 4001), (k + 0.5) * h, (k + 0.49999) * h, (k - 0.49999) * h, k * h,
                        rng.uniform(-1.0e3, 1.0e3, 3000), rng.uniform(-1.0e5, 1.0e5, 3000),
                        np.array([1e3, -5e4, 7.0e5, 1e-300, 0.0, -0.0, 1e-9])])
    s = np.zeros_like(x)
    c = np.zeros_like(x)
    dp = C.POINTER(C.c_double)
    emul.lib.emul_sincos(x.ctypes.data_as(dp), len(x), s.ctypes.data_as(dp), c.ctypes.data_as(dp))
    assert np.max(np.abs(s - np.sin(x))) < 4e-16
    assert np.max(np.abs(c - np.cos(x))) < 4e-16
    assert np.max(np.abs(s * s + c * c - 1.0)) < 6e-16
    try:
        import mpmath
    except ImportError:
        return
    mpmath.mp.dps = 40
    sub = rng.choice(len(x), 1500, replace=False)
    es = max(abs(mpmath.sin(mpmath.mpf(float(x[i]))) - mpmath.mpf(float(s[i]))) for i in sub)
    ec = max(abs(mpmath.cos(mpmath.mpf(float(x[i]))) - mpmath.mpf(float(c[i]))) for i in sub)
    assert es < 3e-16 and ec < 3e-16


def test_cores_match_oracle_all_classes(emul, oracle):
    from astroz_b200 import synth

    jd, fr = synth.time_grid(10080)
    jd, fr = jd[::197].copy(), fr[::197].copy()
    for tles in (synth.near_earth_catalog(400), synth.mixed_catalog(300, n_geo=60, n_molniya=40, n_gps=40),
                 [G.ISS, G.GEO28626, G.SAT55909, G.GPS20413, G.SAT55910, G.HEO09880]):
        po, vo, err, klass = oracle.constellation_propagate(tles, jd, fr)
        shapes = {}
        for lanes in (1, 2, 3):  # one epoch per thread, the two-epoch form of both cores, the shipped 3 x 32-apart grouping
            emul.lib.emul_set_lanes(lanes)
            pe, ve, st = emul(tles, jd, fr)
            assert np.max(np.abs(po - pe)) < 1e-6 and np.max(np.abs(vo - ve)) < 1e-9
            deep = klass > 0
            assert np.array_equal(st[deep], err[deep])
            shapes[lanes] = (pe, ve)
        # the series a cell takes (speculative two-step Kepler solve, micro rotations) is chosen per thread over its
        # epochs, so two groupings may solve a cell differently: they must agree to the reference's own
        # layout-equivalence bound (src/Constellation.zig:869: 1e-10 km)
        for lanes in (2, 3):
            assert np.max(np.abs(shapes[lanes][0] - shapes[1][0])) < 1e-10
            assert np.max(np.abs(shapes[lanes][1] - shapes[1][1])) < 1e-13
        emul.lib.emul_set_lanes(1)


def test_cores_error_cells(emul, oracle):
    from astroz_b200 import synth

    bad = synth.tle_lines(42000, 24, 120.0, 63.4, 0.0, 0.755, 0.0, 10.0, 2.006, 1e-3)
    jd = np.full(40, 2460430.5)
    fr = np.linspace(0.0, 2000.0, 40)
    po, vo, err, _ = oracle.constellation_propagate([G.GPS20413, bad], jd, fr)
    for lanes in (1, 2):         # a failing lane must not disturb the cell sharing its thread
        emul.lib.emul_set_lanes(lanes)
        pe, ve, st = emul([G.GPS20413, bad], jd, fr)
        assert err[1].any() and np.array_equal(st, err)
        assert np.all(pe[1][err[1] != 0] == 0.0)
        ok = err == 0
        assert np.max(np.abs(po[ok] - pe[ok])) < 1e-5
    emul.lib.emul_set_lanes(1)


def test_geodetic_epilogue_matches_reference_iteration(emul, oracle):
    # src/WorldCoordinateSystem.zig:98-121 (radians, km)
    rng = np.random.default_rng(3)
    r = 6378.137 + rng.uniform(150.0, 42000.0, 4000)
    lat = rng.uniform(-1.55, 1.55, 4000)
    lon = rng.uniform(-np.pi, np.pi, 4000)
    ecef = np.stack([r * np.cos(lat) * np.cos(lon), r * np.cos(lat) * np.sin(lon), r * np.sin(lat)], axis=1).copy()
    out = np.zeros_like(ecef)
    dp = C.POINTER(C.c_double)
    emul.lib.emul_ecef_to_geodetic(ecef.ctypes.data_as(dp), len(ecef), out.ctypes.data_as(dp))
    ref = np.array([oracle.ecef_to_geodetic(e) for e in ecef])
    assert np.max(np.abs(out[:, :2] - ref[:, :2])) < 1e-12
    assert np.max(np.abs(out[:, 2] - ref[:, 2])) < 1e-7


def test_geodetic_epilogue_branch_cut_and_axis(emul, oracle):
    # the longitude branch cut (x < 0, y = +-0 and tiny), the equator, the polar axis
    dp = C.POINTER(C.c_double)
    cases = np.array([
        [-7000.0, 0.0, 100.0], [-7000.0, -0.0, 100.0], [-7000.0, 1e-9, -50.0], [-7000.0, -1e-9, -50.0],
        [-7000.0, 1e-300, 3.0], [7000.0, 0.0, 0.0], [0.0, 7000.0, 0.0], [0.0, -7000.0, 1.0e-12],
        [1e-3, -1e-3, 7000.0], [1e-3, 1e-3, -7000.0], [0.0, 0.0, 7000.0], [0.0, 0.0, -6900.0],
        [4000.0, 4000.0, 42164.0], [-30000.0, 29000.0, -500.0]])
    out = np.zeros_like(cases)
    emul.lib.emul_ecef_to_geodetic(np.ascontiguousarray(cases).ctypes.data_as(dp), len(cases), out.ctypes.data_as(dp))
    ref = np.array([oracle.ecef_to_geodetic(e) for e in cases])
    assert np.max(np.abs(out[:, :2] - ref[:, :2])) < 1e-12
    polar = np.abs(np.abs(ref[:, 0]) - np.pi / 2) < 1e-5      # the reference's p / cos(lat) - N loses digits there
    assert np.max(np.abs(out[~polar, 2] - ref[~polar, 2])) < 1e-7
    b = 6378.137 * (1.0 - 1.0 / 298.257223563)
    assert np.allclose(out[polar, 2], np.abs(cases[polar, 2]) - b, atol=1e-6)


def test_angle_table_is_in_sync_and_dense_angles_round_trip(emul):
    """astroz_b200/csrc/az_angle_table.inc is what tools/gen_angle_table.py writes, and angle_of_unit (through the
    geodetic epilogue's longitude) returns atan2 to 3e-16 over a dense sweep of the circle, lattice midpoints included."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert subprocess.run([sys.executable, os.path.join(root, "tools", "gen_angle_table.py"), "--check"]).returncode == 0
    k = np.arange(-402, 403)
    ang = np.concatenate([k / 128.0, (k + 0.5) / 128.0, (k + 0.4999) / 128.0, np.linspace(-np.pi, np.pi, 20001)])
    ang = ang[(ang > -np.pi) & (ang < np.pi)]
    ecef = np.stack([7000.0 * np.cos(ang), 7000.0 * np.sin(ang), np.full_like(ang, 123.0)], axis=1).copy()
    out = np.zeros_like(ecef)
    dp = C.POINTER(C.c_double)
    emul.lib.emul_ecef_to_geodetic(ecef.ctypes.data_as(dp), len(ecef), out.ctypes.data_as(dp))
    assert np.max(np.abs(out[:, 1] - np.arctan2(ecef[:, 1], ecef[:, 0]))) < 4e-16

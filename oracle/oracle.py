"""ctypes loader for the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (astroz_b200) never does.

The oracle is the scalar C restatement in oracle/astroz_oracle.c (parity pinned against the
reference's golden vectors by tests/test_oracle_golden.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")

WGS84, WGS72 = 0, 1
OK, DECAYED, INVALID_ECC, DEEP_SPACE, OOM, BAD_TLE = 0, 1, 2, 3, 4, 5

SGP4_FIELDS = [
    "epochJd", "noKozai", "ecco", "inclo", "nodeo", "argpo", "mo", "bstar", "noUnkozai", "a",
    "sinio", "cosio", "cosio2", "cosio4", "con41", "con42", "x1mth2", "x7thm1",
    "mdot", "argpdot", "nodedot",
    "cc1", "cc4", "cc5", "t2cof", "omgcof", "xnodcf", "xlcof", "xmcof", "aycof", "eta", "delmo", "sinmao",
    "d2", "d3", "d4", "t3cof", "t4cof", "t5cof", "aBase", "vkmpersec", "isimp",
]
_PERT = ["e2", "e3", "i2", "i3", "l2", "l3", "l4", "gh2", "gh3", "gh4", "h2", "h3"]
SDP4_FIELDS = (
    ["solar_" + p for p in _PERT] + ["lunar_" + p for p in _PERT]
    + ["zmol", "zmos", "dedt", "didt", "dmdt", "domdt", "dnodt", "irez",
       "d2201", "d2211", "d3210", "d3222", "d4410", "d4422", "d5220", "d5232", "d5421", "d5433",
       "del1", "del2", "del3", "xlamo", "xfact", "gsto"]
)


def build(force: bool = False) -> None:
    """Compile oracle/_build/*.so with the committed Makefile (gcc only, a few seconds)."""
    targets = [os.path.join(_BUILD, "libastroz_oracle.so"), os.path.join(_BUILD, "libastroz_simd_baseline.so")]
    srcs = [os.path.join(_HERE, f) for f in ("astroz_oracle.c", "astroz_oracle.h", "simd_baseline.c", "Makefile")]
    newest_src = max(os.path.getmtime(s) for s in srcs if os.path.exists(s))
    if not force and all(os.path.exists(t) and os.path.getmtime(t) >= newest_src for t in targets):
        return
    subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)


class _Tle(C.Structure):
    _fields_ = [("satnum", C.c_uint32), ("epochYear", C.c_int), ("epochDay", C.c_double), ("epochJd", C.c_double),
                ("ndot", C.c_double), ("bstar", C.c_double), ("inclDeg", C.c_double), ("raanDeg", C.c_double),
                ("ecc", C.c_double), ("argpDeg", C.c_double), ("maDeg", C.c_double), ("nRevDay", C.c_double)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = os.path.join(_BUILD, "libastroz_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        dp = C.POINTER(C.c_double)
        L.azo_tle_parse.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(_Tle)]
        L.azo_tle_parse.restype = C.c_int
        L.azo_year_doy_to_jd.argtypes = [C.c_int, C.c_double]
        L.azo_year_doy_to_jd.restype = C.c_double
        for name in ("azo_sgp4_init", "azo_sdp4_init"):
            getattr(L, name).argtypes = [C.POINTER(_Tle), C.c_int, C.c_void_p]
            getattr(L, name).restype = C.c_int
        L.azo_sgp4_propagate.argtypes = [C.c_void_p, C.c_double, dp, dp]
        L.azo_sgp4_propagate.restype = None
        L.azo_sdp4_propagate.argtypes = [C.c_void_p, C.c_double, dp, dp]
        L.azo_sdp4_propagate.restype = C.c_int
        L.azo_sdp4_propagate_carry.argtypes = [C.c_void_p, C.c_double, dp, dp, dp]
        L.azo_sdp4_propagate_carry.restype = C.c_int
        L.azo_gstime.argtypes = [C.c_double]
        L.azo_gstime.restype = C.c_double
        L.azo_julian_to_gmst.argtypes = [C.c_double]
        L.azo_julian_to_gmst.restype = C.c_double
        L.azo_ecef_to_geodetic.argtypes = [dp, dp]
        L.azo_sgp4_export.argtypes = [C.c_void_p, dp]
        L.azo_sdp4_export.argtypes = [C.c_void_p, dp]
        L.azo_constellation_propagate.argtypes = [
            C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_size_t, C.c_int, dp, dp, C.c_size_t,
            dp, dp, C.c_int, C.c_int, C.POINTER(C.c_uint8), C.POINTER(C.c_int)]
        L.azo_constellation_propagate.restype = C.c_int
        L.azo_constellation_propagate_mt.argtypes = L.azo_constellation_propagate.argtypes + [C.c_int]
        L.azo_constellation_propagate_mt.restype = C.c_int
        L.azo_satrec_array_sgp4.argtypes = [
            C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_size_t, C.c_int, dp, dp, C.c_size_t, dp, dp]
        L.azo_satrec_array_sgp4.restype = C.c_int
        L.azo_screen_constellation.argtypes = [
            C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_size_t, C.c_int, dp, C.c_size_t, dp, C.c_size_t,
            C.c_double, C.c_double, dp, C.POINTER(C.c_uint32)]
        L.azo_screen_constellation.restype = C.c_int
        L.azo_coarse_screen.argtypes = [dp, C.c_size_t, C.c_size_t, C.c_double, C.POINTER(C.c_uint8),
                                        C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_size_t]
        L.azo_coarse_screen.restype = C.c_size_t
        _lib = L
    return _lib


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _lines(tles):
    n = len(tles)
    a1 = (C.c_char_p * n)(*[t[0].encode() for t in tles])
    a2 = (C.c_char_p * n)(*[t[1].encode() for t in tles])
    return a1, a2


def parse_tle(line1: str, line2: str) -> dict:
    t = _Tle()
    rc = lib().azo_tle_parse(line1.encode(), line2.encode(), C.byref(t))
    if rc != OK:
        raise ValueError("bad TLE")
    return {k: getattr(t, k) for k, _ in _Tle._fields_}


class Sgp4:
    """Scalar SGP4 oracle for one satellite (src/Sgp4.zig:99-106)."""

    def __init__(self, line1: str, line2: str, grav: int = WGS72):
        t = _Tle()
        if lib().azo_tle_parse(line1.encode(), line2.encode(), C.byref(t)) != OK:
            raise ValueError("bad TLE")
        self._buf = C.create_string_buffer(1024)
        self.rc = lib().azo_sgp4_init(C.byref(t), grav, self._buf)
        self.epochJd = t.epochJd
        if self.rc == OK:
            out = np.zeros(48)
            lib().azo_sgp4_export(self._buf, _dp(out))
            self.el = dict(zip(SGP4_FIELDS, out.tolist()))

    def propagate(self, tsince: float):
        r = np.zeros(3)
        v = np.zeros(3)
        lib().azo_sgp4_propagate(self._buf, float(tsince), _dp(r), _dp(v))
        return r, v


class Sdp4:
    """Scalar SDP4 oracle for one satellite (src/Sdp4.zig:152-172)."""

    def __init__(self, line1: str, line2: str, grav: int = WGS72):
        t = _Tle()
        if lib().azo_tle_parse(line1.encode(), line2.encode(), C.byref(t)) != OK:
            raise ValueError("bad TLE")
        self._buf = C.create_string_buffer(2048)
        self.rc = lib().azo_sdp4_init(C.byref(t), grav, self._buf)
        self.epochJd = t.epochJd
        if self.rc == OK:
            out = np.zeros(112)
            lib().azo_sdp4_export(self._buf, _dp(out))
            self.el = dict(zip(SGP4_FIELDS, out[:len(SGP4_FIELDS)].tolist()))
            self.el.update(dict(zip(SDP4_FIELDS, out[48:48 + len(SDP4_FIELDS)].tolist())))
        self.reset_carry()

    def reset_carry(self):
        self._carry = np.array([0.0, self.el["xlamo"], self.el["noUnkozai"]]) if self.rc == OK else np.zeros(3)

    def propagate(self, tsince: float):
        r = np.zeros(3)
        v = np.zeros(3)
        rc = lib().azo_sdp4_propagate(self._buf, float(tsince), _dp(r), _dp(v))
        return rc, r, v

    def propagate_carry(self, tsince: float):
        r = np.zeros(3)
        v = np.zeros(3)
        rc = lib().azo_sdp4_propagate_carry(self._buf, float(tsince), _dp(self._carry), _dp(r), _dp(v))
        return rc, r, v


def gstime(jd: float) -> float:
    return lib().azo_gstime(float(jd))


def julian_to_gmst(jd: float) -> float:
    return lib().azo_julian_to_gmst(float(jd))


def ecef_to_geodetic(ecef) -> np.ndarray:
    e = np.ascontiguousarray(ecef, dtype=np.float64)
    out = np.zeros(3)
    lib().azo_ecef_to_geodetic(_dp(e), _dp(out))
    return out


def host_threads() -> int:
    """Threads the host can really run: affinity mask narrowed by the cgroup CPU quota."""
    n = min(os.cpu_count() or 1, len(os.sched_getaffinity(0)))
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period) + 0.5)))
    except (OSError, ValueError):
        pass
    return max(1, n)


def constellation_propagate(tles, jd, fr, grav: int = WGS72, mode: int = 0, layout: int = 0, velocities: bool = True,
                            threads: int = 1):
    """Scalar oracle of Constellation.init + propagate (src/Constellation.zig:101-308).

    Returns (pos, vel, err[n, nt], klass[n]); pos/vel are shaped by `layout`
    (0: (n, nt, 3) satellite-major, 1: (nt, n, 3) time-major).  threads > 1 deals satellites to pthreads
    (identical arithmetic per cell; used for the full-size BASELINE grids); threads = 0 uses every usable CPU.
    """
    if threads == 0:
        threads = host_threads()
    jd = np.ascontiguousarray(jd, dtype=np.float64)
    fr = np.ascontiguousarray(fr, dtype=np.float64)
    n, nt = len(tles), len(jd)
    shape = (n, nt, 3) if layout == 0 else (nt, n, 3)
    pos = np.zeros(shape)
    vel = np.zeros(shape) if velocities else None
    err = np.zeros((n, nt), dtype=np.uint8)
    klass = np.zeros(n, dtype=np.int32)
    a1, a2 = _lines(tles)
    rc = lib().azo_constellation_propagate_mt(
        a1, a2, n, grav, _dp(jd), _dp(fr), nt, _dp(pos), _dp(vel) if velocities else None, mode, layout,
        err.ctypes.data_as(C.POINTER(C.c_uint8)), klass.ctypes.data_as(C.POINTER(C.c_int)), int(threads))
    if rc != 0:
        raise ValueError(f"oracle constellation init failed rc={rc}")
    return pos, vel, err, klass


def satrec_array_sgp4(tles, jd, fr, grav: int = WGS72):
    """Scalar oracle of SatrecArray.sgp4 for near-earth satellites (bindings/python/astroz/api.py:249-320)."""
    jd = np.ascontiguousarray(jd, dtype=np.float64)
    fr = np.ascontiguousarray(fr, dtype=np.float64)
    n, nt = len(tles), len(jd)
    pos = np.zeros((n, nt, 3))
    vel = np.zeros((n, nt, 3))
    a1, a2 = _lines(tles)
    rc = lib().azo_satrec_array_sgp4(a1, a2, n, grav, _dp(jd), _dp(fr), nt, _dp(pos), _dp(vel))
    if rc != 0:
        raise ValueError(f"oracle satrec_array init failed rc={rc}")
    return pos, vel


def screen_constellation(tles, times, epoch_offsets, target: int, threshold: float, reference_jd: float = 0.0,
                         grav: int = WGS72):
    """Scalar oracle of Constellation.screenConstellation (src/Constellation.zig:683-756)."""
    times = np.ascontiguousarray(times, dtype=np.float64)
    off = np.ascontiguousarray(epoch_offsets, dtype=np.float64)
    n = len(tles)
    dist = np.zeros(n)
    tidx = np.zeros(n, dtype=np.uint32)
    a1, a2 = _lines(tles)
    rc = lib().azo_screen_constellation(a1, a2, n, grav, _dp(times), len(times), _dp(off), target, float(threshold),
                                        float(reference_jd), _dp(dist), tidx.ctypes.data_as(C.POINTER(C.c_uint32)))
    if rc != 0:
        raise ValueError(f"oracle screen failed rc={rc}")
    return dist, tidx


def coarse_screen(positions_sat_major, threshold: float, valid_mask=None, max_results: int = 10_000_000):
    """Oracle of coarseScreen (bindings/python/src/conjunction.zig:11-149) on a (n_sats, n_times, 3) block.
    Returns (pairs[n, 2], t_indices[n]) sorted by (t, s, other)."""
    pos = np.ascontiguousarray(positions_sat_major, dtype=np.float64)
    ns, nt = pos.shape[0], pos.shape[1]
    pairs = np.zeros((max_results, 2), dtype=np.uint32)
    tidx = np.zeros(max_results, dtype=np.uint32)
    m = None if valid_mask is None else np.ascontiguousarray(valid_mask, dtype=np.uint8)
    k = lib().azo_coarse_screen(_dp(pos), ns, nt, float(threshold),
                                m.ctypes.data_as(C.POINTER(C.c_uint8)) if m is not None else None,
                                pairs.ctypes.data_as(C.POINTER(C.c_uint32)), tidx.ctypes.data_as(C.POINTER(C.c_uint32)),
                                max_results)
    pairs, tidx = pairs[:k], tidx[:k]
    order = np.lexsort((pairs[:, 1], pairs[:, 0], tidx)) if k else np.zeros(0, dtype=np.int64)
    return pairs[order], tidx[order]


# ---------------------------------------------------------------------------------------------------
# CPU SIMD baseline (oracle/simd_baseline.c): restatement of the reference's 8-lane batch path, timed
# by bench.py as cpu_baseline / --impl reference.  Never used by the product.
# ---------------------------------------------------------------------------------------------------
_simd = None


def simd_lib() -> C.CDLL:
    global _simd
    if _simd is None:
        path = os.path.join(_BUILD, "libastroz_simd_baseline.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        dp = C.POINTER(C.c_double)
        L.azo_simd_create.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_size_t, C.c_int]
        L.azo_simd_create.restype = C.c_void_p
        L.azo_simd_free.argtypes = [C.c_void_p]
        L.azo_simd_propagate.argtypes = [C.c_void_p, dp, dp, C.c_size_t, dp, dp, C.c_int, C.c_int]
        L.azo_simd_propagate.restype = C.c_int
        L.azo_simd_propagate2.argtypes = [C.c_void_p, dp, dp, C.c_size_t, dp, dp, C.c_int, C.c_int, C.c_int]
        L.azo_simd_propagate2.restype = C.c_int
        L.azo_simd_counts.argtypes = [C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.azo_simd_counts.restype = None
        L.azo_simd_isa.restype = C.c_char_p
        _simd = L
    return _simd


class SimdConstellation:
    """Mixed SGP4/SDP4 constellation on the CPU SIMD baseline (src/Constellation.zig:101-200,245-476)."""

    def __init__(self, tles, grav: int = WGS72):
        a1, a2 = _lines(tles)
        self.n = len(tles)
        self._h = simd_lib().azo_simd_create(a1, a2, self.n, grav)
        if not self._h:
            raise ValueError("SIMD baseline: init failed (invalid element set in the catalog)")
        n, ns, nd = C.c_size_t(), C.c_size_t(), C.c_size_t()
        simd_lib().azo_simd_counts(self._h, C.byref(n), C.byref(ns), C.byref(nd))
        self.numSgp4, self.numSdp4 = ns.value, nd.value

    def propagate(self, jd, fr, layout: int = 1, velocities: bool = True, threads: int | None = None, out=None,
                  sdp4_threads: int = 0):
        """sdp4_threads = 0: the reference's own rule (the deep-space phase gets only the threads the near-earth
        phase left over -- one, when that phase used them all, src/Constellation.zig:358-364); > 0: that many."""
        jd = np.ascontiguousarray(jd, dtype=np.float64)
        fr = np.ascontiguousarray(fr, dtype=np.float64)
        nt = len(jd)
        shape = (self.n, nt, 3) if layout == 0 else (nt, self.n, 3)
        if out is None:
            pos = np.empty(shape)
            vel = np.empty(shape) if velocities else None
        else:
            pos, vel = out
        threads = threads or (os.cpu_count() or 1)   # getMaxThreads, Constellation.zig:61-74
        env = os.environ.get("ASTROZ_THREADS")
        if env:
            threads = int(env)
        simd_lib().azo_simd_propagate2(self._h, _dp(jd), _dp(fr), nt, _dp(pos), _dp(vel) if vel is not None else None,
                                       layout, threads, int(sdp4_threads))
        return pos, vel

    def __del__(self):
        if getattr(self, "_h", None):
            simd_lib().azo_simd_free(self._h)
            self._h = None


def simd_isa() -> str:
    return simd_lib().azo_simd_isa().decode()

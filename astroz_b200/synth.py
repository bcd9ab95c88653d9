"""Synthetic workloads of BASELINE.json (recipes in SURVEY.md section 8d / BASELINE.md section 4).

The reference's 13,478-satellite benchmark catalog is a live CelesTrak download
(benchmarks/sgp4_compat_test.py:79-97), not a fixture, so the bench and the tests use a seeded
synthetic catalog with the same size and orbit-class mix, emitted as real 69-column TLE lines so the
parser (src/Tle.zig:49-101) is exercised end to end.
"""
from __future__ import annotations

import math

import numpy as np

MU_KM3_S2 = 398600.8          # WGS72, src/constants.zig:41-50
RE_KM = 6378.135
HEADLINE_SATS = 13478         # README.md:35-45
HEADLINE_TIMES = 1440
BENCH_JD0 = 2460437.5         # 2024-05-07 00:00 UTC


def _checksum(line68: str) -> int:
    s = 0
    for ch in line68:
        if ch.isdigit():
            s += int(ch)
        elif ch == "-":
            s += 1
    return s % 10


def _exp_field(x: float) -> str:
    """TLE 'assumed decimal point' field, 8 chars: sign, 5 mantissa digits, signed exponent digit."""
    if x == 0.0:
        return " 00000+0"
    sign = "-" if x < 0 else " "
    ax = abs(x)
    e = int(math.floor(math.log10(ax))) + 1
    m = int(round(ax / 10.0 ** e * 1e5))
    if m >= 100000:
        m //= 10
        e += 1
    if e < -9:
        return " 00000+0"
    if e > 9:
        raise ValueError("bstar out of TLE range")
    return f"{sign}{m:05d}{'+' if e >= 0 else '-'}{abs(e)}"


def tle_lines(satnum: int, epoch_yy: int, epoch_doy: float, incl_deg: float, raan_deg: float, ecc: float,
              argp_deg: float, ma_deg: float, n_rev_day: float, bstar: float, ndot: float = 0.0) -> tuple[str, str]:
    """Format one element set as two 69-column TLE lines (columns as parsed by src/Tle.zig:49-101)."""
    nd = f"{ndot:.8f}"
    nd = (nd[0] if nd[0] == "-" else " ") + nd.lstrip("-").lstrip("0")  # ' .00015698'
    l1 = (f"1 {satnum % 100000:05d}U 24001A   {epoch_yy:02d}{epoch_doy:012.8f} {nd:>10s}  00000+0 "
          f"{_exp_field(bstar)} 0  999")
    assert len(l1) == 68, (len(l1), l1)
    e7 = int(round(ecc * 1e7))
    e7 = min(max(e7, 0), 9999999)
    l2 = (f"2 {satnum % 100000:05d} {incl_deg % 180.0:8.4f} {raan_deg % 360.0:8.4f} {e7:07d} "
          f"{argp_deg % 360.0:8.4f} {ma_deg % 360.0:8.4f} {n_rev_day:11.8f}{satnum % 100000:5d}")
    assert len(l2) == 68, (len(l2), l2)
    return l1 + str(_checksum(l1)), l2 + str(_checksum(l2))


def _rev_per_day(a_km):
    return 86400.0 / (2.0 * math.pi) * np.sqrt(MU_KM3_S2 / a_km ** 3)


def _leo_elements(n: int, rng: np.random.Generator):
    """Config-2 shell mix: arrays (incl, raan, ecc, argp, ma, n_rev_day, bstar, epoch_doy)."""
    kind = rng.choice(6, size=n, p=[0.60, 0.15, 0.10, 0.10, 0.03, 0.02])
    incl = np.empty(n)
    a = np.empty(n)
    ecc = np.exp(rng.uniform(math.log(1e-5), math.log(3e-3), n))
    for k in range(6):
        m = kind == k
        c = int(m.sum())
        if k == 0:
            incl[m] = 53.0 + rng.uniform(-0.2, 0.2, c)
            a[m] = RE_KM + rng.uniform(540.0, 570.0, c)
        elif k == 1:
            incl[m] = rng.uniform(97.4, 98.0, c)
            a[m] = RE_KM + rng.uniform(480.0, 620.0, c)
        elif k == 2:
            incl[m] = 87.9 + rng.uniform(-0.1, 0.1, c)
            a[m] = RE_KM + 1200.0 + rng.uniform(-5.0, 5.0, c)
        elif k == 3:
            incl[m] = rng.uniform(0.0, 100.0, c)
            a[m] = RE_KM + rng.uniform(300.0, 2000.0, c)
        elif k == 4:  # low perigee -> simplified drag branch (isimp, src/Sgp4.zig:400)
            incl[m] = rng.uniform(28.0, 99.0, c)
            e = rng.uniform(0.001, 0.02, c)
            ecc[m] = e
            a[m] = (RE_KM + rng.uniform(160.0, 215.0, c)) / (1.0 - e)
        else:         # eccentric, still near-earth (period <= 225 min)
            incl[m] = rng.uniform(5.0, 110.0, c)
            e = rng.uniform(0.02, 0.25, c)
            ecc[m] = e
            a[m] = (RE_KM + rng.uniform(300.0, 800.0, c)) / (1.0 - e)
    raan = rng.uniform(0.0, 360.0, n)
    argp = rng.uniform(0.0, 360.0, n)
    ma = rng.uniform(0.0, 360.0, n)
    bstar = np.exp(rng.uniform(math.log(1e-6), math.log(1e-3), n))
    doy = 127.5 - rng.uniform(0.0, 5.0, n)  # 2024-05-06 12:00 UTC minus U[0,5] days
    return incl, raan, ecc, argp, ma, _rev_per_day(a), bstar, doy


def near_earth_catalog(n: int = HEADLINE_SATS, seed: int = 13478) -> list[tuple[str, str]]:
    """BASELINE config 2: n near-earth element sets."""
    rng = np.random.default_rng(seed)
    cols = _leo_elements(n, rng)
    return [tle_lines(10000 + i, 24, *(float(c[i]) for c in (cols[7], cols[0], cols[1], cols[2], cols[3], cols[4],
                                                             cols[5], cols[6]))) for i in range(n)]


def mixed_catalog(n: int = HEADLINE_SATS, seed: int = 28626, n_geo: int = 1024, n_molniya: int = 256,
                  n_gps: int = 256) -> list[tuple[str, str]]:
    """BASELINE config 3: config-2 slots with GEO / Molniya / GPS-like deep-space objects interleaved."""
    base = near_earth_catalog(n, 13478)
    rng = np.random.default_rng(seed)
    n_ds = min(n_geo + n_molniya + n_gps, n)
    scale = n_ds / float(n_geo + n_molniya + n_gps)
    n_geo, n_molniya = int(n_geo * scale), int(n_molniya * scale)
    n_gps = n_ds - n_geo - n_molniya
    ds = []
    for i in range(n_geo):       # irez = 1
        ds.append(tle_lines(40000 + i, 24, 127.5 - rng.uniform(0, 5), rng.uniform(0.0, 15.0), rng.uniform(0, 360),
                            rng.uniform(0.0, 1e-3), rng.uniform(0, 360), rng.uniform(0, 360),
                            1.0027 + rng.uniform(-0.0005, 0.0005), 0.0))
    for i in range(n_molniya):   # irez = 2
        ds.append(tle_lines(42000 + i, 24, 127.5 - rng.uniform(0, 5), 63.4 + rng.uniform(-0.3, 0.3),
                            rng.uniform(0, 360), rng.uniform(0.68, 0.74), rng.uniform(0, 360), rng.uniform(0, 360),
                            2.006 + rng.uniform(-0.002, 0.002), 1e-4))
    for i in range(n_gps):       # irez = 0
        ds.append(tle_lines(43000 + i, 24, 127.5 - rng.uniform(0, 5), 55.0 + rng.uniform(-1.0, 1.0),
                            rng.uniform(0, 360), rng.uniform(0.0, 0.02), rng.uniform(0, 360), rng.uniform(0, 360),
                            2.0056 + rng.uniform(-0.0005, 0.0005), 1e-4))
    out = list(base)
    out[:n_ds] = ds
    perm = rng.permutation(n)
    return [out[j] for j in perm]


def time_grid(n_times: int = HEADLINE_TIMES, jd0: float = BENCH_JD0) -> tuple[np.ndarray, np.ndarray]:
    """1-minute steps: jd = jd0, fr = i/1440 (README.md:35)."""
    return np.full(n_times, jd0, dtype=np.float64), np.arange(n_times, dtype=np.float64) / 1440.0


def monte_carlo_catalog(n: int = 10000, seed: int = 12345, base: tuple[str, str] | None = None) -> list[tuple[str, str]]:
    """BASELINE config 5: n perturbed element sets of one object (ISS TLE of src/Sgp4.zig:909-910 by default).
    Additive Gaussian draws, seed 12345 (the seed of the reference's Monte-Carlo test, src/MonteCarlo.zig:293):
    sigma_i = sigma_RAAN = sigma_argp = sigma_M = 0.01 deg, sigma_e = 1e-5 (floored at 1e-7), sigma_n = 1e-5 rev/day,
    sigma_B* = 10 %.  There is no reference code path for perturbed-TLE draws (MonteCarlo.zig samples Hohmann
    transfers); the workload is defined here."""
    l1, l2 = base or ("1 25544U 98067A   24127.82853009  .00015698  00000+0  27310-3 0  9995",
                      "2 25544  51.6393 160.4574 0003580 140.6673 205.7250 15.50957674452123")
    incl, raan, ecc = float(l2[8:16]), float(l2[17:25]), float(l2[26:33]) / 1e7
    argp, ma, nn = float(l2[34:42]), float(l2[43:51]), float(l2[52:63])
    bstar = float(l1[53:59]) * 1e-5 * 10.0 ** int(l1[59:61])
    yy, doy = int(l1[18:20]), float(l1[20:32])
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        out.append(tle_lines(
            20000 + i, yy, doy, incl + rng.normal(0, 0.01), raan + rng.normal(0, 0.01),
            max(ecc + rng.normal(0, 1e-5), 1e-7), argp + rng.normal(0, 0.01), ma + rng.normal(0, 0.01),
            nn + rng.normal(0, 1e-5), bstar * (1.0 + rng.normal(0, 0.1))))
    return out


def elements_from_tles(tles) -> np.ndarray:
    """(8, n) float64 element columns in the order of `Constellation.from_elements` -- epoch_jd, mean_motion_rev_day,
    ecc, incl_deg, raan_deg, argp_deg, ma_deg, bstar -- read from the fixed TLE columns (src/Tle.zig:49-101)."""
    out = np.empty((8, len(tles)))
    for i, (l1, l2) in enumerate(tles):
        yy = int(l1[18:20])
        year = 2000 + yy if yy < 57 else 1900 + yy
        jan0 = 367.0 * year - math.floor(7 * (year + 0) / 4) + 30 + 1721013.5  # JD of <year> Jan 0.0 (1901..2099)
        out[0, i] = jan0 + float(l1[20:32])
        out[1, i] = float(l2[52:63])
        out[2, i] = float(l2[26:33]) * 1e-7
        out[3, i] = float(l2[8:16])
        out[4, i] = float(l2[17:25])
        out[5, i] = float(l2[34:42])
        out[6, i] = float(l2[43:51])
        out[7, i] = float(l1[53:59]) * 1e-5 * 10.0 ** int(l1[59:61])
    return out

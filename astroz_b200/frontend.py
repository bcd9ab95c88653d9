"""High-level entry points of the reference's Python package over the CUDA path:

    from astroz_b200.frontend import Constellation, propagate, screen     # was: from astroz import ...

Mirrors `astroz.Constellation`, `astroz.propagate` and `astroz.screen`
(bindings/python/astroz/__init__.py:305-660): same arguments, same output shapes (time-major
`(n_times, n_satellites, 3)`), same row order (near-earth satellites first, deep-space after, :373-375), the
same time convention (`times` in minutes from `start_time`, near-earth tsince = times[t] + (start - epoch) * 1440,
:513-526) and the same loaders for TLE text and OMM JSON.  Differences, all deliberate:

* sources that need the network (CelesTrak group names, URLs, `norad_id=`) raise: this build has no egress and the
  catalogue download is outside the propagation path;
* the deep-space rows of `propagate()` are filled (the reference leaves them uninitialised);
* geodetic output is what the core produces -- latitude / longitude in radians, altitude in km
  (src/Constellation.zig:497) -- like the reference's actual return value, not its docstring.
"""
from __future__ import annotations

import json
import math
from datetime import datetime, timezone
from pathlib import Path

import numpy as np

from ._lib import WGS72
from .constellation import Constellation as DeviceConstellation
from .constellation import OutputMode

_OUTPUT = {"teme": OutputMode.teme, "ecef": OutputMode.ecef, "geodetic": OutputMode.geodetic}


def _start_jd(start_time) -> float:
    """`start_time` (datetime, default: now, UTC) as a Julian date (__init__.py:282-286)."""
    if start_time is None:
        start_time = datetime.now(timezone.utc)
    return 2440587.5 + start_time.timestamp() / 86400.0


def parse_tle_pairs(text: str) -> list[tuple[str, str]]:
    """(line1, line2) pairs of a 2- or 3-line element-set listing; name lines and orphans are skipped
    (__init__.py:184-200)."""
    rows = [ln.strip() for ln in text.strip().splitlines() if ln.strip()]
    pairs, i = [], 0
    while i < len(rows):
        if rows[i].startswith("1 ") and i + 1 < len(rows) and rows[i + 1].startswith("2 "):
            pairs.append((rows[i], rows[i + 1]))
            i += 2
        else:
            i += 1
    return pairs


def _checksum(body: str) -> str:
    return str(sum(int(ch) if ch.isdigit() else (1 if ch == "-" else 0) for ch in body) % 10)


def _implied_decimal(x: float) -> str:
    """8-column TLE field with an implied leading decimal point and a signed power-of-ten digit."""
    if x == 0:
        return " 00000+0"
    mag = abs(x)
    exp = math.floor(math.log10(mag)) + 1
    digits = int(round(mag * 10.0 ** (5 - exp)))
    return f"{'-' if x < 0 else ' '}{digits:05d}{exp:+d}"


def omm_to_tle_pairs(json_text: str) -> list[tuple[str, str]]:
    """OMM JSON (one record or an array) rendered as TLE line pairs, with TLE column precision -- the reference
    converts OMM input the same way before initialising (__init__.py:203-279), so an OMM catalogue propagates to
    the same values here as there."""
    data = json.loads(json_text)
    records = [data] if isinstance(data, dict) else data
    pairs = []
    for rec in records:
        norad = int(rec["NORAD_CAT_ID"])
        cls = (rec.get("CLASSIFICATION_TYPE") or "U")[0]
        obj = rec.get("OBJECT_ID") or "00000A"
        if "-" in obj:  # "1998-067A" -> "98067A"
            launch_year, piece = obj.split("-", 1)
            designator = f"{launch_year[-2:]}{piece:<6s}"
        else:
            designator = f"{obj:<8s}"
        when = datetime.fromisoformat(rec["EPOCH"]).replace(tzinfo=None)
        doy = (when - datetime(when.year, 1, 1)).total_seconds() / 86400.0 + 1.0
        ndot = rec.get("MEAN_MOTION_DOT") or 0
        ndot_txt = ("-" if ndot < 0 else " ") + f"{abs(ndot):.8f}"[1:]
        head = (f"1 {norad:05d}{cls} {designator} {when.year % 100:02d}{doy:012.8f} {ndot_txt} "
                f"{_implied_decimal(rec.get('MEAN_MOTION_DDOT') or 0)} {_implied_decimal(rec['BSTAR'])} "
                f"{rec.get('EPHEMERIS_TYPE') or 0} {rec.get('ELEMENT_SET_NO') or 0:4d}")
        head = head[:68].ljust(68)
        ecc_txt = f"{rec['ECCENTRICITY']:.7f}"[2:]
        tail = (f"2 {norad:05d} {rec['INCLINATION']:8.4f} {rec['RA_OF_ASC_NODE']:8.4f} {ecc_txt} "
                f"{rec['ARG_OF_PERICENTER']:8.4f} {rec['MEAN_ANOMALY']:8.4f} {rec['MEAN_MOTION']:11.8f}"
                f"{rec.get('REV_AT_EPOCH') or 0:5d}")
        tail = tail[:68].ljust(68)
        pairs.append((head + _checksum(head), tail + _checksum(tail)))
    return pairs


def _load(source, norad_id):
    """(text, "tle" | "json") for the offline sources of __init__.py:163-181."""
    if norad_id is not None:
        raise RuntimeError("norad_id= needs a CelesTrak download; this build has no network access")
    if source is None:
        raise ValueError("Must specify 'source' or 'norad_id'")
    if source.startswith(("http://", "https://")):
        raise RuntimeError("URL sources need network access, which this build does not have")
    if "1 " in source and "2 " in source:
        return source, "tle"
    if source.lstrip().startswith(("[", "{")):
        return source, "json"
    if Path(source).exists():
        text = Path(source).read_text()
        return text, ("json" if text.lstrip().startswith(("[", "{")) else "tle")
    raise RuntimeError(f"'{source}' is not TLE text, OMM JSON or a file; CelesTrak group names need network access")


class Constellation:
    """Pre-parsed element sets for repeated propagation / screening (`astroz.Constellation`, __init__.py:305-408).
    Near-earth members occupy rows [0, n_sgp4), deep-space members the rows after them."""

    def __init__(self, source=None, *, norad_id=None, device: int = 0):
        text, fmt = _load(source, norad_id)
        pairs = omm_to_tle_pairs(text) if fmt == "json" else parse_tle_pairs(text)
        self._total_sats = len(pairs)
        if not pairs:
            raise ValueError("no element sets found in source")
        probe = DeviceConstellation(pairs, WGS72, device)
        deep = probe.classes != 0
        probe.deinit()
        ordered = [p for p, d in zip(pairs, deep) if not d] + [p for p, d in zip(pairs, deep) if d]
        self._dev = DeviceConstellation(ordered, WGS72, device)
        self._n_sgp4 = self._dev.numSgp4
        self._n_sdp4 = self._dev.numSdp4

    @property
    def num_satellites(self) -> int:
        return self._total_sats

    @property
    def epochs(self) -> list:
        return self._dev.epochs.tolist()


def _as_constellation(source, norad_id) -> Constellation:
    return source if isinstance(source, Constellation) else Constellation(source, norad_id=norad_id)


def propagate(source, times, *, start_time=None, output="ecef", velocities=False, norad_id=None):
    """`astroz.propagate(source, times, start_time=None, output="ecef", velocities=False)` (__init__.py:411-532):
    positions `(n_times, n_satellites, 3)` (and velocities, km/s) at `times` minutes from `start_time`."""
    const = _as_constellation(source, norad_id)
    times = np.ascontiguousarray(times, dtype=np.float64)
    if output not in _OUTPUT:
        raise ValueError("output must be 'ecef', 'teme' or 'geodetic'")  # sgp4.zig:198-206
    mode = _OUTPUT[output]
    n_sats, nt = const.num_satellites, len(times)
    start = _start_jd(start_time)
    pos = np.empty((nt, n_sats, 3))
    vel = np.empty((nt, n_sats, 3)) if velocities else None
    dev = const._dev
    if const._n_sgp4 and nt:
        offsets = (start - dev.epochs[:const._n_sgp4]) * 1440.0  # __init__.py:513-514
        dev.propagate_into(times, pos, vel, epoch_offsets=offsets, outputMode=mode, reference_jd=start,
                           time_major=True, output_stride=n_sats, want_velocities=velocities)
    if const._n_sdp4 and nt:
        jd = np.full(nt, start)
        dev.propagate_sdp4_into(jd, times / 1440.0, pos, vel, outputMode=mode, time_major=True, output_stride=n_sats,
                                sat_offset=const._n_sgp4)
    return (pos, vel) if velocities else pos


def screen(source, times, threshold=10.0, *, target=None, start_time=None, norad_id=None):
    """`astroz.screen(source, times, threshold=10.0, target=None, start_time=None)` (__init__.py:535-660).
    With `target`: (min_distances[n], min_t_indices[n]) from the fused propagate + screen kernel.  Without: all-vs-all
    (pairs[n, 2], t_indices[n]) from the cell-list kernel; the position block never leaves the GPU."""
    const = _as_constellation(source, norad_id)
    times = np.ascontiguousarray(times, dtype=np.float64)
    dev = const._dev
    start = _start_jd(start_time)
    if const._n_sdp4 == 0:
        offsets = (start - dev.epochs) * 1440.0
        if target is not None:
            dist, tidx = dev.screen_conjunction(times, int(target), float(threshold), epoch_offsets=offsets,
                                                reference_jd=start)
            return np.asarray(dist), np.asarray(tidx, dtype=np.uint32)
        return dev.screen_all(times, float(threshold), epoch_offsets=offsets)
    # mixed catalogue: propagate everything (TEME), then the cell-list screen on the device copy of the block
    import torch
    pos = propagate(const, times, start_time=start_time, output="teme")
    block = torch.from_numpy(pos).cuda(dev.device)
    return dev.coarse_screen_device(block, float(threshold))


__all__ = ["Constellation", "propagate", "screen", "parse_tle_pairs", "omm_to_tle_pairs"]

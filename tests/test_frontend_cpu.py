"""Host-side logic of the high-level front end (no GPU): element-set text handling of
bindings/python/astroz/__init__.py:163-279."""
import json

import pytest

from astroz_b200 import frontend
from tests.golden import tles as T


def test_parse_tle_pairs_skips_names_and_orphans():
    text = "\n".join(["ISS (ZARYA)", T.ISS[0], T.ISS[1], "", "2 99999 orphan line two", "DEB", T.SAT55909[0],
                      T.SAT55909[1], "1 00001 dangling line one"])
    pairs = frontend.parse_tle_pairs(text)
    assert pairs == [T.ISS, T.SAT55909]


def test_omm_record_renders_the_same_tle_columns():
    rec = {"OBJECT_NAME": "ISS (ZARYA)", "OBJECT_ID": "1998-067A", "EPOCH": "2024-05-06T19:53:05.000000",
           "MEAN_MOTION": 15.50957674, "ECCENTRICITY": 0.000358, "INCLINATION": 51.6393, "RA_OF_ASC_NODE": 160.4574,
           "ARG_OF_PERICENTER": 140.6673, "MEAN_ANOMALY": 205.725, "EPHEMERIS_TYPE": 0, "CLASSIFICATION_TYPE": "U",
           "NORAD_CAT_ID": 25544, "ELEMENT_SET_NO": 999, "REV_AT_EPOCH": 45212, "BSTAR": 0.0002731,
           "MEAN_MOTION_DOT": 0.00015698, "MEAN_MOTION_DDOT": 0}
    for payload in (json.dumps(rec), json.dumps([rec, rec])):
        pairs = frontend.omm_to_tle_pairs(payload)
        l1, l2 = pairs[0]
        assert len(l1) == 69 and len(l2) == 69
        # the columns the propagator reads (src/Tle.zig:49-101) carry the OMM values at TLE precision
        assert l1[2:7] == "25544" and l1[18:20] == "24" and abs(float(l1[20:32]) - 127.82853009) < 1e-7
        assert l1[33:43].strip() == ".00015698" and l1[53:61] == " 27310-3"
        assert l2[8:16] == T.ISS[1][8:16] and l2[17:25] == T.ISS[1][17:25] and l2[26:33] == "0003580"
        assert l2[34:42] == T.ISS[1][34:42] and l2[43:51] == T.ISS[1][43:51] and l2[52:63] == T.ISS[1][52:63]
        assert l1[68] == frontend._checksum(l1[:68]) and l2[68] == frontend._checksum(l2[:68])
    assert len(frontend.omm_to_tle_pairs(json.dumps([rec, rec]))) == 2


def test_network_sources_are_refused_loudly():
    for kwargs in ({"source": "starlink"}, {"source": "https://celestrak.org/x"}, {"source": None, "norad_id": 25544}):
        with pytest.raises(RuntimeError):
            frontend._load(kwargs.get("source"), kwargs.get("norad_id"))
    with pytest.raises(ValueError):
        frontend._load(None, None)

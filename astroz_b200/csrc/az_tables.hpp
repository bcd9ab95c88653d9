// az_tables.hpp -- host code that turns prepared elements into the device tables (product code).
//
// Replaces the reflection transposes of the reference: src/Sgp4Batch.zig:78-110 (initBatchElements) and
// src/Sdp4Batch.zig:147-194 (initFromElements), plus the classification loop of
// src/Constellation.zig:101-200.
#pragma once

#include <algorithm>
#include <cstdint>
#include <string>
#include <thread>
#include <vector>

#include "az_device.cuh"
#include "az_elements.hpp"

namespace az {

AZ_EHD inline GravConsts grav_consts(const Gravity &g) {
    return GravConsts{g.j2, g.radiusEarthKm, g.xke * g.radiusEarthKm / 60.0, g.j3oj2, g.xke, 0.5 * g.j2};
}

// column values of one near-earth satellite, in Sgp4Col order
AZ_EHD inline void sgp4_columns(const NearEarth &e, double *c) {
    c[kMo] = e.mo; c[kMdot] = e.mdot; c[kArgpo] = e.argpo; c[kArgpdot] = e.argpdot;
    c[kNodeo] = e.nodeo; c[kNodedot] = e.nodedot; c[kXnodcf] = e.xnodcf; c[kCc1] = e.cc1;
    c[kBc4] = e.bstar * e.cc4;  // only ever used as bstar*cc4 (src/Sgp4Batch.zig:126)
    c[kT2cof] = e.t2cof; c[kOmgcof] = e.omgcof; c[kEta] = e.eta; c[kXmcof] = e.xmcof; c[kDelmo] = e.delmo;
    c[kD2] = e.d2; c[kD3] = e.d3; c[kD4] = e.d4;
    c[kBc5] = e.bstar * e.cc5;  // src/Sgp4Batch.zig:144
    c[kSinmao] = e.sinmao; c[kT3cof] = e.t3cof; c[kT4cof] = e.t4cof; c[kT5cof] = e.t5cof;
    c[kAbase] = e.aBase; c[kEcco] = e.ecco; c[kNo] = e.no; c[kAycof] = e.aycof; c[kXlcof] = e.xlcof;
    c[kCon41] = e.con41; c[kX1mth2] = e.x1mth2; c[kX7thm1] = e.x7thm1; c[kSinio] = e.sinio; c[kCosio] = e.cosio;
    c[kIsimp] = e.isimp ? 1.0 : 0.0;
    c[kMrtA] = -1.5 * e.con41; c[kMrtB] = 0.5 * e.x1mth2; c[kDsuK] = -0.25 * e.x7thm1;
    c[kNodeK] = 1.5 * e.cosio; c[kDincK] = 1.5 * e.cosio * e.sinio; c[kRvK] = 1.5 * e.con41;
}

AZ_EHD inline Sdp4Sat sdp4_record(const DeepSpace &d) {
    const NearEarth &e = d.ne;
    Sdp4Sat r{};
    r.mo = e.mo; r.mdot = e.mdot; r.argpo = e.argpo; r.argpdot = e.argpdot; r.nodeo = e.nodeo; r.nodedot = e.nodedot;
    r.xnodcf = e.xnodcf; r.cc1 = e.cc1; r.bc4 = e.bstar * e.cc4; r.t2cof = e.t2cof; r.ecco = e.ecco; r.no = e.no;
    r.inclo = e.inclo;
    r.se2 = d.sun.e2; r.se3 = d.sun.e3; r.si2 = d.sun.i2; r.si3 = d.sun.i3; r.sl2 = d.sun.l2; r.sl3 = d.sun.l3;
    r.sl4 = d.sun.l4; r.sgh2 = d.sun.gh2; r.sgh3 = d.sun.gh3; r.sgh4 = d.sun.gh4; r.sh2 = d.sun.h2; r.sh3 = d.sun.h3;
    r.ee2 = d.moon.e2; r.e3 = d.moon.e3; r.xi2 = d.moon.i2; r.xi3 = d.moon.i3; r.xl2 = d.moon.l2; r.xl3 = d.moon.l3;
    r.xl4 = d.moon.l4; r.xgh2 = d.moon.gh2; r.xgh3 = d.moon.gh3; r.xgh4 = d.moon.gh4; r.xh2 = d.moon.h2;
    r.xh3 = d.moon.h3;
    r.zmol = d.zmol; r.zmos = d.zmos; r.dedt = d.dedt; r.didt = d.didt; r.dmdt = d.dmdt; r.domdt = d.domdt;
    r.dnodt = d.dnodt;
    {   // resonance coefficients folded on their basis angles (resonance_accel, az_device.cuh); phases of src/Sdp4.zig:826-857
        const double g22 = 5.7686396, g32 = 0.95240898, g44 = 1.8014998, g52 = 1.0508330, g54 = 4.4108898;
        const double fasx2 = 0.13130908, fasx4 = 2.8843198, fasx6 = 0.37448087;
        for (int b = 0; b < 8; ++b) r.rp[b] = r.rq[b] = 0.0;
        if (d.irez == 2) {
            r.rp[0] = d.d2201 * std::cos(g22);  r.rq[0] = -d.d2201 * std::sin(g22);   // 2w + l
            r.rp[1] = d.d2211 * std::cos(g22);  r.rq[1] = -d.d2211 * std::sin(g22);   // l
            r.rp[2] = d.d3210 * std::cos(g32) + d.d5220 * std::cos(g52);              // w + l
            r.rq[2] = -(d.d3210 * std::sin(g32) + d.d5220 * std::sin(g52));
            r.rp[3] = d.d3222 * std::cos(g32) + d.d5232 * std::cos(g52);              // l - w
            r.rq[3] = -(d.d3222 * std::sin(g32) + d.d5232 * std::sin(g52));
            r.rp[4] = d.d4410 * std::cos(g44);  r.rq[4] = -d.d4410 * std::sin(g44);   // 2w + 2l
            r.rp[5] = d.d4422 * std::cos(g44);  r.rq[5] = -d.d4422 * std::sin(g44);   // 2l
            r.rp[6] = d.d5421 * std::cos(g54);  r.rq[6] = -d.d5421 * std::sin(g54);   // w + 2l
            r.rp[7] = d.d5433 * std::cos(g54);  r.rq[7] = -d.d5433 * std::sin(g54);   // 2l - w
        } else if (d.irez == 1) {
            r.rp[0] = d.del1 * std::cos(fasx2);        r.rq[0] = -d.del1 * std::sin(fasx2);        // l
            r.rp[1] = d.del2 * std::cos(2.0 * fasx4);  r.rq[1] = -d.del2 * std::sin(2.0 * fasx4);  // 2l
            r.rp[2] = d.del3 * std::cos(3.0 * fasx6);  r.rq[2] = -d.del3 * std::sin(3.0 * fasx6);  // 3l
        }
    }
    r.xlamo = d.xlamo; r.xfact = d.xfact; r.gsto = d.gsto;
    r.abase = e.aBase; r.invNo = 1.0 / e.no; r.sinio = e.sinio; r.cosio = e.cosio;
    r.epochJd = e.epochJd;
    r.irez = d.irez;
    return r;
}

// Host image of a classified constellation (src/Constellation.zig:78-96): near-earth satellites in
// padded 8-wide tiles, deep-space satellites as records, each list with its original catalog index.
struct CatalogTables {
    uint32_t n = 0, nSgp4 = 0, nSdp4 = 0;
    Gravity grav{};
    double referenceEpochJd = 0.0;             // epoch of the first near-earth satellite (:139-140)
    std::vector<double> sgp4Tiles;             // [tiles][kSgp4Cols][8]
    std::vector<double> sgp4Epoch;             // padded, per near-earth satellite
    std::vector<uint32_t> sgp4Orig;            // padded; padding lanes repeat the last real satellite (:146)
    std::vector<Sdp4Sat> sdp4;                 // one record per deep-space satellite
    std::vector<uint32_t> sdp4Orig;
    std::vector<double> epochs;                // per catalog satellite
    std::vector<int32_t> classes;              // 0 SGP4, 1 + irez for SDP4
    uint32_t sgp4Tiles_count() const { return (nSgp4 + kTileSats - 1) / kTileSats; }
    uint32_t sgp4Padded() const { return sgp4Tiles_count() * kTileSats; }
};

// Classify and tabulate parsed element sets.  Returns kOk or the first non-deep-space init failure
// (src/Constellation.zig:115-126).
inline int build_catalog_records(const TleRecord *recs, uint32_t n, int gravSel, CatalogTables &out) {
    out = CatalogTables{};
    out.n = n;
    out.grav = gravity(gravSel);
    out.epochs.resize(n);
    out.classes.resize(n);

    // Element initialisation is embarrassingly parallel and ~0.6 us per satellite on one core; large ingests
    // (Monte-Carlo draws, OMM streams) are cut into contiguous chunks, one host thread each, and concatenated in
    // chunk order so the catalog order -- and with it the reference epoch and every output row -- is unchanged.
    struct Chunk {
        std::vector<NearEarth> near;
        std::vector<uint32_t> nearIdx;
        std::vector<Sdp4Sat> deep;
        std::vector<uint32_t> deepIdx;
        int rc = kOk;
        uint32_t failAt = 0xffffffffu;
    };
    uint32_t nThreads = 1;
    if (n >= 20000) nThreads = std::min<uint32_t>(std::max(1u, std::thread::hardware_concurrency()), std::min<uint32_t>(64, n / 5000));
    std::vector<Chunk> chunks(nThreads);
    auto work = [&](uint32_t k) {
        Chunk &c = chunks[k];
        const uint32_t b = (uint32_t)((uint64_t)n * k / nThreads), e = (uint32_t)((uint64_t)n * (k + 1) / nThreads);
        c.near.reserve(e - b);
        c.nearIdx.reserve(e - b);
        for (uint32_t i = b; i < e; ++i) {
            const TleRecord &t = recs[i];
            out.epochs[i] = t.epochJd;
            NearEarth ne;
            int rc = build_near_earth(t, out.grav, ne);
            if (rc == kOk) {
                c.near.push_back(ne);
                c.nearIdx.push_back(i);
                out.classes[i] = 0;
            } else if (rc == kDeepSpace) {
                DeepSpace ds;
                rc = build_deep_space(t, out.grav, ds);
                if (rc != kOk) { c.rc = rc; c.failAt = i; return; }
                c.deep.push_back(sdp4_record(ds));
                c.deepIdx.push_back(i);
                out.classes[i] = 1 + ds.irez;
            } else {
                c.rc = rc;
                c.failAt = i;
                return;
            }
        }
    };
    if (nThreads == 1) {
        work(0);
    } else {
        std::vector<std::thread> pool;
        for (uint32_t k = 0; k < nThreads; ++k) pool.emplace_back(work, k);
        for (auto &th : pool) th.join();
    }
    for (const Chunk &c : chunks)  // the first failing element set (in catalog order) aborts, like the reference
        if (c.rc != kOk) return c.rc;
    std::vector<NearEarth> near;
    near.reserve(n);
    for (Chunk &c : chunks) {
        near.insert(near.end(), c.near.begin(), c.near.end());
        out.sgp4Orig.insert(out.sgp4Orig.end(), c.nearIdx.begin(), c.nearIdx.end());
        out.sdp4.insert(out.sdp4.end(), c.deep.begin(), c.deep.end());
        out.sdp4Orig.insert(out.sdp4Orig.end(), c.deepIdx.begin(), c.deepIdx.end());
        Chunk().near.swap(c.near);
    }
    out.nSgp4 = (uint32_t)near.size();
    out.nSdp4 = (uint32_t)out.sdp4.size();
    if (out.nSgp4) out.referenceEpochJd = near[0].epochJd;

    const uint32_t tiles = out.sgp4Tiles_count();
    out.sgp4Tiles.assign((size_t)tiles * kSgp4TileDoubles, 0.0);
    out.sgp4Epoch.resize((size_t)tiles * kTileSats);
    out.sgp4Orig.resize((size_t)tiles * kTileSats);
    double cols[kSgp4Cols];
    for (uint32_t s = 0; s < tiles * kTileSats; ++s) {
        const uint32_t src = s < out.nSgp4 ? s : out.nSgp4 - 1;
        sgp4_columns(near[src], cols);
        double *tile = out.sgp4Tiles.data() + (size_t)(s / kTileSats) * kSgp4TileDoubles;
        for (int c = 0; c < kSgp4Cols; ++c) tile[c * kTileSats + (s % kTileSats)] = cols[c];
        out.sgp4Epoch[s] = near[src].epochJd;
        if (s >= out.nSgp4) out.sgp4Orig[s] = out.sgp4Orig[out.nSgp4 - 1];
    }
    return kOk;
}

// Rows [r0, r1) of a classified catalog as a catalog of their own (one shard of a multi-device handle): the same
// element values, output rows renumbered from 0, and the WHOLE catalog's reference epoch, so tsince -- and with it
// every result bit -- is what the unsharded catalog computes.
inline void slice_catalog(const CatalogTables &full, uint32_t r0, uint32_t r1, CatalogTables &out) {
    out = CatalogTables{};
    out.n = r1 - r0;
    out.grav = full.grav;
    out.referenceEpochJd = full.referenceEpochJd;
    out.epochs.assign(full.epochs.begin() + r0, full.epochs.begin() + r1);
    out.classes.assign(full.classes.begin() + r0, full.classes.begin() + r1);
    // both index lists are ascending (catalog order), so a row range is an index range of each
    const uint32_t *so = full.sgp4Orig.data();
    const uint32_t a = (uint32_t)(std::lower_bound(so, so + full.nSgp4, r0) - so);
    const uint32_t b = (uint32_t)(std::lower_bound(so, so + full.nSgp4, r1) - so);
    const uint32_t *dorig = full.sdp4Orig.data();
    const uint32_t da = (uint32_t)(std::lower_bound(dorig, dorig + full.nSdp4, r0) - dorig);
    const uint32_t db = (uint32_t)(std::lower_bound(dorig, dorig + full.nSdp4, r1) - dorig);
    out.nSgp4 = b - a;
    out.nSdp4 = db - da;
    out.sdp4.assign(full.sdp4.begin() + da, full.sdp4.begin() + db);
    out.sdp4Orig.resize(out.nSdp4);
    for (uint32_t i = 0; i < out.nSdp4; ++i) out.sdp4Orig[i] = full.sdp4Orig[da + i] - r0;
    const uint32_t tiles = out.sgp4Tiles_count();
    out.sgp4Tiles.assign((size_t)tiles * kSgp4TileDoubles, 0.0);
    out.sgp4Epoch.resize((size_t)tiles * kTileSats);
    out.sgp4Orig.resize((size_t)tiles * kTileSats);
    for (uint32_t s = 0; s < tiles * kTileSats; ++s) {
        const uint32_t src = a + (s < out.nSgp4 ? s : out.nSgp4 - 1);  // padding repeats the last real satellite
        const double *from = full.sgp4Tiles.data() + (size_t)(src / kTileSats) * kSgp4TileDoubles + (src % kTileSats);
        double *to = out.sgp4Tiles.data() + (size_t)(s / kTileSats) * kSgp4TileDoubles + (s % kTileSats);
        for (int c = 0; c < kSgp4Cols; ++c) to[c * kTileSats] = from[c * kTileSats];
        out.sgp4Epoch[s] = full.sgp4Epoch[src];
        out.sgp4Orig[s] = full.sgp4Orig[src] - r0;
    }
}

// From TLE text lines (src/Tle.zig:49-101).
inline int build_catalog(const char *const *l1, const char *const *l2, uint32_t n, int gravSel, CatalogTables &out) {
    std::vector<TleRecord> recs(n);
    for (uint32_t i = 0; i < n; ++i)
        if (parse_tle(l1[i], l2[i], recs[i]) != kOk) return kBadTle;
    return build_catalog_records(recs.data(), n, gravSel, out);
}

// Split a text blob into element-set line pairs (src/Tle.zig:103-132 MultiIterator semantics).
inline void split_tle_text(const char *text, size_t len, std::vector<std::string> &l1, std::vector<std::string> &l2) {
    std::string cand;
    bool have = false;
    size_t i = 0;
    while (i <= len) {
        size_t j = i;
        while (j < len && text[j] != '\n' && text[j] != '\r') ++j;
        size_t b = i, e = j;
        while (b < e && (text[b] == ' ' || text[b] == '\t')) ++b;
        while (e > b && (text[e - 1] == ' ' || text[e - 1] == '\t')) --e;
        if (e - b >= 69) {
            if (text[b] == '1') {
                cand.assign(text + b, e - b);
                have = true;
            } else if (text[b] == '2') {
                if (have) {
                    l1.push_back(cand);
                    l2.emplace_back(text + b, e - b);
                    have = false;
                }
            } else {
                have = false;
            }
        }
        i = j + 1;
    }
}

}  // namespace az

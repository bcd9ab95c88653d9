// az_elements.hpp -- host-side element preparation for the device tables (product code, C++17).
//
// TLE text -> per-satellite propagation constants.  This is the cold O(n_sats) part of the path that
// stays on the host: src/Tle.zig:49-101, src/Datetime.zig:222-231,
// src/Sgp4.zig:108-417 (near earth) and src/Sdp4.zig:174-657 (deep space: gstime, dscom, dsinit).
// Written as templates over the arithmetic so a device-side builder (Monte-Carlo ingest) can reuse it.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>

// The element builders below are plain arithmetic over <cmath>; under nvcc they are also compiled for the
// device, where the ingest kernels (az_ingest.cu) run them one satellite per thread.
#ifdef __CUDACC__
#define AZ_EHD __host__ __device__
#else
#define AZ_EHD
#endif

namespace az {

// kernel-level status codes (src/simdKernels.zig:30-37)
enum Status : int { kOk = 0, kDecayed = 1, kInvalidEcc = 2, kDeepSpace = 3, kOom = 4, kBadTle = 5 };

struct Gravity {  // src/constants.zig:30-64
    double radiusEarthKm, mu, j2, j3, j4, xke, tumin, j3oj2;
};

AZ_EHD inline Gravity gravity(int which) {
    if (which == 1) return {6378.135, 398600.8, 0.001082616, -0.00000253881, -0.00000165597, 0.0743669161331734132,
                            13.44683969695931, -0.00234506972242078};
    return {6378.137, 398600.5, 0.00108262998905, -0.00000253215306, -0.00000161098761, 0.07436685316871385,
            13.446851082044981, -0.00233899967218727};
}

struct TleRecord {  // the fields of src/Tle.zig:8-29 the propagators read
    uint32_t satnum = 0;
    double epochJd = 0, bstar = 0, inclDeg = 0, raanDeg = 0, ecc = 0, argpDeg = 0, maDeg = 0, revPerDay = 0;
};

namespace detail {
constexpr double kHPi = 3.14159265358979323846264338327950288;
constexpr double kHTwoPi = 2.0 * kHPi;
constexpr double kDeg = kHPi / 180.0;

AZ_EHD inline double wrap(double x, double m) {  // Zig @mod: sign of the divisor
    double r = std::fmod(x, m);
    if (r != 0.0 && ((r < 0.0) != (m < 0.0))) r += m;
    return r;
}

struct Col {  // fixed-column field of a TLE line, blanks trimmed (src/Tle.zig:277-279)
    const char *p;
    size_t n;
    Col(const char *line, size_t b, size_t e) {
        while (b < e && line[b] == ' ') ++b;
        while (e > b && line[e - 1] == ' ') --e;
        p = line + b;
        n = e - b;
    }
    bool number(double &out) const {
        if (n == 0 || n > 31) return false;
        char buf[32];
        std::memcpy(buf, p, n);
        buf[n] = 0;
        char *end = nullptr;
        out = std::strtod(buf, &end);
        return end == buf + n;
    }
    bool integer(long &out) const {
        if (n == 0 || n > 31) return false;
        char buf[32];
        std::memcpy(buf, p, n);
        buf[n] = 0;
        char *end = nullptr;
        out = std::strtol(buf, &end, 10);
        return end == buf + n;
    }
};

inline size_t usable_length(const char *s) {
    size_t n = std::strlen(s);
    while (n && (s[n - 1] == ' ' || s[n - 1] == '\t' || s[n - 1] == '\r' || s[n - 1] == '\n')) --n;
    return n;
}
}  // namespace detail

// Jan-1 based day-of-year epoch -> single-f64 Julian date (src/Datetime.zig:222-231)
inline double epoch_to_jd(int fullYear, double doy) {
    const double a = std::floor(13.0 / 12.0);
    const double yy = double(fullYear) + 4800.0 - a;
    const double mm = 1.0 + 12.0 * a - 3.0;
    const double noonJan1 = 1.0 + std::floor((153.0 * mm + 2.0) / 5.0) + 365.0 * yy + std::floor(yy / 4.0) -
                            std::floor(yy / 100.0) + std::floor(yy / 400.0) - 32045.0;
    return noonJan1 + doy - 1.5;
}

// src/Tle.zig:49-101: fixed columns, no checksum, only len >= 69
inline int parse_tle(const char *l1, const char *l2, TleRecord &t) {
    using detail::Col;
    while (*l1 == ' ' || *l1 == '\t') ++l1;
    while (*l2 == ' ' || *l2 == '\t') ++l2;
    if (detail::usable_length(l1) < 69 || detail::usable_length(l2) < 69) return kBadTle;
    double mant, day, e7, dummy;
    long expo, yy;
    if (!Col(l1, 53, 59).number(mant) || !Col(l1, 59, 61).integer(expo)) return kBadTle;
    if (!Col(l1, 18, 20).integer(yy) || !Col(l1, 20, 32).number(day)) return kBadTle;
    if (!Col(l1, 33, 43).number(dummy)) return kBadTle;
    t.bstar = (mant * 1e-5) * std::pow(10.0, double(expo));
    t.epochJd = epoch_to_jd(yy < 57 ? 2000 + int(yy) : 1900 + int(yy), day);
    Col sn(l1, 2, 7);
    if (sn.n == 0) return kBadTle;
    {
        char buf[8] = {0};
        std::memcpy(buf, sn.p, sn.n < 7 ? sn.n : 7);
        if (buf[0] >= 'A' && buf[0] <= 'Z') t.satnum = uint32_t(buf[0] - 'A' + 10) * 10000u + uint32_t(std::strtoul(buf + 1, nullptr, 10));
        else t.satnum = uint32_t(std::strtoul(buf, nullptr, 10));
    }
    if (!Col(l2, 26, 33).number(e7)) return kBadTle;
    t.ecc = e7 / 1e7;
    if (!Col(l2, 8, 16).number(t.inclDeg) || !Col(l2, 17, 25).number(t.raanDeg) || !Col(l2, 34, 42).number(t.argpDeg) ||
        !Col(l2, 43, 51).number(t.maDeg) || !Col(l2, 52, 63).number(t.revPerDay))
        return kBadTle;
    return kOk;
}

// Everything the near-earth kernel needs for one satellite (+ what the deep-space builder reuses).
struct NearEarth {
    double epochJd;
    double no_kozai, ecco, inclo, nodeo, argpo, mo, bstar;
    double no, a;  // un-Kozai'd mean motion (rad/min), semi-major axis (ER)
    double sinio, cosio, cosio2, cosio4, con41, con42, x1mth2, x7thm1;
    double mdot, argpdot, nodedot;
    double cc1, cc4, cc5, t2cof, omgcof, xnodcf, xlcof, xmcof, aycof, eta, delmo, sinmao;
    double d2, d3, d4, t3cof, t4cof, t5cof;
    double aBase;
    bool isimp;
};

// Steps shared by SGP4 and SDP4 init (src/Sgp4.zig:192-382).  Returns kOk / kInvalidEcc / kDecayed,
// and reports the orbital period so the caller can classify (src/Sgp4.zig:120-123).
AZ_EHD inline int build_common(const TleRecord &t, const Gravity &g, NearEarth &o, double &periodMin, double &perigeeKm) {
    using namespace detail;
    o = NearEarth{};
    o.epochJd = t.epochJd;
    o.no_kozai = t.revPerDay * kHTwoPi / 1440.0;
    o.ecco = t.ecc;
    o.inclo = t.inclDeg * kDeg;
    o.nodeo = t.raanDeg * kDeg;
    o.argpo = t.argpDeg * kDeg;
    o.mo = t.maDeg * kDeg;
    o.bstar = t.bstar;
    if (o.ecco < 0.0 || o.ecco >= 1.0) return kInvalidEcc;

    // un-Kozai the mean motion (src/Sgp4.zig:206-228)
    const double ci = std::cos(o.inclo);
    const double th2 = ci * ci;
    const double x3thm1 = 3.0 * th2 - 1.0;
    const double beta2 = 1.0 - o.ecco * o.ecco;
    const double beta = std::sqrt(beta2);
    const double a1 = std::pow(g.xke / o.no_kozai, 2.0 / 3.0);
    const double del1 = 0.75 * g.j2 * x3thm1 / (a1 * a1 * beta * beta2);
    const double ao = a1 * (1.0 - del1 * (1.0 / 3.0 + del1 * (1.0 + 134.0 / 81.0 * del1)));
    const double delo = 0.75 * g.j2 * x3thm1 / (ao * ao * beta * beta2);
    o.no = o.no_kozai / (1.0 + delo);
    o.a = std::pow(g.xke / o.no, 2.0 / 3.0);
    if (o.a * (1.0 - o.ecco) < 1.0) return kDecayed;
    periodMin = kHTwoPi / o.no;

    o.sinio = std::sin(o.inclo);
    o.cosio = std::cos(o.inclo);
    o.cosio2 = o.cosio * o.cosio;
    o.cosio4 = o.cosio2 * o.cosio2;
    o.con41 = 3.0 * o.cosio2 - 1.0;
    o.con42 = 1.0 - 5.0 * o.cosio2;
    o.x1mth2 = 1.0 - o.cosio2;
    o.x7thm1 = 7.0 * o.cosio2 - 1.0;

    // secular rates (src/Sgp4.zig:253-284)
    const double omeosq = 1.0 - o.ecco * o.ecco;
    const double rteosq = std::sqrt(omeosq);
    const double pinvsq = 1.0 / std::pow(o.a * omeosq, 2.0);
    const double t1 = 1.5 * g.j2 * pinvsq * o.no;
    const double t2 = 0.5 * t1 * g.j2 * pinvsq;
    const double t3 = -0.46875 * g.j4 * pinvsq * pinvsq * o.no;
    o.mdot = o.no + 0.5 * t1 * rteosq * o.con41 + 0.0625 * t2 * rteosq * (13.0 - 78.0 * o.cosio2 + 137.0 * o.cosio4);
    o.argpdot = -0.5 * t1 * o.con42 + 0.0625 * t2 * (7.0 - 114.0 * o.cosio2 + 395.0 * o.cosio4) +
                t3 * (3.0 - 36.0 * o.cosio2 + 49.0 * o.cosio4);
    const double xhdot1 = -t1 * o.cosio;
    o.nodedot = xhdot1 + (0.5 * t2 * (4.0 - 19.0 * o.cosio2) + 2.0 * t3 * (3.0 - 7.0 * o.cosio2)) * o.cosio;

    // drag terms (src/Sgp4.zig:301-382)
    perigeeKm = (o.a * (1.0 - o.ecco) - 1.0) * g.radiusEarthKm;
    double s = 78.0;
    if (perigeeKm < 156.0) s = (perigeeKm < 98.0) ? 20.0 : perigeeKm - 78.0;
    const double qtemp = (120.0 - s) / g.radiusEarthKm;
    const double sfour = s / g.radiusEarthKm + 1.0;
    const double qzms24 = qtemp * qtemp * qtemp * qtemp;
    const double tsi = 1.0 / (o.a - sfour);
    o.eta = o.a * o.ecco * tsi;
    const double etasq = o.eta * o.eta;
    const double eeta = o.ecco * o.eta;
    const double psisq = std::fabs(1.0 - etasq);
    const double coef = qzms24 * std::pow(tsi, 4.0);
    const double coef1 = coef / std::pow(psisq, 3.5);
    const double cc2 = coef1 * o.no *
                       (o.a * (1.0 + 1.5 * etasq + eeta * (4.0 + etasq)) +
                        0.375 * g.j2 * tsi / psisq * o.con41 * (8.0 + 3.0 * etasq * (8.0 + etasq)));
    o.cc1 = o.bstar * cc2;
    const double cc3 = (o.ecco > 1.0e-4) ? -2.0 * coef * tsi * g.j3oj2 * o.no * o.sinio / o.ecco : 0.0;
    o.cc4 = 2.0 * o.no * coef1 * o.a * omeosq *
            (o.eta * (2.0 + 0.5 * etasq) + o.ecco * (0.5 + 2.0 * etasq) -
             g.j2 * tsi / (o.a * psisq) *
                 (-3.0 * o.con41 * (1.0 - 2.0 * eeta + etasq * (1.5 - 0.5 * eeta)) +
                  0.75 * o.x1mth2 * (2.0 * etasq - eeta * (1.0 + etasq)) * std::cos(2.0 * o.argpo)));
    o.cc5 = 2.0 * coef1 * o.a * omeosq * (1.0 + 2.75 * (etasq + eeta) + eeta * etasq);
    o.xnodcf = 3.5 * omeosq * xhdot1 * o.cc1;
    o.t2cof = 1.5 * o.cc1;
    const double xlnum = -0.25 * g.j3oj2 * o.sinio * (3.0 + 5.0 * o.cosio);
    o.xlcof = xlnum / ((std::fabs(o.cosio + 1.0) > 1.5e-12) ? 1.0 + o.cosio : 1.5e-12);
    o.aycof = -0.5 * g.j3oj2 * o.sinio;
    const double dm = 1.0 + o.eta * std::cos(o.mo);
    o.delmo = dm * dm * dm;
    o.sinmao = std::sin(o.mo);
    o.xmcof = (o.ecco > 1.0e-4) ? -(2.0 / 3.0) * coef * o.bstar / eeta : 0.0;
    o.omgcof = o.bstar * cc3 * std::cos(o.argpo);
    const double ratio = g.xke / o.no;
    o.aBase = std::cbrt(ratio * ratio);
    o.isimp = true;
    return kOk;
}

// src/Sgp4.zig:108-180: kOk, kInvalidEcc, kDecayed or kDeepSpace (period > 225 min)
AZ_EHD inline int build_near_earth(const TleRecord &t, const Gravity &g, NearEarth &o) {
    double period = 0, perigee = 0;
    int rc = build_common(t, g, o, period, perigee);
    if (rc != kOk) return rc;
    if (period > 225.0) return kDeepSpace;
    if (perigee >= 220.0) {  // higher-order drag, src/Sgp4.zig:394-417
        const double s = 78.0 / g.radiusEarthKm + 1.0;
        const double tsi = 1.0 / (o.a - s);
        const double c1sq = o.cc1 * o.cc1;
        o.d2 = 4.0 * o.a * tsi * c1sq;
        const double tmp = o.d2 * tsi * o.cc1 / 3.0;
        o.d3 = (17.0 * o.a + s) * tmp;
        o.d4 = 0.5 * tmp * o.a * tsi * (221.0 * o.a + 31.0 * s) * o.cc1;
        o.t3cof = o.d2 + 2.0 * c1sq;
        o.t4cof = 0.25 * (3.0 * o.d3 + o.cc1 * (12.0 * o.d2 + 10.0 * c1sq));
        o.t5cof = 0.2 * (3.0 * o.d4 + 12.0 * o.cc1 * o.d3 + 6.0 * o.d2 * o.d2 + 15.0 * c1sq * (2.0 * o.d2 + c1sq));
        o.isimp = false;
    }
    return kOk;
}

// ---- deep space -----------------------------------------------------------------------------------
struct LuniSolar {  // src/Sdp4.zig:54-67
    double e2, e3, i2, i3, l2, l3, l4, gh2, gh3, gh4, h2, h3;
};

struct DeepSpace {
    NearEarth ne;
    LuniSolar sun, moon;
    double zmol, zmos, dedt, didt, dmdt, domdt, dnodt;
    int irez;  // 0 none, 1 synchronous, 2 half-day
    double d2201, d2211, d3210, d3222, d4410, d4422, d5220, d5232, d5421, d5433;
    double del1, del2, del3;
    double xlamo, xfact, gsto;
};

AZ_EHD inline double gstime(double jdut1) {  // src/Sdp4.zig:277-285
    const double tu = (jdut1 - 2451545.0) / 36525.0;
    double sec = -6.2e-6 * tu * tu * tu + 0.093104 * tu * tu + (876600.0 * 3600.0 + 8640184.812866) * tu + 67310.54841;
    double th = detail::wrap(sec * detail::kDeg / 240.0, detail::kHTwoPi);
    if (th < 0.0) th += detail::kHTwoPi;
    return th;
}

namespace detail {
struct ThirdBody {  // one pass of the dscom loop (src/Sdp4.zig:391-436)
    double s1, s2, s3, s4, s5, s6, s7;
    double z1, z2, z3, z11, z12, z13, z21, z22, z23, z31, z32, z33;
};

AZ_EHD inline ThirdBody third_body(double zcosg, double zsing, double zcosi, double zsini, double zcosh, double zsinh,
                            double cc, double sinim, double cosim, double sinomm, double cosomm, double emsq,
                            double ecco, double rtemsq, double xnoi) {
    const double a1 = zcosg * zcosh + zsing * zcosi * zsinh;
    const double a3 = -zsing * zcosh + zcosg * zcosi * zsinh;
    const double a7 = -zcosg * zsinh + zsing * zcosi * zcosh;
    const double a8 = zsing * zsini;
    const double a9 = zsing * zsinh + zcosg * zcosi * zcosh;
    const double a10 = zcosg * zsini;
    const double a2 = cosim * a7 + sinim * a8;
    const double a4 = cosim * a9 + sinim * a10;
    const double a5 = -sinim * a7 + cosim * a8;
    const double a6 = -sinim * a9 + cosim * a10;
    const double x1 = a1 * cosomm + a2 * sinomm;
    const double x2 = a3 * cosomm + a4 * sinomm;
    const double x3 = -a1 * sinomm + a2 * cosomm;
    const double x4 = -a3 * sinomm + a4 * cosomm;
    const double x5 = a5 * sinomm;
    const double x6 = a6 * sinomm;
    const double x7 = a5 * cosomm;
    const double x8 = a6 * cosomm;
    ThirdBody b;
    b.z31 = 12.0 * x1 * x1 - 3.0 * x3 * x3;
    b.z32 = 24.0 * x1 * x2 - 6.0 * x3 * x4;
    b.z33 = 12.0 * x2 * x2 - 3.0 * x4 * x4;
    const double z1 = 3.0 * (a1 * a1 + a2 * a2) + b.z31 * emsq;
    const double z2 = 6.0 * (a1 * a3 + a2 * a4) + b.z32 * emsq;
    const double z3 = 3.0 * (a3 * a3 + a4 * a4) + b.z33 * emsq;
    b.z11 = -6.0 * a1 * a5 + emsq * (-24.0 * x1 * x7 - 6.0 * x3 * x5);
    b.z12 = -6.0 * (a1 * a6 + a3 * a5) + emsq * (-24.0 * (x2 * x7 + x1 * x8) - 6.0 * (x3 * x6 + x4 * x5));
    b.z13 = -6.0 * a3 * a6 + emsq * (-24.0 * x2 * x8 - 6.0 * x4 * x6);
    b.z21 = 6.0 * a2 * a5 + emsq * (24.0 * x1 * x5 - 6.0 * x3 * x7);
    b.z22 = 6.0 * (a4 * a5 + a2 * a6) + emsq * (24.0 * (x2 * x5 + x1 * x6) - 6.0 * (x4 * x7 + x3 * x8));
    b.z23 = 6.0 * a4 * a6 + emsq * (24.0 * x2 * x6 - 6.0 * x4 * x8);
    const double betasq = 1.0 - emsq;
    b.z1 = z1 + z1 + betasq * b.z31;
    b.z2 = z2 + z2 + betasq * b.z32;
    b.z3 = z3 + z3 + betasq * b.z33;
    b.s3 = cc * xnoi;
    b.s2 = -0.5 * b.s3 / rtemsq;
    b.s4 = b.s3 * rtemsq;
    b.s1 = -15.0 * ecco * b.s4;
    b.s5 = x1 * x3 + x2 * x4;
    b.s6 = x2 * x3 + x1 * x4;
    b.s7 = x2 * x4 - x1 * x3;
    return b;
}

AZ_EHD inline LuniSolar periodic_coeffs(const ThirdBody &b, double emsq, double ze) {  // src/Sdp4.zig:69-105
    LuniSolar p;
    p.e2 = 2.0 * b.s1 * b.s6;
    p.e3 = 2.0 * b.s1 * b.s7;
    p.i2 = 2.0 * b.s2 * b.z12;
    p.i3 = 2.0 * b.s2 * (b.z13 - b.z11);
    p.l2 = -2.0 * b.s3 * b.z2;
    p.l3 = -2.0 * b.s3 * (b.z3 - b.z1);
    p.l4 = -2.0 * b.s3 * (-21.0 - 9.0 * emsq) * ze;
    p.gh2 = 2.0 * b.s4 * b.z32;
    p.gh3 = 2.0 * b.s4 * (b.z33 - b.z31);
    p.gh4 = -18.0 * b.s4 * ze;
    p.h2 = -2.0 * b.s2 * b.z22;
    p.h3 = -2.0 * b.s2 * (b.z23 - b.z21);
    return p;
}

AZ_EHD inline double horner_up(double x, double c0, double c1, double c2, double c3 = 0.0) {
    // src/Sdp4.zig:671-679: ascending powers, term by term (a cubic, or a quadratic with c3 = 0)
    double acc = c0;
    double xn = x;
    acc += c1 * xn;
    xn *= x;
    acc += c2 * xn;
    xn *= x;
    acc += c3 * xn;
    return acc;
}
}  // namespace detail

// src/Sdp4.zig:174-274 (initElements), :344-499 (dscom), :525-657 (dsinit)
AZ_EHD inline int build_deep_space(const TleRecord &t, const Gravity &g, DeepSpace &d) {
    using namespace detail;
    constexpr double zes = 0.01675, zel = 0.05490, c1ss = 2.9864797e-6, c1l = 4.7968065e-7;
    constexpr double zsinis = 0.39785416, zcosis = 0.91744867, zcosgs = 0.1945905, zsings = -0.98088458;
    constexpr double zns = 1.19459e-5, znl = 1.5835218e-4;
    constexpr double q22 = 1.7891679e-6, q31 = 2.1460748e-6, q33 = 2.2123015e-7;
    constexpr double root22 = 1.7891679e-6, root32 = 3.7393792e-7, root44 = 7.3636953e-9, root52 = 1.1428639e-7,
                     root54 = 2.1765803e-9;
    constexpr double rptim = 4.37526908801129966e-3;
    constexpr double nearEq = 5.2359877e-2;

    d = DeepSpace{};
    double period = 0, perigee = 0;
    int rc = build_common(t, g, d.ne, period, perigee);
    if (rc != kOk) return rc;
    const NearEarth &e = d.ne;

    d.gsto = gstime(t.epochJd);
    const double day = t.epochJd - 2415020.0;

    // --- dscom
    const double snodm = std::sin(e.nodeo), cnodm = std::cos(e.nodeo);
    const double sinomm = std::sin(e.argpo), cosomm = std::cos(e.argpo);
    const double emsq = e.ecco * e.ecco;
    const double rtemsq = std::sqrt(1.0 - emsq);
    const double xnodce = wrap(4.5236020 - 9.2422029e-4 * day, kHTwoPi);
    const double stem = std::sin(xnodce), ctem = std::cos(xnodce);
    const double zcosil = 0.91375164 - 0.03568096 * ctem;
    const double zsinil = std::sqrt(1.0 - zcosil * zcosil);
    const double zsinhl = 0.089683511 * stem / zsinil;
    const double zcoshl = std::sqrt(1.0 - zsinhl * zsinhl);
    const double gam = 5.8351514 + 0.0019443680 * day;
    double zx = 0.39785416 * stem / zsinil;
    const double zy = zcoshl * ctem + 0.91744867 * zsinhl * stem;
    zx = std::atan2(zx, zy);
    zx += gam - xnodce;
    const double zcosgl = std::cos(zx), zsingl = std::sin(zx);
    const double xnoi = 1.0 / e.no;

    const ThirdBody S = third_body(zcosgs, zsings, zcosis, zsinis, cnodm, snodm, c1ss, e.sinio, e.cosio, sinomm, cosomm,
                                   emsq, e.ecco, rtemsq, xnoi);
    const ThirdBody L = third_body(zcosgl, zsingl, zcosil, zsinil, zcoshl * cnodm + zsinhl * snodm,
                                   snodm * zcoshl - cnodm * zsinhl, c1l, e.sinio, e.cosio, sinomm, cosomm, emsq, e.ecco,
                                   rtemsq, xnoi);
    d.sun = periodic_coeffs(S, emsq, zes);
    d.moon = periodic_coeffs(L, emsq, zel);
    d.zmol = wrap(4.7199672 + 0.22997150 * day - gam, kHTwoPi);
    d.zmos = wrap(6.2565837 + 0.017201977 * day, kHTwoPi);

    // --- dsinit: secular rates
    const double ses = S.s1 * zns * S.s5;
    const double sis = S.s2 * zns * (S.z11 + S.z13);
    const double sls = -zns * S.s3 * (S.z1 + S.z3 - 14.0 - 6.0 * emsq);
    const double sghs = S.s4 * zns * (S.z31 + S.z33 - 6.0);
    double shs = -zns * S.s2 * (S.z21 + S.z23);
    const bool equatorial = (e.inclo < nearEq) || (e.inclo > kHPi - nearEq);
    if (equatorial) shs = 0.0;
    if (e.sinio != 0.0) shs = shs / e.sinio;
    const double sgs = sghs - e.cosio * shs;
    d.dedt = ses + L.s1 * znl * L.s5;
    d.didt = sis + L.s2 * znl * (L.z11 + L.z13);
    d.dmdt = sls - znl * L.s3 * (L.z1 + L.z3 - 14.0 - 6.0 * emsq);
    const double sghl = L.s4 * znl * (L.z31 + L.z33 - 6.0);
    double shll = -znl * L.s2 * (L.z21 + L.z23);
    if (equatorial) shll = 0.0;
    d.domdt = sgs + sghl;
    d.dnodt = shs;
    if (e.sinio != 0.0) {
        d.domdt -= e.cosio / e.sinio * shll;
        d.dnodt += shll / e.sinio;
    }

    // --- resonance class and coefficients
    if (e.no >= 0.00826 && e.no <= 0.00924 && e.ecco >= 0.5) d.irez = 2;
    else if (e.no >= 0.0034906585 && e.no <= 0.0052359877) d.irez = 1;
    else d.irez = 0;

    const double eosq = e.ecco * e.ecco;
    const double sini2 = e.sinio * e.sinio;
    const double cosisq = e.cosio2;
    const double aonv = 1.0 / e.a;
    if (d.irez == 1) {
        const double g200 = 1.0 + eosq * (-2.5 + 0.8125 * eosq);
        const double g310 = 1.0 + 2.0 * eosq;
        const double g300 = 1.0 + eosq * (-6.0 + 6.60937 * eosq);
        const double f220 = 0.75 * (1.0 + e.cosio) * (1.0 + e.cosio);
        const double f311 = 0.9375 * sini2 * (1.0 + 3.0 * e.cosio) - 0.75 * (1.0 + e.cosio);
        double f330 = 1.0 + e.cosio;
        f330 = 1.875 * f330 * f330 * f330;
        const double tg = 3.0 * e.no * e.no * aonv * aonv;
        d.del2 = 2.0 * tg * f220 * g200 * q22;
        d.del3 = 3.0 * tg * f330 * g300 * q33 * aonv;
        d.del1 = tg * f311 * g310 * q31 * aonv;
        d.xlamo = wrap(e.mo + e.nodeo + e.argpo - d.gsto, kHTwoPi);
        d.xfact = e.mdot + (e.argpdot + e.nodedot) - rptim + d.dmdt + d.domdt + d.dnodt - e.no;
    } else if (d.irez == 2) {
        const double ec = e.ecco;
        const bool lo65 = ec <= 0.65, lo70 = ec < 0.7;
        const double g201 = -0.306 - (ec - 0.64) * 0.440;
        const double g211 = lo65 ? horner_up(ec, 3.616, -13.2470, 16.2900) : horner_up(ec, -72.099, 331.819, -508.738, 266.724);
        const double g310 = lo65 ? horner_up(ec, -19.302, 117.3900, -228.4190, 156.591) : horner_up(ec, -346.844, 1582.851, -2415.925, 1246.113);
        const double g322 = lo65 ? horner_up(ec, -18.9068, 109.7927, -214.6334, 146.5816) : horner_up(ec, -342.585, 1554.908, -2366.899, 1215.972);
        const double g410 = lo65 ? horner_up(ec, -41.122, 242.6940, -471.0940, 313.953) : horner_up(ec, -1052.797, 4758.686, -7193.992, 3651.957);
        const double g422 = lo65 ? horner_up(ec, -146.407, 841.8800, -1629.014, 1083.435) : horner_up(ec, -3581.690, 16178.110, -24462.770, 12422.520);
        double g520;
        if (lo65) g520 = horner_up(ec, -532.114, 3017.977, -5740.032, 3708.276);
        else if (ec > 0.715) g520 = horner_up(ec, -5149.66, 29936.92, -54087.36, 31324.56);
        else g520 = 1464.74 - 4664.75 * ec + 3763.64 * ec * ec;
        const double g521 = lo70 ? horner_up(ec, -822.71072, 4568.6173, -8491.4146, 5337.524) : horner_up(ec, -51752.104, 218913.95, -309468.16, 146349.42);
        const double g532 = lo70 ? horner_up(ec, -853.66600, 4690.2500, -8624.7700, 5341.400) : horner_up(ec, -40023.880, 170470.89, -242699.48, 115605.82);
        const double g533 = lo70 ? horner_up(ec, -919.22770, 4988.6100, -9064.7700, 5542.21) : horner_up(ec, -37995.780, 161616.52, -229838.20, 109377.94);
        const double ci = e.cosio, si = e.sinio;
        const double f220 = 0.75 * (1.0 + 2.0 * ci + cosisq);
        const double f221 = 1.5 * sini2;
        const double f321 = 1.875 * si * (1.0 - 2.0 * ci - 3.0 * cosisq);
        const double f322 = -1.875 * si * (1.0 + 2.0 * ci - 3.0 * cosisq);
        const double f441 = 35.0 * sini2 * f220;
        const double f442 = 39.3750 * sini2 * sini2;
        const double f522 = 9.84375 * si * (sini2 * (1.0 - 2.0 * ci - 5.0 * cosisq) + 0.33333333 * (-2.0 + 4.0 * ci + 6.0 * cosisq));
        const double f523 = si * (4.92187512 * sini2 * (-2.0 - 4.0 * ci + 10.0 * cosisq) + 6.56250012 * (1.0 + 2.0 * ci - 3.0 * cosisq));
        const double f542 = 29.53125 * si * (2.0 - 8.0 * ci + cosisq * (-12.0 + 8.0 * ci + 10.0 * cosisq));
        const double f543 = 29.53125 * si * (-2.0 - 8.0 * ci + cosisq * (12.0 + 8.0 * ci - 10.0 * cosisq));
        double t1 = 3.0 * e.no * e.no * aonv * aonv;
        double tt = t1 * root22;
        d.d2201 = tt * f220 * g201;
        d.d2211 = tt * f221 * g211;
        t1 = t1 * aonv;
        tt = t1 * root32;
        d.d3210 = tt * f321 * g310;
        d.d3222 = tt * f322 * g322;
        t1 = t1 * aonv;
        tt = 2.0 * t1 * root44;
        d.d4410 = tt * f441 * g410;
        d.d4422 = tt * f442 * g422;
        t1 = t1 * aonv;
        tt = t1 * root52;
        d.d5220 = tt * f522 * g520;
        d.d5232 = tt * f523 * g532;
        tt = 2.0 * t1 * root54;
        d.d5421 = tt * f542 * g521;
        d.d5433 = tt * f543 * g533;
        d.xlamo = wrap(e.mo + e.nodeo + e.nodeo - d.gsto - d.gsto, kHTwoPi);
        d.xfact = e.mdot + d.dmdt + 2.0 * (e.nodedot + d.dnodt - rptim) - e.no;
    }
    return kOk;
}

// GMST for the ECEF / geodetic output modes (src/WorldCoordinateSystem.zig:146-154)
AZ_EHD inline double julian_to_gmst(double jd) {
    const double d = jd - 2451545.0;
    const double t = d / 36525.0;
    double gmst = 280.46061837 + 360.98564736629 * d + 0.000387933 * t * t - t * t * t / 38710000.0;
    gmst = detail::wrap(gmst, 360.0);
    if (gmst < 0) gmst += 360.0;
    return gmst * detail::kDeg;
}

}  // namespace az

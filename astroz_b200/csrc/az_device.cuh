// az_device.cuh -- device-side data layout and the per-cell propagation cores (sm_100a, fp64).
//
// Replaces src/Sgp4Batch.zig (BatchElements :15-75, propagateBatchDirect :113-157) and the shared
// src/Sgp4.zig keplerAndPosVel (:646-750) of the reference.  One *cell* = one (satellite, epoch) pair.
#pragma once

#include "az_math.cuh"

namespace az {

// ---- HBM layout of the near-earth element table ----------------------------------------------------
// Satellite-major tiles: tile k holds satellites [8k, 8k+8) as a [kSgp4Cols][8] block of doubles
// (2,112 contiguous bytes): SoA inside the tile, so one cp.async.bulk (TMA) lands the whole tile in
// shared memory and every column chunk is a 64-byte, 128-bit-aligned run.  Only values the kernel
// reads are stored (the reference's 40-column BatchElements(8) carries 4 splatted constants and a
// host-only epoch column, src/Sgp4Batch.zig:21-24,71-73); bstar is folded into cc4/cc5 on the host.
constexpr int kTileSats = 8;
enum Sgp4Col : int {
    kMo, kMdot, kArgpo, kArgpdot, kNodeo, kNodedot, kXnodcf, kCc1, kBc4, kT2cof,
    kOmgcof, kEta, kXmcof, kDelmo, kD2, kD3, kD4, kBc5, kSinmao, kT3cof, kT4cof, kT5cof,
    kAbase, kEcco, kNo, kAycof, kXlcof, kCon41, kX1mth2, kX7thm1, kSinio, kCosio, kIsimp,
    kSgp4Cols
};
constexpr int kSgp4TileDoubles = kSgp4Cols * kTileSats;
constexpr int kSgp4TileBytes = kSgp4TileDoubles * 8;
static_assert(kSgp4TileBytes % 16 == 0, "TMA bulk copies move multiples of 16 bytes");

struct GravConsts {  // per-model scalars (kernel parameter -> constant bank)
    double j2, radiusEarthKm, vkmpersec, j3oj2, xke;
};

struct CellOut {
    double rx, ry, rz, vx, vy, vz;
    double mrt;  // radius in earth radii after short-period terms (decay diagnostic)
};

// ---- Kepler solve + short-period terms + orientation -> r, v  (src/Sgp4.zig:646-750) ---------------
// Differences from the reference's SIMD formulation, none of which changes the value beyond rounding:
//   * no mod-2pi anywhere: every angle goes straight into a range-reducing sincos;
//   * Newton iterates on the offset eps = E - u (|eps| <= e); sin/cos(E) come from rotating
//     sin/cos(u) by eps with the pi/4 kernels -- one full sincos for the whole solve;
//   * no atan2: (sinu, cosu) is already a unit vector, the J2 short-period angle is applied as a
//     rotation, and so is the inclination correction (sinio/cosio are per-satellite constants);
//   * sqrt(pl) = sqrt(am)*betal, 1/pl and am^-1.5 come from the two rsqrt seeds already needed.
AZ_HD void kepler_posvel(double am, double em, double mm, double argpm, double nodem,
                                              double sinio, double cosio, double aycof, double xlcof, double con41,
                                              double x1mth2, double x7thm1, const GravConsts &g, CellOut &o) {
    const double ya = rsqrt_nr(am);                        // am^-1/2
    const double inv_am = ya * ya;
    const double temp = inv_am * rcp(fma(-em, em, 1.0));  // 1 / (am (1 - em^2))
    double sa, ca;
    sincos_full(argpm, sa, ca);
    const double axnl = em * ca;
    const double aynl = fma(em, sa, temp * aycof);
    const double u = mm + argpm + temp * xlcof * axnl;  // xl - nodem, src/Sgp4.zig:680-682

    double s, c;
    sincos_full(u, s, c);
    double eps = 0.0;
#pragma unroll 1
    for (int it = 0; it < 10; ++it) {  // src/Sgp4.zig:687-694
        const double esine = fma(axnl, s, -(aynl * c));
        const double ecose = fma(axnl, c, aynl * s);
        double delta = (esine - eps) * rcp_fast(1.0 - ecose);
        delta = fmin(fmax(delta, -AZK(clamp)), AZK(clamp));
        eps += delta;
        // rotate (sin E, cos E) by the Newton step; a step below 0.05 rad uses the 9-op series
        double sd, cd;
        if (fabs(delta) <= AZK(tinyLimit)) sincos_tiny(delta, sd, cd);
        else sincos_full(delta, sd, cd);
        const double sn = fma(s, cd, c * sd);
        c = fma(c, cd, -(s * sd));
        s = sn;
        // Newton's residual after this step is ~ (e/2) delta^2 (f'' = e sin E): stop once that is below
        // 1e-15 rad -- tighter than the reference's |delta| < 1e-12 exit (src/Sgp4.zig:693)
        if (delta * delta * em < AZK(keplerTol)) break;
    }

    const double ecose = fma(axnl, c, aynl * s);
    const double esine = fma(axnl, s, -(aynl * c));
    const double omel2 = 1.0 - fma(axnl, axnl, aynl * aynl);
    const double yb = rsqrt_nr(omel2);
    const double betal = sqrt_from_rsqrt(omel2, yb);
    const double sqa = sqrt_from_rsqrt(am, ya);
    const double rl = am * (1.0 - ecose);
    const double irl = rcp(rl);
    const double rdotl = sqa * esine * irl;
    const double rvdotl = sqa * betal * irl;  // sqrt(pl) / rl
    const double aor = am * irl;
    const double est = esine * rcp(1.0 + betal);
    const double sinu = aor * (s - aynl - axnl * est);
    const double cosu = aor * (c - axnl + aynl * est);
    const double sin2u = 2.0 * sinu * cosu;
    const double cos2u = fma(-2.0 * sinu, sinu, 1.0);

    const double ipl = inv_am * (yb * yb);  // 1 / pl
    const double temp1 = 0.5 * g.j2 * ipl;
    const double temp2 = temp1 * ipl;
    const double w = inv_am * ya;  // nm / xke = am^-3/2

    const double mrt = fma(rl, fma(-1.5 * temp2 * betal, con41, 1.0), 0.5 * temp1 * x1mth2 * cos2u);
    const double dsu = -0.25 * temp2 * x7thm1 * sin2u;
    const double t2c = 1.5 * temp2 * cosio;
    const double dnode = t2c * sin2u;
    const double dinc = t2c * sinio * cos2u;
    const double wt1 = w * temp1;
    const double mvt = fma(-wt1 * x1mth2, sin2u, rdotl);
    const double rvdot = fma(wt1, fma(x1mth2, cos2u, 1.5 * con41), rvdotl);

    double sinsu, cossu, snod, cnod, sini, cosi;
    rotate_small(sinu, cosu, dsu, sinsu, cossu);
    sincos_full(nodem + dnode, snod, cnod);
    rotate_small(sinio, cosio, dinc, sini, cosi);

    const double xmx = -snod * cosi;
    const double xmy = cnod * cosi;
    const double ux = fma(xmx, sinsu, cnod * cossu);
    const double uy = fma(xmy, sinsu, snod * cossu);
    const double uz = sini * sinsu;
    const double vx = fma(xmx, cossu, -(cnod * sinsu));
    const double vy = fma(xmy, cossu, -(snod * sinsu));
    const double vz = sini * cossu;

    const double rs = mrt * g.radiusEarthKm;
    o.rx = rs * ux;
    o.ry = rs * uy;
    o.rz = rs * uz;
    o.vx = fma(mvt, ux, rvdot * vx) * g.vkmpersec;
    o.vy = fma(mvt, uy, rvdot * vy) * g.vkmpersec;
    o.vz = fma(mvt, uz, rvdot * vz) * g.vkmpersec;
    o.mrt = mrt;
}

// ---- near-earth secular + drag update, then the shared core (src/Sgp4Batch.zig:113-157) -----------
// `col(i)` returns column i of the satellite handled by this warp (shared-memory broadcast read).
template <typename ColFn>
AZ_HD void sgp4_cell(ColFn col, double t, const GravConsts &g, CellOut &o) {
    const double t2 = t * t;
    const double xmdf = fma(col(kMdot), t, col(kMo));
    const double argpdf = fma(col(kArgpdot), t, col(kArgpo));
    const double nodem = fma(col(kXnodcf), t2, fma(col(kNodedot), t, col(kNodeo)));
    double tempa = fma(-col(kCc1), t, 1.0);
    double tempe = col(kBc4) * t;
    double templ = col(kT2cof) * t2;
    double mm = xmdf, argpm = argpdf;

    if (col(kIsimp) == 0.0) {  // warp-uniform: a warp works on one satellite (src/Sgp4Batch.zig:133-145)
        double sm, cm;
        sincos_full(xmdf, sm, cm);
        const double dm = fma(col(kEta), cm, 1.0);
        const double delm = col(kXmcof) * (dm * dm * dm - col(kDelmo));
        const double tho = fma(col(kOmgcof), t, delm);
        mm = xmdf + tho;
        argpm = argpdf - tho;
        // sin(mm) = sin(xmdf + tho): tho is a drag-sized angle, rotate instead of a second reduction
        double sd, cd;
        if (fabs(tho) <= AZK(quarterLimit)) sincos_quarter(tho, sd, cd);
        else sincos_full(tho, sd, cd);
        const double sinmm = fma(sm, cd, cm * sd);
        const double t3 = t2 * t;
        const double t4 = t3 * t;
        tempa = tempa - col(kD2) * t2 - col(kD3) * t3 - col(kD4) * t4;
        tempe = fma(col(kBc5), sinmm - col(kSinmao), tempe);
        templ = templ + col(kT3cof) * t3 + t4 * fma(t, col(kT5cof), col(kT4cof));
    }

    const double am = col(kAbase) * tempa * tempa;
    const double em = fmax(col(kEcco) - tempe, AZK(emFloor));
    mm = fma(col(kNo), templ, mm);

    kepler_posvel(am, em, mm, argpm, nodem, col(kSinio), col(kCosio), col(kAycof), col(kXlcof), col(kCon41),
                  col(kX1mth2), col(kX7thm1), g, o);
}

// ---- deep space (src/Sdp4Batch.zig:16-125,199-526; src/Sdp4.zig:681-866) --------------------------------
// One record per deep-space satellite, read with warp-uniform loads (a warp works on one satellite).
struct Sdp4Sat {
    double mo, mdot, argpo, argpdot, nodeo, nodedot, xnodcf, cc1, bc4, t2cof, ecco, no, inclo;
    double se2, se3, si2, si3, sl2, sl3, sl4, sgh2, sgh3, sgh4, sh2, sh3;  // solar periodics
    double ee2, e3, xi2, xi3, xl2, xl3, xl4, xgh2, xgh3, xgh4, xh2, xh3;  // lunar periodics
    double zmol, zmos, dedt, didt, dmdt, domdt, dnodt;
    double d2201, d2211, d3210, d3222, d4410, d4422, d5220, d5232, d5421, d5433;
    double del1, del2, del3, xlamo, xfact, gsto;
    double epochJd;
    int irez, pad_;
};

constexpr double kStepp = 720.0;      // src/Sdp4.zig:51-52
constexpr double kStep2 = 259200.0;
constexpr double kRptim = 4.37526908801129966e-3;

// resonance accelerations at (xli, xni, atime)  (src/Sdp4.zig:824-866)
AZ_HD void resonance_accel(const Sdp4Sat &e, double xli, double xni, double atime, double &xndt, double &xnddt,
                           double &xldot) {
    xldot = xni + e.xfact;
    if (e.irez == 2) {
        constexpr double g22 = 5.7686396, g32 = 0.95240898, g44 = 1.8014998, g52 = 1.0508330, g54 = 4.4108898;
        const double xomi = fma(e.argpdot, atime, e.argpo);
        const double x2omi = xomi + xomi;
        const double x2li = xli + xli;
        double s1, c1, s2, c2, s3, c3, s4, c4, s5, c5, s6, c6, s7, c7, s8, c8, s9, c9, s10, c10;
        sincos_full(x2omi + xli - g22, s1, c1);
        sincos_full(xli - g22, s2, c2);
        sincos_full(xomi + xli - g32, s3, c3);
        sincos_full(-xomi + xli - g32, s4, c4);
        sincos_full(x2omi + x2li - g44, s5, c5);
        sincos_full(x2li - g44, s6, c6);
        sincos_full(xomi + xli - g52, s7, c7);
        sincos_full(-xomi + xli - g52, s8, c8);
        sincos_full(xomi + x2li - g54, s9, c9);
        sincos_full(-xomi + x2li - g54, s10, c10);
        xndt = e.d2201 * s1 + e.d2211 * s2 + e.d3210 * s3 + e.d3222 * s4 + e.d4410 * s5 + e.d4422 * s6 +
               e.d5220 * s7 + e.d5232 * s8 + e.d5421 * s9 + e.d5433 * s10;
        xnddt = (e.d2201 * c1 + e.d2211 * c2 + e.d3210 * c3 + e.d3222 * c4 + e.d5220 * c7 + e.d5232 * c8 +
                 2.0 * (e.d4410 * c5 + e.d4422 * c6 + e.d5421 * c9 + e.d5433 * c10)) *
                xldot;
    } else {
        constexpr double fasx2 = 0.13130908, fasx4 = 2.8843198, fasx6 = 0.37448087;
        double s1, c1, s2, c2, s3, c3;
        sincos_full(xli - fasx2, s1, c1);
        sincos_full(2.0 * (xli - fasx4), s2, c2);
        sincos_full(3.0 * (xli - fasx6), s3, c3);
        xndt = e.del1 * s1 + e.del2 * s2 + e.del3 * s3;
        xnddt = (e.del1 * c1 + 2.0 * e.del2 * c2 + 3.0 * e.del3 * c3) * xldot;
    }
}

// one 720-minute Euler-Maclaurin step of the resonance integrator (src/Sdp4.zig:796-801)
AZ_HD void resonance_step(const Sdp4Sat &e, double &xli, double &xni, double &atime, double delt) {
    double xndt, xnddt, xldot;
    resonance_accel(e, xli, xni, atime, xndt, xnddt, xldot);
    xli += xldot * delt + xndt * kStep2;
    xni += xndt * delt + xnddt * kStep2;
    atime += delt;
}

// number of whole 720-minute steps the reference's loop `while |t - atime| >= 720` takes from atime = 0
AZ_HD int resonance_node(double t) {
    const double a = fabs(t);
    int n = (int)floor(a / kStepp);
    if ((double)n * kStepp > a) --n;
    if (a - (double)n * kStepp >= kStepp) ++n;
    return n;
}

// One deep-space cell.  (xli, xni) is the integrator state at lattice node atime = +-720*n nearest
// below |t| (unused when irez == 0).  Returns a kernel-level status (0 ok, 1 decayed, 2 invalid ecc)
// with the scalar path's checks (src/Sdp4.zig:913-967).
AZ_HD int sdp4_cell(const Sdp4Sat &e, double t, double xli, double xni, double atime, const GravConsts &g,
                    CellOut &o) {
    constexpr double zns = 1.19459e-5, znl = 1.5835218e-4, zes = 0.01675, zel = 0.05490;
    const double t2 = t * t;
    const double tempa = fma(-e.cc1, t, 1.0);
    const double tempe = e.bc4 * t;
    const double templ = e.t2cof * t2;

    // secular gravity + drag, then luni-solar secular rates (src/Sdp4Batch.zig:212-236)
    double mm = fma(e.mdot, t, e.mo);
    double argpm = fma(e.argpdot, t, e.argpo);
    double nodem = fma(e.xnodcf, t2, fma(e.nodedot, t, e.nodeo));
    double em = fma(e.dedt, t, e.ecco);
    double inclm = fma(e.didt, t, e.inclo);
    argpm = fma(e.domdt, t, argpm);
    nodem = fma(e.dnodt, t, nodem);
    mm = fma(e.dmdt, t, mm);
    double nm = e.no;

    if (e.irez != 0) {  // final partial step from the lattice node (src/Sdp4.zig:803-819)
        const double ft = t - atime;
        double xndt, xnddt, xldot;
        resonance_accel(e, xli, xni, atime, xndt, xnddt, xldot);
        const double nmr = xni + xndt * ft + xnddt * ft * ft * 0.5;
        const double xl = xli + xldot * ft + xndt * ft * ft * 0.5;
        const double theta = mod_twopi(fma(t, kRptim, e.gsto));
        mm = (e.irez == 2) ? xl - 2.0 * nodem + 2.0 * theta : xl - nodem - argpm + theta;
        nm = e.no + (nmr - e.no);
    }

    if (nm <= 0.0) return 1;
    const double cr = cbrt(g.xke / nm);
    const double am = cr * cr * tempa * tempa;  // (xke/nm)^(2/3) * tempa^2
    em -= tempe;
    if (em >= 1.0 || em < -0.001) return 2;
    em = fmax(em, 1.0e-6);
    if (am < 0.95) return 1;
    mm = fma(e.no, templ, mm);

    // luni-solar periodics, dpper (src/Sdp4.zig:681-759)
    double sz, cz, sinzf, coszf;
    double zm = fma(zns, t, e.zmos);
    sincos_full(zm, sz, cz);
    sincos_full(fma(2.0 * zes, sz, zm), sinzf, coszf);
    double f2 = fma(0.5 * sinzf, sinzf, -0.25);
    double f3 = -0.5 * sinzf * coszf;
    double pe = e.se2 * f2 + e.se3 * f3;
    double pinc = e.si2 * f2 + e.si3 * f3;
    double pl = e.sl2 * f2 + e.sl3 * f3 + e.sl4 * sinzf;
    double pgh = e.sgh2 * f2 + e.sgh3 * f3 + e.sgh4 * sinzf;
    double ph = e.sh2 * f2 + e.sh3 * f3;
    zm = fma(znl, t, e.zmol);
    sincos_full(zm, sz, cz);
    sincos_full(fma(2.0 * zel, sz, zm), sinzf, coszf);
    f2 = fma(0.5 * sinzf, sinzf, -0.25);
    f3 = -0.5 * sinzf * coszf;
    pe += e.ee2 * f2 + e.e3 * f3;
    pinc += e.xi2 * f2 + e.xi3 * f3;
    pl += e.xl2 * f2 + e.xl3 * f3 + e.xl4 * sinzf;
    pgh += e.xgh2 * f2 + e.xgh3 * f3 + e.xgh4 * sinzf;
    ph += e.xh2 * f2 + e.xh3 * f3;

    inclm += pinc;
    em += pe;
    double sinip, cosip;
    sincos_full(inclm, sinip, cosip);
    if (inclm >= 0.2) {
        ph = ph * rcp(sinip);
        pgh = fma(-cosip, ph, pgh);
        argpm += pgh;
        nodem += ph;
        mm += pl;
    } else {  // Lyddane modification for near-equatorial orbits (src/Sdp4.zig:735-758)
        nodem = mod_twopi(nodem);
        double sinop, cosop;
        sincos_full(nodem, sinop, cosop);
        const double alfdp = sinip * sinop + (ph * cosop + pinc * cosip * sinop);
        const double betdp = sinip * cosop + (-ph * sinop + pinc * cosip * cosop);
        const double xls = mm + argpm + cosip * nodem;
        const double dls = pl + pgh - pinc * nodem * sinip;
        const double xnoh = nodem;
        nodem = atan2(alfdp, betdp);
        if (fabs(xnoh - nodem) > kPi) nodem += (nodem < xnoh) ? kTwoPi : -kTwoPi;
        mm += pl;
        argpm = xls + dls - mm - cosip * nodem;
    }
    if (inclm < 0.0) {  // src/Sdp4.zig:932-936 (sin flips sign with the inclination, cos does not)
        sinip = -sinip;
        nodem += kPi;
        argpm -= kPi;
    }
    em = fmax(em, 1.0e-6);
    if (em >= 1.0) return 2;

    // inclination-dependent terms re-derived per cell (src/Sdp4Batch.zig:326-339)
    const double cosip2 = cosip * cosip;
    const double aycof = -0.5 * g.j3oj2 * sinip;
    const double den = 1.0 + cosip;
    const double xlcof = -0.25 * g.j3oj2 * sinip * fma(5.0, cosip, 3.0) * rcp(fabs(den) > 1.5e-12 ? den : 1.5e-12);
    kepler_posvel(am, em, mm, argpm, nodem, sinip, cosip, aycof, xlcof, fma(3.0, cosip2, -1.0), 1.0 - cosip2,
                  fma(7.0, cosip2, -1.0), g, o);
    return (o.mrt < 1.0) ? 1 : 0;
}

// ---- output-mode epilogue (src/Constellation.zig:478-509, src/WorldCoordinateSystem.zig:98-121) ----
AZ_HD void eci_to_ecef(double &x, double &y, double sinG, double cosG) {
    const double ex = fma(x, cosG, y * sinG);
    const double ey = fma(y, cosG, -(x * sinG));
    x = ex;
    y = ey;
}

AZ_HD void ecef_to_geodetic(double &x, double &y, double &z) {
    constexpr double a = 6378.137;
    constexpr double f = 1.0 / 298.257223563;
    constexpr double e2 = 2.0 * f - f * f;
    const double lon = atan2(y, x);
    const double p = sqrt(fma(x, x, y * y));
    double lat = atan2(z, p * (1.0 - e2));
#pragma unroll 1
    for (int i = 0; i < 10; ++i) {
        const double prev = lat;
        const double sl = sin(lat);
        const double N = a / sqrt(fma(-e2 * sl, sl, 1.0));
        lat = atan2(fma(e2 * N, sl, z), p);
        if (fabs(lat - prev) < 1e-12) break;
    }
    const double sl = sin(lat), cl = cos(lat);
    const double N = a / sqrt(fma(-e2 * sl, sl, 1.0));
    x = lat;
    y = lon;
    z = p / cl - N;
}

}  // namespace az
